"""Image-parallel execution across the GPUs of one box.

Images are independent units of the path (no cross-image op in the forward; the frequency-bias
lookup is per image, `/root/reference/model/egtr.py:408-411`), so a batch is split into contiguous
image ranges, one per rank, weights are replicated, and the only exchange is ONE all-gather of
fixed-size per-image result records at the end of the batch.  The reference has no counterpart:
its evaluation is single-GPU (`/root/reference/train_egtr.py:897-899`).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.distributed as dist

RECORD_FIELDS = ("logits", "pred_boxes", "pred_rel", "pred_connectivity")


def shard_range(n_images: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) image range of `rank`; the first `n % world` ranks take one extra image."""
    base, extra = divmod(n_images, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def record_layout(num_queries: int, num_labels: int, num_rel_labels: int) -> Dict[str, Tuple[int, Tuple[int, ...]]]:
    """field -> (offset in floats, per-image shape) of the flat fp32 per-image record."""
    N, K, P = num_queries, num_labels, num_rel_labels
    shapes = {"logits": (N, K), "pred_boxes": (N, 4), "pred_rel": (N, N, P), "pred_connectivity": (N, N, 1)}
    out, off = {}, 0
    for f in RECORD_FIELDS:
        n = 1
        for s in shapes[f]:
            n *= s
        out[f] = (off, shapes[f])
        off += n
    out["_size"] = (off, ())
    return out


def pack_records(outputs, layout) -> torch.Tensor:
    """[B_local, record] fp32 — one row per image."""
    B = outputs["logits"].shape[0]
    return torch.cat([outputs[f].reshape(B, -1) for f in RECORD_FIELDS], dim=1).contiguous()


def unpack_records(flat: torch.Tensor, layout) -> Dict[str, torch.Tensor]:
    out = {}
    for f in RECORD_FIELDS:
        off, shp = layout[f]
        n = 1
        for s in shp:
            n *= s
        out[f] = flat[:, off:off + n].reshape(flat.shape[0], *shp)
    return out


def all_gather_records(local: torch.Tensor, per_rank: int) -> torch.Tensor:
    """One all-gather of the records of every rank -> [world * per_rank, record] in image order.
    `local` may hold fewer than `per_rank` rows on the last ranks (ragged shards): it is zero-padded
    so the collective stays fixed-size, and the caller trims with `shard_range`."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if local.shape[0] < per_rank:
        pad = local.new_zeros(per_rank - local.shape[0], local.shape[1])
        local = torch.cat([local, pad], 0)
    out = local.new_empty(world * per_rank, local.shape[1])
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(out, local.contiguous())
    else:  # gloo (CPU tests)
        parts: List[torch.Tensor] = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local.contiguous())
        out = torch.cat(parts, 0)
    return out


def gather_batch(local_outputs, n_images: int, layout) -> Dict[str, torch.Tensor]:
    """Every rank ends with the results of all `n_images` images, in order."""
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    per_rank = max(shard_range(n_images, world, r)[1] - shard_range(n_images, world, r)[0] for r in range(world))
    flat = all_gather_records(pack_records(local_outputs, layout), per_rank)
    if world > 1:
        keep = []
        for r in range(world):
            lo, hi = shard_range(n_images, world, r)
            keep.append(flat[r * per_rank: r * per_rank + (hi - lo)])
        flat = torch.cat(keep, 0)
    return unpack_records(flat, layout)
