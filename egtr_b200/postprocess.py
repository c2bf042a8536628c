"""Device-side triplet extraction: the step that follows the model in the reference's evaluation loop
(`/root/reference/train_egtr.py:43-173`), SURVEY.md §8f row 1.

`extract_triplets(outputs, num_labels, single=..., topk=100)` returns, per image, what `evaluate_batch` builds
on the CPU before handing it to the evaluators: `obj_scores`, `pred_classes`, `pred_rel_inds`, `rel_scores`
— without moving the N x N x P relation tensor to the host or sorting all of it.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import _lib


def extract_triplets(outputs, num_labels: int, single: bool = False, topk: int = 100) -> Dict[str, torch.Tensor]:
    logits, rel = outputs["logits"], outputs["pred_rel"]
    conn = outputs["pred_connectivity"] if "pred_connectivity" in outputs else None
    if not logits.is_cuda:
        raise _lib.EgtrError("extract_triplets runs on CUDA tensors only (no CPU fallback)")
    B, N, K = logits.shape
    P = rel.shape[-1]
    dev = logits.device
    logits, rel = logits.contiguous().float(), rel.contiguous().float()
    conn = conn.contiguous().float() if conn is not None else None
    with torch.cuda.device(dev):
        nbytes = int(_lib.call("egtr_triplets_scratch_bytes", B, N, P, int(single), topk))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        obj = torch.empty(B, N, dtype=torch.float32, device=dev)
        cls = torch.empty(B, N, dtype=torch.int32, device=dev)
        inds = torch.empty(B, topk, 2 if single else 3, dtype=torch.int32, device=dev)
        scores = torch.empty((B, topk, P) if single else (B, topk), dtype=torch.float32, device=dev)
        _lib.call("egtr_triplets_f32", logits.data_ptr(), rel.data_ptr(), conn.data_ptr() if conn is not None else None, B, N, K,
                  num_labels, P, int(single), topk, scratch.data_ptr(), obj.data_ptr(), cls.data_ptr(), inds.data_ptr(),
                  scores.data_ptr(), torch.cuda.current_stream().cuda_stream)
    return dict(obj_scores=obj, pred_classes=cls, pred_rel_inds=inds, rel_scores=scores)


class TripletLayout:
    """Flat per-rank record of what `evaluate_batch` (`/root/reference/train_egtr.py:56-94, 120-128`) needs of one forward —
    boxes, object scores / classes, the top-k (s, o, p) indices and their scores — as 4-byte words:

        [ pred_boxes B*N*4 f32 | obj_scores B*N f32 | pred_classes B*N i32 | pred_rel_inds B*k*W i32 | rel_scores B*k*(1|P) f32 ]

    Field-major inside a rank; a gathered [world, words] buffer decodes to image order because ranks hold contiguous image
    ranges.  ~6.4 KB per image at N = 200, k = 100 instead of 8.3 MB of raw logits / pred_rel / pred_connectivity."""

    def __init__(self, B: int, N: int, P: int, topk: int = 100, single: bool = False):
        self.B, self.N, self.P, self.topk, self.single = B, N, P, topk, single
        W = 2 if single else 3
        sizes = [("pred_boxes", B * N * 4, torch.float32, (B, N, 4)), ("obj_scores", B * N, torch.float32, (B, N)),
                 ("pred_classes", B * N, torch.int32, (B, N)), ("pred_rel_inds", B * topk * W, torch.int32, (B, topk, W)),
                 ("rel_scores", B * topk * (P if single else 1), torch.float32, (B, topk, P) if single else (B, topk))]
        self.fields, off = {}, 0
        for name, n, dt, shp in sizes:
            self.fields[name] = (off, n, dt, shp)
            off += (n + 3) // 4 * 4  # 16-byte aligned fields
        self.words = off

    def view(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        """Field `name` of a flat int32 record buffer [words], or of a gathered one [world, words] (-> [world*B, ...])."""
        off, n, dt, shp = self.fields[name]
        if flat.dim() == 1:
            return flat[off:off + n].view(dt).view(*shp)
        return flat[:, off:off + n].contiguous().view(dt).view(flat.shape[0] * shp[0], *shp[1:])

    def decode(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        return {k: self.view(flat, k) for k in self.fields}


class TripletRecords(TripletLayout):
    """Static-buffer triplet extraction (CUDA-graph capturable): `enqueue(outputs)` fills `self.flat` on the current stream."""

    def __init__(self, B: int, N: int, K: int, P: int, num_labels: int, device, topk: int = 100, single: bool = False):
        super().__init__(B, N, P, topk, single)
        self.K, self.num_labels = K, num_labels
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.EgtrError("triplet extraction runs on CUDA devices only (no CPU fallback)")
        with torch.cuda.device(self.device):
            self.flat = torch.zeros(self.words, dtype=torch.int32, device=self.device)
            nbytes = int(_lib.call("egtr_triplets_scratch_bytes", B, N, P, int(single), topk))
            self.scratch = torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    def enqueue(self, outputs) -> torch.Tensor:
        """Launch the extraction of `outputs` (device tensors of one forward) into `self.flat` on the current stream."""
        logits, rel, conn = outputs["logits"], outputs["pred_rel"], outputs["pred_connectivity"]
        if not logits.is_cuda:
            raise _lib.EgtrError("triplet extraction runs on CUDA tensors only (no CPU fallback)")
        v = lambda k: self.view(self.flat, k)  # noqa: E731
        v("pred_boxes").copy_(outputs["pred_boxes"])
        _lib.call("egtr_triplets_f32", logits.data_ptr(), rel.data_ptr(), conn.data_ptr(), self.B, self.N, self.K, self.num_labels,
                  self.P, int(self.single), self.topk, self.scratch.data_ptr(), v("obj_scores").data_ptr(), v("pred_classes").data_ptr(),
                  v("pred_rel_inds").data_ptr(), v("rel_scores").data_ptr(), torch.cuda.current_stream().cuda_stream)
        return self.flat
