"""CPU, world_size 2 over gloo: the image-parallel shard/gather logic of egtr_b200.parallel."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from egtr_b200.parallel import gather_batch, record_layout, shard_range


def test_shard_range_covers_batch_contiguously():
    for n in (1, 2, 5, 32, 33):
        for w in (1, 2, 4, 8):
            rs = [shard_range(n, w, r) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in rs) - min(h - l for l, h in rs) <= 1


def _worker(rank, world, port, n_images, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N, K, P = 5, 7, 3
    layout = record_layout(N, K, P)
    lo, hi = shard_range(n_images, world, rank)
    g = torch.Generator().manual_seed(123)
    full = dict(logits=torch.randn(n_images, N, K, generator=g), pred_boxes=torch.rand(n_images, N, 4, generator=g),
                pred_rel=torch.rand(n_images, N, N, P, generator=g), pred_connectivity=torch.rand(n_images, N, N, 1, generator=g))
    local = {k: v[lo:hi] for k, v in full.items()}
    got = gather_batch(local, n_images, layout)
    ok = all(torch.equal(got[k], full[k]) for k in full)
    q.put((rank, ok))
    dist.destroy_process_group()


def test_all_gather_of_per_image_records_world2():
    ctx = mp.get_context("spawn")
    for n_images in (4, 5):  # even and ragged shards
        q = ctx.Queue()
        port = 29500 + os.getpid() % 2000 + n_images
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_images, q)) for r in range(2)]
        [p.start() for p in procs]
        res = [q.get(timeout=120) for _ in procs]
        [p.join(60) for p in procs]
        assert sorted(res) == [(0, True), (1, True)], res
