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


def _triplet_worker(rank, world, port, q):
    """Each rank fills the flat triplet record of its own images; ONE all-gather; every rank decodes all images in order."""
    from egtr_b200.postprocess import TripletLayout
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, N, P, k = 3, 5, 4, 7
    ok = True
    for single in (False, True):
        lay = TripletLayout(B, N, P, topk=k, single=single)
        g = torch.Generator().manual_seed(77)
        full = dict(pred_boxes=torch.rand(world * B, N, 4, generator=g), obj_scores=torch.rand(world * B, N, generator=g),
                    pred_classes=torch.randint(0, 150, (world * B, N), generator=g, dtype=torch.int32),
                    pred_rel_inds=torch.randint(0, N, (world * B, k, 2 if single else 3), generator=g, dtype=torch.int32),
                    rel_scores=torch.rand(*((world * B, k, P) if single else (world * B, k)), generator=g))
        flat = torch.zeros(lay.words, dtype=torch.int32)
        for name, v in full.items():
            lay.view(flat, name).copy_(v[rank * B:(rank + 1) * B])
        parts = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(parts, flat)
        got = lay.decode(torch.stack(parts, 0))
        ok = ok and all(torch.equal(got[kk], full[kk]) for kk in full)
        ok = ok and all(torch.equal(lay.view(flat, kk), full[kk][rank * B:(rank + 1) * B]) for kk in full)
    q.put((rank, ok))
    dist.destroy_process_group()


def test_all_gather_of_triplet_records_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_triplet_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(60) for p in procs]
    assert sorted(res) == [(0, True), (1, True)], res
