"""Golden vectors for the triplet extraction, produced with the reference's own `argsort_desc`
(`/root/reference/lib/pytorch_misc.py:27-34`, imported unmodified).  Run in the build container only."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from lib.pytorch_misc import argsort_desc  # noqa: E402  (the reference's)

from oracle.postprocess_oracle import extract  # noqa: E402


def make(name, B, N, K, P, seed):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, N, K, generator=g) * 2
    rel = torch.rand(B, N, N, P, generator=g) ** 3
    conn = torch.rand(B, N, N, 1, generator=g)
    save = {}
    for single in (False, True):
        mine = extract(logits, rel, conn, K, single)
        for j in range(B):
            obj, cls = torch.max(logits[j].softmax(-1)[:, :K], -1)
            so = torch.outer(obj, obj)
            so[torch.arange(N), torch.arange(N)] = 0.0
            r = torch.clamp(rel[j], 0, 1) * torch.clamp(conn[j], 0, 1)
            sc = (r.max(-1)[0] * so) if single else (r * so.unsqueeze(-1))
            ref_inds = argsort_desc(sc.numpy())[:100]
            # scores are distinct with probability one, so the reference's unstable sort and the stable oracle agree
            assert np.array_equal(ref_inds, mine[j]["pred_rel_inds"]), (name, single, j)
            tag = f"{'single' if single else 'multi'}_{j}"
            save[f"inds_{tag}"] = ref_inds.astype(np.int32)
            save[f"relscores_{tag}"] = mine[j]["rel_scores"].astype(np.float32)
            save[f"obj_{j}"] = obj.numpy()
            save[f"cls_{j}"] = cls.numpy().astype(np.int32)
    meta = dict(name=name, B=B, N=N, K=K, P=P, seed=seed)
    save["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, f"triplets_{name}.npz"), **save)
    print("wrote", name)


if __name__ == "__main__":
    make("small", 2, 24, 20, 12, 5)
    make("vg", 1, 100, 150, 50, 6)
