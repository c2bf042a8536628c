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
