"""CPU oracle of the triplet extraction.  TEST INFRASTRUCTURE ONLY.

numpy/torch restatement of the model-output half of `evaluate_batch` (`/root/reference/train_egtr.py:56-69,
84-94, 120-128`) with the reference's `argsort_desc` (`/root/reference/lib/pytorch_misc.py:27-34`) restated as a
STABLE descending argsort (the reference's `np.argsort` uses quicksort, whose order among exactly equal scores is
unspecified; ties are broken here by ascending flat index).  Pinned by `tests/golden/triplets_*.npz`, generated in the
build container with the reference's own `argsort_desc` (tests/golden/make_golden_triplets.py).
"""
import numpy as np
import torch


def argsort_desc_stable(scores: np.ndarray) -> np.ndarray:
    order = np.argsort(-scores.ravel(), kind="stable")
    return np.column_stack(np.unravel_index(order, scores.shape))


def extract(logits: torch.Tensor, pred_rel: torch.Tensor, pred_conn, num_labels: int, single: bool, topk: int = 100):
    out = []
    for j in range(logits.shape[0]):
        obj_scores, pred_classes = torch.max(logits[j].softmax(-1)[:, :num_labels], -1)      # 56-58
        sub_ob = torch.outer(obj_scores, obj_scores)                                           # 59
        n = logits.shape[1]
        sub_ob[torch.arange(n), torch.arange(n)] = 0.0                                         # 60-62
        rel = torch.clamp(pred_rel[j], 0.0, 1.0)                                               # 65
        if pred_conn is not None:
            rel = rel * torch.clamp(pred_conn[j], 0.0, 1.0)                                    # 66-68
        if single:
            scores = rel.max(-1)[0] * sub_ob                                                   # 120
            inds = argsort_desc_stable(scores.numpy())[:topk]                                  # 121-123
            rs = rel.numpy()[inds[:, 0], inds[:, 1]]                                           # 124-126
        else:
            scores = rel * sub_ob.unsqueeze(-1)                                                # 84
            inds = argsort_desc_stable(scores.numpy())[:topk]                                  # 85-87
            rs = rel.numpy()[inds[:, 0], inds[:, 1], inds[:, 2]]                               # 88-92
        out.append(dict(obj_scores=obj_scores.numpy(), pred_classes=pred_classes.numpy(), pred_rel_inds=inds,
                        rel_scores=rs, scores=scores.numpy()))
    return out
