"""Drop-in for the two helpers of `lib/pytorch_misc.py` the evaluation loop imports (train_egtr.py:30)."""
import numpy as np


def intersect_2d(x1, x2):
    """[m1, n], [m2, n] -> bool [m1, m2]: rows that are equal."""
    if x1.shape[1] != x2.shape[1]:
        raise ValueError("Input arrays must have same #columns")
    return (x1[:, None, :] == x2[None, :, :]).all(-1)


def argsort_desc(scores):
    """Indices [numel, ndim] that sort `scores` descending."""
    return np.stack(np.unravel_index(np.argsort(-scores, axis=None), scores.shape), 1)
