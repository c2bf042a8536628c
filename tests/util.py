"""Shared helpers for the parity tests."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Parity bar of BASELINE.json north_star: 1e-3 relative.  Measured as max-norm relative error,
# max|a-b| / max|b|, per output tensor.
TOL = 1e-3


def relerr(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}, meta


def case_inputs(meta):
    """Rebuild config / weights / images of a golden forward case from its recorded seeds."""
    from egtr_b200.config import WORKLOADS, workload_config
    from egtr_b200.synth import synth_images, synth_state_dict

    cfg = workload_config(meta["workload"], **meta["overrides"])
    H, W = WORKLOADS[meta["workload"]]["image"]
    sd = synth_state_dict(cfg, seed=meta["weight_seed"])
    pad = [tuple(p) for p in meta["pad_to"]] if meta["pad_to"] else None
    px, mask = synth_images(meta["batch"], H, W, seed=meta["image_seed"], pad_to=pad)
    return cfg, sd, px, mask


def compare_forward(out, ref, tol=TOL, argmax_guard=None):
    """Compare model outputs with golden/oracle tensors; returns {name: err}."""
    errs = {}
    for k in ("logits", "pred_boxes", "pred_connectivity", "last_hidden_state"):
        if k in ref:
            errs[k] = relerr(out[k], ref[k])
    if "encoder_last_hidden_state" in ref:
        errs["encoder_last_hidden_state"] = relerr(out["encoder_last_hidden_state"], ref["encoder_last_hidden_state"])
    if "encoder_last_hidden_state_sample" in ref:
        errs["encoder_last_hidden_state"] = relerr(out["encoder_last_hidden_state"][:, ::37, :], ref["encoder_last_hidden_state_sample"])
    rel = out["pred_rel"]
    if "pred_rel" in ref:
        errs["pred_rel"] = relerr(rel, ref["pred_rel"])
    else:
        errs["pred_rel"] = relerr(rel[:, ::3, ::7, :], ref["pred_rel_sample"])
        errs["pred_rel_sum_p"] = relerr(rel.sum(-1), ref["pred_rel_sum_p"])
        errs["pred_rel_sum_ij"] = relerr(rel.sum((1, 2)), ref["pred_rel_sum_ij"])
    return errs
