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
    H, W = meta.get("image") or WORKLOADS[meta["workload"]]["image"]
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
        errs["pred_rel_sum_p"] = relerr(rel.double().sum(-1), ref["pred_rel_sum_p"])
        errs["pred_rel_sum_ij"] = relerr(rel.double().sum((1, 2)), ref["pred_rel_sum_ij"])
    return errs


def p32_encode(x: torch.Tensor) -> torch.Tensor:
    """fp32 [rows, C] -> the P32 row format of include/egtr_b200.h, returned as a float32-typed [rows, C] tensor
    (same bytes): per 32-channel group 32 bf16 hi then 32 bf16 lo, hi = bf16_rn(x), lo = bf16_rn(x - hi)."""
    rows, C = x.shape
    assert C % 32 == 0
    x = x.float()
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    st = torch.stack([hi.view(rows, C // 32, 32), lo.view(rows, C // 32, 32)], 2).contiguous()  # [rows, G, 2, 32]
    return st.view(rows, 2 * C).view(torch.float32)


def p32_decode(p: torch.Tensor) -> torch.Tensor:
    rows, C = p.shape
    u = p.contiguous().view(torch.bfloat16).view(rows, C // 32, 2, 32).float()
    return (u[:, :, 0] + u[:, :, 1]).reshape(rows, C)
