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


def class_flips(out_logits: torch.Tensor, ref_logits: torch.Tensor, tol: float = TOL) -> torch.Tensor:
    """[B,N] bool: queries whose arg-max class differs from the reference's.  `torch.argmax(logits)` selects the frequency-bias row
    of every pair the query takes part in (model/egtr.py:405-413) — a discontinuous step: when the reference's own two best
    logits are closer than the tolerance, either class is a correct answer at that tolerance, and `pred_rel` of the query's pairs
    then differs by a whole `triplet_dist` row.  Such a flip is accepted ONLY if the reference's margin between the two classes is
    below 2 * tol * max|logits| (i.e. inside the logits' own error bar) and at most 2 % of the queries (or two) are affected; the pairs of
    flipped queries are then left out of the pred_rel comparison.  Anything else fails."""
    out_logits, ref_logits = out_logits.detach().double().cpu(), ref_logits.detach().double().cpu()
    got, want = out_logits.argmax(-1), ref_logits.argmax(-1)
    flipped = got != want
    if flipped.any():
        margin = (ref_logits.gather(-1, want[..., None]) - ref_logits.gather(-1, got[..., None])).squeeze(-1)
        worst = float(margin[flipped].max() / ref_logits.abs().max())
        assert worst <= 2 * tol, f"arg-max class differs where the reference margin is {worst:.2e} of max|logits| (> {2 * tol:.0e})"
        assert int(flipped.sum()) <= max(2, 0.02 * flipped.numel()), f"{int(flipped.sum())} of {flipped.numel()} queries changed class"
    return flipped


def pair_mask(flipped: torch.Tensor, si: slice = slice(None), sj: slice = slice(None)) -> torch.Tensor:
    """[B,Ni,Nj] bool: pairs (i, j) none of whose queries changed class (the frequency bias of the others is a different row)."""
    return ~(flipped[:, si][:, :, None] | flipped[:, sj][:, None, :])


def pred_rel_err(out_rel: torch.Tensor, ref_rel: torch.Tensor, flipped: torch.Tensor, si: slice = slice(None), sj: slice = slice(None)) -> float:
    """Max-norm relative error of pred_rel (or a [.., si, sj, ..] sample of it) over the pairs of unflipped queries."""
    a, b = out_rel.detach().double().cpu(), ref_rel.detach().double().cpu()
    keep = pair_mask(flipped, si, sj)
    while keep.dim() < a.dim():
        keep = keep[..., None]
    return float(((a - b).abs() * keep).max() / b.abs().max().clamp_min(1e-30))


def forward_errors(out, ref, keys=("logits", "pred_boxes", "pred_rel", "pred_connectivity"), freq_bias: bool = True, tol: float = TOL):
    """{name: max-norm relative error} of one forward against a reference forward (full tensors), arg-max ties handled as above."""
    flipped = class_flips(out["logits"], ref["logits"], tol) if freq_bias else torch.zeros(ref["logits"].shape[:2], dtype=torch.bool)
    errs = {k: (pred_rel_err(out[k], ref[k], flipped) if k == "pred_rel" else relerr(out[k], ref[k])) for k in keys}
    if flipped.any():
        errs["_class_flips"] = int(flipped.sum())
    return errs


def worst(errs) -> float:
    return max(v for k, v in errs.items() if not k.startswith("_"))


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
    flipped = class_flips(out["logits"], ref["logits"], tol)  # near-tie arg-max classes (see class_flips): their pairs are left out
    if "pred_rel" in ref:
        errs["pred_rel"] = pred_rel_err(rel, ref["pred_rel"], flipped)
    else:
        errs["pred_rel"] = pred_rel_err(rel[:, ::3, ::7, :], ref["pred_rel_sample"], flipped, slice(None, None, 3), slice(None, None, 7))
        errs["pred_rel_sum_p"] = pred_rel_err(rel.double().sum(-1), ref["pred_rel_sum_p"], flipped)
        if not flipped.any():  # a sum over ALL pairs cannot leave the flipped ones out
            errs["pred_rel_sum_ij"] = relerr(rel.double().sum((1, 2)), ref["pred_rel_sum_ij"])
    if flipped.any():
        print(f"note: {int(flipped.sum())} of {flipped.numel()} queries resolve a near-tie arg-max the other way (reference margin < {2 * tol:.0e})")
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
