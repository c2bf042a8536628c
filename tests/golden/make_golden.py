"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For each case it (1) builds the reference `DetrForSceneGraphGeneration` through the import
stubs in `tools/ref_harness.py`, (2) strict-loads the seeded synthetic state dict
(`egtr_b200/synth.py`), (3) runs the reference's own CPU forward
(`evaluate_egtr.py:30-36` call signature) and stores its outputs, and (4) prints the oracle's
deviation from them.  Kernel-level vectors for MSDeformAttn come from the reference's
`ms_deform_attn_core_pytorch` (`model/deformable_detr.py:925-960`).

Weights and inputs are NOT stored: they are regenerated bit-identically from the seeds recorded
in each file (numpy PCG64).
"""
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from egtr_b200.config import WORKLOADS, workload_config  # noqa: E402
from egtr_b200.synth import synth_images, synth_msda_inputs, synth_state_dict  # noqa: E402
from oracle import egtr_oracle as orc  # noqa: E402
from tools.ref_harness import build_reference_model, import_reference  # noqa: E402

CASES = [
    # name, workload, batch, pad_to, config overrides, weight seed, image seed
    ("tiny_b1", "tiny", 1, None, {}, 10, 11),
    ("tiny_b2_ragged", "tiny", 2, [(96, 128), (70, 101)], {}, 10, 12),
    ("small_b2_ragged_logitadj", "small", 2, [(131, 224), (160, 180)], dict(logit_adjustment=True), 20, 21),
    ("small_nofreq", "small", 1, None, dict(use_freq_bias=False), 20, 22),
    ("A", "A", 1, None, {}, 30, 31),
    # BASELINE.json configs[1] at full size, and the D / E label spaces (601 classes / 30 predicates; N_q = 300 / 200 predicates)
    # at a reduced image size (optional 8th field) so the fixtures stay small
    ("B", "B", 1, None, {}, 32, 33),
    ("D_small_b2_ragged", "D", 2, [(224, 320), (192, 272)], {}, 34, 35, (224, 320)),
    ("E_small", "E", 1, None, {}, 36, 37, (256, 256)),
]


def relerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def run_case(name, wl, batch, pad_to, over, wseed, iseed, image=None):
    cfg = workload_config(wl, **over)
    H, W = image or WORKLOADS[wl]["image"]
    sd = synth_state_dict(cfg, seed=wseed)
    px, mask = synth_images(batch, H, W, seed=iseed, pad_to=pad_to)
    dd, eg, rcfg, model = build_reference_model(cfg.to_dict())
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    t0 = time.perf_counter()
    with torch.no_grad():
        out = model(pixel_values=px, pixel_mask=mask, output_attentions=False,
                    output_attention_states=True, output_hidden_states=True)
    t_ref = time.perf_counter() - t0
    t0 = time.perf_counter()
    mine = orc.forward(sd, cfg, px, mask)
    t_orc = time.perf_counter() - t0

    ref = dict(
        logits=out.logits, pred_boxes=out.pred_boxes, pred_rel=out.pred_rel,
        pred_connectivity=out.pred_connectivity, last_hidden_state=out.last_hidden_state,
        encoder_last_hidden_state=out.encoder_last_hidden_state,
    )
    errs = {k: relerr(mine[k], v) for k, v in ref.items()}
    print(f"[{name}] ref {t_ref:.2f}s oracle {t_orc:.2f}s  oracle-vs-reference max-norm rel err:")
    for k, v in errs.items():
        print(f"    {k:28s} {v:.3e}")
    assert max(errs.values()) < 2e-4, errs

    save = {k: v.numpy() for k, v in ref.items()}
    rel = save["pred_rel"]
    if rel.size > 300_000:  # cfg A and larger: keep a strided sample + reductions of the multi-MB tensor
        save["pred_rel_sample"] = rel[:, ::3, ::7, :].copy()
        # reductions accumulate in float64: numpy's float32 sum over the two leading axes of a 10^7-element tensor is a plain
        # running sum (2e-4 off at the E label space), which would make the fixture less accurate than what it checks
        save["pred_rel_sum_p"] = rel.sum(-1, dtype=np.float64).astype(np.float32)
        save["pred_rel_sum_ij"] = rel.sum((1, 2), dtype=np.float64).astype(np.float32)
        del save["pred_rel"]
    if save["encoder_last_hidden_state"].size > 300_000:
        e = save.pop("encoder_last_hidden_state")
        save["encoder_last_hidden_state_sample"] = e[:, ::37, :].copy()
    meta = dict(case=name, workload=wl, batch=batch, pad_to=pad_to, overrides=over, weight_seed=wseed,
                image_seed=iseed, image=[H, W], reference_commit="7f87450", torch=torch.__version__,
                oracle_vs_reference=errs)
    save["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, f"forward_{name}.npz"), **save)


def msda_kernel_vectors():
    dd, _ = import_reference()
    cases = [
        ("msda_enc_small", 2, [(12, 17), (6, 9), (3, 5), (2, 3)], None, 3),
        ("msda_dec_small", 2, [(20, 27), (10, 14), (5, 7), (3, 4)], 19, 4),
        ("msda_edge", 1, [(1, 1), (2, 1), (1, 3), (5, 5)], 7, 5),
    ]
    for name, B, shapes, nq, seed in cases:
        S = sum(h * w for h, w in shapes)
        value, spatial, start, loc, w = synth_msda_inputs(B, shapes, nq or S, seed=seed)
        if name == "msda_edge":  # exact grid lines, far outside, boundary values
            loc.view(-1)[::5] = 0.0
            loc.view(-1)[1::7] = 1.0
            loc.view(-1)[2::11] = -0.75
            loc.view(-1)[3::13] = 1.9
        ref = dd.ms_deform_attn_core_pytorch(value, [tuple(s) for s in spatial.tolist()], loc, w)
        mine = orc.msda_core(value, [tuple(s) for s in spatial.tolist()], loc, w)
        e = relerr(mine, ref)
        print(f"[{name}] oracle msda_core vs reference ms_deform_attn_core_pytorch: {e:.3e}")
        assert e < 1e-5
        meta = dict(case=name, batch=B, shapes=shapes, n_query=nq or S, seed=seed, edge=(name == "msda_edge"))
        np.savez_compressed(
            os.path.join(HERE, f"{name}.npz"), out=ref.numpy(),
            loc=loc.numpy() if name == "msda_edge" else np.zeros(0, np.float32),
            meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
        )


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    msda_kernel_vectors()
    only = sys.argv[1:] or None
    for c in CASES:
        if only is None or c[0] in only:
            run_case(*c)
