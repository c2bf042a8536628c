"""Bisect a deviation of the throughput configuration at full size (workload B).

Runs the eager forward under a matrix of (split-K cap, grid divisor, balanced grids), logs a checksum of every GEMM's output
and reports, per configuration, the final-output error against the latency configuration (which is pinned by the reference's
golden) and the first GEMM call whose output deviates.  One process, a few forwards: seconds on the GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import egtr_b200.engine as E
from bench import build_case
from egtr_b200 import _lib
from egtr_b200.model.egtr import DetrForSceneGraphGeneration

OUT = open(os.path.join("gpurun_out", "diag.txt"), "w") if os.path.isdir("gpurun_out") else sys.stdout


def say(*a):
    msg = " ".join(str(x) for x in a)
    print(msg, flush=True)
    if OUT is not sys.stdout:
        OUT.write(msg + "\n")
        OUT.flush()


def relerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


if len(sys.argv) > 1 and sys.argv[1] == "golden":  # the inputs of tests/golden/forward_B.npz (weight seed 32, image seed 33)
    from egtr_b200.config import workload_config
    from egtr_b200.synth import synth_images, synth_state_dict
    cfg = workload_config("B")
    sd = synth_state_dict(cfg, seed=32)
    px, mask = synth_images(1, 800, 1333, seed=33)
else:
    cfg, sd, px, mask, _ = build_case("B", 1)
model = DetrForSceneGraphGeneration(cfg)
model.load_state_dict(sd)
model.cuda().eval()
eng = model.engine()
px, mask = px.cuda(), mask.cuda()

knobs = {"sk": 64, "div": 1}
orig_call = E.call


def call(name, *args):
    if name == "egtr_set_splitk_max":
        args = (knobs["sk"],)
    elif name == "egtr_set_grid_div":
        args = (knobs["div"],)
    return orig_call(name, *args)


E.call = call
rec = []
orig_gemm = E.Engine.gemm


def decode(t, fmt):
    if fmt == 1:
        return t.contiguous().view(torch.bfloat16).float().view(-1, 2, 32).sum(1).reshape(-1)
    return t.reshape(-1)


def gemm(self, lin, M, out, **kw):
    orig_gemm(self, lin, M, out, **kw)
    try:
        _checksum(lin, M, out, kw)
    except Exception as e:  # noqa: BLE001  (the bisect must not die on a checksum)
        rec.append((f"M={M} N={lin.N} K={lin.K} checksum failed: {e!r}", float("nan"), float("nan"), torch.zeros(1)))


def _checksum(lin, M, out, kw):
    ldo = kw.get("ldo") or lin.N
    n = out.numel() if kw.get("remap") else min(out.numel(), kw.get("out_col", 0) + M * ldo)
    n -= n % 64
    x = decode(out.reshape(-1)[:n], kw.get("out_fmt", 0)).double()
    desc = f"M={M} N={lin.N} K={lin.K}" + (" conv%dx%d/s%d" % (kw["conv"]["KH"], kw["conv"]["KW"], kw["conv"]["stride"]) if kw.get("conv") else "") \
        + (" ln" if kw.get("ln") is not None else "") + (" res" if kw.get("res") is not None else "") + (" relu" if kw.get("relu") else "") \
        + (" remap" if kw.get("remap") else "") + f" fmt a{kw.get('a_fmt', 0)} o{kw.get('out_fmt', 0)}"
    rec.append((desc, float(x.sum()), float(x.abs().sum()), x[:: max(1, x.numel() // 4096)].float().clone()))
    if kw.get("ln_out2") is not None:
        o2 = kw["ln_out2"]
        y = decode(o2.reshape(-1)[: M * 256], 1).double()
        rec.append((desc + " (out2)", float(y.sum()), float(y.abs().sum()), y[:: max(1, y.numel() // 4096)].float().clone()))


E.Engine.gemm = gemm
KEYS = ("logits", "pred_boxes", "pred_rel", "pred_connectivity")


def run(sk, div, bal):
    knobs["sk"], knobs["div"] = sk, div
    orig_call("egtr_set_grid_balance", bal)
    rec.clear()
    taps = {}
    out = eng.forward(px, mask, taps=taps, throughput=True)
    torch.cuda.synchronize()
    return {k: out[k].clone() for k in KEYS}, {k: v.clone() for k, v in taps.items() if v.dtype == torch.float32}, list(rec)


ref_out, ref_taps, ref_rec = run(64, 1, 0)
say("reference run (split-K 64, full grids): GEMM calls", len(ref_rec))
again_out, _, again_rec = run(64, 1, 0)
say("  repeat of the reference run:", {k: f"{relerr(again_out[k], ref_out[k]):.1e}" for k in KEYS})
CONFIGS = [(1, 2, 1), (1, 2, 0), (1, 1, 0), (64, 2, 0), (64, 2, 1), (1, 2, 1), (1, 4, 1), (1, 4, 0)]
for sk, div, bal in CONFIGS:
    try:
        out, taps, r = run(sk, div, bal)
    except Exception as e:  # noqa: BLE001
        say(f"config sk={sk} div={div} bal={bal}: EXCEPTION {e!r}")
        continue
    errs = {k: relerr(out[k], ref_out[k]) for k in KEYS}
    terr = {k: relerr(taps[k], ref_taps[k]) for k in ("c3", "c4", "c5", "source_flatten", "enc0_out") if k in taps and k in ref_taps}
    bad = max(errs.values()) > 1e-3
    say(f"config sk={sk} div={div} bal={bal}: {'FAIL' if bad else 'ok'}  out {({k: f'{v:.1e}' for k, v in errs.items()})}  taps {({k: f'{v:.1e}' for k, v in terr.items()})}")
    if len(r) != len(ref_rec):
        say(f"   GEMM call count differs: {len(r)} vs {len(ref_rec)}")
    dev = []
    for i, (a, b) in enumerate(zip(r, ref_rec)):
        e_abs = abs(a[2] - b[2]) / max(abs(b[2]), 1e-30)
        e_smp = relerr(a[3], b[3]) if a[3].shape == b[3].shape else float("nan")
        if e_abs > 1e-3 or not (e_smp < 1e-2):
            dev.append((i, a[0], e_abs, e_smp))
    for i, d, e1, e2 in dev[:6]:
        say(f"   deviating GEMM #{i}: {d}: |x|-sum rel diff {e1:.2e}, sampled max-norm err {e2:.2e}")
    if dev:
        say(f"   ({len(dev)} deviating GEMM outputs of {len(r)}; first at #{dev[0][0]})")
say("done")
