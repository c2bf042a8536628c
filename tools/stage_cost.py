"""Marginal cost of each stage in the TIMED configuration (workload B, 8 forwards in flight, half grids).

A lone forward's per-kernel times do not say what a stage costs once eight forwards overlap: small kernels hide behind other
images' GEMMs, persistent grids hold SMs.  This tool captures the forward with one group of launches REMOVED (outputs are then
garbage — dev only), replays eight such graphs on eight streams exactly like bench.py's `value` leg, and reports images/s and
the ms per image the removed group was worth.  It decides where kernel work pays."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import egtr_b200.engine as E
from bench import build_case
from egtr_b200.model.egtr import DetrForSceneGraphGeneration

WORKLOAD = os.environ.get("STAGE_COST_WORKLOAD", "B")
BATCH = int(os.environ.get("STAGE_COST_BATCH", "1"))
INFLIGHT = int(os.environ.get("STAGE_COST_INFLIGHT", "8"))
STEPS = int(os.environ.get("STAGE_COST_STEPS", "160"))
OUT = open(os.path.join("gpurun_out", "stage_cost.txt"), "a") if os.path.isdir("gpurun_out") else None


def say(*a):
    msg = " ".join(str(x) for x in a)
    print(msg, flush=True)
    if OUT:
        OUT.write(msg + "\n")
        OUT.flush()


cfg, sd, px, mask, _ = build_case(WORKLOAD, BATCH)
model = DetrForSceneGraphGeneration(cfg)
model.load_state_dict(sd)
model.cuda().eval()
eng = model.engine()
px, mask = px.cuda(), mask.cuda()
B, _, H, W = px.shape
Md = B * cfg.num_queries

state = {"spans": [], "skip": set()}
orig_call, orig_span = E.call, E.Engine.span


class Span:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        state["spans"].append(self.name)
        return self

    def __exit__(self, *a):
        state["spans"].pop()
        return False


def span(self, name):
    return Span(name)


PASS = ("egtr_set_scratch_slot", "egtr_set_splitk_max", "egtr_set_grid_div", "egtr_groupnorm_scratch_doubles")


def gemm_m(args):  # egtr_gemm_sbf16(src, planes, M, N, Npad, K, ep, stream)
    return int(args[2])


def call(name, *args):
    if name in PASS:
        return orig_call(name, *args)
    sk, sp = state["skip"], state["spans"]
    stage = next((s for s in sp if s.startswith("stage_")), "stage_none")
    drop = False
    if "dec_small" in sk and stage == "stage_decoder" and not (name == "egtr_gemm_sbf16" and gemm_m(args) > 512):
        drop = True
    if "dec_value" in sk and stage == "stage_decoder" and name == "egtr_gemm_sbf16" and gemm_m(args) > 512:
        drop = True
    if "msda_enc" in sk and "msda_enc" in sp:
        drop = True
    if "relation" in sk and stage == "stage_relation":
        drop = True
    if "enc_gemm" in sk and stage == "stage_encoder" and name == "egtr_gemm_sbf16":
        drop = True
    if "enc_ffn" in sk and stage == "stage_encoder" and name == "egtr_gemm_sbf16" and (int(args[3]) == 1024 or int(args[5]) == 1024):
        drop = True
    if "bb_gemm" in sk and stage == "stage_backbone" and name == "egtr_gemm_sbf16" and "gemm_p32" in sp:
        drop = True
    if "stem" in sk and stage == "stage_backbone" and (name in ("egtr_pad_nchw3_to_nhwc4_f32", "egtr_maxpool3x3s2_nhwc_ex", "egtr_stem_pad_split_bf16",
                                                                "egtr_stem_conv7x7s2_bf16x3") or
                                                        (name == "egtr_gemm_sbf16" and "gemm_p32" not in sp)):
        drop = True
    if "groupnorm" in sk and name == "egtr_groupnorm_ex":
        drop = True
    if "geometry" in sk and name == "egtr_levels_geometry_f32":
        drop = True
    if "p32_to_rows" in sk and name == "egtr_p32_to_rows":
        drop = True
    if drop:
        return 0
    return orig_call(name, *args)


E.call = call
E.Engine.span = span


def measure(skip):
    state["skip"] = set(skip)
    eng._ws.clear() if False else None
    runners = [E.GraphRunner(eng, B, H, W, slot=i, throughput=INFLIGHT > 1) for i in range(INFLIGHT)]
    for r in runners:
        r.px.copy_(px)
        r.pm.copy_(mask)
    streams = [torch.cuda.Stream() for _ in range(INFLIGHT)]
    torch.cuda.synchronize()

    def loop(n):
        for i in range(n):
            with torch.cuda.stream(streams[i % INFLIGHT]):
                runners[i % INFLIGHT].graph.replay()

    loop(2 * INFLIGHT)
    torch.cuda.synchronize()
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        loop(STEPS)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / STEPS
        best = dt if best is None or dt < best else best
    # a lone replay's latency
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        runners[0].graph.replay()
    torch.cuda.synchronize()
    lone = (time.perf_counter() - t0) / 10
    del runners
    return best * 1e3, lone * 1e3


VARIANTS = [(), ("dec_small",), ("msda_enc",), ("relation",), ("enc_ffn",), ("enc_gemm",), ("bb_gemm",), ("stem",), ("groupnorm",),
            ("geometry",), ("dec_value",), ("p32_to_rows",)]
if len(sys.argv) > 1:
    VARIANTS = [()] + [tuple(v.split("+")) for v in sys.argv[1:]]
say(f"workload {WORKLOAD} batch {BATCH}, {INFLIGHT} forwards in flight, {STEPS} steps x 3 (best)")
base = None
for v in VARIANTS:
    ms, lone = measure(v)
    if base is None:
        base = (ms, lone)
    say(f"skip {'+'.join(v) or '(nothing)':<14} step {ms:7.3f} ms  ({B * 1e3 / ms:7.1f} images/s)  worth {base[0] - ms:6.3f} ms/step = {100 * (base[0] - ms) / base[0]:5.1f} %"
        f"   lone replay {lone:6.3f} ms (worth {base[1] - lone:6.3f})")
