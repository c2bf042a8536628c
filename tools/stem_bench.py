"""Stem kernels alone at workload B: gather-GEMM path vs the TMA path (CUDA events, L2 flushed between iterations)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import build_case
from egtr_b200._lib import call
from egtr_b200.engine import Engine, _ptr, _stream

cfg, sd, px, mask, _ = build_case("B", 1)
px = px.cuda()
B, _, H, W = px.shape
res = {}
for mode in ("gather", "tma"):
    os.environ["EGTR_STEM"] = mode
    eng = Engine(cfg, sd, torch.device("cuda:0"))
    ws = eng._workspace(B, H, W)
    h1, w1 = ws["stem_hw"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def pad():
        if mode == "tma":
            call("egtr_stem_pad_split_bf16", _ptr(px), B, H, W, _ptr(ws["px_planes"]), _stream())
        else:
            if ws["px4"] is None:
                ws["px4"] = torch.empty(B * (H + 6) * (W + 6) * 4, dtype=torch.float32, device="cuda")
            call("egtr_pad_nchw3_to_nhwc4_f32", _ptr(px), B, H, W, 3, _ptr(ws["px4"]), _stream())

    def conv():
        if mode == "tma":
            call("egtr_stem_conv7x7s2_bf16x3", _ptr(ws["px_planes"]), B, H, W, _ptr(eng.stem_w_planes), _ptr(eng.stem_bias), _ptr(ws["stem"]), _stream())
        else:
            eng.gemm(eng.stem, B * h1 * w1, ws["stem"], relu=True,
                     conv=dict(x=ws["px4"], mode=3, H=H + 6, W=W + 6, C=4, OH=h1, OW=w1, KH=7, KW=7, stride=2, pad=0))

    for name, fn in (("pad", pad), ("conv", conv)):
        ts = []
        for i in range(12):
            flush.fill_(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res[(mode, name)] = sorted(ts)[len(ts) // 2]
        print(mode, name, f"{res[(mode, name)]:.1f} us (median of 12, incl. ~5 us of event overhead)")
