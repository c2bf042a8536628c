"""Dev micro-benchmark of the tcgen05 GEMM on the shapes the model runs (plain rows and 3x3 conv gather).
Prints us / algorithmic TFLOP/s per shape; `--profile i` wraps shape i in cudaProfilerStart/Stop for ncu."""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from egtr_b200 import _lib
from egtr_b200._lib import ASrc, Epilogue
from egtr_b200.engine import Lin, _conv_mat, _ptr, _stream

SHAPES = [  # name, M, N, K, kind
    ("enc_fc1", 22223, 1024, 256, "plain"),
    ("enc_fc2+res", 22223, 256, 1024, "res"),
    ("enc_offaw+pos", 22223, 384, 256, "a2"),
    ("enc_value", 22223, 256, 256, "plain"),
    ("dec_value6", 22223, 1536, 256, "plain"),
    ("l1_conv1", 66800, 64, 256, "plain"),
    ("l1_conv3+res", 66800, 256, 64, "res"),
    ("l1_conv2_3x3", 66800, 64, 576, "conv200x334x64"),
    ("l2_conv2_3x3", 16700, 128, 1152, "conv100x167x128"),
    ("l3_conv2_3x3", 4200, 256, 2304, "conv50x84x256"),
    ("l3_conv3+res", 4200, 1024, 256, "res"),
    ("l4_conv2_3x3", 1050, 512, 4608, "conv25x42x512"),
    ("rel_w2", 40000, 256, 256, "plain"),
    ("dec_fc2", 200, 256, 1024, "res"),
    ("dec_qk", 200, 512, 256, "a2"),
    ("tiny_1tile", 128, 64, 64, "plain"),
    ("tiny_1pair", 256, 256, 256, "plain"),
    ("one_round_n256", 18944, 256, 256, "plain"),
    ("two_rounds_n256", 37888, 256, 256, "plain"),
    ("one_round_k1024", 18944, 256, 1024, "plain"),
]

ap = argparse.ArgumentParser()
ap.add_argument("--profile", type=int, default=-1)
ap.add_argument("--only", type=str, default="")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--p32", action="store_true", help="feed plain-row shapes as P32 rows (TMA-fed kernel); --p32out also writes P32")
ap.add_argument("--p32out", action="store_true")
ap.add_argument("--p32prof", action="store_true", help="with an EGTR_P32_PROF build: phase timestamps of CTA 0")
ap.add_argument("--noflush", action="store_true", help="do not flush L2 between iterations (operands stay L2-resident as inside the forward)")
ap.add_argument("--prof", action="store_true", help="with an EGTR_GEMM_PROF build: print per-role cycle accounting")
args = ap.parse_args()
dev = torch.device("cuda:0")
_lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator().manual_seed(0)
for idx, (name, M, N, K, kind) in enumerate(SHAPES):
    if args.only and args.only not in name:
        continue
    if args.profile >= 0 and idx != args.profile:
        continue
    src, ep = ASrc(), Epilogue()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    lin = Lin(w, torch.randn(N, generator=g).to(dev), dev)
    out = torch.empty(M, N, device=dev)
    keep = []
    if kind.startswith("conv"):
        h, wd, c = [int(v) for v in kind[4:].split("x")]
        x = torch.randn(1, h, wd, c, generator=g).to(dev)
        keep.append(x)
        src.a, src.mode = _ptr(x), 1
        src.H, src.W, src.C, src.OH, src.OW, src.KH, src.KW, src.stride, src.pad = h, wd, c, h, wd, 3, 3, 1, 1
        assert h * wd == M and 9 * c == K
    else:
        a = torch.randn(M, K, generator=g)
        if args.p32 and kind != "a2":
            from tests.util import p32_encode
            a = p32_encode(a)
            src.fmt = 1
            ep.out_fmt = int(args.p32out and N % 32 == 0)
        a = a.to(dev)
        keep.append(a)
        src.a, src.mode, src.lda = _ptr(a), 0, K
        if kind == "a2":
            a2 = torch.randn(M, K, generator=g).to(dev)
            keep.append(a2)
            src.a2 = _ptr(a2)
    ep.bias, ep.out, ep.ldo, ep.ldr, ep.relu = _ptr(lin.b), _ptr(out), N, N, 1
    if kind == "res":
        r = torch.randn(M, N, generator=g).to(dev)
        keep.append(r)
        ep.res = _ptr(r)

    def run():
        _lib.call("egtr_gemm_sbf16", C.byref(src), _ptr(lin.planes), M, N, lin.Npad, K, C.byref(ep), _stream())

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    if args.profile >= 0:
        torch.cuda.profiler.start()
        run()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print("profiled", name)
        break
    ts = []
    for _ in range(args.iters):
        if not args.noflush:
            flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    if args.p32prof:
        lib = _lib.load()
        lib.egtr_debug_p32_prof.argtypes = [C.c_void_p]
        buf = (C.c_ulonglong * 8)()
        flush.fill_(0); run(); torch.cuda.synchronize()
        lib.egtr_debug_p32_prof(buf)
        t = [int(v) for v in buf[:6]]
        names = ["setup+pdl", "first stage full", "first acc full", "epilogue w0 done", "exit"]
        print("   CTA0 phases (us): " + " | ".join(f"{n} +{(t[i + 1] - t[i]) / 1e3:.2f}" for i, n in enumerate(names)) + f" | total {(t[5] - t[0]) / 1e3:.2f}")
    if args.prof:
        import numpy as np
        buf = (C.c_ulonglong * (148 * 16))()
        lib = _lib.load()
        lib.egtr_debug_gemm_prof.argtypes = [C.c_void_p]
        run(); torch.cuda.synchronize()
        lib.egtr_debug_gemm_prof(buf)
        a = np.array(buf[:], dtype=np.float64).reshape(148, 16)
        a = a[a[:, 6] > 0]  # CTAs that ran
        m = a.mean(0)
        print(f"   cycles/CTA: total {m[6]:9.0f} | TMA wait_empty {m[0]:8.0f} | MMA wait_tmem {m[4]:8.0f} wait_full {m[5]:8.0f} | "
              f"producer(w6) wait_empty {m[8]:8.0f} store {m[9]:8.0f} | epilogue wait_full {m[12]:8.0f} work {m[13]:8.0f} (tmem_ld {m[14]:8.0f} stage+store {m[15]:8.0f})")
    print(f"{name:16s} M={M:6d} N={N:5d} K={K:5d}  {us:8.1f} us  {2 * M * N * K / us / 1e6:7.1f} TFLOP/s (x3 bf16 = {6 * M * N * K / us / 1e6:7.1f})  BN={os.environ.get('EGTR_GEMM_BLOCK_N', 'auto')}")
