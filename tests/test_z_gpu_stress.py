"""GPU: repeatability of the serving (throughput) configuration at BASELINE.json's full size.

Round 1 found that persistent GEMM grids restricted to a share of the SMs (`egtr_set_grid_div` > 1) make 10-28 % of the
800x1333 forwards deviate — a race that a single pass of the parity tests rarely sees (profiles/r01_throughput_race_matrix.txt).
The shipped default is full grids; these tests run the shipped configuration many times, alone and in flight, and keep a
reproducer of the partial-grid race as an expected failure.  (File name: runs after the other GPU tests.)"""
import pytest
import torch

from tests.util import TOL, relerr

pytestmark = pytest.mark.gpu
KEYS = ("logits", "pred_boxes", "pred_rel", "pred_connectivity")
H, W = 800, 1333


@pytest.fixture(scope="module")
def case(cuda):
    from egtr_b200.config import workload_config
    from egtr_b200.model.egtr import DetrForSceneGraphGeneration
    from egtr_b200.synth import synth_images, synth_state_dict
    cfg = workload_config("B")
    model = DetrForSceneGraphGeneration(cfg)
    model.load_state_dict(synth_state_dict(cfg, 32))
    model.cuda().eval()
    px, mask = synth_images(1, H, W, seed=33)
    px, mask = px.to(cuda), mask.to(cuda)
    eng = model.engine()
    out = eng.forward(px, mask)  # latency configuration: pinned against the reference golden by test_gpu_forward (forward_B)
    torch.cuda.synchronize()
    return eng, px, mask, {k: out[k].clone() for k in KEYS}


def _worst(out, ref):
    from tests.util import forward_errors, worst
    return worst(forward_errors(out, ref, KEYS))  # near-tie arg-max classes handled as in tests/util.py::class_flips


def test_throughput_configuration_repeats_at_full_size(case):
    eng, px, mask, ref = case
    errs = [_worst(eng.forward(px, mask, throughput=True), ref) for _ in range(40)]
    print("throughput forwards vs the lone forward: worst", f"{max(errs):.2e}")
    assert max(errs) < TOL, sorted(errs)[-5:]


def test_forwards_in_flight_reproduce_the_lone_forward_at_full_size(case):
    """Four captured forwards in flight on four streams (the serving runner's and bench.py's arrangement), five rounds."""
    eng, px, mask, ref = case
    conc = 4
    runners = [eng.graph_runner(1, H, W, slot=i, throughput=True) for i in range(conc)]
    streams = [torch.cuda.Stream() for _ in range(conc)]
    main = torch.cuda.current_stream()
    errs = []
    for _ in range(5):
        for st in streams:
            st.wait_stream(main)
        for i in range(conc):
            with torch.cuda.stream(streams[i]):
                runners[i](px, mask)
        for st in streams:
            main.wait_stream(st)
        torch.cuda.synchronize()
        errs += [_worst(r.out, ref) for r in runners]
    print("forwards in flight vs the lone forward: worst", f"{max(errs):.2e}")
    assert max(errs) < TOL, sorted(errs)[-5:]


@pytest.mark.parametrize("div", [2, 4])  # half grids (batches of 4 - 8 in flight) and quarter grids (batch 1, the shipped B configuration)
def test_partial_grid_race_reproducer(case, div):
    """Round 1's race (LayerNorm epilogue: ld.shared of a residual box not ordered before the TMA request that overwrites it,
    visible only with partial grids) — fixed by a proxy fence; 0 / 1188 forwards deviated in round 2's matrix
    (profiles/r02_race_matrix.txt).  Regular test since."""
    eng, px, mask, ref = case
    keep = eng.throughput_grid_div
    eng.throughput_grid_div = div
    try:
        errs = [_worst(eng.forward(px, mask, throughput=True), ref) for _ in range(60)]
    finally:
        eng.throughput_grid_div = keep
    assert max(errs) < TOL, f"{sum(e > TOL for e in errs)}/60 forwards deviate"


KERNELS = {  # the five GEMMs of one encoder layer at workload B (M = 22 223 tokens)
    "offaw": dict(N=384, K=256),
    "value": dict(N=256, K=256, keep=True),
    "out_proj_ln": dict(N=256, K=256, ln=True, res=True),
    "fc1": dict(N=1024, K=256, relu=True),
    "fc2_ln_out2": dict(N=256, K=1024, ln=True, res=True, out2=True),
}


@pytest.mark.parametrize("kind", list(KERNELS))
def test_partial_grid_race_per_kernel(cuda, kind):
    """One kernel at a time, same inputs: the launch with full grids is the reference, 40 launches on half of the SMs must be
    bit-identical to it (no split-K at these shapes: the tiles and their arithmetic are the same, only the CTA that runs them
    differs)."""
    import ctypes as C
    import warnings

    from egtr_b200 import _lib
    from egtr_b200._lib import ASrc, Epilogue
    from egtr_b200.engine import Lin
    from tests.util import p32_encode
    k = KERNELS[kind]
    M, N, K = 22223, k["N"], k["K"]
    g = torch.Generator().manual_seed(N * 7 + K)
    a = p32_encode(torch.randn(M, K, generator=g)).to(cuda)
    lin = Lin((torch.randn(N, K, generator=g) / K ** 0.5).to(cuda), torch.randn(N, generator=g).to(cuda), cuda)
    res = p32_encode(torch.randn(M, 256, generator=g)).to(cuda) if k.get("res") else None
    gamma, beta = (1 + 0.2 * torch.randn(256, generator=g)).to(cuda), torch.randn(256, generator=g).to(cuda)
    addend = torch.randn(M, 256, generator=g).to(cuda)
    keep = (torch.rand(M, generator=g) > 0.1).to(torch.uint8).to(cuda) if k.get("keep") else None
    st = torch.cuda.current_stream().cuda_stream

    def launch(div):
        out = torch.full((M, N), float("nan"), device=cuda)
        out2 = torch.full((M, 256), float("nan"), device=cuda)
        src, ep = ASrc(), Epilogue()
        src.a, src.mode, src.lda, src.fmt = a.data_ptr(), 0, K, 1
        ep.bias, ep.out, ep.ldo, ep.ldr = lin.b.data_ptr(), out.data_ptr(), N, N
        ep.relu, ep.out_fmt = int(bool(k.get("relu"))), (1 if k.get("ln") else 0)
        if res is not None:
            ep.res, ep.res_fmt = res.data_ptr(), 1
        if keep is not None:
            ep.row_keep = keep.data_ptr()
        if k.get("ln"):
            ep.ln_gamma, ep.ln_beta = gamma.data_ptr(), beta.data_ptr()
            if k.get("out2"):
                ep.ln_out2, ep.ln_addend = out2.data_ptr(), addend.data_ptr()
        _lib.call("egtr_set_splitk_max", 1)
        _lib.call("egtr_set_grid_div", div)
        try:
            _lib.call("egtr_gemm_sbf16", C.byref(src), lin.planes.data_ptr(), M, N, lin.Npad, K, C.byref(ep), st)
            torch.cuda.synchronize()
        finally:
            _lib.call("egtr_set_grid_div", 1)
            _lib.call("egtr_set_splitk_max", 64)
        return out.view(torch.int32), out2.view(torch.int32)

    ref, ref2 = launch(1)
    again, _ = launch(1)
    assert torch.equal(ref, again), "full grids are not repeatable either"
    bad_runs, bad_rows, bad_cols = 0, set(), set()
    for _ in range(40):
        o, o2 = launch(2)
        diff = (o != ref) | ((o2 != ref2) if k.get("out2") else False)
        if bool(diff.any()):
            bad_runs += 1
            rows, cols = diff.nonzero(as_tuple=True)
            bad_rows.update((rows // 32).unique().tolist()[:64])
            bad_cols.update((cols // 32).unique().tolist())
    msg = (f"partial-grid repeatability, {kind} (M={M} N={N} K={K}): {bad_runs}/40 launches differ from the full-grid launch; "
           f"32-row groups hit {sorted(bad_rows)[:24]}, 32-column chunks hit {sorted(bad_cols)}")
    warnings.warn(msg)
    assert bad_runs == 0, msg


# Last in the last file: a device-side fault here cannot take other tests with it.
@pytest.mark.parametrize("hw,pad", [((33, 47), None), ((17, 23), None), ((64, 64), [(64, 64), (40, 33)]),
                                    ((97, 131), [(97, 131), (50, 131), (97, 60)])])
def test_forward_edge_sizes_vs_live_oracle(cuda, hw, pad):
    """Degenerate geometries: images so small that the coarse levels shrink to 1x1 / 2x2 maps (patch tiles, TMA boxes and the
    MSDA patches are then larger than the map), sizes that are not multiples of the strides, and images padded to less than
    half of the batch canvas."""
    from egtr_b200.config import workload_config
    from egtr_b200.synth import synth_images, synth_state_dict
    from oracle import egtr_oracle as orc
    cfg = workload_config("tiny")
    sd = synth_state_dict(cfg, 70)
    batch = len(pad) if pad else 1
    px, mask = synth_images(batch, hw[0], hw[1], seed=71, pad_to=pad)
    want = orc.forward(sd, cfg, px, mask)
    from egtr_b200.model.egtr import DetrForSceneGraphGeneration
    model = DetrForSceneGraphGeneration(cfg)
    model.load_state_dict(sd)
    model.cuda().eval()
    out = model(pixel_values=px.to(cuda), pixel_mask=mask.to(cuda), output_attentions=False, output_attention_states=True,
                output_hidden_states=True)
    torch.cuda.synchronize()
    from tests.util import compare_forward
    errs = compare_forward(out, want)
    print(hw, batch, {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < TOL, errs
