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
    return max(relerr(out[k], ref[k]) for k in KEYS)


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


@pytest.mark.xfail(reason="known race with partial persistent grids (egtr_set_grid_div > 1), round-1 finding; default is full grids",
                   strict=False)
def test_partial_grid_race_reproducer(case):
    eng, px, mask, ref = case
    keep = eng.throughput_grid_div
    eng.throughput_grid_div = 2
    try:
        errs = [_worst(eng.forward(px, mask, throughput=True), ref) for _ in range(60)]
    finally:
        eng.throughput_grid_div = keep
    assert max(errs) < TOL, f"{sum(e > TOL for e in errs)}/60 forwards deviate"
