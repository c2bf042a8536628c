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


# Last in the last file: a device-side fault here cannot take other tests with it.
@pytest.mark.xfail(strict=False, reason="written after round 1's GPU minutes were spent: not yet run on hardware (geometries far "
                                        "below anything the reference is used with); an XPASS promotes it to a regular test next round")
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
