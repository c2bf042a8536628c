"""CPU: the oracle against the golden fixtures produced by the unmodified reference."""
import pytest
import torch

from oracle import egtr_oracle as orc
from tests.util import case_inputs, compare_forward, load_golden, relerr

FORWARD_CASES = ["forward_tiny_b1", "forward_tiny_b2_ragged", "forward_small_b2_ragged_logitadj", "forward_small_nofreq",
                 "forward_D_small_b2_ragged", "forward_E_small", "forward_B"]  # D / E: the Open-Images and stress label spaces (BASELINE.json configs 3, 4); B: configs[1] at full size


@pytest.mark.parametrize("name", FORWARD_CASES)
def test_oracle_matches_reference_forward(name):
    ref, meta = load_golden(name)
    cfg, sd, px, mask = case_inputs(meta)
    out = orc.forward(sd, cfg, px, mask)
    errs = compare_forward(out, ref)
    assert max(errs.values()) < 5e-5, errs


@pytest.mark.parametrize("name", ["msda_enc_small", "msda_dec_small", "msda_edge"])
def test_oracle_msda_core_matches_reference_kernel_twin(name):
    from egtr_b200.synth import synth_msda_inputs

    ref, meta = load_golden(name)
    value, spatial, start, loc, w = synth_msda_inputs(meta["batch"], [tuple(s) for s in meta["shapes"]], meta["n_query"], seed=meta["seed"])
    if meta["edge"]:
        loc = ref["loc"]
    out = orc.msda_core(value, [tuple(s) for s in meta["shapes"]], loc, w)
    assert relerr(out, ref["out"]) < 1e-5


def test_nearest_mask_matches_interpolate():
    torch.manual_seed(0)
    m = (torch.rand(2, 37, 53) > 0.3).long()
    for size in [(5, 7), (10, 14), (19, 27), (3, 4)]:
        ref = torch.nn.functional.interpolate(m[None].float(), size=size).to(torch.bool)[0]
        assert torch.equal(orc.nearest_mask(m, size), ref)


def test_both_msda_restatements_agree():
    from egtr_b200.synth import synth_msda_inputs

    shapes = [(11, 14), (6, 7), (3, 4), (2, 2)]
    value, spatial, start, loc, w = synth_msda_inputs(2, shapes, 23, seed=77)
    a = orc.msda_core(value, shapes, loc, w)
    b = orc.msda_core_grid_sample(value, shapes, loc, w)
    assert relerr(a, b) < 1e-5
