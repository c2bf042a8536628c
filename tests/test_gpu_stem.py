"""GPU: the TMA-fed stem (stem.cu: 7x7/2 convolution + folded FrozenBN + ReLU over two bf16 planes of the image, overlapping-row
tensor map) against the gather-GEMM it replaces and against a plain fp64 convolution."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import relerr

pytestmark = pytest.mark.gpu


def _engines(cuda, monkeypatch, workload="tiny"):
    from egtr_b200.config import workload_config
    from egtr_b200.engine import Engine
    from egtr_b200.synth import synth_state_dict
    cfg = workload_config(workload)
    sd = synth_state_dict(cfg, 60)
    monkeypatch.setenv("EGTR_STEM", "gather")
    old = Engine(cfg, sd, cuda)
    monkeypatch.setenv("EGTR_STEM", "tma")
    new = Engine(cfg, sd, cuda)
    assert old.stem_mode == "gather" and new.stem_mode == "tma"
    return cfg, sd, old, new


@pytest.mark.parametrize("batch,H,W", [(1, 96, 128), (2, 97, 131), (1, 64, 70), (1, 800, 1333), (3, 130, 258)])
def test_stem_matches_gather_gemm_and_fp64_conv(cuda, monkeypatch, batch, H, W):
    from egtr_b200.engine import _fold_bn
    from egtr_b200.synth import synth_images
    cfg, sd, old, new = _engines(cuda, monkeypatch)
    px, mask = synth_images(batch, H, W, seed=61)
    px, mask = px.to(cuda), mask.to(cuda)
    old.forward(px, mask)
    new.forward(px, mask)
    torch.cuda.synchronize()
    a = old._workspace(batch, H, W)["stem"].view(batch, -1)
    b = new._workspace(batch, H, W)["stem"].view(batch, -1)
    bb = "model.backbone.conv_encoder.model."
    w, shift = _fold_bn({k: v.double() for k, v in sd.items() if k.startswith(bb + "conv1") or k.startswith(bb + "bn1")}, bb + "conv1", bb + "bn1")
    want = F.relu(F.conv2d(px.double().cpu(), w, shift, stride=2, padding=3)).permute(0, 2, 3, 1).reshape(batch, -1)
    e_old, e_new, e_pair = relerr(a, want), relerr(b, want), relerr(b, a)
    print(batch, H, W, f"gather vs fp64 {e_old:.1e}  tma vs fp64 {e_new:.1e}  tma vs gather {e_pair:.1e}")
    assert e_new < 2e-5 and e_pair < 2e-5


def test_forward_with_tma_stem_matches_gather_stem(cuda, monkeypatch):
    from egtr_b200.synth import synth_images
    from tests.util import TOL, compare_forward
    cfg, sd, old, new = _engines(cuda, monkeypatch, "small")
    px, mask = synth_images(2, 160, 224, seed=62, pad_to=[(160, 224), (120, 190)])
    want = old.forward(px.to(cuda), mask.to(cuda))
    got = new.forward(px.to(cuda), mask.to(cuda))
    torch.cuda.synchronize()
    errs = compare_forward(got, {k: v for k, v in want.items() if isinstance(v, torch.Tensor)})
    print({k: f"{v:.1e}" for k, v in errs.items()})
    assert max(errs.values()) < TOL
