"""Input staging: oracle pinned bit-exactly against the installed Pillow; host tap tables vs the oracle;
device kernels vs the oracle (GPU)."""
import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as po

SIZES = [(480, 640, 800, 1066), (375, 500, 800, 1066), (600, 800, 300, 400), (37, 53, 20, 29), (37, 53, 111, 160),
         (500, 333, 800, 533), (64, 48, 64, 48), (200, 300, 200, 150), (333, 500, 800, 1201)]


@pytest.mark.parametrize("h,w,oh,ow", SIZES)
def test_oracle_resize_is_bit_exact_vs_pillow(h, w, oh, ow):
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(h * 7 + w)
    im = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(im).resize((ow, oh), Image.BILINEAR))
    assert np.array_equal(po.pil_bilinear_resize(im, oh, ow), ref)


def test_target_size_rule():
    assert po.target_size(480, 640) == (800, 1066)
    assert po.target_size(500, 1500) == (444, 1332)   # longer side capped at max_size
    assert po.target_size(800, 1200) == (800, 1200)   # already at size
    assert po.target_size(640, 480) == (1066, 800)
    from egtr_b200.preprocess import target_size
    for hw in [(480, 640), (500, 1500), (800, 1200), (640, 480), (333, 500), (1024, 1024), (97, 1031)]:
        assert target_size(*hw) == po.target_size(*hw)
        assert target_size(*hw, size=600, max_size=1000) == po.target_size(*hw, size=600, max_size=1000)


@pytest.mark.parametrize("n_in,n_out", [(640, 1066), (500, 1066), (800, 400), (53, 29), (53, 160), (48, 48), (1500, 1332), (7, 3), (3, 11)])
def test_host_tap_tables_match_oracle(n_in, n_out):
    from egtr_b200.preprocess import tap_tables
    b, k, ks = tap_tables(n_in, n_out)
    if n_in == n_out:
        assert ks == 1 and (k == 1 << 22).all() and (b[:, 0] == np.arange(n_out)).all()
        return
    bo, ko, kso = po.bilinear_coeffs(n_in, n_out)
    assert ks == kso and np.array_equal(b, bo) and np.array_equal(k, ko)


def test_stage_batch_oracle_shapes():
    rng = np.random.default_rng(1)
    ims = [rng.integers(0, 256, (60, 80, 3), dtype=np.uint8), rng.integers(0, 256, (90, 50, 3), dtype=np.uint8)]
    px, mask, sizes = po.stage_batch(ims, size=96, max_size=160)
    assert sizes == [(96, 128), (160, 89)] and px.shape == (2, 3, 160, 128) and mask.shape == (2, 160, 128)
    assert mask[0, :96, :128].all() and not mask[0, 96:].any() and (px[0, :, 96:] == 0).all()
    assert mask[1, :, :89].all() and not mask[1, :, 89:].any()


@pytest.mark.gpu
def test_device_stager_matches_oracle_bit_exact(cuda):
    from egtr_b200.preprocess import DeviceImageStager
    rng = np.random.default_rng(2)
    ims = [rng.integers(0, 256, hw + (3,), dtype=np.uint8) for hw in [(120, 160), (150, 100), (97, 131), (200, 200)]]
    want_px, want_mask, want_sizes = po.stage_batch(ims, size=192, max_size=320)
    px, mask, sizes = DeviceImageStager(size=192, max_size=320, device=cuda).stage(ims)
    torch.cuda.synchronize()
    assert sizes == want_sizes
    assert torch.equal(mask.cpu(), torch.from_numpy(want_mask))
    assert torch.equal(px.cpu(), torch.from_numpy(want_px))  # integer resampling + IEEE fp32 normalisation: exact


@pytest.mark.gpu
def test_device_stager_workload_b_size(cuda):
    """A VG-sized image (480x640 -> 800x1066) through the stager feeds the model without further host work."""
    from egtr_b200.preprocess import DeviceImageStager
    rng = np.random.default_rng(3)
    im = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    px, mask, sizes = DeviceImageStager(device=cuda).stage([im])
    want_px, want_mask, _ = po.stage_batch([im])
    assert sizes == [(800, 1066)] and torch.equal(px.cpu(), torch.from_numpy(want_px)) and bool(mask.all())


def test_size_rule_and_constants_against_installed_transformers():
    """f2 remainder: transformers 4.18 — the version the reference pins — is not installed, but its successor is.
    The DETR size rule of the installed version equals ours whenever the `max_size` cap is not hit; when it is, newer versions
    derive the long side from the UNROUNDED short side (`raw_size`, a later upstream fix) while 4.18 — restated here — rounds the
    short side first.  The normalisation constants are the same objects in both."""
    tr = pytest.importorskip("transformers.models.detr.image_processing_detr")
    from transformers.utils.constants import IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD
    from egtr_b200.preprocess import IMAGE_MEAN, IMAGE_STD, target_size
    assert tuple(IMAGENET_DEFAULT_MEAN) == IMAGE_MEAN and tuple(IMAGENET_DEFAULT_STD) == IMAGE_STD
    assert tuple(IMAGENET_DEFAULT_MEAN) == tuple(po.IMAGE_MEAN) and tuple(IMAGENET_DEFAULT_STD) == tuple(po.IMAGE_STD)
    rng = np.random.default_rng(4)
    n_uncapped = n_capped = n_capped_equal = 0
    for _ in range(4000):
        h, w = int(rng.integers(50, 2000)), int(rng.integers(50, 2000))
        size, mx = int(rng.choice([480, 600, 800])), int(rng.choice([800, 1000, 1333]))
        theirs, ours = tr.get_size_with_aspect_ratio((h, w), size, mx), target_size(h, w, size, mx)
        assert ours == po.target_size(h, w, size, mx)
        if max(h, w) / min(h, w) * size > mx:
            n_capped += 1
            n_capped_equal += ours == theirs
            # same short side; the long sides differ by at most the rounding of the short side times the aspect ratio
            assert min(ours) == min(theirs) and abs(max(ours) - max(theirs)) <= 0.5 * max(h, w) / min(h, w) + 1
        else:
            n_uncapped += 1
            assert ours == theirs, (h, w, size, mx)
    assert n_uncapped > 1000 and n_capped > 1000
    # the VG evaluation size (800 / 1333) on the dataset's common 4:3 and 3:2 images never hits the cap
    for h, w in ((480, 640), (600, 800), (768, 1024), (500, 375), (333, 500)):
        assert target_size(h, w, 800, 1333) == tr.get_size_with_aspect_ratio((h, w), 800, 1333)
