"""On-disk artifact formats: HF config.json + Lightning checkpoint -> model, as evaluate_egtr.py loads them."""
import pytest
import torch

from egtr_b200.checkpoint import latest_checkpoint, load_artifact, save_artifact, strip_lightning_prefix
from egtr_b200.config import DeformableDetrConfig, workload_config
from egtr_b200.synth import synth_images, synth_state_dict


def test_artifact_roundtrip_cpu(tmp_path):
    cfg = workload_config("tiny")
    cfg.logit_adjustment = False
    sd = synth_state_dict(cfg, 3)
    save_artifact(str(tmp_path), cfg, sd, epoch=2, step=10)
    save_artifact(str(tmp_path), cfg, {k: v * 0 for k, v in sd.items()}, epoch=1, step=5)  # older epoch must lose
    save_artifact(str(tmp_path), cfg, sd, epoch=11, step=99)
    assert latest_checkpoint(str(tmp_path)).endswith("epoch=11-step=99.ckpt")  # numeric, not lexicographic, order
    cfg2 = DeformableDetrConfig.from_pretrained(str(tmp_path))
    assert cfg2.to_dict() == cfg.to_dict()
    model = load_artifact(str(tmp_path), logit_adjustment=True, logit_adj_tau=0.5, device=None)
    assert model.config.logit_adjustment is True and model.config.logit_adj_tau == 0.5
    got = model.state_dict()
    assert set(got) == set(sd)
    for k in sd:
        assert torch.equal(got[k].cpu(), sd[k]), k
    with pytest.raises(KeyError):
        strip_lightning_prefix({"backbone.x": torch.zeros(1)})
    with pytest.raises(FileNotFoundError):
        latest_checkpoint(str(tmp_path / "nowhere"))


@pytest.mark.gpu
def test_artifact_model_matches_direct_load(cuda, tmp_path):
    from egtr_b200.model.egtr import DetrForSceneGraphGeneration
    cfg = workload_config("tiny")
    sd = synth_state_dict(cfg, 5)
    save_artifact(str(tmp_path), cfg, sd)
    px, mask = synth_images(1, 96, 128, seed=6)
    m1 = load_artifact(str(tmp_path), device="cuda")
    m2 = DetrForSceneGraphGeneration(cfg)
    m2.load_state_dict(sd)
    m2.cuda().eval()
    o1 = m1(pixel_values=px.to(cuda), pixel_mask=mask.to(cuda))
    o2 = m2(pixel_values=px.to(cuda), pixel_mask=mask.to(cuda))
    for k in ("logits", "pred_boxes", "pred_rel", "pred_connectivity"):
        assert torch.equal(o1[k], o2[k]), k
