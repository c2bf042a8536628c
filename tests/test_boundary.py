"""The reference-facing boundary B1: `dropin/` on sys.path makes the reference's own import statements and
loading / FPS-loop code (`/root/reference/evaluate_egtr.py:21-36, 225-242`) resolve to this package."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# the statements of evaluate_egtr.py (imports 21-22, model construction 225-242, calculate_fps 26-36), verbatim in shape
DROPIN_SCRIPT = textwrap.dedent('''
    import sys, torch
    from glob import glob
    sys.path.insert(0, {dropin!r})
    from model.deformable_detr import DeformableDetrConfig, DeformableDetrFeatureExtractor
    from model.egtr import DetrForSceneGraphGeneration
    from lib.evaluation.coco_eval import CocoEvaluator   # evaluate_egtr.py:15-20
    from lib.evaluation.oi_eval import OIEvaluator
    from lib.evaluation.sg_eval import BasicSceneGraphEvaluator, calculate_mR_from_evaluator_list
    from lib.pytorch_misc import argsort_desc
    import model.egtr, egtr_b200.model.egtr
    assert model.egtr.DetrForSceneGraphGeneration is egtr_b200.model.egtr.DetrForSceneGraphGeneration
    artifact_path = {artifact!r}
    config = DeformableDetrConfig.from_pretrained(artifact_path)
    config.logit_adjustment = False
    config.logit_adj_tau = 0.3
    model = DetrForSceneGraphGeneration.from_pretrained("SenseTime/deformable-detr", config=config, ignore_mismatched_sizes=True)
    ckpt_path = sorted(glob(f"{{artifact_path}}/checkpoints/epoch=*.ckpt"), key=lambda x: int(x.split("epoch=")[1].split("-")[0]))[-1]
    state_dict = torch.load(ckpt_path, map_location="cpu")["state_dict"]
    for k in list(state_dict.keys()):
        state_dict[k[6:]] = state_dict.pop(k)  # "model."
    model.load_state_dict(state_dict)
    feature_extractor = DeformableDetrFeatureExtractor.from_pretrained("SenseTime/deformable-detr", size=96, max_size=128)
    if {gpu}:
        model.cuda()
        model.eval()
        from egtr_b200.synth import synth_images
        loader = []
        for i in range(3):
            px, pm = synth_images(2, 96, 128, seed=40 + i, pad_to=[(96, 128), (80, 100)])
            loader.append(feature_extractor.pad_and_create_pixel_mask([px[0], px[1, :, :80, :100]]))
            assert torch.equal(loader[-1]["pixel_mask"], pm)
        with torch.no_grad():
            for batch in loader:  # calculate_fps
                outputs = model(pixel_values=batch["pixel_values"].cuda(), pixel_mask=batch["pixel_mask"].cuda(), output_attentions=False,
                                output_attention_states=True, output_hidden_states=True)
        assert "pred_connectivity" in outputs and outputs["pred_rel"].shape == (2, 24, 24, 12) and outputs.logits.shape == (2, 24, 20)
        det = feature_extractor.post_process(outputs, torch.tensor([[96, 128], [80, 100]]).cuda())
        assert len(det) == 2 and det[0]["boxes"].shape == (100, 4)
        torch.save({{k: outputs[k].cpu() for k in ("logits", "pred_boxes", "pred_rel", "pred_connectivity")}}, {out!r})
    print("DROPIN-OK")
''')


def _run_dropin(tmp_path, gpu):
    from egtr_b200.checkpoint import save_artifact
    from egtr_b200.config import workload_config
    from egtr_b200.synth import synth_state_dict
    cfg = workload_config("tiny")
    sd = synth_state_dict(cfg, 21)
    save_artifact(str(tmp_path), cfg, sd, epoch=3)
    out = str(tmp_path / "out.pt")
    script = DROPIN_SCRIPT.format(dropin=os.path.join(ROOT, "dropin"), artifact=str(tmp_path), gpu=gpu, out=out)
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, cwd=str(tmp_path), env=env, timeout=600)
    assert r.returncode == 0 and "DROPIN-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    return cfg, sd, out


def test_dropin_import_path_and_artifact_loading(tmp_path):
    _run_dropin(tmp_path, gpu=False)


@pytest.mark.gpu
def test_dropin_fps_loop_matches_oracle(cuda, tmp_path):
    from egtr_b200.synth import synth_images
    from oracle import egtr_oracle as orc
    from tests.util import forward_errors, worst
    cfg, sd, out = _run_dropin(tmp_path, gpu=True)
    got = torch.load(out)
    px, pm = synth_images(2, 96, 128, seed=42, pad_to=[(96, 128), (80, 100)])  # the loop's last batch
    want = orc.forward(sd, cfg, px, pm)
    errs = forward_errors(got, want)
    assert worst(errs) < 1e-3, errs


def test_post_process_matches_plain_loops():
    from egtr_b200.model.deformable_detr import DeformableDetrFeatureExtractor
    from egtr_b200.model.outputs import DetrSceneGraphGenerationOutput
    g = torch.Generator().manual_seed(4)
    B, N, K = 2, 30, 11
    logits, boxes = torch.randn(B, N, K, generator=g), torch.rand(B, N, 4, generator=g)
    sizes = torch.tensor([[480, 640], [333, 500]])
    res = DeformableDetrFeatureExtractor().post_process(DetrSceneGraphGenerationOutput(logits=logits, pred_boxes=boxes), sizes)
    for b in range(B):
        p = 1 / (1 + np.exp(-logits[b].double().numpy()))
        order = np.argsort(-p.ravel(), kind="stable")[:100]
        q, lab = order // K, order % K
        assert np.array_equal(res[b]["labels"].numpy(), lab)
        assert np.allclose(res[b]["scores"].numpy(), p.ravel()[order], rtol=1e-6)
        h, w = sizes[b].tolist()
        cx, cy, bw, bh = boxes[b].double().numpy()[q].T
        want = np.stack([(cx - bw / 2) * w, (cy - bh / 2) * h, (cx + bw / 2) * w, (cy + bh / 2) * h], 1)
        assert np.allclose(res[b]["boxes"].numpy(), want, rtol=1e-5, atol=1e-4)
    with pytest.raises(ValueError):
        DeformableDetrFeatureExtractor().post_process(DetrSceneGraphGenerationOutput(logits=logits, pred_boxes=boxes), sizes[:1])


def test_deformable_detr_model_state_dict_is_the_model_subtree():
    from egtr_b200.config import workload_config
    from egtr_b200.model.deformable_detr import DeformableDetrModel
    from egtr_b200.synth import synth_state_dict
    cfg = workload_config("tiny")
    sd = synth_state_dict(cfg, 9)
    sub = {k[6:]: v for k, v in sd.items() if k.startswith("model.")}
    m = DeformableDetrModel(cfg)
    assert set(m.state_dict()) == set(sub)
    m.load_state_dict(sub, strict=True)
    with pytest.raises(Exception):
        m(torch.zeros(1, 3, 32, 32))  # CPU model: no fallback


@pytest.mark.gpu
def test_deformable_detr_model_forward_matches_oracle(cuda):
    from egtr_b200.config import workload_config
    from egtr_b200.model.deformable_detr import DeformableDetrModel
    from egtr_b200.synth import synth_images, synth_state_dict
    from oracle import egtr_oracle as orc
    from tests.util import relerr
    cfg = workload_config("tiny")
    sd = synth_state_dict(cfg, 9)
    m = DeformableDetrModel(cfg)
    m.load_state_dict({k[6:]: v for k, v in sd.items() if k.startswith("model.")})
    m.cuda().eval()
    px, pm = synth_images(2, 96, 128, seed=10, pad_to=[(96, 128), (64, 100)])
    o = m(pixel_values=px.to(cuda), pixel_mask=pm.to(cuda), output_attention_states=True, output_hidden_states=True)
    want = orc.forward(sd, cfg, px, pm)
    assert relerr(o.last_hidden_state, want["last_hidden_state"]) < 1e-3
    assert relerr(o.intermediate_hidden_states, want["intermediate_hidden_states"]) < 1e-3
    assert relerr(o.encoder_last_hidden_state, want["encoder_last_hidden_state"]) < 1e-3
    assert relerr(o.init_reference_points, want["init_reference_points"]) < 1e-5
    assert o.intermediate_reference_points.shape == (2, cfg.decoder_layers, cfg.num_queries, 2)
    assert len(o.decoder_attention_queries) == cfg.decoder_layers and len(o.decoder_hidden_states) == cfg.decoder_layers + 1
    for l in (0, cfg.decoder_layers - 1):
        assert relerr(o.decoder_attention_queries[l], want["decoder_attention_queries"][l]) < 1e-3
        assert relerr(o.decoder_attention_keys[l], want["decoder_attention_keys"][l]) < 1e-3
    assert "logits" not in o and "pred_rel" not in o
