"""GPU: the whole hot path through the reference-facing model API, vs golden fixtures (made by the
unmodified reference) and vs the CPU oracle run live on the same seeded inputs."""
import os

import pytest
import torch

from tests.util import TOL, case_inputs, compare_forward, load_golden

pytestmark = pytest.mark.gpu

CASES = ["forward_tiny_b1", "forward_tiny_b2_ragged", "forward_small_b2_ragged_logitadj", "forward_small_nofreq", "forward_A",
         "forward_B", "forward_D_small_b2_ragged", "forward_E_small"]  # B: BASELINE.json configs[1] at full size, from the reference itself


def _run(cfg, sd, px, mask, cuda):
    from egtr_b200.model.egtr import DetrForSceneGraphGeneration
    model = DetrForSceneGraphGeneration(cfg)
    model.load_state_dict(sd)
    model.cuda().eval()
    out = model(pixel_values=px.to(cuda), pixel_mask=mask.to(cuda), output_attentions=False,
                output_attention_states=True, output_hidden_states=True)
    torch.cuda.synchronize()
    return model, out


@pytest.mark.parametrize("throughput", [False, True])  # lone forward (split-K on) / one of several in flight (split-K off)
@pytest.mark.parametrize("name", CASES)
def test_forward_matches_reference_golden(cuda, name, throughput):
    ref, meta = load_golden(name)
    cfg, sd, px, mask = case_inputs(meta)
    model, out = _run(cfg, sd, px, mask, cuda)
    if throughput:
        out = model.engine().forward(px.to(cuda), mask.to(cuda), throughput=True)
        torch.cuda.synchronize()
    assert out["logits"].shape == (meta["batch"], cfg.num_queries, cfg.num_labels)
    assert "pred_connectivity" in out and out["pred_rel"].shape[-1] == cfg.num_rel_labels
    errs = compare_forward(out, ref)
    print(name, os.environ.get("EGTR_B200_GEMM", "tc"), {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < TOL, errs


def test_forward_matches_live_oracle_with_taps(cuda):
    """Stage-by-stage check on a ragged batch: where would a deviation first appear?"""
    from egtr_b200.config import workload_config
    from egtr_b200.engine import Engine
    from egtr_b200.synth import synth_images, synth_state_dict
    from oracle import egtr_oracle as orc
    from tests.util import relerr
    cfg = workload_config("small")
    sd = synth_state_dict(cfg, 40)
    px, mask = synth_images(2, 160, 224, seed=41, pad_to=[(160, 224), (120, 190)])
    taps_o, taps_g = {}, {}
    want = orc.forward(sd, cfg, px, mask, taps=taps_o)
    eng = Engine(cfg, sd, cuda)
    got = eng.forward(px.to(cuda), mask.to(cuda), taps=taps_g)
    torch.cuda.synchronize()
    stage = {}
    for k in ("c3", "c4", "c5", "source_flatten", "valid_ratios", "enc0_out"):
        stage[k] = relerr(taps_g[k], taps_o[k])
    m = taps_o["mask_flatten"][..., None]  # position embeddings are compared on real (unmasked) tokens
    stage["lvl_pos_embed_flatten"] = relerr(taps_g["lvl_pos_embed_flatten"].cpu() * m, taps_o["lvl_pos_embed_flatten"] * m)
    assert torch.equal(taps_g["mask_flatten"].cpu(), taps_o["mask_flatten"])
    for i in (0, 5):
        stage[f"q{i}"] = relerr(got["decoder_attention_queries"][i], want["decoder_attention_queries"][i])
        stage[f"k{i}"] = relerr(got["decoder_attention_keys"][i], want["decoder_attention_keys"][i])
    stage["intermediate"] = relerr(got["intermediate_hidden_states"], want["intermediate_hidden_states"])
    stage["init_ref"] = relerr(got["init_reference_points"], want["init_reference_points"])
    errs = compare_forward(got, want)
    print({k: f"{v:.2e}" for k, v in {**stage, **errs}.items()})
    assert max(stage.values()) < TOL and max(errs.values()) < TOL, (stage, errs)


def test_model_api_contract(cuda):
    from egtr_b200.config import workload_config
    from egtr_b200.model.egtr import DetrForSceneGraphGeneration
    from egtr_b200.synth import synth_images, synth_state_dict
    cfg = workload_config("tiny")
    model = DetrForSceneGraphGeneration.from_pretrained("SenseTime/deformable-detr", config=cfg, ignore_mismatched_sizes=True)
    model.load_state_dict(synth_state_dict(cfg, 1))
    model.cuda()
    model.eval()
    assert model.device.type == "cuda"
    px, mask = synth_images(1, 96, 128)
    out = model(px.cuda(), mask.cuda(), output_attention_states=True, output_hidden_states=True)
    out2 = model(px.cuda())  # pixel_mask defaults to all ones (deformable_detr.py:2208-2211)
    assert torch.equal(out.logits, out2.logits) and torch.equal(out.pred_rel, out2.pred_rel)
    assert len(out.decoder_hidden_states) == cfg.decoder_layers + 1
    tup = model(px.cuda(), return_dict=False)
    assert isinstance(tup, tuple) and torch.equal(tup[0], out.logits)
    with pytest.raises(NotImplementedError):
        model(px.cuda(), labels=[{}])


def test_pipelined_runner_matches_direct_forward(cuda):
    """Host-in / host-out pipelining over three streams returns exactly what the plain forward returns."""
    from egtr_b200.config import workload_config
    from egtr_b200.model.egtr import DetrForSceneGraphGeneration
    from egtr_b200.serving import PipelinedRunner
    from egtr_b200.synth import synth_images, synth_state_dict
    cfg = workload_config("tiny")
    model = DetrForSceneGraphGeneration(cfg)
    model.load_state_dict(synth_state_dict(cfg, 3))
    model.cuda().eval()
    batches = [synth_images(2, 96, 128, seed=100 + i, pad_to=[(96, 128), (64 + 8 * i, 100)]) for i in range(5)]
    want = []
    for px, pm in batches:
        o = model(px.cuda(), pm.cuda())
        want.append({k: o[k].cpu() for k in ("logits", "pred_boxes", "pred_rel", "pred_connectivity")})
    pipe = PipelinedRunner(model, 2, 96, 128, depth=2)
    got = list(pipe.run([(px.pin_memory(), pm.pin_memory()) for px, pm in batches]))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        for k in w:
            assert torch.equal(g[k], w[k]), k
    model.use_cuda_graph = True  # the single-call graph option of the model API gives the same answers too
    o = model(batches[0][0].cuda(), batches[0][1].cuda())
    assert torch.equal(o.pred_rel.cpu(), want[0]["pred_rel"])


@pytest.mark.parametrize("wl,batch", [("tiny", 2), ("small", 1), ("A", 1)])
def test_fused_relation_stage_matches_unfused_kernels(cuda, wl, batch, monkeypatch):
    """tc backend (pair-gating producer + dot/finish epilogues) vs the simt cross-check path (separate kernels)."""
    from egtr_b200.config import WORKLOADS, workload_config
    from egtr_b200.engine import Engine
    from egtr_b200.synth import synth_images, synth_state_dict
    from tests.util import relerr
    cfg = workload_config(wl, logit_adjustment=(wl == "small"))
    H, W = WORKLOADS[wl]["image"]
    sd = synth_state_dict(cfg, 77)
    px, mask = synth_images(batch, H, W, seed=78)
    monkeypatch.setenv("EGTR_B200_GEMM", "tc")
    got = Engine(cfg, sd, cuda).forward(px.to(cuda), mask.to(cuda))
    monkeypatch.setenv("EGTR_B200_GEMM", "simt")
    want = Engine(cfg, sd, cuda).forward(px.to(cuda), mask.to(cuda))
    torch.cuda.synchronize()
    # (the unfused cross-check path keeps `value` in fp32 rows; the product path stores it as fp16 pair records: 5e-4, not 2e-4)
    from tests.util import forward_errors, worst
    errs = forward_errors(got, want)
    print(wl, {k: f"{v:.2e}" for k, v in errs.items()})
    assert worst(errs) < 5e-4, errs


@pytest.mark.parametrize("wl,hw,batch", [("D", (224, 320), 2), ("E", (256, 256), 2)])
def test_forward_other_label_spaces_vs_live_oracle(cuda, wl, hw, batch):
    """BASELINE.json configs D (601 classes / 30 predicates, N_q=200) and E (N_q=300, 200 predicates) at a reduced image size the
    CPU oracle finishes in seconds: query / class / predicate counts are the full ones, so every N-, K- and P-dependent
    kernel shape (pair tiles, padded predicate columns, frequency-bias gather) is the production one."""
    from egtr_b200.config import workload_config
    from egtr_b200.synth import synth_images, synth_state_dict
    from oracle import egtr_oracle as orc
    cfg = workload_config(wl)
    sd = synth_state_dict(cfg, 50)
    px, mask = synth_images(batch, hw[0], hw[1], seed=51, pad_to=[hw, (hw[0] - 32, hw[1] - 48)][:batch])
    want = orc.forward(sd, cfg, px, mask)
    _, out = _run(cfg, sd, px, mask, cuda)
    errs = compare_forward(out, want)
    print(wl, {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < TOL, errs


FULL = {  # BASELINE.json configs at their full per-GPU size: (workload, per-GPU batch, smaller image of the ragged batch)
    "B": ("B", 1, None),                 # configs[1]: VG, batch 1
    "C": ("B", 4, (736, 1216)),          # configs[2]: VG, 32 images over 8 GPUs
    "D": ("D", 4, (768, 1024)),          # configs[3]: Open Images label space, 16 images over 4 GPUs
    "E": ("E", 8, (960, 992)),           # configs[4]: stress, N_q=300 / 200 predicates, 64 images over 8 GPUs
}


@pytest.mark.parametrize("name", list(FULL))
def test_forward_full_size_configs(cuda, name):
    """Full-size parity.  One image of the batch (the ragged one, so pixel_mask has zeros) is checked against the live CPU oracle;
    the rest of the batch through the size-independent property the domain offers — images are independent units,
    so a batched forward must reproduce the forward of each image alone (same padded tensor, same mask)."""
    from egtr_b200.config import WORKLOADS, workload_config
    from egtr_b200.synth import synth_images, synth_state_dict
    from oracle import egtr_oracle as orc
    from tests.util import relerr
    wl, batch, small = FULL[name]
    cfg = workload_config(wl)
    H, W = WORKLOADS[wl]["image"]
    sd = synth_state_dict(cfg, 60)
    pad = None if small is None else [(H, W)] * (batch - 1) + [small]
    px, mask = synth_images(batch, H, W, seed=61, pad_to=pad)
    model, out = _run(cfg, sd, px, mask, cuda)
    keys = ("logits", "pred_boxes", "pred_rel", "pred_connectivity")
    for k in keys:
        assert torch.isfinite(out[k]).all(), k
    assert out["pred_rel"].shape == (batch, cfg.num_queries, cfg.num_queries, cfg.num_rel_labels)
    assert float(out["pred_rel"].min()) >= 0.0 and float(out["pred_rel"].max()) <= 1.0
    last = batch - 1
    want = orc.forward(sd, cfg, px[last:], mask[last:])
    from tests.util import forward_errors, worst
    errs = forward_errors({k: out[k][last:] for k in keys}, want, keys)
    print(name, "vs oracle", {k: f"{v:.2e}" for k, v in errs.items()})
    assert worst(errs) < TOL, errs
    if batch > 1:
        for b in sorted({0, last}):
            alone = model(pixel_values=px[b:b + 1].to(cuda), pixel_mask=mask[b:b + 1].to(cuda), output_attentions=False,
                          output_attention_states=True, output_hidden_states=True)
            torch.cuda.synchronize()
            inv = forward_errors({k: out[k][b:b + 1] for k in keys}, alone, keys)
            print(name, f"image {b} batched vs alone", {k: f"{v:.2e}" for k, v in inv.items()})
            # two runs of the same arithmetic on different tilings: fp32 summation-order noise, amplified where a `value` element
            # rounds to the other fp16 neighbour (EGTR_FMT_H16PAIR) — a fifth of the parity bar
            assert worst(inv) < TOL / 5, (b, inv)


def test_workspace_cache_is_bounded_and_eviction_is_safe(cuda, monkeypatch):
    """the per-shape workspace cache must not grow without bound over a dataset pass (every batch pads to its own
    H, W).  With a small byte cap, old shapes are evicted (their graph runners too) and re-created on demand with identical results;
    a runner held by the caller survives the eviction of its workspace."""
    from egtr_b200.config import workload_config
    from egtr_b200.engine import Engine, GraphRunner
    from egtr_b200.synth import synth_images, synth_state_dict
    cfg = workload_config("tiny")
    sd = synth_state_dict(cfg, 31)
    monkeypatch.setenv("EGTR_WS_CAP_GB", "0.02")  # ~20 MB: two or three tiny workspaces
    eng = Engine(cfg, sd, cuda)
    shapes = [(96, 128), (64, 96), (128, 96), (80, 112), (96, 160), (112, 112)]
    first = {}
    held = GraphRunner(eng, 1, 96, 128)  # captured on the (1, 96, 128, 0) workspace
    px0, pm0 = synth_images(1, 96, 128, seed=900)
    want_held = {k: v.clone() for k, v in held(px0.to(cuda), pm0.to(cuda)).items() if k in ("logits", "pred_rel")}
    for rnd in range(2):
        for i, (h, w) in enumerate(shapes):
            px, pm = synth_images(1, h, w, seed=900 + i)
            o = eng.forward(px.to(cuda), pm.to(cuda))
            torch.cuda.synchronize()
            if rnd == 0:
                first[(h, w)] = o["pred_rel"].clone()
            else:
                assert torch.equal(first[(h, w)], o["pred_rel"]), (h, w)
            assert eng.workspace_bytes() <= eng.ws_cap_bytes + max(eng._ws_bytes.values())
    assert len([k for k in eng._ws if k[0] != "graph"]) < len(shapes)  # something was evicted
    got = held(px0.to(cuda), pm0.to(cuda))
    torch.cuda.synchronize()
    for k in want_held:
        assert torch.equal(got[k], want_held[k]), k
