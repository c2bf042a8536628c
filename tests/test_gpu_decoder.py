"""GPU: the one-kernel decoder stack (decoder.cu, `egtr_decoder_fused_f32`) against the layer-by-layer launch sequence it
replaces (itself pinned by the reference's goldens and the oracle's Q/K taps) and against the CPU oracle.

Reference semantics: model/deformable_detr.py:1390-1489 (layer), 1149-1262 (self-attention with Q/K capture), 1774-1968 (stack)."""
import ctypes as C

import pytest
import torch

from tests.util import TOL, compare_forward, relerr

pytestmark = pytest.mark.gpu


def _sync():
    """synchronize; a trap inside the decoder kernel (a barrier wait that timed out) is reported with its fault code"""
    from egtr_b200._lib import call
    try:
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        raise AssertionError(f"launch failed; decoder fault code {call('egtr_decoder_fault')}: {e}") from e


def _engine(cuda, workload, seed, mode, monkeypatch):
    from egtr_b200.config import workload_config
    from egtr_b200.engine import Engine
    from egtr_b200.synth import synth_state_dict
    monkeypatch.setenv("EGTR_DECODER", mode)
    cfg = workload_config(workload)
    sd = synth_state_dict(cfg, seed)
    return cfg, sd, Engine(cfg, sd, cuda)


# N = 40 (one 128-row tile, two key groups), N = 100, N = 200 (two tiles, seven key groups; one image and two)
# ... each N = 200 case on the 8-CTA cluster (forwards in flight) and on the 16-CTA cluster that splits the rows (a lone forward)
CASES = [("small", 2, (160, 224), [(160, 224), (120, 190)], 0), ("small", 1, (128, 160), None, 0), ("A", 1, (192, 256), None, 0),
         ("B", 1, (160, 224), None, 8), ("B", 1, (160, 224), None, 16), ("B", 2, (160, 224), [(160, 224), (128, 200)], 8),
         ("B", 2, (160, 224), [(160, 224), (128, 200)], 16)]


@pytest.mark.parametrize("workload,batch,hw,pad_to,cluster", CASES)
def test_fused_stack_matches_layer_sequence(cuda, monkeypatch, workload, batch, hw, pad_to, cluster):
    if cluster:
        monkeypatch.setenv("EGTR_DECODER_CLUSTER", str(cluster))
    """Every output of the decoder — the stacked intermediate states (per layer: where would a deviation first appear?), the
    captured self-attention queries / keys of every layer, and everything downstream of them."""
    from egtr_b200.synth import synth_images
    px, mask = synth_images(batch, hw[0], hw[1], seed=51, pad_to=pad_to)
    cfg, sd, eng_l = _engine(cuda, workload, 50, "layers", monkeypatch)
    want = eng_l.forward(px.to(cuda), mask.to(cuda))
    _sync()
    _, _, eng_f = _engine(cuda, workload, 50, "fused", monkeypatch)
    assert eng_f.dec_fused_ok
    got = eng_f.forward(px.to(cuda), mask.to(cuda))
    _sync()
    stage = {}
    nl = cfg.decoder_layers
    for l in range(nl):
        stage[f"q{l}"] = relerr(got["decoder_attention_queries"][l], want["decoder_attention_queries"][l])
        stage[f"k{l}"] = relerr(got["decoder_attention_keys"][l], want["decoder_attention_keys"][l])
        stage[f"h{l}"] = relerr(got["intermediate_hidden_states"][:, l], want["intermediate_hidden_states"][:, l])
    print(workload, batch, {k: f"{v:.1e}" for k, v in stage.items()})
    assert max(stage.values()) < TOL, stage
    errs = compare_forward(got, {k: v for k, v in want.items() if isinstance(v, torch.Tensor)})
    print(workload, batch, {k: f"{v:.1e}" for k, v in errs.items()})
    assert max(errs.values()) < TOL, errs
    # bit-identical across repeats (no atomics, no arrival-order dependence)
    again = eng_f.forward(px.to(cuda), mask.to(cuda))
    torch.cuda.synchronize()
    assert torch.equal(again["intermediate_hidden_states"], got["intermediate_hidden_states"])
    assert torch.equal(again["pred_rel"], got["pred_rel"])


def test_fused_stack_with_external_attention_core(cuda, monkeypatch):
    """Bring-up split of the kernel by phase ranges: [init, qkv] -> egtr_mha_core_f32 (the validated stand-alone kernel) ->
    [out_proj .. ln3] per layer.  Separates the GEMM / LayerNorm / MSDA phases from the in-kernel attention core."""
    from egtr_b200._lib import call
    from egtr_b200.engine import _ptr, _stream
    from egtr_b200.synth import synth_images
    px, mask = synth_images(1, 128, 160, seed=53)
    cfg, sd, eng_l = _engine(cuda, "small", 52, "layers", monkeypatch)
    want = eng_l.forward(px.to(cuda), mask.to(cuda))
    _, _, eng = _engine(cuda, "small", 52, "fused", monkeypatch)
    got_full = eng.forward(px.to(cuda), mask.to(cuda))  # fills dec_value_h16, valid_ratios, ... of the workspace
    torch.cuda.synchronize()
    B, _, H, W = px.shape
    ws = eng._workspace(B, H, W, 0)
    N, nl, S = cfg.num_queries, cfg.decoder_layers, ws["S"]
    Md = B * N
    f32 = dict(dtype=torch.float32, device=cuda)
    qkv_all, inter = torch.zeros(nl, Md, 768, **f32), torch.zeros(B, nl, N, 256, **f32)
    attn = torch.empty(Md, 256, **f32)
    scratch = ws["dec_scratch"]
    attn_p = scratch[7 * Md * 1024: 8 * Md * 1024]
    st = _stream()

    def launch(l, p0, p1):
        call("egtr_decoder_fused_f32", C.byref(eng.dec_fused_w), _ptr(scratch), _ptr(ws["dec_value_h16"]), B * S + 1, ws["shapes_c"], 4,
             _ptr(ws["valid_ratios"]), B, S, _ptr(qkv_all), _ptr(inter), l, l + 1, p0, p1, 0, st)

    for l in range(nl):
        launch(l, 0 if l == 0 else 1, 2)
        call("egtr_mha_core_f32", _ptr(qkv_all[l]), 768, B, N, 8, 32, _ptr(attn), st)
        call("egtr_rows_to_p32", _ptr(attn), None, Md, 256, 256, _ptr(attn_p), st)
        launch(l, 3, 12)
    _sync()
    errs = {f"h{l}": relerr(inter[:, l], want["intermediate_hidden_states"][:, l]) for l in range(nl)}
    errs.update({f"q{l}": relerr(qkv_all[l][:, :256], want["decoder_attention_queries"][l].permute(0, 2, 1, 3).reshape(Md, 256)) for l in range(nl)})
    print("external attention core:", {k: f"{v:.1e}" for k, v in errs.items()})
    assert max(errs.values()) < TOL, errs
    # and the in-kernel attention core agrees with the stand-alone one
    full = {f"h{l}": relerr(got_full["intermediate_hidden_states"][:, l], inter[:, l]) for l in range(nl)}
    print("in-kernel attention vs stand-alone:", {k: f"{v:.1e}" for k, v in full.items()})
    assert max(full.values()) < TOL, full


def test_fused_stack_matches_oracle(cuda, monkeypatch):
    from egtr_b200.synth import synth_images
    from oracle import egtr_oracle as orc
    px, mask = synth_images(2, 160, 224, seed=55, pad_to=[(160, 224), (120, 190)])
    cfg, sd, eng = _engine(cuda, "small", 54, "fused", monkeypatch)
    want = orc.forward(sd, cfg, px, mask)
    got = eng.forward(px.to(cuda), mask.to(cuda))
    torch.cuda.synchronize()
    stage = {"intermediate": relerr(got["intermediate_hidden_states"], want["intermediate_hidden_states"])}
    for i in (0, cfg.decoder_layers - 1):
        stage[f"q{i}"] = relerr(got["decoder_attention_queries"][i], want["decoder_attention_queries"][i])
        stage[f"k{i}"] = relerr(got["decoder_attention_keys"][i], want["decoder_attention_keys"][i])
    errs = compare_forward(got, want)
    print({k: f"{v:.1e}" for k, v in {**stage, **errs}.items()})
    assert max(stage.values()) < TOL and max(errs.values()) < TOL, (stage, errs)
