"""CPU: the host-side operand layouts of the round-2 kernels, checked against the arithmetic they stand for in plain torch (fp64).

* `fused_decoder_matrices` (decoder.cu's `egtr_decoder_weights_t`): head-major q | k | v and offsets | logits rows, the `query_pos`
  halves of the position-added projections as per-layer row biases, the per-layer vector block — against
  `model/deformable_detr.py:1160-1168` (q scaled after the projection), `1404-1409` and `1040-1049`.
* `stem_weight_rows` (stem.cu): the window geometry the overlapping-row tensor map delivers (8 pixels x 4 channels starting at padded
  pixel 2*ox of padded row 2*oy + ky) times the laid-out weights = conv 7x7 / stride 2 / pad 3, for both image layouts."""
import torch
import torch.nn.functional as F

from egtr_b200.engine import fused_decoder_matrices, stem_weight_rows


def _decoder_sd(layers, d=256, ffn=1024, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g) * 0.1  # noqa: E731
    sd = {}
    for i in range(layers):
        p = f"model.decoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + f"self_attn.{n}.weight"], sd[p + f"self_attn.{n}.bias"] = r(d, d), r(d)
        sd[p + "encoder_attn.sampling_offsets.weight"], sd[p + "encoder_attn.sampling_offsets.bias"] = r(256, d), r(256)
        sd[p + "encoder_attn.attention_weights.weight"], sd[p + "encoder_attn.attention_weights.bias"] = r(128, d), r(128)
        sd[p + "encoder_attn.output_proj.weight"], sd[p + "encoder_attn.output_proj.bias"] = r(d, d), r(d)
        sd[p + "fc1.weight"], sd[p + "fc1.bias"], sd[p + "fc2.weight"], sd[p + "fc2.bias"] = r(ffn, d), r(ffn), r(d, ffn), r(d)
        for n in ("self_attn_layer_norm", "encoder_attn_layer_norm", "final_layer_norm"):
            sd[p + n + ".weight"], sd[p + n + ".bias"] = 1 + r(d), r(d)
    return sd


def test_fused_decoder_matrices_reproduce_the_position_added_projections():
    L, N, d, heads = 2, 11, 256, 8
    sd = _decoder_sd(L)
    g = torch.Generator().manual_seed(1)
    qpos, h, t1 = torch.randn(N, d, generator=g), torch.randn(N, d, generator=g), torch.randn(N, d, generator=g)
    m = fused_decoder_matrices(sd, L, qpos, heads)
    assert m["w_qkv"].shape == (L * 768, d) and m["w_offaw"].shape == (L * 384, d) and m["w_fc2"].shape == (L * d, 1024)
    assert m["qkv_pos"].shape == (L, N, 768) and m["off_pos"].shape == (L, N, 384) and m["vec"].shape == (L, 9 * 256 + 1024)
    for l in range(L):
        p = f"model.decoder.layers.{l}."
        hq = (h + qpos).double()
        want_q = (hq @ sd[p + "self_attn.q_proj.weight"].double().t() + sd[p + "self_attn.q_proj.bias"].double()) * (d // heads) ** -0.5
        want_k = hq @ sd[p + "self_attn.k_proj.weight"].double().t() + sd[p + "self_attn.k_proj.bias"].double()
        want_v = h.double() @ sd[p + "self_attn.v_proj.weight"].double().t() + sd[p + "self_attn.v_proj.bias"].double()
        tq = (t1 + qpos).double()
        want_off = tq @ sd[p + "encoder_attn.sampling_offsets.weight"].double().t() + sd[p + "encoder_attn.sampling_offsets.bias"].double()
        want_aw = tq @ sd[p + "encoder_attn.attention_weights.weight"].double().t() + sd[p + "encoder_attn.attention_weights.bias"].double()
        for r in range(heads):  # what CTA r of the cluster computes: its 96 (48) weight rows times the plain activations + its row-bias columns
            cols = h.double() @ m["w_qkv"][l * 768 + r * 96: l * 768 + (r + 1) * 96].double().t()
            pos = m["qkv_pos"][l].double()
            assert torch.allclose(cols[:, 0:32] + pos[:, 32 * r: 32 * r + 32], want_q[:, 32 * r: 32 * r + 32], atol=1e-5)
            assert torch.allclose(cols[:, 32:64] + pos[:, 256 + 32 * r: 256 + 32 * r + 32], want_k[:, 32 * r: 32 * r + 32], atol=1e-5)
            assert torch.allclose(cols[:, 64:96] + pos[:, 512 + 32 * r: 512 + 32 * r + 32], want_v[:, 32 * r: 32 * r + 32], atol=1e-5)
            oc = t1.double() @ m["w_offaw"][l * 384 + r * 48: l * 384 + (r + 1) * 48].double().t()
            op = m["off_pos"][l].double()
            assert torch.allclose(oc[:, :32] + op[:, 32 * r: 32 * r + 32], want_off[:, 32 * r: 32 * r + 32], atol=1e-5)
            assert torch.allclose(oc[:, 32:] + op[:, 256 + 16 * r: 256 + 16 * r + 16], want_aw[:, 16 * r: 16 * r + 16], atol=1e-5)
        v = m["vec"][l]
        for off, key in ((0, "self_attn.out_proj.bias"), (256, "encoder_attn.output_proj.bias"), (512, "fc2.bias"), (768, "self_attn_layer_norm.weight"),
                         (1024, "self_attn_layer_norm.bias"), (1280, "encoder_attn_layer_norm.weight"), (1536, "encoder_attn_layer_norm.bias"),
                         (1792, "final_layer_norm.weight"), (2048, "final_layer_norm.bias"), (2304, "fc1.bias")):
            assert torch.equal(v[off: off + sd[p + key].numel()], sd[p + key])
        assert torch.equal(m["w_fc1"][l * 1024: (l + 1) * 1024], sd[p + "fc1.weight"]) and torch.equal(m["w_o"][l * d: (l + 1) * d], sd[p + "self_attn.out_proj.weight"])


def _windows(img_rows: torch.Tensor, OH: int, OW: int, elems_per_px: int, krow: int) -> torch.Tensor:
    """What the stem's tensor map delivers: for output pixel (oy, ox) and filter row ky, `krow` consecutive elements of padded row
    2*oy + ky starting at padded pixel 2*ox.  img_rows: [Hp, Wp * elems_per_px] (+ slack) -> [OH*OW, 7 * krow]."""
    flat = img_rows.reshape(-1)
    pitch = img_rows.shape[1]
    out = torch.zeros(OH * OW, 7 * krow, dtype=img_rows.dtype)
    for oy in range(OH):
        for ox in range(OW):
            for ky in range(7):
                s = (2 * oy + ky) * pitch + 2 * ox * elems_per_px
                seg = flat[s: s + krow]
                out[oy * OW + ox, ky * krow: ky * krow + seg.numel()] = seg
    return out


def test_stem_weight_rows_times_tensor_map_windows_is_the_convolution():
    g = torch.Generator().manual_seed(2)
    w = torch.randn(64, 3, 7, 7, generator=g) * 0.05
    for H, W in ((20, 26), (17, 23)):
        x = torch.randn(1, 3, H, W, generator=g)
        OH, OW = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
        want = F.conv2d(x.double(), w.double(), stride=2, padding=3)[0].permute(1, 2, 0).reshape(OH * OW, 64)
        Hp, Wp = H + 6, (W + 6 + 1) & ~1
        pad = torch.zeros(Hp, Wp, 4, dtype=torch.float64)
        pad[3: 3 + H, 3: 3 + W, :3] = x[0].permute(1, 2, 0).double()
        # layout 1: two planes hi / lo of [Hp][Wp][4]; the three split products sum to the fp32 product up to the dropped lo*lo term
        rows = stem_weight_rows(w, 1, 32).double()  # [64, 7*32]
        got = _windows(pad.reshape(Hp, Wp * 4), OH, OW, 4, 32) @ rows.t()
        assert torch.allclose(got, want, atol=1e-9), float((got - want).abs().max())
        # layout 2: one plane, hi | lo interleaved per pixel; set B13 (hi weights at hi and lo positions) + set B2 (lo weights at hi positions)
        hi = pad.float().to(torch.bfloat16)
        lo = (pad.float() - hi.float()).to(torch.bfloat16)
        inter = torch.cat([hi, lo], -1).double()  # [Hp, Wp, 8]
        sets = stem_weight_rows(w, 2, 64).double()  # [2, 64, 7*64]
        win = _windows(inter.reshape(Hp, Wp * 8), OH, OW, 8, 64)
        got2 = win @ sets[0].t() + win @ sets[1].t()
        err = float((got2 - want).abs().max() / want.abs().max())
        assert err < 5e-5, err  # bf16x3: everything but the lo*lo term (2^-16 relative)
