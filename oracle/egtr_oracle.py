"""CPU oracle for the EGTR inference hot path.  TEST INFRASTRUCTURE ONLY.

A functional fp32 restatement (torch CPU ops, no nn.Module tree) of the
reference forward `DetrForSceneGraphGeneration.forward`
(`/root/reference/model/egtr.py:241-540`) down through `DeformableDetrModel.forward`
(`/root/reference/model/deformable_detr.py:2161-2390`).  Each function cites the
reference lines it follows.  It exists to CHECK the CUDA path; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs
may import it.  The product (`egtr_b200/`) never does.

Parity pin: the reference ships no golden vectors, so this
oracle is pinned against the reference ITSELF, imported unmodified in the build
container through four import stubs (`tools/ref_harness.py`) — see
`tests/golden/make_golden.py`, which wrote `tests/golden/*.npz`, and
`tests/test_oracle.py`, which re-checks the oracle against those files everywhere.

Third-party arithmetic restated here (absent from /root/reference):
timm==0.5.4 `resnet50` (`requirements.txt:8`, call site `deformable_detr.py:749-755`):
ResNet-50 v1.5 — 7x7/2 stem, 3x3/2 max-pool, bottlenecks [3,4,6,3] with the
stride on the 3x3 conv, expansion 4, projection shortcut in the first block of
every stage.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


# --------------------------------------------------------------------------- backbone
def frozen_bn(x: Tensor, sd: StateDict, p: str) -> Tensor:
    """deformable_detr.py:704-714 (eps=1e-5, scale = w * rsqrt(var + eps))."""
    scale = sd[p + ".weight"] * torch.rsqrt(sd[p + ".running_var"] + 1e-5)
    shift = sd[p + ".bias"] - sd[p + ".running_mean"] * scale
    return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)


def _bottleneck(x: Tensor, sd: StateDict, p: str, stride: int) -> Tensor:
    y = F.relu(frozen_bn(F.conv2d(x, sd[p + "conv1.weight"]), sd, p + "bn1"))
    y = F.relu(frozen_bn(F.conv2d(y, sd[p + "conv2.weight"], stride=stride, padding=1), sd, p + "bn2"))
    y = frozen_bn(F.conv2d(y, sd[p + "conv3.weight"]), sd, p + "bn3")
    if (p + "downsample.0.weight") in sd:
        # stays nn.BatchNorm2d in the reference (replace_batch_norm skips Sequential
        # children, deformable_detr.py:718-730); eval-mode BN is the same affine map.
        x = frozen_bn(F.conv2d(x, sd[p + "downsample.0.weight"], stride=stride), sd, p + "downsample.1")
    return F.relu(y + x)


def resnet50_c3c4c5(sd: StateDict, x: Tensor, prefix: str = "model.backbone.conv_encoder.model.") -> List[Tensor]:
    """timm resnet50 features_only, out_indices (2,3,4) (deformable_detr.py:748-755, 778)."""
    x = F.relu(frozen_bn(F.conv2d(x, sd[prefix + "conv1.weight"], stride=2, padding=3), sd, prefix + "bn1"))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    feats = []
    for li, nblk in enumerate((3, 4, 6, 3), start=1):
        for bi in range(nblk):
            x = _bottleneck(x, sd, f"{prefix}layer{li}.{bi}.", 2 if (bi == 0 and li > 1) else 1)
        if li >= 2:
            feats.append(x)
    return feats


def nearest_mask(pixel_mask: Tensor, size: Tuple[int, int]) -> Tensor:
    """`F.interpolate(mask[None].float(), size)` then `.to(bool)` (deformable_detr.py:783-785,
    2235-2237): legacy 'nearest' picks src = floor(dst * in/out) with a float32 scale."""
    B, H, W = pixel_mask.shape
    h, w = size
    sy = torch.floor(torch.arange(h, dtype=torch.float32) * torch.tensor(H / h, dtype=torch.float32)).long().clamp_(max=H - 1)
    sx = torch.floor(torch.arange(w, dtype=torch.float32) * torch.tensor(W / w, dtype=torch.float32)).long().clamp_(max=W - 1)
    return pixel_mask[:, sy][:, :, sx].to(torch.bool)


def sine_position_embedding(mask: Tensor, d_model: int = 256) -> Tensor:
    """deformable_detr.py:850-876 with normalize=True, scale=2*pi, T=10000 (910-916)."""
    n = d_model // 2
    m = mask.to(torch.float32)
    y = m.cumsum(1)
    x = m.cumsum(2)
    y = (y - 0.5) / (y[:, -1:, :] + 1e-6) * (2 * math.pi)
    x = (x - 0.5) / (x[:, :, -1:] + 1e-6) * (2 * math.pi)
    i = torch.arange(n, dtype=torch.float32)
    dim_t = 10000.0 ** (2 * torch.div(i, 2, rounding_mode="trunc") / n)
    px = x[..., None] / dim_t
    py = y[..., None] / dim_t
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).flatten(3)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).flatten(3)
    return torch.cat((py, px), dim=3).permute(0, 3, 1, 2)


def _input_proj(sd: StateDict, level: int, x: Tensor, stride: int = 1, pad: int = 0) -> Tensor:
    """1x1 (or 3x3/2) conv + GroupNorm(32) (deformable_detr.py:1991-2011)."""
    p = f"model.input_proj.{level}."
    y = F.conv2d(x, sd[p + "0.weight"], sd[p + "0.bias"], stride=stride, padding=pad)
    return F.group_norm(y, 32, sd[p + "1.weight"], sd[p + "1.bias"], eps=1e-5)


# --------------------------------------------------------------------------- MSDeformAttn
def msda_core(value: Tensor, shapes: Sequence[Tuple[int, int]], loc: Tensor, weight: Tensor) -> Tensor:
    """K1, restated from the CUDA kernel's arithmetic
    (`model/custom_kernel/cuda/ms_deform_im2col_cuda.cuh:237-299`, bilinear 33-84):
    pixel coords `x = loc_x*W - 0.5`, a sample contributes only if
    `-1 < y < H and -1 < x < W`, out-of-range corners read as zero.
    value [B,S,M,D], loc [B,Lq,M,L,P,2] (x,y), weight [B,Lq,M,L,P] -> [B,Lq,M*D]."""
    B, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    out = torch.zeros(B, Lq, M, D, dtype=value.dtype)
    start = 0
    bidx = torch.arange(B).view(B, 1, 1, 1)
    midx = torch.arange(M).view(1, 1, M, 1)
    for l, (H, W) in enumerate(shapes):
        v = value[:, start : start + H * W]  # [B,HW,M,D]
        start += H * W
        xim = loc[:, :, :, l, :, 0] * W - 0.5  # [B,Lq,M,P]
        yim = loc[:, :, :, l, :, 1] * H - 0.5
        inside = (yim > -1) & (xim > -1) & (yim < H) & (xim < W)
        y0 = torch.floor(yim)
        x0 = torch.floor(xim)
        ly, lx = yim - y0, xim - x0
        hy, hx = 1 - ly, 1 - lx
        y0, x0 = y0.long(), x0.long()
        acc = torch.zeros(B, Lq, M, P, D, dtype=value.dtype)
        for dy, dx, cw in ((0, 0, hy * hx), (0, 1, hy * lx), (1, 0, ly * hx), (1, 1, ly * lx)):
            yy, xx = y0 + dy, x0 + dx
            ok = inside & (yy >= 0) & (yy <= H - 1) & (xx >= 0) & (xx <= W - 1)
            idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1))  # [B,Lq,M,P]
            g = v[bidx, idx, midx]  # [B,Lq,M,P,D]
            acc = acc + g * (cw * ok.to(value.dtype))[..., None]
        out = out + (acc * weight[:, :, :, l, :, None]).sum(3)
    return out.reshape(B, Lq, M * D)


def msda_core_grid_sample(value: Tensor, shapes: Sequence[Tuple[int, int]], loc: Tensor, weight: Tensor) -> Tensor:
    """The reference's own CPU path for K1 — `ms_deform_attn_core_pytorch`, deformable_detr.py:925-960:
    per level a bilinear `grid_sample(align_corners=False, padding_mode="zeros")` on grids 2*loc-1.  Same
    result as `msda_core` (tests/test_oracle.py); used by bench.py's CPU legs because it is what the
    reference actually executes on a CPU (the fallback at deformable_detr.py:1096-1101)."""
    B, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    grids = 2 * loc - 1
    start, sampled = 0, []
    for l, (H, W) in enumerate(shapes):
        v = value[:, start : start + H * W].flatten(2).transpose(1, 2).reshape(B * M, D, H, W)
        start += H * W
        g = grids[:, :, :, l].transpose(1, 2).flatten(0, 1)  # [B*M, Lq, P, 2]
        sampled.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    w = weight.transpose(1, 2).reshape(B * M, 1, Lq, L * P)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * w).sum(-1).view(B, M * D, Lq)
    return out.transpose(1, 2).contiguous()


MSDA_IMPL = {"gather": msda_core, "grid_sample": msda_core_grid_sample}
_msda_impl = "gather"


def set_msda_impl(name: str) -> None:
    """`gather` (default; the CUDA kernel's arithmetic, cuh:237-299) or `grid_sample` (the reference's CPU path)."""
    global _msda_impl
    assert name in MSDA_IMPL
    _msda_impl = name


def _linear(sd: StateDict, p: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def msda_module(sd, p, query, pos, enc, enc_mask, ref, shapes, heads, points, taps=None):
    """deformable_detr.py:1026-1104 (2-d reference-point branch 1066-1073)."""
    B, Lq, C = query.shape
    S = enc.shape[1]
    L = len(shapes)
    q = query if pos is None else query + pos
    value = _linear(sd, p + ".value_proj", enc)
    if enc_mask is not None:
        value = value.masked_fill(~enc_mask[..., None], 0.0)
    value = value.view(B, S, heads, C // heads)
    off = _linear(sd, p + ".sampling_offsets", q).view(B, Lq, heads, L, points, 2)
    aw = _linear(sd, p + ".attention_weights", q).view(B, Lq, heads, L * points)
    aw = F.softmax(aw, -1).view(B, Lq, heads, L, points)
    norm = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32)  # (W_l, H_l)
    loc = ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    core = MSDA_IMPL[_msda_impl](value, shapes, loc, aw)
    if taps is not None:
        taps.update(value=value, sampling_locations=loc, attention_weights=aw, core=core)
    return _linear(sd, p + ".output_proj", core)


# --------------------------------------------------------------------------- encoder / decoder
def valid_ratio(mask: Tensor) -> Tensor:
    """deformable_detr.py:2064-2073 — (w_ratio, h_ratio)."""
    _, h, w = mask.shape
    vh = mask[:, :, 0].sum(1).float() / h
    vw = mask[:, 0, :].sum(1).float() / w
    return torch.stack([vw, vh], -1)


def encoder_reference_points(shapes, valid_ratios: Tensor) -> Tensor:
    """deformable_detr.py:1616-1648."""
    refs = []
    for l, (H, W) in enumerate(shapes):
        ry, rx = torch.meshgrid(
            torch.linspace(0.5, H - 0.5, H, dtype=torch.float32),
            torch.linspace(0.5, W - 0.5, W, dtype=torch.float32),
            indexing="ij",
        )
        ry = ry.reshape(-1)[None] / (valid_ratios[:, None, l, 1] * H)
        rx = rx.reshape(-1)[None] / (valid_ratios[:, None, l, 0] * W)
        refs.append(torch.stack((rx, ry), -1))
    ref = torch.cat(refs, 1)
    return ref[:, :, None] * valid_ratios[:, None]


def encoder_layer(sd, p, x, pos, mask, ref, shapes, cfg, taps=None):
    """deformable_detr.py:1283-1358 (eval: dropout inert)."""
    a = msda_module(sd, p + "self_attn", x, pos, x, mask, ref, shapes, cfg.encoder_attention_heads, cfg.encoder_n_points, taps)
    x = F.layer_norm(x + a, (x.shape[-1],), sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"])
    f = _linear(sd, p + "fc2", F.relu(_linear(sd, p + "fc1", x)))
    return F.layer_norm(x + f, (x.shape[-1],), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"])


def decoder_self_attention(sd, p, h, pos, heads):
    """deformable_detr.py:1149-1262; returns (out, q_scaled[B,M,N,D], k[B,M,N,D])."""
    B, N, C = h.shape
    D = C // heads
    hp = h + pos
    q = _linear(sd, p + ".q_proj", hp) * (D ** -0.5)
    k = _linear(sd, p + ".k_proj", hp)
    v = _linear(sd, p + ".v_proj", h)  # value has no position embedding (1168)
    q = q.view(B, N, heads, D).transpose(1, 2)
    k = k.view(B, N, heads, D).transpose(1, 2)
    v = v.view(B, N, heads, D).transpose(1, 2)
    att = F.softmax(q @ k.transpose(-1, -2), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, N, C)
    return _linear(sd, p + ".out_proj", o), q.contiguous(), k.contiguous()


def decoder_layer(sd, p, h, pos, ref_in, enc, enc_mask, shapes, cfg, taps=None):
    """deformable_detr.py:1390-1489."""
    C = h.shape[-1]
    a, q, k = decoder_self_attention(sd, p + "self_attn", h, pos, cfg.decoder_attention_heads)
    h = F.layer_norm(h + a, (C,), sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"])
    c = msda_module(sd, p + "encoder_attn", h, pos, enc, enc_mask, ref_in, shapes, cfg.decoder_attention_heads, cfg.decoder_n_points, taps)
    h = F.layer_norm(h + c, (C,), sd[p + "encoder_attn_layer_norm.weight"], sd[p + "encoder_attn_layer_norm.bias"])
    f = _linear(sd, p + "fc2", F.relu(_linear(sd, p + "fc1", h)))
    h = F.layer_norm(h + f, (C,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"])
    return h, q, k


def mlp3(sd, p, x):
    """DeformableDetrMLPPredictionHead, 3 layers, ReLU between (deformable_detr.py:2864-2883)."""
    x = F.relu(_linear(sd, p + ".layers.0", x))
    x = F.relu(_linear(sd, p + ".layers.1", x))
    return _linear(sd, p + ".layers.2", x)


def inverse_sigmoid(x: Tensor, eps: float = 1e-5) -> Tensor:
    """deformable_detr.py:658-662."""
    x = x.clamp(0, 1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


# --------------------------------------------------------------------------- relation head
def relation_head(sd, cfg, queries: Sequence[Tensor], keys: Sequence[Tensor], h_last: Tensor, logits: Tensor,
                  row_chunk: int = 16) -> Tuple[Tensor, Tensor]:
    """egtr.py:322-418, 507-516.  queries/keys: per decoder layer [B,M,N,D] (queries are the
    SCALED q captured at deformable_detr.py:1179-1185).  i = subject = query side, j = object =
    key side.  The N x N x 7 x 512 pair tensor is built `row_chunk` subject rows at a time; the
    arithmetic per pair is the reference's, in the reference's order."""
    B, M, N, D = queries[0].shape
    C = M * D
    unscale = D ** 0.5
    pq = [_linear(sd, f"proj_q.{l}", q.transpose(1, 2).reshape(B, N, C) * unscale) for l, q in enumerate(queries)]
    pk = [_linear(sd, f"proj_k.{l}", k.transpose(1, 2).reshape(B, N, C)) for l, k in enumerate(keys)]
    sub = torch.stack(pq + [_linear(sd, "final_sub_proj", h_last)], dim=2)  # [B,N,7,C]
    obj = torch.stack(pk + [_linear(sd, "final_obj_proj", h_last)], dim=2)  # [B,N,7,C]
    Lr = sub.shape[2]
    P = cfg.num_rel_labels
    pred_rel = torch.empty(B, N, N, P)
    pred_con = torch.empty(B, N, N, 1)
    for i0 in range(0, N, row_chunk):
        i1 = min(N, i0 + row_chunk)
        src = torch.cat(
            [sub[:, i0:i1, None].expand(B, i1 - i0, N, Lr, C), obj[:, None, :].expand(B, i1 - i0, N, Lr, C)], dim=-1
        )  # [B,r,N,7,2C]
        gate = torch.sigmoid(_linear(sd, "rel_predictor_gate", src))
        gated = (gate * src).sum(dim=-2)
        pred_rel[:, i0:i1] = mlp3(sd, "rel_predictor", gated)
        pred_con[:, i0:i1] = mlp3(sd, "connectivity_layer", gated)
    if cfg.use_freq_bias:  # egtr.py:405-413
        node = torch.argmax(logits, dim=-1)
        for b in range(B):
            pred_rel[b] += sd["triplet_dist"][node[b]][:, node[b]]
    if cfg.logit_adjustment:  # egtr.py:509-512
        pred_rel = pred_rel - cfg.logit_adj_tau * sd["rel_dist"].log()
    return pred_rel.sigmoid(), pred_con.sigmoid()


# --------------------------------------------------------------------------- whole forward
@torch.no_grad()
def forward(sd: StateDict, cfg, pixel_values: Tensor, pixel_mask: Optional[Tensor] = None,
            taps: Optional[dict] = None, relation_row_chunk: int = 16) -> Dict[str, Tensor]:
    """DetrForSceneGraphGeneration.forward at inference (labels=None)."""
    B, _, H, W = pixel_values.shape
    d = cfg.d_model
    if pixel_mask is None:
        pixel_mask = torch.ones(B, H, W, dtype=torch.long)
    feats = resnet50_c3c4c5(sd, pixel_values)
    if taps is not None:
        taps["c3"], taps["c4"], taps["c5"] = feats

    # input_proj + masks + position embeddings (deformable_detr.py:2216-2241)
    sources, masks = [], []
    for l, f in enumerate(feats):
        sources.append(_input_proj(sd, l, f))
        masks.append(nearest_mask(pixel_mask, f.shape[-2:]))
    for l in range(len(feats), cfg.num_feature_levels):
        src = _input_proj(sd, l, feats[-1] if l == len(feats) else sources[-1], stride=2, pad=1)
        sources.append(src)
        masks.append(nearest_mask(pixel_mask, src.shape[-2:]))
    poss = [sine_position_embedding(m, d) for m in masks]

    # flatten (2249-2278)
    shapes = [tuple(s.shape[-2:]) for s in sources]
    src_flat = torch.cat([s.flatten(2).transpose(1, 2) for s in sources], 1)
    mask_flat = torch.cat([m.flatten(1) for m in masks], 1)
    pos_flat = torch.cat(
        [p.flatten(2).transpose(1, 2) + sd["model.level_embed"][l].view(1, 1, -1) for l, p in enumerate(poss)], 1
    )
    vr = torch.stack([valid_ratio(m) for m in masks], 1).float()
    if taps is not None:
        taps.update(source_flatten=src_flat, lvl_pos_embed_flatten=pos_flat, mask_flatten=mask_flat, valid_ratios=vr)

    # encoder (1650-1744)
    ref = encoder_reference_points(shapes, vr)
    x = src_flat
    for i in range(cfg.encoder_layers):
        t = {} if (taps is not None and i == 0) else None
        x = encoder_layer(sd, f"model.encoder.layers.{i}.", x, pos_flat, mask_flat, ref, shapes, cfg, t)
        if t is not None:
            taps["enc0_msda"] = t
            taps["enc0_out"] = x
    enc = x

    # queries (2339-2343) and decoder (1774-1968; no box refine -> constant reference points)
    qpe = sd["model.query_position_embeddings.weight"]
    query_pos = qpe[:, :d].unsqueeze(0).expand(B, -1, -1)
    h = qpe[:, d:].unsqueeze(0).expand(B, -1, -1)
    ref_pts = _linear(sd, "model.reference_points", query_pos).sigmoid()
    ref_in = ref_pts[:, :, None] * vr[:, None]
    qs, ks, inter = [], [], []
    for i in range(cfg.decoder_layers):
        t = {} if (taps is not None and i == 0) else None
        h, q, k = decoder_layer(sd, f"model.decoder.layers.{i}.", h, query_pos, ref_in, enc, mask_flat, shapes, cfg, t)
        if t is not None:
            taps["dec0_msda"] = t
        qs.append(q)
        ks.append(k)
        inter.append(h)

    # detection heads (egtr.py:283-314): only the last level is returned
    logits = _linear(sd, f"class_embed.{cfg.decoder_layers - 1}", h)
    delta = mlp3(sd, f"bbox_embed.{cfg.decoder_layers - 1}", h)
    delta[..., :2] += inverse_sigmoid(ref_pts)
    boxes = delta.sigmoid()

    pred_rel, pred_con = relation_head(sd, cfg, qs, ks, h, logits, row_chunk=relation_row_chunk)
    return dict(
        logits=logits,
        pred_boxes=boxes,
        pred_rel=pred_rel,
        pred_connectivity=pred_con,
        last_hidden_state=h,
        encoder_last_hidden_state=enc,
        intermediate_hidden_states=torch.stack(inter, 1),
        init_reference_points=ref_pts,
        decoder_attention_queries=tuple(qs),
        decoder_attention_keys=tuple(ks),
    )
