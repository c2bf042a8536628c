"""Host-side driver of the EGTR forward: weight preparation and the launch sequence.

PyTorch is used for device memory, streams and the tensors handed back to the caller; every
arithmetic step of the path is a kernel of `libegtr_b200.so` reached through the C ABI
(`include/egtr_b200.h`).  The sequence mirrors `DeformableDetrModel.forward`
(`/root/reference/model/deformable_detr.py:2161-2390`) and the inference part of
`DetrForSceneGraphGeneration.forward` (`/root/reference/model/egtr.py:241-418, 507-540`).

Layouts: images are consumed NCHW as given; every activation after the stem is NHWC, i.e. a
row-major [rows, C] matrix — which is also the `[B, S, 256]` token layout of the transformer, so
`flatten(2).transpose(1, 2)` (deformable_detr.py:2259) costs nothing.
"""
from __future__ import annotations

import collections
import ctypes as C
import math
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib
from ._lib import ASrc, DecoderWeights, Epilogue, RelheadWeights, call

RESNET_BLOCKS = (3, 4, 6, 3)
SKINNY_M = 512  # GEMMs with at most this many rows may take the fp32 skinny kernel (decoder, detection heads)


def _skinny(M: int, N: int, groups: int) -> bool:
    """Latency-bound shapes: few rows AND few enough 16x64 tiles to fit about two waves of the 148 SMs."""
    return M <= SKINNY_M and groups * ((M + 15) // 16) * ((N + 63) // 64) <= 2 * 148


def _ptr(t: Optional[torch.Tensor], col: int = 0) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr() + col * t.element_size()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def gemm_backend() -> str:
    """`tc` (tcgen05, the product path) or `simt` — the fp32 CUDA-core kernel, a bring-up/debug knob
    used by tests to cross-check the tensor-core path.  Both are kernels of this library."""
    return os.environ.get("EGTR_B200_GEMM", "tc")


class Lin:
    """A prepared weight: fp32 [N,K] plus bf16 hi/lo planes [2,Npad,K] for the tensor-core GEMM."""

    def __init__(self, w: torch.Tensor, b: Optional[torch.Tensor], device):
        w = w.detach().to(device=device, dtype=torch.float32).contiguous()
        self.N, self.K = w.shape
        self.Npad = ((self.N + 63) // 64) * 64
        self.w = w
        self.b = None if b is None else b.detach().to(device=device, dtype=torch.float32).contiguous()
        self.planes = torch.empty(2 * self.Npad * self.K, dtype=torch.bfloat16, device=device)
        call("egtr_split_weight_bf16", _ptr(w), self.N, self.K, self.Npad, _ptr(self.planes), _stream())


class LinStack:
    """`groups` weights of identical [N,K] stacked for one grouped launch: planes [2][G*Npad][K], bias [G*Npad]."""

    def __init__(self, ws: List[torch.Tensor], bs: List[Optional[torch.Tensor]], device):
        self.G = len(ws)
        self.N, self.K = ws[0].shape
        self.Npad = ((self.N + 63) // 64) * 64
        self.plane_rows = self.G * self.Npad
        w = torch.zeros(self.plane_rows, self.K, dtype=torch.float32, device=device)
        b = torch.zeros(self.plane_rows, dtype=torch.float32, device=device)
        for g, (wg, bg) in enumerate(zip(ws, bs)):
            w[g * self.Npad: g * self.Npad + self.N] = wg.to(device=device, dtype=torch.float32)
            if bg is not None:
                b[g * self.Npad: g * self.Npad + self.N] = bg.to(device=device, dtype=torch.float32)
        self.w, self.b = w, b
        self.planes = torch.empty(2 * self.plane_rows * self.K, dtype=torch.bfloat16, device=device)
        call("egtr_split_weight_bf16", _ptr(w), self.plane_rows, self.K, self.plane_rows, _ptr(self.planes), _stream())
        self.n_base = (C.c_int * self.G)(*[g * self.Npad for g in range(self.G)])


def _conv_mat(w: torch.Tensor, k_pad: Optional[int] = None) -> torch.Tensor:
    """[Cout,Cin,KH,KW] -> [Cout, KH*KW*Cin] (k = (ky*KW + kx)*Cin + c), optionally zero-padded in K."""
    cout = w.shape[0]
    m = w.permute(0, 2, 3, 1).reshape(cout, -1)
    if k_pad is not None and k_pad > m.shape[1]:
        m = torch.cat([m, m.new_zeros(cout, k_pad - m.shape[1])], 1)
    return m.contiguous()


def _fold_bn(sd, conv: str, bn: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """FrozenBN folded into the conv: scale = w*rsqrt(var+1e-5) (deformable_detr.py:704-714)."""
    scale = sd[bn + ".weight"] * torch.rsqrt(sd[bn + ".running_var"] + 1e-5)
    shift = sd[bn + ".bias"] - sd[bn + ".running_mean"] * scale
    return sd[conv + ".weight"] * scale.view(-1, 1, 1, 1), shift


def conv_out(n: int, k: int, s: int, p: int) -> int:
    return (n + 2 * p - k) // s + 1


def level_shapes(H: int, W: int, num_levels: int = 4) -> List[Tuple[int, int]]:
    """Feature-map sizes of C3, C4, C5 and the extra stride-2 levels for an H x W input."""
    h, w = conv_out(H, 7, 2, 3), conv_out(W, 7, 2, 3)
    h, w = conv_out(h, 3, 2, 1), conv_out(w, 3, 2, 1)  # max-pool -> C2 resolution
    shapes = []
    for _ in range(3):
        h, w = conv_out(h, 3, 2, 1), conv_out(w, 3, 2, 1)
        shapes.append((h, w))
    for _ in range(num_levels - 3):
        h, w = conv_out(h, 3, 2, 1), conv_out(w, 3, 2, 1)
        shapes.append((h, w))
    return shapes[:num_levels]


def fused_decoder_matrices(sd: Dict[str, torch.Tensor], layers: int, query_pos: torch.Tensor, heads: int = 8) -> Dict[str, torch.Tensor]:
    """fp32 operands of the one-kernel decoder stack (decoder.cu, `egtr_decoder_weights_t`) from the reference's state-dict keys —
    pure torch, any device.  Every projection is stacked over the layers; the q | k | v rows are head-major (CTA r of the cluster
    owns head r: q_r | k_r | v_r, q pre-scaled by head_dim^-0.5 as `deformable_detr.py:1166` does after the projection); the
    offsets / logits rows likewise (head r: its 32 sampling-offset rows | its 16 attention-logit rows); the `query_pos` halves of
    `(h + query_pos) . W` (deformable_detr.py:1404-1409, 1040) are weight-only terms composed in fp64 into per-layer row biases in
    the STANDARD column order (`qkv_pos` [L, N, 3d]: q | k | v incl. biases, `off_pos` [L, N, 384])."""
    d = query_pos.shape[1]
    hd = d // heads
    scaling = hd ** -0.5
    qpos = query_pos.double()
    N = qpos.shape[0]
    wq, wo, woff, wout, w1, w2, vec, qkv_pos, off_pos = [], [], [], [], [], [], [], [], []
    for i in range(layers):
        p = f"model.decoder.layers.{i}."
        q_w, q_b = sd[p + "self_attn.q_proj.weight"] * scaling, sd[p + "self_attn.q_proj.bias"] * scaling
        k_w, k_b = sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.k_proj.bias"]
        v_w, v_b = sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"]
        wq.append(torch.stack([q_w.view(heads, hd, d), k_w.view(heads, hd, d), v_w.view(heads, hd, d)], 1).reshape(3 * d, d))
        qkv_pos.append(torch.cat([qpos @ q_w.double().t() + q_b.double(), qpos @ k_w.double().t() + k_b.double(),
                                  v_b.double().expand(N, -1)], 1).float())
        wo.append(sd[p + "self_attn.out_proj.weight"])
        ow = torch.cat([sd[p + "encoder_attn.sampling_offsets.weight"], sd[p + "encoder_attn.attention_weights.weight"]], 0)
        ob = torch.cat([sd[p + "encoder_attn.sampling_offsets.bias"], sd[p + "encoder_attn.attention_weights.bias"]], 0)
        n_off = sd[p + "encoder_attn.sampling_offsets.weight"].shape[0]  # heads * levels * points * 2
        po, pl = n_off // heads, (ow.shape[0] - n_off) // heads
        woff.append(torch.cat([torch.cat([ow[po * r: po * (r + 1)], ow[n_off + pl * r: n_off + pl * (r + 1)]], 0) for r in range(heads)], 0))
        off_pos.append((qpos @ ow.double().t() + ob.double()).float())
        wout.append(sd[p + "encoder_attn.output_proj.weight"])
        w1.append(sd[p + "fc1.weight"])
        w2.append(sd[p + "fc2.weight"])
        vec.append(torch.cat([sd[p + "self_attn.out_proj.bias"], sd[p + "encoder_attn.output_proj.bias"], sd[p + "fc2.bias"],
                              sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"],
                              sd[p + "encoder_attn_layer_norm.weight"], sd[p + "encoder_attn_layer_norm.bias"],
                              sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], sd[p + "fc1.bias"]]))
    cat = lambda ws: torch.cat(ws, 0).float().contiguous()  # noqa: E731
    return dict(w_qkv=cat(wq), w_o=cat(wo), w_offaw=cat(woff), w_out=cat(wout), w_fc1=cat(w1), w_fc2=cat(w2), vec=torch.stack(vec).float().contiguous(),
                qkv_pos=torch.stack(qkv_pos).contiguous(), off_pos=torch.stack(off_pos).contiguous())


def stem_weight_rows(w: torch.Tensor, layout: int, krow: int) -> torch.Tensor:
    """The stem's weights laid out like the operand rows its tensor map delivers (stem.cu, include/egtr_b200.h).  `w`: the BN-folded
    conv1 weight [64, 3, 7, 7] (fp32).  layout 1 -> fp32 [64, 7 * krow] with w[o][ky][4 * kx + c] (to be split into hi / lo planes);
    layout 2 (one image plane, hi | lo interleaved per pixel) -> bf16 [2, 64, 7 * 64]: set 0 = hi(w) at elements 8 * kx + c AND
    8 * kx + 4 + c, set 1 = lo(w) at 8 * kx + c; zeros elsewhere."""
    o = w.shape[0]
    w7 = torch.cat([w, w.new_zeros(o, 1, 7, 7)], 1).permute(0, 2, 3, 1).contiguous().float()  # [64, ky, kx, 4]
    if layout == 2:
        hi = w7.to(torch.bfloat16)
        lo = (w7 - hi.float()).to(torch.bfloat16)
        sets = torch.zeros(2, o, 7, 8, 8, dtype=torch.bfloat16, device=w.device)
        sets[0, :, :, :7, 0:4] = hi
        sets[0, :, :, :7, 4:8] = hi
        sets[1, :, :, :7, 0:4] = lo
        return sets.reshape(2, o, 7 * 64).contiguous()
    wt = torch.zeros(o, 7, krow, dtype=torch.float32, device=w.device)
    wt[:, :, :28] = w7.reshape(o, 7, 28)
    return wt.reshape(o, 7 * krow).contiguous()


class Engine:
    def __init__(self, config, state_dict: Dict[str, torch.Tensor], device, model_only: bool = False):
        """`model_only`: the bare `DeformableDetrModel` (backbone -> encoder -> decoder; `state_dict` holds the `model.*` keys
        only) — no detection or relation heads are prepared or run."""
        _lib.load()
        self.model_only = model_only
        self.cfg = config
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.EgtrError("the EGTR hot path runs on CUDA devices only (no CPU fallback)")
        c = config
        if c.two_stage or c.with_box_refine:
            raise _lib.EgtrError("two_stage / with_box_refine are outside the EGTR inference path (SURVEY.md §2.1)")
        if not (c.d_model == 256 and c.encoder_attention_heads == 8 and c.decoder_attention_heads == 8
                and c.num_feature_levels == 4 and c.encoder_n_points == 4 and c.decoder_n_points == 4
                and c.backbone == "resnet50" and not c.dilation and c.position_embedding_type == "sine"
                and c.activation_function == "relu"):
            raise _lib.EgtrError("kernels are built for the shipped EGTR architecture (d_model 256, 8 heads, 4 levels x 4 points, resnet50)")
        # activation workspaces (and the model API's graph runners) per input shape, least-recently-used first.  The reference's
        # evaluation loop pads every batch to its own max H, W, so a dataset pass sees hundreds of shapes at ~1 GB per 800x1333
        # image: the cache is capped in bytes (EGTR_WS_CAP_GB, default 48) and evicts whole shapes, oldest first.
        self._ws: "collections.OrderedDict[tuple, object]" = collections.OrderedDict()
        self._ws_bytes: Dict[tuple, int] = {}
        self.ws_cap_bytes = int(float(os.environ.get("EGTR_WS_CAP_GB", "48")) * (1 << 30))
        # forwards in flight (serving): each persistent GEMM takes a share of the SMs (egtr_set_grid_div(2)) so that GEMMs of
        # different images run side by side, +10 % at workload B.  Round 1 found a race in this mode (LayerNorm epilogue, a missing
        # proxy fence); fixed, and re-measured in round 2: 0 / 1188 full-size forwards deviate (profiles/r02_race_matrix.txt).
        # The share: 0 = by batch size — a quarter of the SMs at batch 1 (with sixteen images in flight: +5 % over halves with eight),
        # half of them for larger batches, whose launches are long enough (quarters measured -6 % at C, -24 % at E).
        self.throughput_grid_div = int(os.environ.get("EGTR_THROUGHPUT_GRID_DIV", "0"))
        self.throughput_splitk = int(os.environ.get("EGTR_THROUGHPUT_SPLITK", "1"))  # split-K cap of forwards in flight (1 = off)
        # MSDeformAttn `value` storage: "h16" = fp16 pair records written by the value_proj GEMM's epilogue (half the L1 wavefronts
        # of the bilinear gather, include/egtr_b200.h EGTR_FMT_H16PAIR), "f32" = fp32 rows [S, 256] (round 1; dev A/B)
        self.msda_value = os.environ.get("EGTR_MSDA_VALUE", "h16")
        # decoder stack for small query sets: "fused" (default) = ONE cluster kernel for all layers (decoder.cu: 16 CTAs for a lone
        # forward, 0.57 ms; 8 CTAs per image for forwards in flight), "layers" = the round-1/2 sequence of skinny CUDA-core GEMMs
        # (ten launches per layer over the whole GPU: 0.76 ms alone, 10 % of the step with eight forwards in flight; kept as the
        # cross-check of the fused kernel and for shapes it does not take)
        self.decoder_mode = os.environ.get("EGTR_DECODER", "fused")
        self.probe: Optional[Dict[str, list]] = None  # bench.py: name -> [(start_event, end_event), ...]
        self.probe_flops: Dict[str, int] = {}         # bench.py: name -> algorithmic FLOPs issued under that span
        with torch.cuda.device(self.device):
            self._prepare(state_dict)

    # ------------------------------------------------------------------ per-kernel timing (bench.py roofline)
    class _Span:
        def __init__(self, eng, name):
            self.eng, self.name = eng, name

        def __enter__(self):
            if self.eng.probe is not None:
                self.e0 = torch.cuda.Event(enable_timing=True)
                self.e0.record()

        def __exit__(self, *a):
            if self.eng.probe is not None:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record()
                self.eng.probe.setdefault(self.name, []).append((self.e0, e1))

    def span(self, name):
        return Engine._Span(self, name)

    # ------------------------------------------------------------------ weights
    def _prepare(self, sd_in):
        dev = self.device
        sd = {k: v.detach().to(dev, torch.float32) if v.is_floating_point() else v for k, v in sd_in.items()}
        cfg = self.cfg
        d = cfg.d_model
        L = lambda w, b=None: Lin(w, b, dev)  # noqa: E731
        bb = "model.backbone.conv_encoder.model."
        w, b = _fold_bn(sd, bb + "conv1", bb + "bn1")
        # stem weights for the NHWC4 gather: k = (ky*7 + kx)*4 + c, 49 taps x 4 -> 196, padded to K = 256
        w4 = torch.cat([w, w.new_zeros(w.shape[0], 1, 7, 7)], 1)
        self.stem = L(_conv_mat(w4, 256), b)
        # TMA-fed stem (stem.cu): weights laid out like the operand rows its tensor map delivers (include/egtr_b200.h)
        self.stem_mode = os.environ.get("EGTR_STEM", "tma")
        krow = int(call("egtr_stem_krow"))
        layout = int(call("egtr_stem_layout"))
        rows = stem_weight_rows(w, layout, krow)
        if layout == 2:
            self.stem_w_planes = rows.reshape(-1).contiguous()
        else:
            self.stem_w_planes = torch.empty(2 * rows.shape[0] * rows.shape[1], dtype=torch.bfloat16, device=dev)
            call("egtr_split_weight_bf16", _ptr(rows), rows.shape[0], rows.shape[1], rows.shape[0], _ptr(self.stem_w_planes), _stream())
        self.stem_bias = b.detach().to(device=dev, dtype=torch.float32).contiguous()
        self.blocks = []
        for li, nblk in enumerate(RESNET_BLOCKS, start=1):
            for bi in range(nblk):
                p = f"{bb}layer{li}.{bi}."
                blk = {"stride": 2 if (bi == 0 and li > 1) else 1, "last": bi == nblk - 1}
                for j in (1, 2, 3):
                    w, b = _fold_bn(sd, p + f"conv{j}", p + f"bn{j}")
                    blk[f"c{j}"] = L(_conv_mat(w), b)
                if (p + "downsample.0.weight") in sd:
                    w, b = _fold_bn(sd, p + "downsample.0", p + "downsample.1")
                    blk["ds"] = L(_conv_mat(w), b)
                self.blocks.append((li, blk))
        self.input_proj = []
        for l in range(cfg.num_feature_levels):
            p = f"model.input_proj.{l}."
            self.input_proj.append((L(_conv_mat(sd[p + "0.weight"]), sd[p + "0.bias"]), sd[p + "1.weight"].contiguous(), sd[p + "1.bias"].contiguous()))
        self.level_embed = sd["model.level_embed"].contiguous()
        i = torch.arange(d // 2, dtype=torch.float32)
        self.dim_t = (10000.0 ** (2 * torch.div(i, 2, rounding_mode="trunc") / (d // 2))).to(dev)  # deformable_detr.py:860-865

        def msda(p):
            return dict(
                offaw=L(torch.cat([sd[p + "sampling_offsets.weight"], sd[p + "attention_weights.weight"]], 0),
                        torch.cat([sd[p + "sampling_offsets.bias"], sd[p + "attention_weights.bias"]], 0)),
                out=L(sd[p + "output_proj.weight"], sd[p + "output_proj.bias"]),
            )

        def ln(p):
            return sd[p + ".weight"].contiguous(), sd[p + ".bias"].contiguous()

        self.enc = []
        for i in range(cfg.encoder_layers):
            p = f"model.encoder.layers.{i}."
            lay = msda(p + "self_attn.")
            lay["value"] = L(sd[p + "self_attn.value_proj.weight"], sd[p + "self_attn.value_proj.bias"])
            lay["ln1"], lay["ln2"] = ln(p + "self_attn_layer_norm"), ln(p + "final_layer_norm")
            lay["fc1"], lay["fc2"] = L(sd[p + "fc1.weight"], sd[p + "fc1.bias"]), L(sd[p + "fc2.weight"], sd[p + "fc2.bias"])
            self.enc.append(lay)

        heads = cfg.decoder_attention_heads
        scaling = (d // heads) ** -0.5
        self.dec = []
        vals_w, vals_b = [], []
        for i in range(cfg.decoder_layers):
            p = f"model.decoder.layers.{i}."
            lay = msda(p + "encoder_attn.")
            # q is scaled after the projection in the reference (deformable_detr.py:1166): fold it in.
            # q | k (on h + pos) and v (on h) as three groups of one launch
            lay["qkv"] = LinStack([sd[p + "self_attn.q_proj.weight"] * scaling, sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.v_proj.weight"]],
                                  [sd[p + "self_attn.q_proj.bias"] * scaling, sd[p + "self_attn.k_proj.bias"], sd[p + "self_attn.v_proj.bias"]], dev)
            lay["o"] = L(sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
            lay["ln1"], lay["ln2"], lay["ln3"] = ln(p + "self_attn_layer_norm"), ln(p + "encoder_attn_layer_norm"), ln(p + "final_layer_norm")
            lay["fc1"], lay["fc2"] = L(sd[p + "fc1.weight"], sd[p + "fc1.bias"]), L(sd[p + "fc2.weight"], sd[p + "fc2.bias"])
            vals_w.append(sd[p + "encoder_attn.value_proj.weight"])
            vals_b.append(sd[p + "encoder_attn.value_proj.bias"])
            self.dec.append(lay)
        # the six cross-attention value projections read the same encoder output: one [S,256]x[256,1536] GEMM
        self.dec_value = L(torch.cat(vals_w, 0), torch.cat(vals_b, 0))

        qpe = sd["model.query_position_embeddings.weight"]
        self.query_pos = qpe[:, :d].contiguous()
        self.query_tgt = qpe[:, d:].contiguous()
        self.refpt_w = sd["model.reference_points.weight"].contiguous()
        self.refpt_b = sd["model.reference_points.bias"].contiguous()

        self._prepare_fused_decoder(sd)

        if self.model_only:
            torch.cuda.current_stream().synchronize()
            return
        last = cfg.decoder_layers - 1
        self.cls = L(sd[f"class_embed.{last}.weight"], sd[f"class_embed.{last}.bias"])
        self.box0 = L(sd[f"bbox_embed.{last}.layers.0.weight"], sd[f"bbox_embed.{last}.layers.0.bias"])
        self.box1 = L(sd[f"bbox_embed.{last}.layers.1.weight"], sd[f"bbox_embed.{last}.layers.1.bias"])
        self.box2_w = sd[f"bbox_embed.{last}.layers.2.weight"].contiguous()
        self.box2_b = sd[f"bbox_embed.{last}.layers.2.bias"].contiguous()

        self._prepare_relation_head(sd)
        torch.cuda.current_stream().synchronize()

    def _prepare_fused_decoder(self, sd):
        """Weights of the one-kernel decoder stack (decoder.cu, `egtr_decoder_weights_t`): every projection stacked over the
        layers as bf16 hi/lo planes; q|k|v rows head-major (CTA r of the cluster owns head r); the `query_pos` halves of
        `(h + query_pos) . W` (deformable_detr.py:1404-1409, 1040) are weight-only terms, composed here in fp64."""
        cfg, dev = self.cfg, self.device
        nl, N, d = cfg.decoder_layers, cfg.num_queries, cfg.d_model
        self.dec_fused_ok = N <= 256 and cfg.decoder_ffn_dim == 1024 and d == 256
        if not self.dec_fused_ok:
            return
        mats = fused_decoder_matrices(sd, nl, self.query_pos, cfg.decoder_attention_heads)

        def planes(w):
            w = w.to(device=dev, dtype=torch.float32).contiguous()
            out = torch.empty(2 * w.shape[0] * w.shape[1], dtype=torch.bfloat16, device=dev)
            call("egtr_split_weight_bf16", _ptr(w), w.shape[0], w.shape[1], w.shape[0], _ptr(out), _stream())
            return out

        keep = {k: planes(mats[k]) for k in ("w_qkv", "w_o", "w_offaw", "w_out", "w_fc1", "w_fc2")}
        keep.update(vec=mats["vec"].to(dev), qkv_pos=mats["qkv_pos"].to(dev), off_pos=mats["off_pos"].to(dev),
                    tgt=self.query_tgt.to(dev, torch.float32).contiguous(), ref_points=torch.empty(N, 2, dtype=torch.float32, device=dev))
        call("egtr_small_linear_f32", _ptr(self.query_pos), 256, _ptr(self.refpt_w), _ptr(self.refpt_b), N, 256, 2, 1,
             None, 0, 0, _ptr(keep["ref_points"]), 2, _stream())
        self._dec_fused_tensors = keep  # the struct holds raw pointers
        wst = DecoderWeights()
        wst.layers, wst.n_queries = nl, N
        for k, t in keep.items():
            setattr(wst, k, _ptr(t))
        self.dec_fused_w = wst

    def _prepare_relation_head(self, sd):
        """Weights of the relation head (egtr.py:322-418, 507-516); also used on its own by `egtr_b200.relation_head.RelationHead`."""
        cfg, dev = self.cfg, self.device
        d, heads = cfg.d_model, cfg.decoder_attention_heads
        L = lambda w, b=None: Lin(w, b, dev)  # noqa: E731
        # relation head (egtr.py:322-418): un-scaling of the captured q folded into proj_q
        unscale = (d // heads) ** 0.5
        r1, c1, g = sd["rel_predictor.layers.0.weight"], sd["connectivity_layer.layers.0.weight"], sd["rel_predictor_gate.weight"]
        w1s = torch.cat([r1[:, :d], c1[:, :d], g[:, :d]], 0).double()  # [513,256] acts on the subject half
        w1o = torch.cat([r1[:, d:], c1[:, d:], g[:, d:]], 0).double()  # [513,256] acts on the object half
        gate_b = torch.cat([torch.zeros(2 * d, device=dev), sd["rel_predictor_gate.bias"]], 0).double()
        # Two stacked Linears with nothing in between (proj_q[l] then layer 1 of both MLPs and the gate,
        # egtr.py:339 -> 400-402/416) are composed once at load, in fp64: U_l = (W1s Wq_l) q_l + W1s b_l.
        ws, bs = [], []
        for l in range(cfg.decoder_layers + 1):
            wq = (sd[f"proj_q.{l}.weight"] * unscale, sd[f"proj_q.{l}.bias"]) if l < cfg.decoder_layers else (sd["final_sub_proj.weight"], sd["final_sub_proj.bias"])
            ws.append((w1s @ wq[0].double()).float())
            bs.append((w1s @ wq[1].double()).float())
        for l in range(cfg.decoder_layers + 1):
            wk = (sd[f"proj_k.{l}.weight"], sd[f"proj_k.{l}.bias"]) if l < cfg.decoder_layers else (sd["final_obj_proj.weight"], sd["final_obj_proj.bias"])
            ws.append((w1o @ wk[0].double()).float())
            bs.append((w1o @ wk[1].double() + gate_b).float())
        self.rel_uv = LinStack(ws, bs, dev)  # groups 0..6 -> U_l (subject side), 7..13 -> V_l (object side)
        self.rel_b1 = torch.cat([sd["rel_predictor.layers.0.bias"], sd["connectivity_layer.layers.0.bias"]], 0).contiguous()
        # layer 2 of both MLPs as one block-diagonal [512,256] weight: output columns 0..255 = relation MLP (reads
        # hidden channels 0..255), columns 256..511 = connectivity MLP (reads channels 256..511)
        self.rel_w2both = L(torch.cat([sd["rel_predictor.layers.1.weight"], sd["connectivity_layer.layers.1.weight"]], 0),
                            torch.cat([sd["rel_predictor.layers.1.bias"], sd["connectivity_layer.layers.1.bias"]], 0))
        self.rel_w2 = L(sd["rel_predictor.layers.1.weight"], sd["rel_predictor.layers.1.bias"])
        self.con_w2 = L(sd["connectivity_layer.layers.1.weight"], sd["connectivity_layer.layers.1.bias"])
        self.rel_w3 = L(sd["rel_predictor.layers.2.weight"], sd["rel_predictor.layers.2.bias"])
        # layer 3 on the TMA-fed kernel wants N % 32 == 0: predicate rows padded with zeros to a multiple of 64
        P_ = sd["rel_predictor.layers.2.weight"].shape[0]
        Pp = ((P_ + 63) // 64) * 64
        w3p = torch.zeros(Pp, 256, device=dev)
        w3p[:P_] = sd["rel_predictor.layers.2.weight"]
        b3p = torch.zeros(Pp, device=dev)
        b3p[:P_] = sd["rel_predictor.layers.2.bias"]
        self.rel_w3p = L(w3p, b3p)
        self.con_w3_w = sd["connectivity_layer.layers.2.weight"].contiguous()
        self.con_w3_b = sd["connectivity_layer.layers.2.bias"].contiguous()
        self.triplet = sd["triplet_dist"].contiguous()
        self.rel_dist = sd["rel_dist"].contiguous()
        self.rel_adj = (float(getattr(cfg, "logit_adj_tau", 0.3)) * self.rel_dist.log()).contiguous()  # egtr.py:509-512
        self.con_w3_b_host = float(sd["connectivity_layer.layers.2.bias"].item())
        # fused relation head (relhead.cu): layer-2 / layer-3 weights as bf16 "P32 group" rows for its TMA boxes
        Lr = cfg.decoder_layers + 1
        self.rel_fused = P_ <= 256 and Lr <= 7
        if self.rel_fused:
            bf16 = dict(dtype=torch.bfloat16, device=dev)
            self.rel_w2g = torch.empty(512, 512, **bf16)
            call("egtr_pack_weight_p32g", _ptr(self.rel_w2both.w), 512, 256, 512, None, _ptr(self.rel_w2g), _stream())
            if P_ <= 64:
                r = torch.arange(64, device=dev)
                # each CTA of a pair feeds 16 weight rows to each of the two N = 32 layer-3 MMAs (include/egtr_b200.h)
                self.rel_w3perm = (32 * ((r % 32) // 16) + 16 * (r // 32) + r % 16).to(torch.int32).contiguous()
                rows3 = 64
            else:  # one N = rows3 layer-3 MMA: predicates in their own order
                self.rel_w3perm, rows3 = None, 64 * ((P_ + 63) // 64)
            self.rel_w3g = torch.empty(rows3, 512, **bf16)
            call("egtr_pack_weight_p32g", _ptr(self.rel_w3.w), P_, 256, rows3, _ptr(self.rel_w3perm), _ptr(self.rel_w3g), _stream())
            hw = RelheadWeights()
            hw.layers, hw.uv_planes, hw.uv_bias, hw.uv_npad = Lr, _ptr(self.rel_uv.planes), _ptr(self.rel_uv.b), self.rel_uv.Npad
            hw.b1, hw.w2g, hw.b2 = _ptr(self.rel_b1), _ptr(self.rel_w2g), _ptr(self.rel_w2both.b)
            hw.w3g, hw.b3, hw.w3c, hw.b3c = _ptr(self.rel_w3g), _ptr(self.rel_w3.b), _ptr(self.con_w3_w), self.con_w3_b_host
            self.rel_head_w = hw

    # ------------------------------------------------------------------ CUDA-graph replay
    def graph_runner(self, B: int, H: int, W: int, slot: int = 0, throughput: bool = False) -> "GraphRunner":
        """The ~330 launches of one forward captured once per input shape and replayed as one graph:
        the host-side launch sequence disappears from the step time (batch 1 is launch-bound otherwise)."""
        key = ("graph", B, H, W, slot, throughput)
        if key not in self._ws:
            runner = GraphRunner(self, B, H, W, slot, throughput)
            self._ws[key] = runner
        self._ws.move_to_end(key)
        if (B, H, W, slot) in self._ws:
            self._ws.move_to_end((B, H, W, slot))
        return self._ws[key]

    def _evict(self, keep: tuple):
        """Drop least-recently-used workspaces (and the graph runners captured on them) until the cache fits its byte cap.
        Runners held elsewhere (serving) keep their own reference to the workspace they were captured on."""
        total = sum(self._ws_bytes.values())
        for key in [k for k in self._ws if k != keep and k[0] != "graph"]:
            if total <= self.ws_cap_bytes:
                break
            total -= self._ws_bytes.pop(key, 0)
            del self._ws[key]
            for gk in [g for g in self._ws if g[0] == "graph" and g[1:5] == key]:
                del self._ws[gk]

    def workspace_bytes(self) -> int:
        return sum(self._ws_bytes.values())

    # ------------------------------------------------------------------ launch helpers
    def gemm(self, lin: Lin, M: int, out: torch.Tensor, *, a: Optional[torch.Tensor] = None, lda: Optional[int] = None,
             a_col: int = 0, a2: Optional[torch.Tensor] = None, conv: Optional[dict] = None, relu: bool = False,
             res: Optional[torch.Tensor] = None, ldr: int = 0, ldo: Optional[int] = None, out_col: int = 0,
             remap: Optional[Tuple[int, int, int]] = None, row_keep: Optional[torch.Tensor] = None,
             a_fmt: int = 0, out_fmt: int = 0, res_fmt: int = 0, ln: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
             ln_out2: Optional[torch.Tensor] = None, ln_addend: Optional[torch.Tensor] = None):
        """`a_fmt`/`out_fmt`/`res_fmt` = 1: the tensor is stored as P32 rows (include/egtr_b200.h); a P32 operand takes the
        TMA-fed tcgen05 kernel (gemm_p32.cu)."""
        src = ASrc()
        if conv is None:
            src.a, src.a2, src.mode = _ptr(a, a_col), _ptr(a2, a_col), 0
            src.lda = lda if lda is not None else lin.K
        else:
            src.a, src.a2, src.mode, src.lda = _ptr(conv["x"]), None, conv.get("mode", 1), 0
            for k in ("H", "W", "C", "OH", "OW", "KH", "KW", "stride", "pad"):
                setattr(src, k, conv[k])
        ep = Epilogue()
        ep.bias, ep.res, ep.out = _ptr(lin.b), _ptr(res), _ptr(out, out_col)
        ep.ldo = ldo if ldo is not None else lin.N
        ep.ldr = ldr if ldr else ep.ldo
        ep.relu = int(relu)
        ep.rows_per_b, ep.bstride, ep.off = remap if remap else (0, 0, 0)
        ep.row_keep = _ptr(row_keep)
        src.fmt, ep.out_fmt, ep.res_fmt = a_fmt, out_fmt, res_fmt
        if ln is not None:  # LayerNorm(acc + bias + res) in the epilogue (P32 operand kernel, N == 256)
            assert a_fmt == 1 and out_fmt == 1
            ep.ln_gamma, ep.ln_beta = _ptr(ln[0]), _ptr(ln[1])
            ep.ln_out2, ep.ln_addend = _ptr(ln_out2), _ptr(ln_addend)
        if a_fmt == 1:
            assert a2 is None
            with self.span("gemm_p32"), self.span(f"gemm_p32:{M}x{lin.N}x{lin.K}" + (":conv" if conv is not None else "")):
                call("egtr_gemm_sbf16", C.byref(src), _ptr(lin.planes), M, lin.N, lin.Npad, lin.K, C.byref(ep), _stream())
            if self.probe is not None:
                self.probe_flops["gemm_p32"] = self.probe_flops.get("gemm_p32", 0) + 2 * M * lin.N * lin.K
            return
        if gemm_backend() == "simt":
            call("egtr_gemm_f32", C.byref(src), _ptr(lin.w), M, lin.N, lin.K, C.byref(ep), _stream())
        elif conv is None and _skinny(M, lin.N, 1):
            # a few hundred rows: latency-bound -> many small fp32 CTAs beat 128-row tensor-core tiles
            one = (C.c_void_p * 1)
            call("egtr_gemm_f32_grouped", one(src.a), one(src.a2), one(ep.out), (C.c_int * 1)(0), 1, (C.c_int * 1)(src.lda),
                 _ptr(lin.w), M, lin.N, lin.K, C.byref(ep), _stream())
        else:
            call("egtr_gemm_sbf16", C.byref(src), _ptr(lin.planes), M, lin.N, lin.Npad, lin.K, C.byref(ep), _stream())

    def gemm_grouped(self, st: "LinStack", M: int, *, a, lda, out, ldo, a2=None, a_col=None, relu: bool = False):
        G = st.G
        a_col = a_col or [0] * G
        a2 = a2 or [None] * G
        ap = (C.c_void_p * G)(*[_ptr(t, c) for t, c in zip(a, a_col)])
        a2p = (C.c_void_p * G)(*[_ptr(t, c) for t, c in zip(a2, a_col)])
        op = (C.c_void_p * G)(*[_ptr(t, c) for t, c in out])
        ldap = (C.c_int * G)(*lda)
        ep = Epilogue()
        ep.bias, ep.res, ep.out, ep.ldo, ep.ldr, ep.relu = _ptr(st.b), None, _ptr(out[0][0], out[0][1]), ldo, ldo, int(relu)
        if gemm_backend() == "simt":  # debug cross-check path: one CUDA-core launch per group
            for g in range(G):
                src, e2 = ASrc(), Epilogue()
                src.a, src.a2, src.mode, src.lda = ap[g], a2p[g], 0, lda[g]
                e2.bias, e2.out, e2.ldo, e2.ldr, e2.relu = _ptr(st.b, g * st.Npad), op[g], ldo, ldo, int(relu)
                call("egtr_gemm_f32", C.byref(src), _ptr(st.w, g * st.Npad * st.K), M, st.N, st.K, C.byref(e2), _stream())
            return
        if _skinny(M, st.N, G):
            call("egtr_gemm_f32_grouped", ap, a2p, op, st.n_base, G, ldap, _ptr(st.w), M, st.N, st.K, C.byref(ep), _stream())
            return
        call("egtr_gemm_sbf16_grouped", ap, a2p, op, st.n_base, G, ldap, _ptr(st.planes), st.plane_rows, M, st.N, st.Npad, st.K,
             C.byref(ep), _stream())

    def layernorm(self, x, res, ln, rows, out):
        call("egtr_add_layernorm_f32", _ptr(x), _ptr(res), _ptr(ln[0]), _ptr(ln[1]), rows, 256, _ptr(out), _stream())

    def _dec_self_attn(self, lay, ws, Md, B, N, h, qkv, t1, offaw):
        """Self-attention half of a decoder layer (deformable_detr.py:1404-1417) + the cross-attention's offsets/weights
        projection: h -> qkv (captured), t1 = LN1(h + out_proj(attn)), offaw = Linear(t1 + query_pos)."""
        st = _stream()
        qpos = ws["qpos"]
        self.gemm_grouped(lay["qkv"], Md, a=[h, h, h], a2=[qpos, qpos, None], lda=[256, 256, 256],
                          out=[(qkv, 0), (qkv, 256), (qkv, 512)], ldo=768)
        call("egtr_mha_core_f32", _ptr(qkv), 768, B, N, 8, 32, _ptr(ws["dattn"]), st)
        call("egtr_gemm_f32_splitk", _ptr(ws["dattn"]), None, 256, _ptr(lay["o"].w), Md, 256, 256, 2, _ptr(ws["dpart"]), st)
        call("egtr_sum_layernorm_f32", _ptr(ws["dpart"]), 2, Md * 256, _ptr(lay["o"].b), _ptr(h), _ptr(lay["ln1"][0]), _ptr(lay["ln1"][1]),
             Md, 256, _ptr(t1), None, 0, 0, st)
        self.gemm(lay["offaw"], Md, offaw, a=t1, a2=qpos, lda=256)

    # ------------------------------------------------------------------ workspace
    def _workspace(self, B: int, H: int, W: int, slot: int = 0) -> dict:
        """Activation buffers of one forward; `slot` > 0 gives an independent set for a forward that may run concurrently
        with slot 0's (a second CUDA graph on another stream)."""
        key = (B, H, W, slot)
        ws = self._ws.get(key)
        if ws is not None:
            self._ws.move_to_end(key)
            return ws
        dev, cfg = self.device, self.cfg
        f32 = dict(dtype=torch.float32, device=dev)
        shapes = level_shapes(H, W, cfg.num_feature_levels)
        S = sum(h * w for h, w in shapes)
        N = cfg.num_queries
        h1, w1 = conv_out(H, 7, 2, 3), conv_out(W, 7, 2, 3)
        h2, w2 = conv_out(h1, 3, 2, 1), conv_out(w1, 3, 2, 1)
        ws = dict(shapes=shapes, S=S, stem_hw=(h1, w1), c2_hw=(h2, w2))
        ws["shapes_c"] = (C.c_int * (2 * len(shapes)))(*[v for hw in shapes for v in hw])
        ws["starts"] = [sum(h * w for h, w in shapes[:l]) for l in range(len(shapes))]
        ws["px4"] = torch.empty(B * (H + 6) * (W + 6) * 4, **f32) if self.stem_mode != "tma" else None
        ws["px_planes"] = torch.zeros(int(call("egtr_stem_planes_bytes", B, H, W)), dtype=torch.uint8, device=dev) if self.stem_mode == "tma" else None
        ws["stem"] = torch.empty(B * h1 * w1, 64, **f32)
        # backbone ping-pong buffers sized for the largest stage output (layer1: h2*w2 x 256)
        big = B * h2 * w2 * 256
        ws["bb"] = [torch.empty(big, **f32) for _ in range(4)]
        ws["c5"] = None
        ws["mask_flat"] = torch.empty(B, S, dtype=torch.uint8, device=dev)
        ws["pos"] = torch.empty(B * S, 256, **f32)
        ws["valid_ratios"] = torch.empty(B, len(shapes), 2, **f32)
        ws["geo_scratch"] = torch.empty(2 * B * S, **f32)
        gn = max(call("egtr_groupnorm_scratch_doubles", B, h * w) for h, w in shapes)
        ws["gn_scratch"] = torch.empty(int(gn), dtype=torch.float64, device=dev)
        ws["x"] = [torch.empty(B * S, 256, **f32) for _ in range(5)]
        ws["offaw"] = torch.empty(B * S, 384, **f32)
        f32_value = self.msda_value != "h16" or gemm_backend() == "simt"  # fp32 value rows only where a path reads them
        ws["value"] = torch.empty(B * S, 256, **f32) if f32_value else None
        ws["attn"] = torch.empty(B * S, 256, **f32)
        ws["ffn"] = torch.empty(B * S, 1024, **f32)
        ws["dec_value"] = torch.empty(B * S, 256 * cfg.decoder_layers, **f32) if f32_value else None
        # H16 pair records [heads][B*S + 1][2][32] fp16: the same bytes per token-head as fp32 rows, plus one record per head;
        # zero-initialised: the two padding slots per head are never written and must stay finite
        if not f32_value:
            ws["value_h16"] = torch.zeros(8 * (B * S + 1) * 64, dtype=torch.float16, device=dev)
            ws["dec_value_h16"] = torch.zeros(8 * cfg.decoder_layers * (B * S + 1) * 64, dtype=torch.float16, device=dev)
        ws["qpos"] = self.query_pos.unsqueeze(0).expand(B, -1, -1).reshape(B * N, 256).contiguous()
        ws["tgt"] = self.query_tgt.unsqueeze(0).expand(B, -1, -1).reshape(B * N, 256).contiguous()
        ws["ref"] = torch.empty(N, 2, **f32)
        ws["dh"] = [torch.empty(B * N, 256, **f32) for _ in range(4)]
        ws["dattn"] = torch.empty(B * N, 256, **f32)
        ws["doffaw"] = torch.empty(B * N, 384, **f32)
        ws["dffn"] = torch.empty(B * N, 1024, **f32)
        ws["dpart"] = torch.empty(8, B * N, 256, **f32)  # split-K partial sums of the decoder's out_proj / fc2
        if self.dec_fused_ok and B * N <= SKINNY_M:
            # fused decoder stack: one zero-initialised, 1 KB aligned scratch block (its first B*N KB are the final hidden state)
            nbytes = int(call("egtr_decoder_scratch_bytes", B, N))
            raw = torch.zeros(nbytes + 1024, dtype=torch.uint8, device=dev)
            off = (-raw.data_ptr()) % 1024
            ws["dec_scratch"] = raw[off: off + nbytes]
            ws["dec_h_last"] = ws["dec_scratch"][: B * N * 1024].view(torch.float32).view(B * N, 256)
        ws["box_h"] = [torch.empty(B * N, 256, **f32) for _ in range(2)]
        Lr = cfg.decoder_layers + 1
        ws["U"] = torch.empty(B * N * Lr, 516, **f32)
        ws["V"] = torch.empty(B * N * Lr, 516, **f32)
        # pair-sized scratch (H1 / H2r / H2c / rel_logits / con_logits) exists only on the paths that materialise it: `_pair_buf`
        ws["_key"] = key
        ws["cls_idx"] = torch.empty(B * N, dtype=torch.int32, device=dev)
        self._ws[key] = ws
        tensors = [t for v in ws.values() for t in (v if isinstance(v, (list, tuple)) else [v]) if isinstance(t, torch.Tensor)]
        self._ws_bytes[key] = sum(t.numel() * t.element_size() for t in tensors)
        self._evict(keep=key)
        return ws

    def _pair_buf(self, ws: dict, name: str, rows: int, cols: int) -> torch.Tensor:
        """Pair-dimension scratch of the unfused relation paths (simt cross-check; P > 64), allocated on first use."""
        if name not in ws:
            ws[name] = torch.empty(rows, cols, dtype=torch.float32, device=self.device)
            if ws["_key"] in self._ws_bytes:
                self._ws_bytes[ws["_key"]] += rows * cols * 4
        return ws[name]

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor, pixel_mask: Optional[torch.Tensor] = None, taps: Optional[dict] = None,
                slot: int = 0, throughput: bool = False) -> Dict[str, torch.Tensor]:
        """`throughput`: this forward is one of several in flight (serving): few-tile GEMMs then skip split-K — other images
        fill the SMs and the partial-sum round trip is pure cost.  A lone forward (the default) uses split-K for latency."""
        cfg, dev = self.cfg, self.device
        if pixel_values.device != dev:
            raise _lib.EgtrError(f"pixel_values on {pixel_values.device}, model on {dev}")
        with torch.cuda.device(dev):
            return self._forward(pixel_values, pixel_mask, taps, slot, throughput)

    def _forward(self, pixel_values, pixel_mask, taps, slot: int = 0, throughput: bool = False, static_out: bool = False):
        """`static_out`: outputs that live in the workspace (encoder output, reference points) are handed out as views instead of
        copies — for CUDA-graph replay, whose outputs are static buffers anyway (GraphRunner)."""
        cfg, dev = self.cfg, self.device
        call("egtr_set_scratch_slot", slot)
        call("egtr_set_splitk_max", self.throughput_splitk if throughput else 64)
        call("egtr_set_grid_div", (self.throughput_grid_div or (4 if pixel_values.shape[0] == 1 else 2)) if throughput else 1)
        st = _stream()
        px = pixel_values.to(torch.float32).contiguous()
        B, Cin, H, W = px.shape
        if Cin != 3:
            raise ValueError(f"pixel_values must have 3 channels, got {Cin}")
        if pixel_mask is None:
            pixel_mask = torch.ones(B, H, W, dtype=torch.long, device=dev)
        pm = pixel_mask.to(torch.long).contiguous()
        ws = self._workspace(B, H, W, slot)
        shapes, S, starts = ws["shapes"], ws["S"], ws["starts"]
        N, d, Lv = cfg.num_queries, cfg.d_model, len(shapes)
        f32 = dict(dtype=torch.float32, device=dev)

        # ---- geometry: masks, position embeddings, valid ratios (deformable_detr.py:783-785, 850-876, 2064-2073)
        # A lone forward runs it on a side stream next to the backbone (it depends on the mask only; its scan kernel is a 30 us
        # chain on four CTAs) and joins before the first GroupNorm, which adds the position embeddings; forwards in flight have
        # other images to fill the GPU with.
        geo_side = None
        if not throughput and os.environ.get("EGTR_GEOMETRY_STREAM", "1") == "1":
            geo_side = ws.get("geo_stream")
            if geo_side is None:
                geo_side = ws["geo_stream"] = torch.cuda.Stream(device=dev)
            geo_side.wait_stream(torch.cuda.current_stream())
        call("egtr_levels_geometry_f32", _ptr(pm), B, H, W, ws["shapes_c"], Lv, _ptr(self.level_embed), _ptr(self.dim_t), d,
             _ptr(ws["mask_flat"]), _ptr(ws["pos"]), _ptr(ws["valid_ratios"]), _ptr(ws["geo_scratch"]),
             geo_side.cuda_stream if geo_side is not None else st)

        # ---- backbone (deformable_detr.py:778): stem 7x7/2 as a gather-GEMM over the NCHW image, max-pool, bottlenecks
        h1, w1 = ws["stem_hw"]
        _sp_bb = self.span("stage_backbone"); _sp_bb.__enter__()
        if self.stem_mode == "tma" and gemm_backend() != "simt":
            # two bf16 planes of the zero-bordered NHWC4 image; a filter row of 128 output pixels' windows is one TMA box (stem.cu)
            call("egtr_stem_pad_split_bf16", _ptr(px), B, H, W, _ptr(ws["px_planes"]), st)
            call("egtr_stem_conv7x7s2_bf16x3", _ptr(ws["px_planes"]), B, H, W, _ptr(self.stem_w_planes), _ptr(self.stem_bias), _ptr(ws["stem"]), st)
        else:
            if ws["px4"] is None:
                ws["px4"] = torch.empty(B * (H + 6) * (W + 6) * 4, **f32)
            call("egtr_pad_nchw3_to_nhwc4_f32", _ptr(px), B, H, W, 3, _ptr(ws["px4"]), st)
            self.gemm(self.stem, B * h1 * w1, ws["stem"], relu=True,
                      conv=dict(x=ws["px4"], mode=3, H=H + 6, W=W + 6, C=4, OH=h1, OW=w1, KH=7, KW=7, stride=2, pad=0))
        h, w = ws["c2_hw"]
        bufs = ws["bb"]
        x = bufs[0]
        # P32 from here on (product path): every conv reads its operand through TMA; 3x3 / strided convs as patch tiles
        p32 = 0 if gemm_backend() == "simt" else 1
        call("egtr_maxpool3x3s2_nhwc_ex", _ptr(ws["stem"]), B, h1, w1, 64, _ptr(x), p32, st)
        fm = dict(a_fmt=p32, out_fmt=p32)
        cur, cin = 0, 64
        for li, blk in self.blocks:
            s = blk["stride"]
            oh, ow = (conv_out(h, 3, 2, 1), conv_out(w, 3, 2, 1)) if s == 2 else (h, w)
            planes = blk["c1"].N
            free = [i for i in range(4) if i != cur]
            y1, y2, idt = bufs[free[0]], bufs[free[1]], bufs[free[2]]
            self.gemm(blk["c1"], B * h * w, y1, a=x, lda=cin, relu=True, **fm)
            self.gemm(blk["c2"], B * oh * ow, y2, relu=True,
                      conv=dict(x=y1, H=h, W=w, C=planes, OH=oh, OW=ow, KH=3, KW=3, stride=s, pad=1), **fm)
            if "ds" in blk:
                if s == 1 and p32:
                    self.gemm(blk["ds"], B * oh * ow, idt, a=x, lda=cin, **fm)
                else:
                    self.gemm(blk["ds"], B * oh * ow, idt,
                              conv=dict(x=x, H=h, W=w, C=cin, OH=oh, OW=ow, KH=1, KW=1, stride=s, pad=0), **fm)
                res = idt
            else:
                res = x
            # conv3 + bn3 + identity + ReLU in one epilogue; output reuses y1's buffer
            self.gemm(blk["c3"], B * oh * ow, y1, a=y2, lda=planes, relu=True, res=res, ldr=planes * 4, res_fmt=p32, **fm)
            x, cur, cin, h, w = y1, free[0], planes * 4, oh, ow
            if blk.get("last") and li >= 2:
                # C3/C4/C5 feed input_proj straight away: 1x1 conv + GroupNorm written into the level's
                # slice of source_flatten [B,S,256] (deformable_detr.py:2221-2241, 2259-2266)
                lvl = li - 2
                if geo_side is not None:
                    torch.cuda.current_stream().wait_stream(geo_side)
                    geo_side = None
                lin, gw, gb = self.input_proj[lvl]
                hw = h * w
                self.gemm(lin, B * hw, ws["x"][0], a=x, lda=cin, ldo=256, remap=(hw, S, starts[lvl]), a_fmt=p32)
                call("egtr_groupnorm_ex", _ptr(ws["x"][0]), B, hw, S, starts[lvl], 256, 32, _ptr(gw), _ptr(gb), _ptr(ws["gn_scratch"]),
                     _ptr(ws["x"][1]) if p32 else None, _ptr(ws["pos"]), _ptr(ws["x"][2]) if p32 else None, st)
                if li == 4:  # extra level: 3x3/2 conv on C5
                    if Lv > 4:
                        raise _lib.EgtrError("more than 4 feature levels are not built")
                    lin2, gw2, gb2 = self.input_proj[3]
                    oh2, ow2 = shapes[3]
                    self.gemm(lin2, B * oh2 * ow2, ws["x"][0], ldo=256, remap=(oh2 * ow2, S, starts[3]),
                              conv=dict(x=x, H=h, W=w, C=cin, OH=oh2, OW=ow2, KH=3, KW=3, stride=2, pad=1), a_fmt=p32)
                    call("egtr_groupnorm_ex", _ptr(ws["x"][0]), B, oh2 * ow2, S, starts[3], 256, 32, _ptr(gw2), _ptr(gb2), _ptr(ws["gn_scratch"]),
                         _ptr(ws["x"][1]) if p32 else None, _ptr(ws["pos"]), _ptr(ws["x"][2]) if p32 else None, st)
                if taps is not None:
                    xf = x[: B * hw * cin]
                    if p32:
                        xf = torch.empty(B * hw, cin, **f32)
                        call("egtr_p32_to_rows", _ptr(x), B * hw, cin, _ptr(xf), cin, st)
                    taps[f"c{li + 1}"] = xf.view(B, h, w, cin).permute(0, 3, 1, 2).clone()
        if taps is not None:
            taps["source_flatten"] = ws["x"][0].view(B, S, 256).clone()
            taps["lvl_pos_embed_flatten"] = ws["pos"].view(B, S, 256).clone()
            taps["mask_flatten"] = ws["mask_flat"].bool().clone()
            taps["valid_ratios"] = ws["valid_ratios"].clone()

        _sp_bb.__exit__()
        # ---- encoder (deformable_detr.py:1283-1358)
        _sp_enc = self.span("stage_encoder"); _sp_enc.__enter__()
        M = B * S
        pos, offaw, value, attn, ffn = ws["pos"], ws["offaw"], ws["value"], ws["attn"], ws["ffn"]
        vr = ws["valid_ratios"]
        if gemm_backend() == "simt":
            xa, xb, xc = ws["x"][:3]
            for i, lay in enumerate(self.enc):
                self.gemm(lay["offaw"], M, offaw, a=xa, a2=pos, lda=256)
                self.gemm(lay["value"], M, value, a=xa, lda=256, row_keep=ws["mask_flat"])
                with self.span("msda_enc"):
                    call("egtr_msda_fused_fwd_f32", _ptr(value), 256, ws["shapes_c"], _ptr(offaw), 384, None, _ptr(vr), 1,
                         B, S, 8, 32, Lv, S, 4, _ptr(attn), st)
                self.gemm(lay["out"], M, xb, a=attn, lda=256, res=xa, ldr=256)
                self.layernorm(xb, None, lay["ln1"], M, xc)
                self.gemm(lay["fc1"], M, ffn, a=xc, lda=256, relu=True)
                self.gemm(lay["fc2"], M, xb, a=ffn, lda=1024, res=xc, ldr=256)
                self.layernorm(xb, None, lay["ln2"], M, xa)
                if taps is not None and i == 0:
                    taps["enc0_out"] = xa.view(B, S, 256).clone()
            enc_f32, enc_p32 = xa, None
        else:
            # Product path: every tensor that feeds a GEMM lives in HBM as P32 rows (split-bf16 at fp32 pitch), written by
            # the kernel that produces it; the GEMMs stream both operands with TMA.  x: layer input, xp: x + pos (operand of
            # the sampling_offsets / attention_weights projections, deformable_detr.py:1040), xc: post-attention LayerNorm.
            x0, x, xp, xc, _ = ws["x"]  # x / xp were written as P32 by the GroupNorm of each level
            h16 = self.msda_value == "h16"
            enc_f32 = x0
            nl_enc = len(self.enc)
            for i, lay in enumerate(self.enc):
                self.gemm(lay["offaw"], M, offaw, a=xp, lda=256, a_fmt=1)
                if h16:
                    self.gemm(lay["value"], M, ws["value_h16"], a=x, lda=256, a_fmt=1, row_keep=ws["mask_flat"], out_fmt=2)
                    with self.span("msda_enc"):
                        call("egtr_msda_fused_fwd_h16", _ptr(ws["value_h16"]), M + 1, 0, 8, ws["shapes_c"], _ptr(offaw), 384, None, _ptr(vr), 1,
                             B, S, 8, 32, Lv, S, 4, _ptr(attn), 1, st)
                else:
                    self.gemm(lay["value"], M, value, a=x, lda=256, a_fmt=1, row_keep=ws["mask_flat"])
                    with self.span("msda_enc"):
                        call("egtr_msda_fused_fwd_ex", _ptr(value), 256, ws["shapes_c"], _ptr(offaw), 384, None, _ptr(vr), 1,
                             B, S, 8, 32, Lv, S, 4, _ptr(attn), 1, st)
                # output_proj + residual + LayerNorm and fc2 + residual + LayerNorm are ONE kernel each: the GEMM's epilogue
                # normalises its 256-wide rows in TMEM and writes P32 rows (and, for the layer output, x + pos as well)
                self.gemm(lay["out"], M, xc, a=attn, lda=256, a_fmt=1, res=x, res_fmt=1, ldr=256, out_fmt=1, ln=lay["ln1"])
                self.gemm(lay["fc1"], M, ffn, a=xc, lda=256, a_fmt=1, relu=True, out_fmt=1)
                last = i == nl_enc - 1
                self.gemm(lay["fc2"], M, x, a=ffn, lda=1024, a_fmt=1, res=xc, res_fmt=1, ldr=256, out_fmt=1, ln=lay["ln2"],
                          ln_out2=None if last else xp, ln_addend=None if last else pos)
                if last or (taps is not None and i == 0):
                    call("egtr_p32_to_rows", _ptr(x), M, 256, _ptr(enc_f32), 256, st)
                if taps is not None and i == 0:
                    taps["enc0_out"] = enc_f32.view(B, S, 256).clone()
            enc_p32 = x
        enc = enc_f32
        # graph replay: a view of the workspace (no 22.8 MB copy per image); eager calls return a fresh tensor like the reference
        enc_out = enc.view(B, S, 256) if static_out else enc.view(B, S, 256).clone()
        _sp_enc.__exit__()
        _sp_dec = self.span("stage_decoder"); _sp_dec.__enter__()

        # ---- decoder (deformable_detr.py:1390-1489, 1774-1968)
        nl = cfg.decoder_layers
        dv = ws["dec_value"]
        dec_h16 = enc_p32 is not None and self.msda_value == "h16"
        if dec_h16:
            self.gemm(self.dec_value, M, ws["dec_value_h16"], a=enc_p32, lda=256, a_fmt=1, row_keep=ws["mask_flat"], out_fmt=2)
        elif enc_p32 is not None:
            self.gemm(self.dec_value, M, dv, a=enc_p32, lda=256, a_fmt=1, row_keep=ws["mask_flat"])
        else:
            self.gemm(self.dec_value, M, dv, a=enc, lda=256, row_keep=ws["mask_flat"])
        Md = B * N
        qpos = ws["qpos"]
        hbuf = ws["dh"]
        dpart = ws["dpart"]
        dec_fused = dec_h16 and "dec_scratch" in ws and Lv == 4 and self.decoder_mode == "fused"
        if dec_fused:
            # ONE launch for the whole stack (decoder.cu): a cluster of eight CTAs per image, every GEMM on tcgen05
            if "ref_done" not in ws:
                call("egtr_small_linear_f32", _ptr(self.query_pos), 256, _ptr(self.refpt_w), _ptr(self.refpt_b), N, 256, 2, 1,
                     None, 0, 0, _ptr(ws["ref"]), 2, st)
                ws["ref_done"] = True
            qkv_all = torch.empty(nl, Md, 768, **f32)
            inter = torch.empty(B, nl, N, 256, **f32)
            with self.span("decoder_fused"):
                call("egtr_decoder_fused_f32", C.byref(self.dec_fused_w), _ptr(ws["dec_scratch"]), _ptr(ws["dec_value_h16"]), M + 1,
                     ws["shapes_c"], Lv, _ptr(vr), B, S, _ptr(qkv_all), _ptr(inter), 0, nl, 0, 12, 1, st)
            qkvs = [qkv_all[l] for l in range(nl)]
            hcur = ws["dec_h_last"]
        if not dec_fused and "dec0" not in ws:
            # Input-independent prefix (weights only): the reference points and the whole self-attention half of decoder
            # layer 0 — q|k|v of the learned queries, attention, out_proj + LayerNorm, and the sampling_offsets /
            # attention_weights projection of its cross-attention — are computed once per workspace, not per image.
            call("egtr_small_linear_f32", _ptr(self.query_pos), 256, _ptr(self.refpt_w), _ptr(self.refpt_b), N, 256, 2, 1,
                 None, 0, 0, _ptr(ws["ref"]), 2, st)
            lay = self.dec[0]
            qkv0 = torch.empty(Md, 768, **f32)
            t1_0 = torch.empty(Md, 256, **f32)
            offaw0 = torch.empty(Md, 384, **f32)
            self._dec_self_attn(lay, ws, Md, B, N, ws["tgt"], qkv0, t1_0, offaw0)
            ws["dec0"] = (qkv0, t1_0, offaw0)
        if not dec_fused:
            hcur = ws["tgt"]
            qkvs = []
            inter = torch.empty(B, nl, N, 256, **f32)
        for i, lay in enumerate(() if dec_fused else self.dec):
            t1, t2, t3 = [b for b in hbuf if b is not hcur][:3]
            if i == 0:
                qkv, t1, offaw_i = ws["dec0"]
            else:
                qkv = torch.empty(Md, 768, **f32)  # captured per layer: q (scaled) | k | v
                offaw_i = ws["doffaw"]
                self._dec_self_attn(lay, ws, Md, B, N, hcur, qkv, t1, offaw_i)
            qkvs.append(qkv)
            with self.span("msda_dec"):
                if dec_h16:
                    call("egtr_msda_fused_fwd_h16", _ptr(ws["dec_value_h16"]), M + 1, i * 8, 8 * nl, ws["shapes_c"], _ptr(offaw_i), 384,
                         _ptr(ws["ref"]), _ptr(vr), 0, B, S, 8, 32, Lv, N, 4, _ptr(ws["dattn"]), 0, st)
                else:
                    call("egtr_msda_fused_fwd_f32", _ptr(dv, i * 256), 256 * nl, ws["shapes_c"], _ptr(offaw_i), 384,
                         _ptr(ws["ref"]), _ptr(vr), 0, B, S, 8, 32, Lv, N, 4, _ptr(ws["dattn"]), st)
            # output_proj / fc2 as split-K sums; bias + residual + LayerNorm consume them (deformable_detr.py:1441-1477)
            call("egtr_gemm_f32_splitk", _ptr(ws["dattn"]), None, 256, _ptr(lay["out"].w), Md, 256, 256, 2, _ptr(dpart), st)
            call("egtr_sum_layernorm_f32", _ptr(dpart), 2, Md * 256, _ptr(lay["out"].b), _ptr(t1), _ptr(lay["ln2"][0]), _ptr(lay["ln2"][1]),
                 Md, 256, _ptr(t2), None, 0, 0, st)
            self.gemm(lay["fc1"], Md, ws["dffn"], a=t2, lda=256, relu=True)
            call("egtr_gemm_f32_splitk", _ptr(ws["dffn"]), None, 1024, _ptr(lay["fc2"].w), Md, 256, 1024, 4, _ptr(dpart), st)
            # the layer output also lands in its slot of the stacked intermediate states [B, layers, N, 256]
            call("egtr_sum_layernorm_f32", _ptr(dpart), 4, Md * 256, _ptr(lay["fc2"].b), _ptr(t2), _ptr(lay["ln3"][0]), _ptr(lay["ln3"][1]),
                 Md, 256, _ptr(t3), _ptr(inter, i * N * 256), N, nl * N * 256, st)
            hcur = t3
        h_last = hcur
        # captured decoder self-attention states as [B, heads, N, 32] views (deformable_detr.py:1179-1185)
        qs = tuple(q.view(B, N, 3, 8, 32)[:, :, 0].permute(0, 2, 1, 3) for q in qkvs)
        ks = tuple(q.view(B, N, 3, 8, 32)[:, :, 1].permute(0, 2, 1, 3) for q in qkvs)
        ref_b = (ws["ref"] if static_out else ws["ref"].clone()).unsqueeze(0).expand(B, -1, -1)  # [B,N,2]: one set for the batch and for every layer (no box refinement)
        model_out = dict(last_hidden_state=inter[:, nl - 1], intermediate_hidden_states=inter, encoder_last_hidden_state=enc_out,
                         init_reference_points=ref_b, intermediate_reference_points=ref_b.unsqueeze(1).expand(B, nl, N, 2),
                         decoder_attention_queries=qs, decoder_attention_keys=ks)
        if self.model_only:
            _sp_dec.__exit__()
            return model_out

        # ---- detection heads (egtr.py:283-314; only the last level is returned at inference)
        K = cfg.num_labels
        logits = torch.empty(B, N, K, **f32)
        boxes = torch.empty(B, N, 4, **f32)
        self.gemm(self.cls, Md, logits, a=h_last, lda=256)
        self.gemm(self.box0, Md, ws["box_h"][0], a=h_last, lda=256, relu=True)
        self.gemm(self.box1, Md, ws["box_h"][1], a=ws["box_h"][0], lda=256, relu=True)
        call("egtr_small_linear_f32", _ptr(ws["box_h"][1]), 256, _ptr(self.box2_w), _ptr(self.box2_b), Md, 256, 4, 2,
             _ptr(ws["ref"]), 2, N, _ptr(boxes), 4, st)

        _sp_dec.__exit__()
        # ---- relation head (egtr.py:322-418, 507-516)
        _sp_rel = self.span("stage_relation"); _sp_rel.__enter__()
        Lr = nl + 1
        P = cfg.num_rel_labels
        # per-query layer-1 partials U_l(i), V_l(j) (+ gate logits in column 512): one grouped launch of 14 GEMMs
        a_list = [qkvs[l] for l in range(nl)] + [h_last] + [qkvs[l] for l in range(nl)] + [h_last]
        a_cols = [0] * nl + [0] + [256] * nl + [0]
        ldas = [768] * nl + [256] + [768] * nl + [256]
        outs = [(ws["U"], l * 516) for l in range(Lr)] + [(ws["V"], l * 516) for l in range(Lr)]
        fused = self.rel_fused and gemm_backend() != "simt" and os.environ.get("EGTR_RELHEAD", "fused") == "fused"
        if not fused:
            self.gemm_grouped(self.rel_uv, Md, a=a_list, a_col=a_cols, lda=ldas, out=outs, ldo=Lr * 516)
        pairs = B * N * N
        pred_rel = torch.empty(B, N, N, P, **f32)
        pred_con = torch.empty(B, N, N, 1, **f32)
        if gemm_backend() == "simt":
            # unfused cross-check path: pair kernel -> H1 -> two layer-2 GEMMs -> layer 3 -> finish kernel
            H1, H2r, H2c = self._pair_buf(ws, "H1", pairs, 512), self._pair_buf(ws, "H2r", pairs, 256), self._pair_buf(ws, "H2c", pairs, 256)
            rel_logits, con_logits = self._pair_buf(ws, "rel_logits", pairs, self.rel_w3p.N), self._pair_buf(ws, "con_logits", pairs, 1)
            call("egtr_relation_pair_hidden_f32", _ptr(ws["U"]), _ptr(ws["V"]), 516, _ptr(self.rel_b1), B, N, Lr, _ptr(H1), st)
            self.gemm(self.rel_w2, pairs, H2r, a=H1, lda=512, a_col=0, relu=True)
            self.gemm(self.con_w2, pairs, H2c, a=H1, lda=512, a_col=256, relu=True)
            self.gemm(self.rel_w3, pairs, rel_logits, a=H2r, lda=256, ldo=P)
            call("egtr_small_linear_f32", _ptr(H2c), 256, _ptr(self.con_w3_w), _ptr(self.con_w3_b), pairs, 256, 1, 0,
                 None, 0, 0, _ptr(con_logits), 1, st)
            call("egtr_relation_finish_f32", _ptr(rel_logits), P, _ptr(con_logits), 1, _ptr(logits), K,
                 _ptr(self.triplet), _ptr(self.rel_dist), float(getattr(cfg, "logit_adj_tau", 0.3)), int(bool(cfg.use_freq_bias)),
                 int(bool(cfg.logit_adjustment)), B, N, P, _ptr(ws["cls_idx"]), _ptr(pred_rel), _ptr(pred_con), st)
        elif fused:
            # ONE entry point: 14 per-query projections -> class argmax -> the fused pair kernel (relhead.cu);
            # nothing of the pair dimension except pred_rel / pred_connectivity reaches HBM
            qp = (C.c_void_p * nl)(*[_ptr(q, 0) for q in qkvs])
            kp = (C.c_void_p * nl)(*[_ptr(q, 256) for q in qkvs])
            call("egtr_relation_head_fwd_f32", qp, kp, 768, _ptr(h_last), 256, _ptr(logits), K, C.byref(self.rel_head_w),
                 _ptr(self.triplet), _ptr(self.rel_dist), float(getattr(cfg, "logit_adj_tau", 0.3)), int(bool(cfg.use_freq_bias)),
                 int(bool(cfg.logit_adjustment)), B, N, P, _ptr(ws["U"]), _ptr(ws["V"]), _ptr(ws["cls_idx"]), _ptr(pred_rel),
                 _ptr(pred_con), st)
        else:
            # EGTR_RELHEAD=unfused (round-1 path, dev A/B) — fused pair stage: the gating + layer 1 is the GEMM's operand producer (never in HBM), layer 2 of both MLPs is
            # one block-diagonal tcgen05 GEMM, the connectivity head's last layer + sigmoid is its epilogue; only the
            # relation MLP's 256-wide hidden goes to HBM for the layer-3 GEMM whose epilogue finishes pred_rel.
            TI, TJ = (N + 7) // 8, (N + 15) // 16
            src, ep = ASrc(), Epilogue()
            src.a, src.a2, src.aux, src.mode, src.lda = _ptr(ws["U"]), _ptr(ws["V"]), _ptr(self.rel_b1), 4, 516
            src.H, src.W, src.C, src.OH, src.OW = N, Lr, 256, TI, TJ
            H2r, rel_logits = self._pair_buf(ws, "H2r", pairs, 256), self._pair_buf(ws, "rel_logits", pairs, self.rel_w3p.N)
            ep.bias, ep.out, ep.ldo, ep.ldr, ep.relu = _ptr(self.rel_w2both.b), _ptr(H2r), 256, 256, 1
            ep.out_fmt = 1  # the relation MLP's hidden goes to HBM as P32 rows: layer 3 streams it by TMA
            ep.pair_n = N
            ep.dot_w, ep.dot_out, ep.dot_b, ep.dot_col0 = _ptr(self.con_w3_w), _ptr(pred_con), self.con_w3_b_host, 256
            lin = self.rel_w2both
            call("egtr_gemm_sbf16", C.byref(src), _ptr(lin.planes), B * TI * TJ * 128, lin.N, lin.Npad, lin.K, C.byref(ep), st)
            call("egtr_argmax_rows_f32", _ptr(logits), K, B * N, _ptr(ws["cls_idx"]), st)
            lin3 = self.rel_w3p
            self.gemm(lin3, pairs, rel_logits, a=H2r, lda=256, a_fmt=1, ldo=lin3.N)
            call("egtr_relation_finish_f32", _ptr(rel_logits), lin3.N, None, 0, None, K,
                 _ptr(self.triplet) if cfg.use_freq_bias else None, _ptr(self.rel_dist), float(getattr(cfg, "logit_adj_tau", 0.3)),
                 int(bool(cfg.use_freq_bias)), int(bool(cfg.logit_adjustment)), B, N, P, _ptr(ws["cls_idx"]), _ptr(pred_rel), None, st)
        _sp_rel.__exit__()
        return dict(logits=logits, pred_boxes=boxes, pred_rel=pred_rel, pred_connectivity=pred_con, **model_out)


class GraphRunner:
    """Static-buffer CUDA graph of `Engine._forward` for one (B, H, W).  Outputs are the graph's own
    buffers and are overwritten by the next replay (callers that keep results must clone them).

    `prologue(runner)` / `epilogue(runner, out) -> dict` (optional) are captured into the same graph before / after the
    forward: input staging that fills `runner.px` / `runner.pm` from the caller's own static buffers (uint8 images,
    egtr_b200/serving.py) and post-processing of the outputs (triplet extraction); the epilogue's result is `runner.extra`."""

    def __init__(self, eng: Engine, B: int, H: int, W: int, slot: int = 0, throughput: bool = False, prologue=None, epilogue=None):
        self.eng = eng
        self.slot = slot
        self.throughput = throughput
        self.extra = None
        dev = eng.device
        with torch.cuda.device(dev):
            self.ws = eng._workspace(B, H, W, slot)  # the captured graph holds raw pointers into it: keep it alive past a cache eviction
            self.px = torch.zeros(B, 3, H, W, dtype=torch.float32, device=dev)
            self.pm = torch.ones(B, H, W, dtype=torch.long, device=dev)

            def run():
                if prologue is not None:
                    prologue(self)
                out = eng._forward(self.px, self.pm, None, slot, throughput, static_out=True)
                return out, (epilogue(self, out) if epilogue is not None else None)

            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):  # warm-up: one-time attribute calls, tensor-map cache, workspace allocation
                    run()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            probe, eng.probe = eng.probe, None
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out, self.extra = run()
            eng.probe = probe

    def __call__(self, pixel_values: torch.Tensor, pixel_mask: Optional[torch.Tensor] = None):
        self.px.copy_(pixel_values, non_blocking=True)
        if pixel_mask is not None:
            self.pm.copy_(pixel_mask, non_blocking=True)
        else:
            self.pm.fill_(1)
        self.graph.replay()
        return self.out
