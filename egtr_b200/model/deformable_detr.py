"""Mirror of the reference's `model.deformable_detr` public surface for the inference hot path.

What `evaluate_egtr.py` imports from the reference module (`/root/reference/evaluate_egtr.py:21-22`):
`DeformableDetrConfig`, `DeformableDetrFeatureExtractor`; plus the native-op wrapper
`MultiScaleDeformableAttentionFunction` (`/root/reference/model/deformable_detr.py:402-455`) whose
`forward` is served here by the C-ABI kernel `egtr_msda_fwd_f32` instead of the JIT-built pybind
module (`/root/reference/model/load_custom.py:23-57`).

Deviation (documented in DESIGN.md): when the native op cannot run, the reference prints and
silently falls back to `grid_sample` (`deformable_detr.py:1096-1101`); this package raises.
"""
from __future__ import annotations

import json
import os
from typing import List, Optional

import torch
from torch import nn

from .. import _lib
from ..config import DeformableDetrConfig  # noqa: F401  (re-export: part of the boundary)
from .outputs import DeformableDetrModelOutput  # noqa: F401

__all__ = [
    "DeformableDetrConfig", "DeformableDetrFeatureExtractor", "DeformableDetrModel",
    "MultiScaleDeformableAttentionFunction", "ms_deform_attn_forward", "inverse_sigmoid",
]


def ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                           attention_weights, im2col_step: int = 64):
    """Signature of the reference pybind op (`model/custom_kernel/vision.cpp:13`,
    `ms_deform_attn.h:20-39`): value [B,S,M,D], spatial_shapes [L,2] int64 (device),
    level_start_index [L] int64 (device), sampling_locations [B,Lq,M,L,P,2],
    attention_weights [B,Lq,M,L,P] -> [B,Lq,M*D].  `im2col_step` only chunked the reference's
    launch (`cu:53-78`); it has no effect on results and is accepted for compatibility."""
    for name, t in (("value", value), ("spatial_shapes", value_spatial_shapes), ("level_start_index", value_level_start_index),
                    ("sampling_loc", sampling_locations), ("attn_weight", attention_weights)):
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")  # AT_ERROR("Not implemented on the CPU"), ms_deform_attn.h:38
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")  # cu:31-35
    if value.dtype != torch.float32:
        raise RuntimeError("ms_deform_attn_forward: only float32 is built (reference also dispatched float64)")
    B, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    out = torch.empty(B, Lq, M * D, dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device):
        _lib.call("egtr_msda_fwd_f32", value.data_ptr(), value_spatial_shapes.to(torch.int64).data_ptr(),
                  value_level_start_index.to(torch.int64).data_ptr(), sampling_locations.data_ptr(),
                  attention_weights.data_ptr(), B, S, M, D, L, Lq, P, out.data_ptr(),
                  torch.cuda.current_stream().cuda_stream)
    return out


class MultiScaleDeformableAttentionFunction(torch.autograd.Function):
    """Forward-only equivalent of `deformable_detr.py:402-455` (backward is training, out of scope)."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights, im2col_step):
        return ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                                      attention_weights, im2col_step)

    @staticmethod
    def backward(ctx, grad_output):
        raise NotImplementedError("MSDeformAttn backward belongs to training; the B200 path is inference-only (SURVEY.md §2.1)")


def inverse_sigmoid(x, eps=1e-5):
    """`deformable_detr.py:658-662`."""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def center_to_corners_format(x):
    cx, cy, w, h = x.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)


class DeformableDetrFeatureExtractor:
    """The part of HF `DetrFeatureExtractor` the evaluation script touches: construction with
    `size` / `max_size` (`evaluate_egtr.py:174-176`), batch padding to a `pixel_mask`
    (`train_egtr.py:176-186`) and `post_process` (`deformable_detr.py:273-319`).  Image decoding /
    resizing of dataset samples is the §8f-2 "next" row and stays on the caller's side."""

    image_mean = [0.485, 0.456, 0.406]
    image_std = [0.229, 0.224, 0.225]

    def __init__(self, size: int = 800, max_size: int = 1333, **kwargs):
        self.size, self.max_size = size, max_size

    @classmethod
    def from_pretrained(cls, name_or_path: str = "", size: int = 800, max_size: int = 1333, **kwargs):
        return cls(size=size, max_size=max_size, **kwargs)

    def pad_and_create_pixel_mask(self, pixel_values_list: List[torch.Tensor], return_tensors: str = "pt"):
        """Zero-pad CHW images to the largest H, W of the batch; mask 1 = real pixel."""
        c = pixel_values_list[0].shape[0]
        H = max(int(im.shape[1]) for im in pixel_values_list)
        W = max(int(im.shape[2]) for im in pixel_values_list)
        px = torch.zeros(len(pixel_values_list), c, H, W, dtype=torch.float32)
        mask = torch.zeros(len(pixel_values_list), H, W, dtype=torch.long)
        for i, im in enumerate(pixel_values_list):
            im = torch.as_tensor(im, dtype=torch.float32)
            px[i, :, : im.shape[1], : im.shape[2]] = im
            mask[i, : im.shape[1], : im.shape[2]] = 1
        return {"pixel_values": px, "pixel_mask": mask}

    def post_process(self, outputs, target_sizes):
        """Detections per image from raw outputs (role of `deformable_detr.py:273-319`): the 100 best (query, class) pairs by
        sigmoid score over all N*K pairs, their boxes as absolute (x0, y0, x1, y1) in the image sizes `target_sizes` [B,2] = (h, w)."""
        logits, cxcywh = outputs.logits, outputs.pred_boxes
        target_sizes = torch.as_tensor(target_sizes, device=logits.device)
        if target_sizes.dim() != 2 or target_sizes.shape != (logits.shape[0], 2):
            raise ValueError(f"target_sizes must be [batch, 2] = (height, width) per image; got {tuple(target_sizes.shape)} for batch {logits.shape[0]}")
        B, N, K = logits.shape
        scores, flat = logits.sigmoid().flatten(1).topk(min(100, N * K), dim=1)
        query, label = flat // K, flat % K
        corners = center_to_corners_format(cxcywh)[torch.arange(B, device=logits.device)[:, None], query]  # [B,100,4]
        wh = target_sizes.to(corners.dtype).flip(1).repeat(1, 2)  # (w, h, w, h)
        corners = corners * wh[:, None, :]
        return [dict(scores=scores[b], labels=label[b], boxes=corners[b]) for b in range(B)]


def _attach(root: nn.Module, dotted: str, tensor: torch.Tensor, buffer: bool):
    """Register `tensor` under the reference's dotted state-dict key, creating bare container modules."""
    parts = dotted.split(".")
    mod = root
    for i, p in enumerate(parts[:-1]):
        if p not in mod._modules:
            # containers indexed by integers are ModuleLists in the reference (layers, input_proj, class_embed, ...)
            mod.add_module(p, nn.ModuleList() if parts[i + 1].isdigit() else nn.Module())
        mod = mod._modules[p]
    if buffer:
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


_BUFFER_SUFFIXES = ("running_mean", "running_var", "num_batches_tracked")


def _is_frozen_bn_key(k: str) -> bool:
    # FrozenBN keeps weight/bias as buffers too (deformable_detr.py:673-678); `downsample.1` stays a BatchNorm2d
    tail = k.rsplit(".", 2)
    return "backbone" in k and (tail[-2].startswith("bn") and tail[-1] in ("weight", "bias"))


class _WeightTree(nn.Module):
    """nn.Module whose state_dict() keys/shapes equal the reference model's."""

    def __init__(self, config, prefix_filter=None):
        super().__init__()
        from ..synth import synth_state_dict

        self.config = config
        sd = synth_state_dict(config, seed=0)
        aliased = {}
        for k, v in sd.items():
            if prefix_filter is not None and not prefix_filter(k):
                continue
            # class_embed.{1..5} / bbox_embed.{1..5} alias module 0 (egtr.py:154-157)
            head = k.split(".", 2)
            if head[0] in ("class_embed", "bbox_embed") and head[1] != "0":
                aliased[(head[0], head[1])] = True
                continue
            _attach(self, k, v.clone(), buffer=k.endswith(_BUFFER_SUFFIXES) or _is_frozen_bn_key(k))
        for (name, idx) in aliased:
            self._modules[name].add_module(idx, self._modules[name]._modules["0"])


class DeformableDetrModel(_WeightTree):
    """The bare encoder-decoder (`/root/reference/model/deformable_detr.py:1978-2390`) as its own entry point: same state-dict
    keys as the reference module (the `model.*` subtree of the scene-graph model without the prefix), same forward keywords,
    `DeformableDetrModelOutput` out — served by the same engine launches as `DetrForSceneGraphGeneration` up to the decoder."""

    config_class = DeformableDetrConfig

    def __init__(self, config):
        nn.Module.__init__(self)
        from ..synth import synth_state_dict
        self.config = config
        for k, v in synth_state_dict(config, seed=0).items():
            if k.startswith("model."):
                k = k[len("model."):]
                _attach(self, k, v.clone(), buffer=k.endswith(_BUFFER_SUFFIXES) or _is_frozen_bn_key(k))
        self._engine = None
        self.use_cuda_graph = False

    @property
    def device(self):
        return self.level_embed.device

    def load_state_dict(self, state_dict, strict: bool = True):
        out = super().load_state_dict(dict(state_dict), strict=strict)
        self._engine = None
        return out

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def engine(self):
        from ..engine import Engine
        if self._engine is None:
            if self.device.type != "cuda":
                raise _lib.EgtrError("model is on %s: call .cuda() first — the B200 path has no CPU fallback" % self.device)
            self._engine = Engine(self.config, {"model." + k: v for k, v in self.state_dict().items()}, self.device, model_only=True)
        return self._engine

    @torch.no_grad()
    def forward(self, pixel_values, pixel_mask=None, decoder_attention_mask=None, encoder_outputs=None, inputs_embeds=None,
                decoder_inputs_embeds=None, output_attentions=None, output_attention_states=None, output_hidden_states=None,
                return_dict=None):
        if encoder_outputs is not None or inputs_embeds is not None or decoder_inputs_embeds is not None or decoder_attention_mask is not None:
            raise NotImplementedError("encoder_outputs / *_embeds / decoder_attention_mask are unused by the EGTR path and not built")
        if output_attentions:
            raise NotImplementedError("output_attentions=True (attention maps) is not built; the EGTR path captures queries / keys instead")
        if self.use_cuda_graph:
            B, _, H, W = pixel_values.shape
            with torch.cuda.device(self.device):
                o = self.engine().graph_runner(B, H, W)(pixel_values, pixel_mask)
        else:
            o = self.engine().forward(pixel_values, pixel_mask)
        inter = o["intermediate_hidden_states"]
        dec_states = None
        if output_hidden_states:
            tgt = self.engine().query_tgt.unsqueeze(0).expand(inter.shape[0], -1, -1)
            dec_states = (tgt,) + tuple(inter[:, i] for i in range(inter.shape[1]))
        out = DeformableDetrModelOutput(
            init_reference_points=o["init_reference_points"], last_hidden_state=o["last_hidden_state"],
            intermediate_hidden_states=inter, intermediate_reference_points=o["intermediate_reference_points"],
            decoder_hidden_states=dec_states, encoder_last_hidden_state=o["encoder_last_hidden_state"],
            decoder_attention_queries=o["decoder_attention_queries"] if output_attention_states else None,
            decoder_attention_keys=o["decoder_attention_keys"] if output_attention_states else None,
        )
        return_dict = return_dict if return_dict is not None else getattr(self.config, "use_return_dict", True)
        return out if return_dict else out.to_tuple()
