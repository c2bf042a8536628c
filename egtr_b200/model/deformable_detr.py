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
        out_logits, out_bbox = outputs.logits, outputs.pred_boxes
        if len(out_logits) != len(target_sizes):
            raise ValueError("Make sure that you pass in as many target sizes as the batch dimension of the logits")
        if target_sizes.shape[1] != 2:
            raise ValueError("Each element of target_sizes must contain the size (h, w) of each image of the batch")
        prob = out_logits.sigmoid()
        topk_values, topk_indexes = torch.topk(prob.view(out_logits.shape[0], -1), 100, dim=1)
        topk_boxes = torch.div(topk_indexes, out_logits.shape[2], rounding_mode="trunc")
        labels = topk_indexes % out_logits.shape[2]
        boxes = center_to_corners_format(out_bbox)
        boxes = torch.gather(boxes, 1, topk_boxes.unsqueeze(-1).repeat(1, 1, 4))
        img_h, img_w = target_sizes.unbind(1)
        scale_fct = torch.stack([img_w, img_h, img_w, img_h], dim=1)
        boxes = boxes * scale_fct[:, None, :]
        return [{"scores": s, "labels": l, "boxes": b} for s, l, b in zip(topk_values, labels, boxes)]


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


class DeformableDetrModel(nn.Module):
    """Placeholder for import compatibility: the bare encoder-decoder is only reachable through
    `DetrForSceneGraphGeneration` on the B200 path (`egtr_b200/model/egtr.py`)."""

    def __init__(self, config):
        super().__init__()
        raise NotImplementedError("use egtr_b200.model.egtr.DetrForSceneGraphGeneration; the bare model is not a separate entry point here")
