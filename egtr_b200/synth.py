"""Deterministic synthetic weights and inputs for the EGTR hot path.

There is no network for checkpoints, so tests, goldens and the bench all use
seeded random weights carrying the reference's state-dict key names and shapes
(`/root/reference/model/egtr.py:123-226`, `model/deformable_detr.py:1978-2048`;
key list in SURVEY.md §8b).  numpy's PCG64 stream is bit-identical on every
machine, so the GPU box regenerates exactly the tensors the goldens were made
from without shipping 170 MB of weights.

The default reference init is degenerate for parity testing (SURVEY.md §8c
"traps": zero sampling_offsets weight, zero attention_weights, zero last
bbox layer, identity FrozenBN, uninitialised triplet_dist/rel_dist), so every
tensor here is drawn with a scale that keeps activations O(1) end to end.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

RESNET50_BLOCKS = (3, 4, 6, 3)
RESNET50_PLANES = (64, 128, 256, 512)


def _normal(rng, shape, std):
    return (rng.standard_normal(shape, dtype=np.float32) * np.float32(std)).astype(np.float32)


def _uniform(rng, shape, lo, hi):
    return rng.uniform(lo, hi, size=shape).astype(np.float32)


def _conv(rng, sd, name, cout, cin, k, bias=False, gain=1.0):
    fan_in = cin * k * k
    sd[name + ".weight"] = _normal(rng, (cout, cin, k, k), gain / math.sqrt(fan_in))
    if bias:
        sd[name + ".bias"] = _normal(rng, (cout,), 0.05)


def _bn(rng, sd, name, c, w_lo=0.8, w_hi=1.2, tracked=False):
    sd[name + ".weight"] = _uniform(rng, (c,), w_lo, w_hi)
    sd[name + ".bias"] = _normal(rng, (c,), 0.1)
    sd[name + ".running_mean"] = _normal(rng, (c,), 0.1)
    sd[name + ".running_var"] = _uniform(rng, (c,), 0.5, 1.5)
    if tracked:  # the reference leaves `downsample.1` as nn.BatchNorm2d (deformable_detr.py:718-730)
        sd[name + ".num_batches_tracked"] = np.zeros((), dtype=np.int64)


def _linear(rng, sd, name, cout, cin, std=None, bias_std=0.05):
    std = (1.0 / math.sqrt(cin)) if std is None else std
    sd[name + ".weight"] = _normal(rng, (cout, cin), std)
    sd[name + ".bias"] = _normal(rng, (cout,), bias_std)


def _ln(rng, sd, name, c):
    sd[name + ".weight"] = _uniform(rng, (c,), 0.8, 1.2)
    sd[name + ".bias"] = _normal(rng, (c,), 0.05)


def _msda(rng, sd, name, d, heads, levels, points):
    _linear(rng, sd, name + ".sampling_offsets", heads * levels * points * 2, d, std=0.06, bias_std=1.5)
    _linear(rng, sd, name + ".attention_weights", heads * levels * points, d, std=0.06, bias_std=0.3)
    _linear(rng, sd, name + ".value_proj", d, d)
    _linear(rng, sd, name + ".output_proj", d, d)


def _mlp3(rng, sd, name, cin, hid, cout):
    _linear(rng, sd, name + ".layers.0", hid, cin)
    _linear(rng, sd, name + ".layers.1", hid, hid)
    _linear(rng, sd, name + ".layers.2", cout, hid, std=2.0 / math.sqrt(hid), bias_std=0.5)


def synth_state_dict(config, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Random state dict with the reference's exact key names/shapes (strict-loadable
    into `/root/reference/model/egtr.py::DetrForSceneGraphGeneration`)."""
    rng = np.random.default_rng(seed)
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    d = config.d_model
    K, P, N = config.num_labels, config.num_rel_labels, config.num_queries
    M, L = config.encoder_attention_heads, config.num_feature_levels

    sd["triplet_dist"] = _normal(rng, (K + 1, K + 1, P), 1.0)
    e = rng.standard_normal(P).astype(np.float64)
    e = np.exp(e - e.max())
    sd["rel_dist"] = (e / e.sum()).astype(np.float32)
    sd["model.level_embed"] = _normal(rng, (L, d), 1.0)

    # ResNet-50 v1.5 (timm 0.5.4 `resnet50`, features_only; deformable_detr.py:748-755)
    bb = "model.backbone.conv_encoder.model."
    _conv(rng, sd, bb + "conv1", 64, 3, 7, gain=1.4)
    _bn(rng, sd, bb + "bn1", 64)
    inplanes = 64
    for li, (nblk, planes) in enumerate(zip(RESNET50_BLOCKS, RESNET50_PLANES), start=1):
        for bi in range(nblk):
            p = f"{bb}layer{li}.{bi}."
            _conv(rng, sd, p + "conv1", planes, inplanes, 1, gain=1.4)
            _bn(rng, sd, p + "bn1", planes)
            _conv(rng, sd, p + "conv2", planes, planes, 3, gain=1.4)
            _bn(rng, sd, p + "bn2", planes)
            _conv(rng, sd, p + "conv3", planes * 4, planes, 1, gain=1.0)
            _bn(rng, sd, p + "bn3", planes * 4, 0.3, 0.6)
            if bi == 0:
                _conv(rng, sd, p + "downsample.0", planes * 4, inplanes, 1, gain=1.0)
                _bn(rng, sd, p + "downsample.1", planes * 4, 0.6, 0.9, tracked=True)
            inplanes = planes * 4

    chans = [512, 1024, 2048]
    for i in range(L):
        if i < len(chans):
            _conv(rng, sd, f"model.input_proj.{i}.0", d, chans[i], 1, bias=True)
        else:
            _conv(rng, sd, f"model.input_proj.{i}.0", d, chans[-1] if i == len(chans) else d, 3, bias=True)
        sd[f"model.input_proj.{i}.1.weight"] = _uniform(rng, (d,), 0.8, 1.2)
        sd[f"model.input_proj.{i}.1.bias"] = _normal(rng, (d,), 0.05)

    sd["model.query_position_embeddings.weight"] = _normal(rng, (N, 2 * d), 1.0)

    for i in range(config.encoder_layers):
        p = f"model.encoder.layers.{i}."
        _msda(rng, sd, p + "self_attn", d, M, L, config.encoder_n_points)
        _ln(rng, sd, p + "self_attn_layer_norm", d)
        _linear(rng, sd, p + "fc1", config.encoder_ffn_dim, d)
        _linear(rng, sd, p + "fc2", d, config.encoder_ffn_dim)
        _ln(rng, sd, p + "final_layer_norm", d)

    for i in range(config.decoder_layers):
        p = f"model.decoder.layers.{i}."
        for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):
            _linear(rng, sd, p + "self_attn." + nm, d, d, std=1.5 / math.sqrt(d) if nm in ("q_proj", "k_proj") else None)
        _ln(rng, sd, p + "self_attn_layer_norm", d)
        _msda(rng, sd, p + "encoder_attn", d, config.decoder_attention_heads, L, config.decoder_n_points)
        _ln(rng, sd, p + "encoder_attn_layer_norm", d)
        _linear(rng, sd, p + "fc1", config.decoder_ffn_dim, d)
        _linear(rng, sd, p + "fc2", d, config.decoder_ffn_dim)
        _ln(rng, sd, p + "final_layer_norm", d)

    _linear(rng, sd, "model.reference_points", 2, d, std=1.0 / math.sqrt(d), bias_std=0.1)

    # class_embed / bbox_embed: six aliases of ONE module when with_box_refine=False (egtr.py:154-157)
    cls, box = OrderedDict(), OrderedDict()
    _linear(rng, cls, "c", K, d, std=2.0 / math.sqrt(d), bias_std=0.5)
    _mlp3(rng, box, "b", d, d, 4)
    for i in range(config.decoder_layers):
        for k, v in cls.items():
            sd[k.replace("c.", f"class_embed.{i}.", 1)] = v
        for k, v in box.items():
            sd[k.replace("b.", f"bbox_embed.{i}.", 1)] = v

    for i in range(config.decoder_layers):
        _linear(rng, sd, f"proj_q.{i}", d, d)
    for i in range(config.decoder_layers):
        _linear(rng, sd, f"proj_k.{i}", d, d)
    _linear(rng, sd, "final_sub_proj", d, d)
    _linear(rng, sd, "final_obj_proj", d, d)
    _linear(rng, sd, "rel_predictor_gate", 1, 2 * d, std=1.0 / math.sqrt(2 * d), bias_std=0.2)
    _mlp3(rng, sd, "rel_predictor", 2 * d, d, P)
    _mlp3(rng, sd, "connectivity_layer", 2 * d, d, 1)
    return OrderedDict((k, torch.from_numpy(np.ascontiguousarray(v))) for k, v in sd.items())


def synth_images(batch, height, width, seed: int = 1, pad_to=None):
    """`pixel_values` ~ N(0,1) (ImageNet-normalised range) and an all-valid `pixel_mask`.

    `pad_to=[(h_i, w_i), ...]` marks image i as valid only in its top-left h_i x w_i corner
    (zero pixels elsewhere), the way `collate_fn` pads ragged batches (train_egtr.py:176-186)."""
    rng = np.random.default_rng(seed)
    px = rng.standard_normal((batch, 3, height, width), dtype=np.float32)
    mask = np.ones((batch, height, width), dtype=np.int64)
    if pad_to is not None:
        for i, (h, w) in enumerate(pad_to):
            mask[i, h:, :] = 0
            mask[i, :, w:] = 0
            px[i, :, h:, :] = 0
            px[i, :, :, w:] = 0
    return torch.from_numpy(px), torch.from_numpy(mask)


def synth_msda_inputs(batch, shapes, n_query, heads=8, head_dim=32, points=4, seed: int = 2):
    """Kernel-level micro-inputs for MSDeformAttn: value ~ N(0,1),
    sampling_loc ~ U(-0.1, 1.1) (~8 % out of bounds -> padding branch, cuh:288),
    attn_weight = softmax(randn) over the L*P samples."""
    rng = np.random.default_rng(seed)
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    value = rng.standard_normal((batch, S, heads, head_dim), dtype=np.float32)
    loc = rng.uniform(-0.1, 1.1, size=(batch, n_query, heads, L, points, 2)).astype(np.float32)
    a = rng.standard_normal((batch, n_query, heads, L * points)).astype(np.float64)
    a = np.exp(a - a.max(-1, keepdims=True))
    w = (a / a.sum(-1, keepdims=True)).astype(np.float32).reshape(batch, n_query, heads, L, points)
    spatial = np.asarray(shapes, dtype=np.int64)
    start = np.concatenate([[0], np.cumsum(spatial[:, 0] * spatial[:, 1])[:-1]]).astype(np.int64)
    return tuple(torch.from_numpy(np.ascontiguousarray(x)) for x in (value, spatial, start, loc, w))
