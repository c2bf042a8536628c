"""TEST INFRASTRUCTURE — CPU restatement of the input staging that precedes the hot path.

The reference stages images with HuggingFace `DetrFeatureExtractor` (transformers==4.18.0, `requirements.txt:5`; call sites
`/root/reference/data/visual_genome.py:64-66`, `/root/reference/train_egtr.py:176-186`, `/root/reference/evaluate_egtr.py:174-176`):
resize so that the shorter side is `size` (capped so the longer side stays <= `max_size`) with PIL bilinear resampling,
rescale to [0,1], normalise with the ImageNet mean / std, zero-pad a batch to its largest height / width and emit a
`pixel_mask` (1 = real pixel).  Neither transformers 4.18 nor its sources are in /root/reference, so:
  * the resampling is restated from Pillow's published algorithm (src/libImaging/Resample.c: precompute_coeffs,
    normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc) and PINNED bit-exactly against the Pillow that is
    installed here (tests/test_preprocess.py);
  * the size rule, rescale / normalise arithmetic and padding are restated from transformers 4.18's
    `feature_extraction_detr.py` (`get_size_with_aspect_ratio`, `_resize`, `_normalize`, `pad_and_create_pixel_mask`) —
    parity UNPINNED for that part (no copy of the library to run).
Only tests/ (and smoke / bench baselines) may import this module.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2  # Resample.c: 8 bits of pixel, 2 bits of headroom for the accumulated coefficients
IMAGE_MEAN = (0.485, 0.456, 0.406)
IMAGE_STD = (0.229, 0.224, 0.225)


def target_size(height: int, width: int, size: int = 800, max_size: int = 1333) -> Tuple[int, int]:
    """(out_h, out_w) of `DetrFeatureExtractor._resize` for an int `size` (get_size_with_aspect_ratio)."""
    w, h = width, height
    if max_size is not None:
        min_orig, max_orig = float(min(w, h)), float(max(w, h))
        if max_orig / min_orig * size > max_size:
            size = int(round(max_size * min_orig / max_orig))
    if (w <= h and w == size) or (h <= w and h == size):
        return h, w
    if w < h:
        ow = size
        oh = int(size * h / w)
    else:
        oh = size
        ow = int(size * w / h)
    return oh, ow


def bilinear_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the triangle filter (support 1.0) over the whole axis.
    Returns bounds [out,2] (first input index, tap count), integer coefficients [out, ksize], ksize."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        x = np.arange(xmax, dtype=np.float64)
        w = np.abs((x + xmin - center + 0.5) * ss)
        w = np.where(w < 1.0, 1.0 - w, 0.0)
        ww = w.sum()
        if ww != 0.0:
            w = w / ww
        kk[xx, :xmax] = w
        bounds[xx] = (xmin, xmax)
    ki = np.where(kk < 0, (-0.5 + kk * (1 << PRECISION_BITS)).astype(np.int64), (0.5 + kk * (1 << PRECISION_BITS)).astype(np.int64))
    return bounds, ki.astype(np.int32), ksize


def _resample_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One 8-bit pass (ImagingResampleHorizontal_8bpc / Vertical_8bpc): ss = 2^(P-1) + sum(pixel * k); clip8(ss >> P)."""
    in_size = img.shape[axis]
    bounds, ki, ksize = bilinear_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)  # [in, ...]
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        xmin, n = bounds[xx]
        acc = np.tensordot(ki[xx, :n].astype(np.int64), src[xmin:xmin + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_bilinear_resize(img_u8: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """uint8 [H,W,C] -> uint8 [out_h,out_w,C], bit-identical to PIL.Image.resize((out_w,out_h), BILINEAR): horizontal pass
    first, then vertical; a pass whose size does not change is skipped (ImagingResample)."""
    out = img_u8
    if out.shape[1] != out_w:
        out = _resample_axis(out, out_w, 1)
    if out.shape[0] != out_h:
        out = _resample_axis(out, out_h, 0)
    return out


def normalize(img_u8: np.ndarray) -> np.ndarray:
    """uint8 [H,W,3] -> float32 [3,H,W]: x / 255 then (x - mean) / std, all in float32 (feature_extraction_detr.py _normalize)."""
    x = img_u8.astype(np.float32) / np.float32(255.0)
    x = x.transpose(2, 0, 1)
    mean = np.array(IMAGE_MEAN, np.float32)[:, None, None]
    std = np.array(IMAGE_STD, np.float32)[:, None, None]
    return ((x - mean) / std).astype(np.float32)


def stage_batch(images: Sequence[np.ndarray], size: int = 800, max_size: int = 1333) -> Tuple[np.ndarray, np.ndarray, List[Tuple[int, int]]]:
    """list of uint8 [H,W,3] -> pixel_values f32 [B,3,Hm,Wm] (zero padded bottom/right), pixel_mask i64 [B,Hm,Wm], sizes."""
    outs = []
    for im in images:
        oh, ow = target_size(im.shape[0], im.shape[1], size, max_size)
        outs.append(normalize(pil_bilinear_resize(im, oh, ow)))
    hm, wm = max(o.shape[1] for o in outs), max(o.shape[2] for o in outs)
    px = np.zeros((len(outs), 3, hm, wm), np.float32)
    mask = np.zeros((len(outs), hm, wm), np.int64)
    for i, o in enumerate(outs):
        px[i, :, : o.shape[1], : o.shape[2]] = o
        mask[i, : o.shape[1], : o.shape[2]] = 1
    return px, mask, [(o.shape[1], o.shape[2]) for o in outs]
