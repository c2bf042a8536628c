"""Input staging on the device: uint8 images in, `pixel_values` / `pixel_mask` out.

The reference prepares every image on the host with HuggingFace `DetrFeatureExtractor` (transformers 4.18;
`/root/reference/data/visual_genome.py:64-66`, `/root/reference/train_egtr.py:176-186`): PIL bilinear resize to shorter side
`size` (longer side capped at `max_size`), rescale to [0,1], ImageNet normalisation, zero padding to the batch's largest size
and a `pixel_mask`.  `DeviceImageStager` produces the same tensors on the GPU from the raw uint8 images: the host only
computes the two small tap tables per image (Pillow's precompute_coeffs arithmetic) and uploads the uint8 pixels — 0.9 MB
for a 480x640 image instead of 21 MB of fp32 pixels + int64 mask — and two kernels of `libegtr_b200.so` do the rest
(bit-identical to Pillow's 8-bit resampling; IEEE fp32 divisions for the normalisation).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from ._lib import call

PRECISION_BITS = 32 - 8 - 2
IMAGE_MEAN = (0.485, 0.456, 0.406)
IMAGE_STD = (0.229, 0.224, 0.225)


def target_size(height: int, width: int, size: int = 800, max_size: Optional[int] = 1333) -> Tuple[int, int]:
    """(out_h, out_w): shorter side -> `size`, unless that pushes the longer side past `max_size` (transformers 4.18
    feature_extraction_detr.py, get_size_with_aspect_ratio)."""
    w, h = width, height
    if max_size is not None:
        lo, hi = float(min(w, h)), float(max(w, h))
        if hi / lo * size > max_size:
            size = int(round(max_size * lo / hi))
    if (w <= h and w == size) or (h <= w and h == size):
        return h, w
    if w < h:
        return int(size * h / w), size
    return size, int(size * w / h)


def tap_tables(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Triangle-filter taps of one axis as Pillow computes them (Resample.c: precompute_coeffs, normalize_coeffs_8bpc):
    bounds int32 [out,2] = (first input index, tap count), taps int32 [out,ksize] in 22-bit fixed point."""
    if in_size == out_size:  # Pillow skips the pass; an identity tap (1.0 in fixed point) reproduces the pixel exactly
        b = np.stack([np.arange(out_size, dtype=np.int32), np.ones(out_size, np.int32)], 1)
        return np.ascontiguousarray(b), np.full((out_size, 1), 1 << PRECISION_BITS, np.int32), 1
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    support = fscale  # bilinear support 1.0 x filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)   # C casts truncate; the operands are non-negative here
    xmin = np.where(center - support + 0.5 < 0, 0, xmin)
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size)
    n = xmax - xmin
    t = np.arange(ksize, dtype=np.float64)[None, :]
    w = np.abs((t + xmin[:, None] - center[:, None] + 0.5) * (1.0 / fscale))
    w = np.where(w < 1.0, 1.0 - w, 0.0)
    w = np.where(t < n[:, None], w, 0.0)
    ww = np.zeros(out_size, np.float64)
    for j in range(ksize):  # Pillow accumulates the normaliser tap by tap, in this order
        ww = ww + w[:, j]
    w = np.where(ww[:, None] != 0.0, w / np.where(ww == 0.0, 1.0, ww)[:, None], w)
    ki = np.where(w < 0, (-0.5 + w * (1 << PRECISION_BITS)).astype(np.int64), (0.5 + w * (1 << PRECISION_BITS)).astype(np.int64))
    bounds = np.stack([xmin, n], 1).astype(np.int32)
    return np.ascontiguousarray(bounds), np.ascontiguousarray(ki.astype(np.int32)), ksize


class DeviceImageStager:
    """`stage(images)` with `images` a list of uint8 [H,W,3] host arrays/tensors (RGB) -> (`pixel_values` f32 [B,3,Hm,Wm],
    `pixel_mask` i64 [B,Hm,Wm]) on `device`, equal to `DetrFeatureExtractor(size, max_size)(images)` + `pad_and_create_pixel_mask`."""

    def __init__(self, size: int = 800, max_size: Optional[int] = 1333, device="cuda"):
        self.size, self.max_size = size, max_size
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceImageStager stages on a CUDA device (the host path is the reference's own feature extractor)")
        self._mean = (C.c_float * 3)(*IMAGE_MEAN)
        self._std = (C.c_float * 3)(*IMAGE_STD)

    def stage(self, images: Sequence) -> Tuple[torch.Tensor, torch.Tensor, List[Tuple[int, int]]]:
        imgs = [torch.as_tensor(np.ascontiguousarray(im)) if not isinstance(im, torch.Tensor) else im.contiguous() for im in images]
        for im in imgs:
            if im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3:
                raise ValueError("images must be uint8 [H, W, 3]")
        sizes = [target_size(im.shape[0], im.shape[1], self.size, self.max_size) for im in imgs]
        hm, wm = max(s[0] for s in sizes), max(s[1] for s in sizes)
        dev = self.device
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream().cuda_stream
            px = torch.zeros(len(imgs), 3, hm, wm, dtype=torch.float32, device=dev)
            mask = torch.zeros(len(imgs), hm, wm, dtype=torch.int64, device=dev)
            keep = []
            for b, (im, (oh, ow)) in enumerate(zip(imgs, sizes)):
                h, w = im.shape[0], im.shape[1]
                bh, kh, ksh = tap_tables(w, ow)
                bv, kv, ksv = tap_tables(h, oh)
                src = im.to(dev, non_blocking=True)
                t_bh, t_kh, t_bv, t_kv = (torch.from_numpy(a).to(dev, non_blocking=True) for a in (bh, kh, bv, kv))
                tmp = torch.empty(h, ow, 3, dtype=torch.uint8, device=dev)
                call("egtr_resample_h_u8", src.data_ptr(), h, w, 3, ow, t_bh.data_ptr(), t_kh.data_ptr(), ksh, tmp.data_ptr(), st)
                call("egtr_resample_v_normalize_f32", tmp.data_ptr(), h, ow, 3, oh, t_bv.data_ptr(), t_kv.data_ptr(), ksv, self._mean,
                     self._std, px[b].data_ptr(), hm * wm, wm, mask[b].data_ptr(), wm, st)
                keep.append((src, t_bh, t_kh, t_bv, t_kv, tmp))
            for tensors in keep:  # allocations stay alive until the kernels that read them are enqueued on this stream
                for t in tensors:
                    t.record_stream(torch.cuda.current_stream())
        return px, mask, sizes


class StaticStager:
    """Fixed-shape staging for CUDA-graph capture (egtr_b200/serving.py): `B` uint8 RGB images of one raw size
    [B, h0, w0, 3] in a static device buffer `u8` -> `pixel_values` / `pixel_mask` of a (Hm, Wm) batch tensor, with the same
    two kernels (and the same Pillow-exact arithmetic) as `DeviceImageStager`.  The tap tables are built once; a step uploads
    only the uint8 pixels (3.2 MB for 800x1333 instead of 21.3 MB of fp32 pixels + int64 mask)."""

    def __init__(self, B: int, raw_hw: Tuple[int, int], pad_hw: Tuple[int, int], device, size: int = 800, max_size: Optional[int] = 1333,
                 out_hw: Optional[Tuple[int, int]] = None):
        self.B, self.raw_hw, self.pad_hw = B, tuple(raw_hw), tuple(pad_hw)
        h0, w0 = self.raw_hw
        self.out_hw = tuple(out_hw) if out_hw is not None else target_size(h0, w0, size, max_size)
        oh, ow = self.out_hw
        if oh > pad_hw[0] or ow > pad_hw[1]:
            raise ValueError(f"resized image {self.out_hw} does not fit the batch tensor {self.pad_hw}")
        self.device = torch.device(device)
        bh, kh, self.ksh = tap_tables(w0, ow)
        bv, kv, self.ksv = tap_tables(h0, oh)
        with torch.cuda.device(self.device):
            self.u8 = torch.zeros(B, h0, w0, 3, dtype=torch.uint8, device=self.device)
            self.t_bh, self.t_kh, self.t_bv, self.t_kv = (torch.from_numpy(a).to(self.device) for a in (bh, kh, bv, kv))
            self.tmp = torch.empty(h0, ow, 3, dtype=torch.uint8, device=self.device) if w0 != ow else None
        self._mean = (C.c_float * 3)(*IMAGE_MEAN)
        self._std = (C.c_float * 3)(*IMAGE_STD)

    def enqueue(self, px: torch.Tensor, pm: torch.Tensor) -> None:
        """Launch the staging kernels on the current stream: self.u8 -> px [B,3,Hm,Wm] f32, pm [B,Hm,Wm] i64."""
        (h0, w0), (oh, ow), (hm, wm) = self.raw_hw, self.out_hw, self.pad_hw
        st = torch.cuda.current_stream().cuda_stream
        if (oh, ow) != (hm, wm):
            px.zero_()
            pm.zero_()
        for b in range(self.B):
            src = self.u8[b]
            if self.tmp is not None:
                call("egtr_resample_h_u8", src.data_ptr(), h0, w0, 3, ow, self.t_bh.data_ptr(), self.t_kh.data_ptr(), self.ksh,
                     self.tmp.data_ptr(), st)
                src = self.tmp
            call("egtr_resample_v_normalize_f32", src.data_ptr(), h0, ow, 3, oh, self.t_bv.data_ptr(), self.t_kv.data_ptr(), self.ksv,
                 self._mean, self._std, px[b].data_ptr(), hm * wm, wm, pm[b].data_ptr(), wm, st)
