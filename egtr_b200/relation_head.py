"""The relation head as a stand-alone operator (boundary B2): the part of
`DetrForSceneGraphGeneration.forward` that turns the captured decoder self-attention queries / keys into `pred_rel` and
`pred_connectivity` (`/root/reference/model/egtr.py:322-418, 507-516`), behind ONE native entry point,
`egtr_relation_head_fwd_f32` (include/egtr_b200.h; relhead.cu).

    head = RelationHead(config, model.state_dict(), device="cuda")      # reference key names; prepared once
    pred_rel, pred_connectivity = head(outputs.decoder_attention_queries, outputs.decoder_attention_keys,
                                       sequence_output, logits)          # the tensors the reference forward has at line 322

A maintainer who keeps the reference's PyTorch backbone / transformer can call this in place of lines 322-418 + 507-516.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Sequence, Tuple

import torch

from . import _lib
from ._lib import call
from .engine import Engine, _ptr


class RelationHead:
    def __init__(self, config, state_dict: Dict[str, torch.Tensor], device="cuda"):
        _lib.load()
        self.cfg = config
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.EgtrError("the relation head runs on CUDA devices only (no CPU fallback)")
        keys = ("rel_predictor", "connectivity_layer", "proj_q", "proj_k", "final_sub_proj", "final_obj_proj", "triplet_dist", "rel_dist")
        sd = {k: v.detach().to(self.device, torch.float32) for k, v in state_dict.items() if k.startswith(keys)}
        # weight preparation is the engine's (fp64 composition of proj_* with layer 1, bf16 hi/lo packing for the TMA boxes)
        self._w = Engine.__new__(Engine)
        self._w.cfg, self._w.device = config, self.device
        with torch.cuda.device(self.device):
            self._w._prepare_relation_head(sd)
            torch.cuda.current_stream().synchronize()
        if not self._w.rel_fused:
            raise _lib.EgtrError("egtr_relation_head_fwd_f32 is built for <= 7 relation layers and <= 256 predicates")

    @torch.no_grad()
    def __call__(self, queries: Sequence[torch.Tensor], keys: Sequence[torch.Tensor], h_last: torch.Tensor,
                 logits: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """queries / keys: per decoder layer [B, heads, N, 32] (queries SCALED, as captured at deformable_detr.py:1179-1185);
        h_last [B, N, 256]; logits [B, N, K].  Returns sigmoid-ed pred_rel [B,N,N,P] and pred_connectivity [B,N,N,1]."""
        cfg, w = self.cfg, self._w
        B, N, K = logits.shape
        P, nl = cfg.num_rel_labels, cfg.decoder_layers
        if len(queries) != nl or len(keys) != nl:
            raise ValueError(f"expected {nl} captured query / key tensors, got {len(queries)} / {len(keys)}")
        dev = self.device
        rows = lambda t: t.to(dev, torch.float32).transpose(1, 2).reshape(B * N, 256).contiguous()  # noqa: E731
        q, k = [rows(t) for t in queries], [rows(t) for t in keys]
        h = h_last.to(dev, torch.float32).reshape(B * N, 256).contiguous()
        lg = logits.to(dev, torch.float32).contiguous()
        with torch.cuda.device(dev):
            U = torch.empty(B * N * (nl + 1) * 516, dtype=torch.float32, device=dev)
            V = torch.empty_like(U)
            cls = torch.empty(B * N, dtype=torch.int32, device=dev)
            pred_rel = torch.empty(B, N, N, P, dtype=torch.float32, device=dev)
            pred_con = torch.empty(B, N, N, 1, dtype=torch.float32, device=dev)
            qp = (C.c_void_p * nl)(*[_ptr(t) for t in q])
            kp = (C.c_void_p * nl)(*[_ptr(t) for t in k])
            call("egtr_relation_head_fwd_f32", qp, kp, 256, _ptr(h), 256, _ptr(lg), K, C.byref(w.rel_head_w), _ptr(w.triplet),
                 _ptr(w.rel_dist), float(getattr(cfg, "logit_adj_tau", 0.3)), int(bool(cfg.use_freq_bias)),
                 int(bool(cfg.logit_adjustment)), B, N, P, _ptr(U), _ptr(V), _ptr(cls), _ptr(pred_rel), _ptr(pred_con),
                 torch.cuda.current_stream().cuda_stream)
        return pred_rel, pred_con
