"""`DetrForSceneGraphGeneration` — the reference's model API (`/root/reference/model/egtr.py:122-540`)
served by the B200 engine.

Kept from the reference: constructor signature `(config, fg_matrix=None)`, `from_pretrained(arch,
config=..., ignore_mismatched_sizes=True)`, the exact state-dict key set (strict `load_state_dict`
of a reference checkpoint works), `.cuda()/.to()/.eval()/.device`, the `forward` keyword list and the
output object.  Not kept: training (`labels` must be None — losses and the Hungarian matcher are out
of scope) and any network access (`from_pretrained` never downloads).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from .. import _lib
from ..config import DeformableDetrConfig
from ..engine import Engine
from .deformable_detr import _WeightTree
from .outputs import DetrSceneGraphGenerationOutput

__all__ = ["DetrForSceneGraphGeneration", "DetrSceneGraphGenerationOutput"]


class DetrForSceneGraphGeneration(_WeightTree):
    config_class = DeformableDetrConfig
    base_model_prefix = "model"
    main_input_name = "pixel_values"

    def __init__(self, config, **kwargs):
        super().__init__(config)
        self._engine: Optional[Engine] = None
        # replay the forward as one CUDA graph per input shape (outputs then alias static buffers that the
        # next call overwrites); off by default to keep the reference's fresh-tensor semantics
        self.use_cuda_graph = False
        fg_matrix = kwargs.get("fg_matrix", None)
        if fg_matrix is not None:  # training-time statistics (egtr.py:169-184)
            eps = config.freq_bias_eps
            fg = torch.as_tensor(fg_matrix, dtype=torch.float32)
            rel_dist = fg.sum((0, 1)) / (fg.sum() + eps)
            triplet = fg + eps / (fg.sum(2, keepdim=True) + eps)
            triplet = F.log_softmax(triplet, dim=-1) if config.use_log_softmax else triplet.log()
            self.rel_dist.data = rel_dist
            self.triplet_dist.data = triplet

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path=None, *args, config=None, **kwargs):
        """The reference calls this with a hub id and then overwrites every tensor from the Lightning
        checkpoint (`evaluate_egtr.py:229-240`).  There is no network here: build from `config` (or from
        `<path>/config.json`) and expect a `load_state_dict` to follow."""
        kwargs.pop("ignore_mismatched_sizes", None)
        if config is None:
            config = DeformableDetrConfig.from_pretrained(pretrained_model_name_or_path)
        return cls(config, **{k: v for k, v in kwargs.items() if k == "fg_matrix"})

    @property
    def device(self):
        return self.triplet_dist.device

    def _invalidate(self):
        self._engine = None

    def load_state_dict(self, state_dict, strict: bool = True):
        state_dict = dict(state_dict)
        out = super().load_state_dict(state_dict, strict=strict)
        self._invalidate()
        return out

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def engine(self) -> Engine:
        if self._engine is None:
            if self.device.type != "cuda":
                raise _lib.EgtrError("model is on %s: call .cuda() first — the B200 path has no CPU fallback" % self.device)
            self._engine = Engine(self.config, self.state_dict(), self.device)
        return self._engine

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, pixel_values, pixel_mask=None, decoder_attention_mask=None, encoder_outputs=None,
                inputs_embeds=None, decoder_inputs_embeds=None, labels=None, output_attentions=None,
                output_hidden_states=None, output_attention_states=None, return_dict=None):
        if labels is not None:
            raise NotImplementedError("labels/losses are training; the B200 path is inference-only (SURVEY.md §2.1)")
        if encoder_outputs is not None or inputs_embeds is not None or decoder_inputs_embeds is not None or decoder_attention_mask is not None:
            raise NotImplementedError("encoder_outputs / *_embeds / decoder_attention_mask are unused by evaluate_egtr.py and not built")
        if output_attentions:
            raise NotImplementedError("output_attentions=True (attention maps) is not built; evaluate_egtr.py passes False")
        if self.use_cuda_graph:
            B, _, H, W = pixel_values.shape
            with torch.cuda.device(self.device):
                o = self.engine().graph_runner(B, H, W)(pixel_values, pixel_mask)
        else:
            o = self.engine().forward(pixel_values, pixel_mask)
        return_dict = return_dict if return_dict is not None else getattr(self.config, "use_return_dict", True)
        inter = o["intermediate_hidden_states"]
        dec_states = None
        if output_hidden_states:
            # embeddings + one per layer + the last again, as DeformableDetrDecoder stacks them (deformable_detr.py:1869-1944)
            B = inter.shape[0]
            tgt = self.engine().query_tgt.unsqueeze(0).expand(B, -1, -1)
            dec_states = (tgt,) + tuple(inter[:, i] for i in range(inter.shape[1]))
        out = DetrSceneGraphGenerationOutput(
            logits=o["logits"], pred_boxes=o["pred_boxes"], pred_rel=o["pred_rel"],
            pred_connectivity=o["pred_connectivity"], last_hidden_state=o["last_hidden_state"],
            decoder_hidden_states=dec_states, encoder_last_hidden_state=o["encoder_last_hidden_state"],
        )
        if not return_dict:
            return out.to_tuple()
        return out
