"""Minimal stand-in for HuggingFace `ModelOutput`: attribute AND key access, `in`, tuple conversion.

The reference returns `DetrSceneGraphGenerationOutput(ModelOutput)` (`/root/reference/model/egtr.py:53-115`);
its callers use `out["logits"]`, `out.logits` and `"pred_connectivity" in out`
(`/root/reference/train_egtr.py:56-69`, `/root/reference/model/deformable_detr.py:288`).
"""
from collections import OrderedDict


class ModelOutput(OrderedDict):
    _fields = ()

    def __init__(self, **kwargs):
        super().__init__()
        for f in self._fields:
            object.__setattr__(self, f, None)
        for k, v in kwargs.items():
            if k not in self._fields:
                raise TypeError(f"{type(self).__name__} has no field {k!r}")
            object.__setattr__(self, k, v)
            if v is not None:  # like HF: None fields are not keys
                super().__setitem__(k, v)

    def __getitem__(self, k):
        if isinstance(k, str):
            return super().__getitem__(k)
        return self.to_tuple()[k]

    def __setitem__(self, k, v):
        if k in self._fields:
            object.__setattr__(self, k, v)
        super().__setitem__(k, v)

    def __setattr__(self, k, v):
        if k in self._fields and v is not None:
            super().__setitem__(k, v)
        object.__setattr__(self, k, v)

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())


class DetrSceneGraphGenerationOutput(ModelOutput):
    _fields = ("loss", "loss_dict", "logits", "pred_boxes", "pred_rel", "pred_connectivity", "auxiliary_outputs",
               "last_hidden_state", "decoder_hidden_states", "decoder_attentions", "cross_attentions",
               "encoder_last_hidden_state", "encoder_hidden_states", "encoder_attentions")


class DeformableDetrModelOutput(ModelOutput):
    _fields = ("init_reference_points", "last_hidden_state", "intermediate_hidden_states",
               "intermediate_reference_points", "decoder_hidden_states", "decoder_attentions", "cross_attentions",
               "encoder_last_hidden_state", "encoder_hidden_states", "encoder_attentions", "enc_outputs_class",
               "enc_outputs_coord_logits", "decoder_attention_queries", "decoder_attention_keys")
