"""`DeformableDetrConfig` — the hyper-parameter carrier of the drop-in boundary.

Mirrors the constructor arguments, defaults, `attribute_map` aliases and
on-disk `config.json` round trip of the reference's config
(`/root/reference/model/deformable_detr.py:72-267`), plus the ad-hoc EGTR
attributes the training script bolts on and inference reads
(`/root/reference/train_egtr.py:230-253`, `/root/reference/model/egtr.py:123-223,
405, 509-512`).  It does not inherit from HuggingFace `PretrainedConfig`: the
hot path needs a plain attribute bag that loads `config.json`, nothing else.
"""
from __future__ import annotations

import copy
import json
import os

_DEFAULTS = dict(
    num_queries=300,
    max_position_embeddings=1024,
    encoder_layers=6,
    encoder_ffn_dim=1024,
    encoder_attention_heads=8,
    decoder_layers=6,
    decoder_ffn_dim=1024,
    decoder_attention_heads=8,
    encoder_layerdrop=0.0,
    decoder_layerdrop=0.0,
    is_encoder_decoder=True,
    activation_function="relu",
    d_model=256,
    dropout=0.1,
    attention_dropout=0.0,
    activation_dropout=0.0,
    init_std=0.02,
    init_xavier_std=1.0,
    return_intermediate=True,
    auxiliary_loss=False,
    position_embedding_type="sine",
    backbone="resnet50",
    dilation=False,
    num_feature_levels=4,
    encoder_n_points=4,
    decoder_n_points=4,
    two_stage=False,
    two_stage_num_proposals=300,
    with_box_refine=False,
    class_cost=1,
    bbox_cost=5,
    giou_cost=2,
    mask_loss_coefficient=1,
    dice_loss_coefficient=1,
    bbox_loss_coefficient=5,
    giou_loss_coefficient=2,
    eos_coefficient=0.1,
    focal_alpha=0.25,
)

# EGTR inference attributes (train_egtr.py:230-253); `num_labels` is HF's own knob.
_EGTR_DEFAULTS = dict(
    num_labels=150,
    num_rel_labels=50,
    use_freq_bias=True,
    use_log_softmax=False,
    freq_bias_eps=1e-12,
    logit_adjustment=False,
    logit_adj_tau=0.3,
    output_attention_states=True,
    output_attentions=False,
    output_hidden_states=False,
    use_return_dict=True,
)


class DeformableDetrConfig:
    model_type = "deformable_detr"
    attribute_map = {"hidden_size": "d_model", "num_attention_heads": "encoder_attention_heads"}

    def __init__(self, **kwargs):
        for k, v in {**_DEFAULTS, **_EGTR_DEFAULTS}.items():
            object.__setattr__(self, k, copy.deepcopy(kwargs.pop(k, v)))
        if self.two_stage is True and self.with_box_refine is False:
            raise ValueError("If two_stage is True, with_box_refine must be True.")  # deformable_detr.py:246-247
        kwargs.pop("return_dict", None)
        for k, v in kwargs.items():  # unknown keys are kept, like HF's PretrainedConfig
            if k in self.attribute_map:
                k = self.attribute_map[k]
            object.__setattr__(self, k, v)

    # --- attribute_map aliases (deformable_detr.py:171-174, 261-267) ------------------
    def __getattr__(self, name):
        amap = type(self).attribute_map
        if name in amap:
            return getattr(self, amap[name])
        raise AttributeError(f"{type(self).__name__} has no attribute {name!r}")

    def __setattr__(self, name, value):
        object.__setattr__(self, type(self).attribute_map.get(name, name), value)

    # --- on-disk format: <dir>/config.json (evaluate_egtr.py:225, train_egtr.py:350-353) --
    def to_dict(self):
        out = {k: v for k, v in self.__dict__.items() if not k.startswith("_")}
        out["model_type"] = self.model_type
        return out

    def to_json_string(self):
        return json.dumps(self.to_dict(), indent=2, sort_keys=True) + "\n"

    def save_pretrained(self, save_directory):
        os.makedirs(save_directory, exist_ok=True)
        with open(os.path.join(save_directory, "config.json"), "w") as f:
            f.write(self.to_json_string())

    @classmethod
    def from_dict(cls, d):
        d = dict(d)
        d.pop("model_type", None)
        d.pop("transformers_version", None)
        d.pop("architectures", None)
        return cls(**d)

    @classmethod
    def from_pretrained(cls, path, **kwargs):
        fname = os.path.join(path, "config.json") if os.path.isdir(path) else path
        if not os.path.isfile(fname):
            raise OSError(
                f"{path!r} is not a directory holding config.json (no network: hub ids cannot be resolved)"
            )
        with open(fname) as f:
            d = json.load(f)
        d.update(kwargs)
        return cls.from_dict(d)

    def __repr__(self):
        return f"{type(self).__name__} {self.to_json_string()}"


# Named workloads of BASELINE.json `configs` (SURVEY.md §8 shape table).
WORKLOADS = {
    "A": dict(image=(480, 640), num_queries=100, num_labels=150, num_rel_labels=50),
    "B": dict(image=(800, 1333), num_queries=200, num_labels=150, num_rel_labels=50),
    "D": dict(image=(800, 1333), num_queries=200, num_labels=601, num_rel_labels=30),
    "E": dict(image=(1024, 1024), num_queries=300, num_labels=150, num_rel_labels=200),
    # small cases the CPU oracle finishes in well under a second
    "tiny": dict(image=(96, 128), num_queries=24, num_labels=20, num_rel_labels=12),
    "small": dict(image=(160, 224), num_queries=40, num_labels=30, num_rel_labels=16),
}


def workload_config(name: str, **overrides) -> "DeformableDetrConfig":
    w = dict(WORKLOADS[name])
    w.pop("image")
    w.update(overrides)
    return DeformableDetrConfig(**w)
