"""On-disk formats either side of the hot path: the reference's training artifacts.

An EGTR run directory (`--artifact_path` of `/root/reference/evaluate_egtr.py:225-242`) holds
  * `config.json`              — HuggingFace-style dump of `DeformableDetrConfig` (+ the EGTR attributes of
                                 `train_egtr.py:230-253`), read by `DeformableDetrConfig.from_pretrained`;
  * `checkpoints/epoch=*.ckpt` — PyTorch-Lightning checkpoints: a pickled dict whose `"state_dict"` maps
                                 `"model." + <DetrForSceneGraphGeneration key>` to tensors (the LightningModule wraps the
                                 network as `self.model`, `train_egtr.py:350-353`); evaluation takes the highest epoch.
`load_artifact` performs exactly the reference's loading steps and returns the B200 model; `save_artifact` writes the
same layout (tests, weight conversion).
"""
from __future__ import annotations

import glob
import os
from typing import Dict, Optional

import torch

from .config import DeformableDetrConfig

PREFIX = "model."


def latest_checkpoint(artifact_path: str) -> str:
    """`sorted(glob(.../checkpoints/epoch=*.ckpt), key=epoch)[-1]` (evaluate_egtr.py:231-234)."""
    paths = glob.glob(os.path.join(artifact_path, "checkpoints", "epoch=*.ckpt"))
    if not paths:
        raise FileNotFoundError(f"no checkpoints/epoch=*.ckpt under {artifact_path!r}")
    return sorted(paths, key=lambda x: int(x.split("epoch=")[1].split("-")[0].split(".")[0]))[-1]


def strip_lightning_prefix(state_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """`state_dict[k[6:]] = state_dict.pop(k)` for every key (evaluate_egtr.py:235-237) — but keys that do not carry the
    `model.` wrapper prefix are an error instead of being silently truncated."""
    out = {}
    for k, v in state_dict.items():
        if not k.startswith(PREFIX):
            raise KeyError(f"checkpoint key {k!r} lacks the Lightning wrapper prefix {PREFIX!r}")
        out[k[len(PREFIX):]] = v
    return out


def load_artifact(artifact_path: str, logit_adjustment: Optional[bool] = None, logit_adj_tau: Optional[float] = None,
                  device: Optional[str] = "cuda", checkpoint: Optional[str] = None):
    """Config + latest Lightning checkpoint -> ready `DetrForSceneGraphGeneration` (eval mode, on `device`)."""
    from .model.egtr import DetrForSceneGraphGeneration

    config = DeformableDetrConfig.from_pretrained(artifact_path)
    if logit_adjustment is not None:
        config.logit_adjustment = logit_adjustment
    if logit_adj_tau is not None:
        config.logit_adj_tau = logit_adj_tau
    model = DetrForSceneGraphGeneration.from_pretrained(None, config=config, ignore_mismatched_sizes=True)
    ckpt = checkpoint or latest_checkpoint(artifact_path)
    blob = torch.load(ckpt, map_location="cpu", weights_only=False)
    if "state_dict" not in blob:
        raise KeyError(f"{ckpt!r} is not a Lightning checkpoint (no 'state_dict')")
    model.load_state_dict(strip_lightning_prefix(blob["state_dict"]))
    if device is not None:
        model.to(device)
    model.eval()
    return model


def save_artifact(artifact_path: str, config: DeformableDetrConfig, state_dict: Dict[str, torch.Tensor], epoch: int = 0,
                  step: int = 0) -> str:
    """Write `config.json` and `checkpoints/epoch=<e>-step=<s>.ckpt` in the reference's layout; returns the ckpt path."""
    config.save_pretrained(artifact_path)
    os.makedirs(os.path.join(artifact_path, "checkpoints"), exist_ok=True)
    path = os.path.join(artifact_path, "checkpoints", f"epoch={epoch}-step={step}.ckpt")
    torch.save({"epoch": epoch, "global_step": step,
                "state_dict": {PREFIX + k: v.detach().cpu() for k, v in state_dict.items()}}, path)
    return path
