"""Import the UNMODIFIED reference (`/root/reference/model`) in this container.

Dev-only (this container): /root/reference does not exist on the GPU box.  The
reference pins transformers==4.18 / timm==0.5.4 which are absent here, so four
in-memory stubs are installed before import.  Nothing in the
reference tree is modified or copied.
"""
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = "/root/reference"


class _FeatureInfo:
    def channels(self):
        return [512, 1024, 2048]

    def reduction(self):
        return [8, 16, 32]


class _FakeTimmResNet50(nn.Module):
    """torchvision ResNet-50 (v1.5, stride on the 3x3) with timm's flat child names;
    emits [C3, C4, C5] like timm features_only(out_indices=(2,3,4))."""

    def __init__(self):
        super().__init__()
        import torchvision

        net = torchvision.models.resnet50(weights=None)
        self.conv1, self.bn1, self.act1, self.maxpool = net.conv1, net.bn1, net.relu, net.maxpool
        self.layer1, self.layer2, self.layer3, self.layer4 = net.layer1, net.layer2, net.layer3, net.layer4
        self.feature_info = _FeatureInfo()

    def forward(self, x):
        x = self.maxpool(self.act1(self.bn1(self.conv1(x))))
        c2 = self.layer1(x)
        c3 = self.layer2(c2)
        c4 = self.layer3(c3)
        c5 = self.layer4(c4)
        return [c3, c4, c5]


def _fake_create_model(name, pretrained=False, features_only=True, out_indices=(2, 3, 4), **kw):
    assert name == "resnet50" and tuple(out_indices) == (2, 3, 4), (name, out_indices)
    return _FakeTimmResNet50()


def import_reference():
    import transformers
    import transformers.modeling_utils as mu
    from transformers.image_transforms import center_to_corners_format

    if not hasattr(transformers, "DetrFeatureExtractor"):
        transformers.DetrFeatureExtractor = type("DetrFeatureExtractor", (), {})
    if not hasattr(mu, "PretrainedConfig"):
        mu.PretrainedConfig = transformers.PretrainedConfig
    modname = "transformers.models.detr.feature_extraction_detr"
    if modname not in sys.modules:
        fake = types.ModuleType(modname)
        fake.center_to_corners_format = center_to_corners_format
        sys.modules[modname] = fake
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import model.deformable_detr as dd  # noqa: E402  (the reference's module)
    import model.egtr as eg  # noqa: E402

    dd.create_model = _fake_create_model
    dd.requires_backends = lambda *a, **k: None
    return dd, eg


def build_reference_model(cfg_kwargs):
    dd, eg = import_reference()
    cfg_kwargs = {k: v for k, v in cfg_kwargs.items()
                  if k not in ("use_return_dict", "model_type", "output_attentions", "output_hidden_states")}
    config = dd.DeformableDetrConfig(**cfg_kwargs)
    model = eg.DetrForSceneGraphGeneration(config)
    model.eval()
    return dd, eg, config, model


if __name__ == "__main__":
    kw = dict(num_queries=100, num_labels=150, num_rel_labels=50, use_freq_bias=True,
              logit_adjustment=False, logit_adj_tau=0.3, auxiliary_loss=False,
              output_attention_states=True)
    dd, eg, config, model = build_reference_model(kw)
    n = 0
    for k, v in model.state_dict().items():
        print(k, tuple(v.shape), v.dtype)
        n += v.numel()
    print("params", n)
