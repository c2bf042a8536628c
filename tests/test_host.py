"""CPU: host-side logic — config round trip, synthetic weights, ABI surface, module key parity."""
import ctypes
import json
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config_roundtrip_and_aliases(tmp_path):
    from egtr_b200.config import DeformableDetrConfig

    c = DeformableDetrConfig(num_queries=200, num_labels=150, num_rel_labels=50, logit_adjustment=True, custom_flag=7)
    assert c.hidden_size == c.d_model == 256 and c.num_attention_heads == 8
    c.logit_adj_tau = 0.25
    c.save_pretrained(tmp_path)
    d = DeformableDetrConfig.from_pretrained(str(tmp_path))
    assert d.num_queries == 200 and d.logit_adjustment is True and d.logit_adj_tau == 0.25 and d.custom_flag == 7
    with pytest.raises(ValueError):
        DeformableDetrConfig(two_stage=True, with_box_refine=False)
    with pytest.raises(OSError):
        DeformableDetrConfig.from_pretrained("SenseTime/deformable-detr")  # hub ids need a network


def test_synth_is_deterministic_and_aliased():
    from egtr_b200.config import workload_config
    from egtr_b200.synth import synth_state_dict

    cfg = workload_config("tiny")
    a, b = synth_state_dict(cfg, 3), synth_state_dict(cfg, 3)
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    assert torch.equal(a["class_embed.0.weight"], a["class_embed.5.weight"])
    assert a["triplet_dist"].shape == (21, 21, 12)


def test_level_shapes_match_survey_table():
    from egtr_b200.engine import level_shapes

    assert level_shapes(480, 640) == [(60, 80), (30, 40), (15, 20), (8, 10)]
    assert level_shapes(800, 1333) == [(100, 167), (50, 84), (25, 42), (13, 21)]
    assert level_shapes(1024, 1024) == [(128, 128), (64, 64), (32, 32), (16, 16)]


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads without a GPU and exports each function `include/egtr_b200.h` declares."""
    from egtr_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "egtr_b200.h")).read()
    declared = set(re.findall(r"\b(egtr_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.call("egtr_abi_version") == 1


def test_model_state_dict_keys_match_reference_layout():
    from egtr_b200.config import workload_config
    from egtr_b200.model.egtr import DetrForSceneGraphGeneration
    from egtr_b200.synth import synth_state_dict

    cfg = workload_config("tiny")
    m = DetrForSceneGraphGeneration(cfg)
    sd = synth_state_dict(cfg, 5)
    assert set(m.state_dict().keys()) == set(sd.keys())
    m.load_state_dict(sd, strict=True)
    assert m.class_embed[0] is m.class_embed[5]
    ck = {"model." + k: v for k, v in sd.items()}  # Lightning checkpoint prefix (evaluate_egtr.py:232-240)
    m.load_state_dict({k[6:]: v for k, v in ck.items()})
    with pytest.raises(Exception):
        m(torch.zeros(1, 3, 32, 32))  # CPU model: no fallback


def test_output_object_supports_attr_key_and_in():
    from egtr_b200.model.outputs import DetrSceneGraphGenerationOutput

    o = DetrSceneGraphGenerationOutput(logits=torch.zeros(1), pred_connectivity=torch.ones(1))
    assert o.logits is o["logits"] and "pred_connectivity" in o and "pred_rel" not in o and o.loss is None
    assert len(o.to_tuple()) == 2


def test_argmax_tie_policy_of_the_parity_comparison():
    """tests/util.py::class_flips excuses an arg-max class change only inside the reference's own error bar."""
    from tests.util import class_flips, forward_errors, worst
    g = torch.Generator().manual_seed(0)
    ref = dict(logits=torch.randn(1, 6, 5, generator=g), pred_boxes=torch.rand(1, 6, 4, generator=g),
               pred_rel=torch.rand(1, 6, 6, 3, generator=g), pred_connectivity=torch.rand(1, 6, 6, 1, generator=g))
    ref["logits"][0, 2, 1] = ref["logits"][0, 2].max() + 1.0
    ref["logits"][0, 2, 3] = ref["logits"][0, 2, 1] - 1e-6  # query 2: classes 1 and 3 tie to 1e-6
    out = {k: v.clone() for k, v in ref.items()}
    out["logits"][0, 2, 3] += 3e-6                            # ... and resolve the other way
    out["pred_rel"][0, 2, :, :] = 0.9                         # its pairs take another frequency-bias row
    out["pred_rel"][0, :, 2, :] = 0.1
    flipped = class_flips(out["logits"], ref["logits"])
    assert flipped.tolist() == [[False, False, True, False, False, False]]
    errs = forward_errors(out, ref)
    assert errs["_class_flips"] == 1 and worst(errs) < 1e-5
    out["pred_rel"][0, 0, 1, 0] += 0.5                        # a pair of unflipped queries is still compared
    assert worst(forward_errors(out, ref)) > 0.1
    bad = {k: v.clone() for k, v in ref.items()}
    bad["logits"][0, 4] = bad["logits"][0, 4].flip(0) * 3     # a class change with a real margin is a failure
    if bad["logits"][0, 4].argmax() != ref["logits"][0, 4].argmax():
        with pytest.raises(AssertionError):
            class_flips(bad["logits"], ref["logits"])
