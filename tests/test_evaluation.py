"""Scene-graph recall evaluation against golden vectors produced by the unmodified reference evaluator
(tests/golden/make_golden_sgeval.py)."""
import os

import numpy as np
import pytest

from egtr_b200.evaluation import (BasicSceneGraphEvaluator, bbox_overlaps, calculate_mR_from_evaluator_list, match_matrix,
                                  recall_at)

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sgeval.npz"))
N_CASES = len(G["cases_mode"])


def test_bbox_overlaps_matches_reference_loops():
    b = G["iou_boxes"]
    got = bbox_overlaps(b[:25], b[15:])
    assert got.shape == G["iou_ref"].shape and np.array_equal(got, G["iou_ref"])
    assert bbox_overlaps(np.array([[0, 0, 9, 9.0]]), np.array([[10, 0, 19, 9.0]]))[0, 0] == 0.0      # disjoint by one pixel
    assert bbox_overlaps(np.array([[0, 0, 9, 9.0]]), np.array([[9, 0, 18, 9.0]]))[0, 0] == 10.0 / 190  # one shared pixel column


@pytest.mark.parametrize("ci", range(N_CASES))
def test_evaluator_matches_reference(ci):
    mode, multiple, iou = str(G["cases_mode"][ci]), bool(G["cases_multiple"][ci]), float(G["cases_iou"][ci])
    ev = BasicSceneGraphEvaluator(mode, multiple_preds=multiple)
    for si in range(4):
        tag = f"c{ci}s{si}_"
        gt = {k: G[tag + k] for k in ("gt_relations", "gt_boxes", "gt_classes")}
        pred = {k: G[tag + k] for k in ("pred_rel_inds", "rel_scores", "pred_boxes", "pred_classes", "obj_scores")}
        p2g, five, scores = ev.evaluate_scene_graph_entry(gt, pred, iou_thresh=iou)
        assert [len(x) for x in p2g] == G[tag + "p2g_len"].tolist()
        assert [g for x in p2g for g in x] == G[tag + "p2g_flat"].tolist()
        assert np.array_equal(five, G[tag + "five"])
        # the rank formulation gives the same recalls as the reference's running union
        pred_rels = pred["pred_rel_inds"] if multiple else np.column_stack((pred["pred_rel_inds"], pred["rel_scores"].argmax(1)))
        boxes = gt["gt_boxes"] if mode in ("sgcls", "predcls") else pred["pred_boxes"]
        classes = gt["gt_classes"] if mode == "predcls" else pred["pred_classes"]
        m = match_matrix(gt["gt_relations"], gt["gt_boxes"], gt["gt_classes"], pred_rels, boxes, classes, iou, phrdet=mode == "phrdet")
        r = recall_at(m, (20, 50, 100))
        for k in (20, 50, 100):
            assert r[k] == ev.result_dict[mode + "_recall"][k][-1]
    for k in (20, 50, 100):
        assert ev.result_dict[mode + "_recall"][k] == G[f"c{ci}_recall{k}"].tolist()


def test_mean_recall_over_predicates(capsys):
    evs = []
    for pid, name, vals in ((0, "on", [1.0, 0.5]), (1, "has", [0.0]), (2, "never", [])):
        e = BasicSceneGraphEvaluator.all_modes(multiple_preds=False)
        for v in vals:
            for k in (20, 50, 100):
                e["sgdet"].result_dict["sgdet_recall"][k].append(v)
        evs.append((pid, name, e))
    with np.errstate(all="ignore"), pytest.warns(RuntimeWarning):
        mr = calculate_mR_from_evaluator_list(evs, "sgdet")
    capsys.readouterr()
    assert mr == {"mR@20": 0.25, "mR@50": 0.25, "mR@100": 0.25}  # the never-seen predicate (NaN) adds 0 but counts in the mean


def test_preddet_matches_reference():
    """`vrd_modes()` builds a preddet evaluator; the mode is pinned against the unmodified reference
    (tests/golden/make_golden_preddet.py), including a scene without predictions."""
    P = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sgeval_preddet.npz"))
    evs = BasicSceneGraphEvaluator.vrd_modes()
    assert set(evs) == {"preddet", "phrdet"} and evs["preddet"].multiple_preds
    ev = evs["preddet"]
    for si in range(int(P["n_scenes"])):
        gt = {k: P[f"s{si}_{k}"] for k in ("gt_relations", "gt_boxes", "gt_classes")}
        pred = {k: P[f"s{si}_{k}"] for k in ("pred_rel_inds", "rel_scores", "pred_boxes", "pred_classes", "obj_scores")}
        assert ev.evaluate_scene_graph_entry(gt, pred) == (None, None, None)
    for k in (20, 50, 100):
        assert ev.result_dict["preddet_recall"][k] == P[f"recall{k}"].tolist()
