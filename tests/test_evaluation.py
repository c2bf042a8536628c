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


# ---------------------------------------------------------------------------------------------- Open-Images / COCO evaluators
def _oi_scenes():
    import json
    import os
    import sys
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gdir)
    import make_golden_oieval as mk  # the scene generator only (the reference is not imported)
    return mk, json.load(open(os.path.join(gdir, "oieval.json")))


def test_oi_relation_metrics_match_reference_golden():
    """w_rel_mAP / w_phr_mAP / microR@50 / score of the unmodified reference (`oi_eval.eval_rel_results`, `ap_eval_rel.ap_eval`) on
    seeded scenes fed through `OIEvaluator.__call__` in `evaluate_batch`'s entry format (train_egtr.py:154-173)."""
    from egtr_b200.oi_evaluation import OIEvaluator, eval_rel_results
    mk, golden = _oi_scenes()
    for seed, want in golden.items():
        ev = OIEvaluator([f"p{i}" for i in range(mk.N_PRED_CLS)], [f"c{i}" for i in range(mk.N_CLS)])
        for gt, pred in mk.scenes(int(seed), want["n_images"]):
            ev(gt, pred)
        got = eval_rel_results(ev.all_result, ev.predicate_cls_list)
        for k in ("w_rel_mAP", "w_phr_mAP", "microR@50", "score"):
            assert abs(got[k] - want[k]) < 1e-12, (seed, k, got[k], want[k])
        agg = ev.aggregate_metrics()
        assert {"w_rel_mAP", "w_phr_mAP", "microR@50", "score", "bbox/AP50"} <= set(agg)


def test_coco_box_ap_known_answers():
    """pycocotools is absent (parity unpinned): the bbox algorithm restated in `CocoBoxEval` against hand-computed cases."""
    from egtr_b200.oi_evaluation import CocoBoxEval
    gts = [dict(image_id=0, category_id=1, bbox=[10, 10, 100, 100]), dict(image_id=0, category_id=1, bbox=[200, 200, 100, 100])]
    # TP (0.9), FP (0.8), TP (0.7): precision 1 up to recall 0.5, 2/3 up to recall 1 -> 101-point AP = (51 + 50 * 2/3) / 101
    ev = CocoBoxEval(gts)
    ev.add_detections([dict(image_id=0, category_id=1, bbox=[10, 10, 100, 100], score=0.9), dict(image_id=0, category_id=1, bbox=[400, 400, 50, 50], score=0.8),
                       dict(image_id=0, category_id=1, bbox=[200, 200, 100, 100], score=0.7)])
    st = ev.summarize()
    assert abs(st[1] - (51 + 50 * 2 / 3) / 101) < 1e-9 and abs(st[0] - st[1]) < 1e-9 and abs(st[8] - 1.0) < 1e-9
    assert abs(st[6] - 0.5) < 1e-9  # AR@1: only the best-scored detection counts
    # one ground truth, one detection at IoU 0.62: a match at the thresholds 0.50, 0.55, 0.60 only
    ev = CocoBoxEval([dict(image_id=0, category_id=1, bbox=[0, 0, 100, 100])])
    ev.add_detections([dict(image_id=0, category_id=1, bbox=[0, 0, 100, 62], score=0.5)])
    st = ev.summarize()
    assert abs(st[0] - 0.3) < 1e-9 and abs(st[1] - 1.0) < 1e-9 and st[2] == 0.0
    # a crowd region absorbs detections without counting them; area ranges: a 20 x 20 box is "small" only
    ev = CocoBoxEval([dict(image_id=0, category_id=1, bbox=[0, 0, 20, 20]), dict(image_id=0, category_id=1, bbox=[100, 100, 200, 200], iscrowd=1)])
    ev.add_detections([dict(image_id=0, category_id=1, bbox=[0, 0, 20, 20], score=0.9), dict(image_id=0, category_id=1, bbox=[120, 120, 50, 50], score=0.8)])
    st = ev.summarize()
    assert abs(st[1] - 1.0) < 1e-9 and abs(st[3] - 1.0) < 1e-9 and st[4] == -1.0 and st[5] == -1.0


def test_coco_evaluator_interface():
    """`CocoEvaluator` as `evaluate_egtr.py:86-103` drives it: xyxy tensors per image id, labels shifted by +1 (coco_eval.py:44-45)."""
    import torch
    from egtr_b200.oi_evaluation import CocoEvaluator
    ds = dict(images=[dict(id=7), dict(id=9)], categories=[dict(id=1), dict(id=2)],
              annotations=[dict(id=0, image_id=7, category_id=1, bbox=[10, 20, 30, 40], area=1200, iscrowd=0),
                           dict(id=1, image_id=9, category_id=2, bbox=[50, 50, 80, 60], area=4800, iscrowd=0)])
    ev = CocoEvaluator(ds, ["bbox"])
    ev.update({7: dict(boxes=torch.tensor([[10., 20., 40., 60.]]), scores=torch.tensor([0.9]), labels=torch.tensor([0]))})
    ev.update({9: dict(boxes=torch.tensor([[50., 50., 130., 110.], [0., 0., 10., 10.]]), scores=torch.tensor([0.8, 0.3]), labels=torch.tensor([1, 1]))})
    ev.synchronize_between_processes()
    ev.accumulate()
    ev.summarize()
    assert abs(ev.coco_eval["bbox"].stats[1] - 1.0) < 1e-9  # AP50, the number evaluate_egtr.py:103 reports


def test_oi_evaluator_edge_cases():
    """Images without ground-truth relations (the reference's evaluator raises KeyError there), without predictions, and a perfect
    prediction: recall and both mAPs are 1 for the latter, the former two only dilute them."""
    import numpy as np
    from egtr_b200.oi_evaluation import OIEvaluator
    ev = OIEvaluator(["on", "under"], ["a", "b", "c"])
    boxes = np.array([[0, 0, 50, 50], [60, 60, 120, 120]], np.float32)
    gt = dict(gt_boxes=boxes, gt_classes=np.array([0, 1]), gt_relations=np.array([[0, 1, 1]]))
    pairs = np.array([(s, o) for s in range(2) for o in range(2)])
    scores = np.zeros((4, 2), np.float32)
    scores[1, 1] = 0.9  # pair (0, 1), predicate 1
    ev(gt, dict(pred_boxes=boxes.copy(), pred_classes=np.array([0, 1]), obj_scores=np.array([0.9, 0.8], np.float32), sbj_obj_inds=pairs, pred_scores=scores))
    m = ev.aggregate_metrics()
    assert abs(m["microR@50"] - 1.0) < 1e-9 and abs(m["w_rel_mAP"] - 1.0) < 1e-9 and abs(m["w_phr_mAP"] - 1.0) < 1e-9 and abs(m["score"] - 1.0) < 1e-9
    # an image with no ground-truth relation and one with all-zero predicate scores
    ev(dict(gt_boxes=boxes, gt_classes=np.array([0, 1]), gt_relations=np.zeros((0, 3), np.int64)),
       dict(pred_boxes=boxes.copy(), pred_classes=np.array([0, 1]), obj_scores=np.array([0.9, 0.8], np.float32), sbj_obj_inds=pairs, pred_scores=scores))
    ev(gt, dict(pred_boxes=boxes.copy(), pred_classes=np.array([0, 1]), obj_scores=np.array([0.9, 0.8], np.float32), sbj_obj_inds=pairs,
                pred_scores=np.zeros((4, 2), np.float32)))
    m2 = ev.aggregate_metrics()
    assert abs(m2["microR@50"] - 0.5) < 1e-9 and 0.0 < m2["w_rel_mAP"] <= 1.0
