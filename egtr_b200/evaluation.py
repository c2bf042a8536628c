"""Scene-graph recall evaluation: R@K / mR@K over the triplets the hot path emits.

Same interface as the reference's `BasicSceneGraphEvaluator` / `calculate_mR_from_evaluator_list`
(`/root/reference/lib/evaluation/sg_eval.py:18-139, 330-372`; call sites `train_egtr.py:100-160`,
`evaluate_egtr.py:56-118`), so `evaluate_batch` keeps working unchanged, but the matching is one vectorised pass: triplets
are hashed to integers, the [gt, pred] match matrix is `same triplet & IoU(subject) >= t & IoU(object) >= t`, and recall@K
follows from the rank of the first matching prediction of every ground-truth relation — no Python loop over ground-truth
triplets and no compiled Cython helper (`lib/fpn/box_intersections_cpu/bbox.pyx` is restated in numpy: pixel-inclusive
"+1" box areas, `bbox.pyx:14-61`).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

MODES = ["sgdet"]


def rows_equal(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """bool [len(a), len(b)]: row i of `a` equals row j of `b` (role of `lib/pytorch_misc.py::intersect_2d`)."""
    a, b = np.asarray(a), np.asarray(b)
    if a.shape[1] != b.shape[1]:
        raise ValueError(f"row width mismatch: {a.shape[1]} vs {b.shape[1]}")
    return (a[:, None, :] == b[None, :, :]).all(-1)


def descending_indices(scores: np.ndarray) -> np.ndarray:
    """[scores.size, scores.ndim] multi-indices of `scores` from the largest value down (role of `lib/pytorch_misc.py::argsort_desc`)."""
    scores = np.asarray(scores)
    return np.stack(np.unravel_index(np.argsort(-scores, axis=None), scores.shape), 1)


def bbox_overlaps(boxes: np.ndarray, query_boxes: np.ndarray) -> np.ndarray:
    """IoU [N, K] of xyxy boxes with pixel-inclusive extents (w = x2 - x1 + 1), zero when the intersection is empty."""
    b = np.asarray(boxes, dtype=np.float64)[:, None, :]
    q = np.asarray(query_boxes, dtype=np.float64)[None, :, :]
    iw = np.minimum(b[..., 2], q[..., 2]) - np.maximum(b[..., 0], q[..., 0]) + 1.0
    ih = np.minimum(b[..., 3], q[..., 3]) - np.maximum(b[..., 1], q[..., 1]) + 1.0
    area_b = (b[..., 2] - b[..., 0] + 1.0) * (b[..., 3] - b[..., 1] + 1.0)
    area_q = (q[..., 2] - q[..., 0] + 1.0) * (q[..., 3] - q[..., 1] + 1.0)
    inter = iw * ih
    ok = (iw > 0) & (ih > 0)
    ua = area_b + area_q - inter
    return np.where(ok, inter / np.where(ok, ua, 1.0), 0.0)


def _union_boxes(pairs: np.ndarray) -> np.ndarray:
    """[n, 8] (subject box | object box) -> [n, 4] enclosing box (phrase detection)."""
    return np.concatenate([np.minimum(pairs[:, :2], pairs[:, 4:6]), np.maximum(pairs[:, 2:4], pairs[:, 6:8])], 1)


def match_matrix(gt_rels, gt_boxes, gt_classes, pred_rels, pred_boxes, pred_classes, iou_thresh=0.5, phrdet=False) -> np.ndarray:
    """bool [n_gt, n_pred]: prediction j (s, o, p) reproduces ground-truth relation i (same classes and predicate, boxes
    overlapping at `iou_thresh`) — `_compute_pred_matches` of sg_eval.py:269-327 as one array expression."""
    gt_rels, pred_rels = np.asarray(gt_rels), np.asarray(pred_rels)
    gt_classes, pred_classes = np.asarray(gt_classes), np.asarray(pred_classes)
    gt_boxes, pred_boxes = np.asarray(gt_boxes, np.float64), np.asarray(pred_boxes, np.float64)
    gt_trip = np.stack([gt_classes[gt_rels[:, 0]], gt_rels[:, 2], gt_classes[gt_rels[:, 1]]], 1).astype(np.int64)
    pr_trip = np.stack([pred_classes[pred_rels[:, 0]], pred_rels[:, 2], pred_classes[pred_rels[:, 1]]], 1).astype(np.int64)
    base = int(max(gt_trip.max(initial=0), pr_trip.max(initial=0))) + 1
    key = lambda t: (t[:, 0] * base + t[:, 1]) * base + t[:, 2]  # noqa: E731
    same = key(gt_trip)[:, None] == key(pr_trip)[None, :]
    gt_pair = np.concatenate([gt_boxes[gt_rels[:, 0]], gt_boxes[gt_rels[:, 1]]], 1)
    pr_pair = np.concatenate([pred_boxes[pred_rels[:, 0]], pred_boxes[pred_rels[:, 1]]], 1)
    if phrdet:
        return same & (bbox_overlaps(_union_boxes(gt_pair), _union_boxes(pr_pair)) >= iou_thresh)
    return same & (bbox_overlaps(gt_pair[:, :4], pr_pair[:, :4]) >= iou_thresh) & (bbox_overlaps(gt_pair[:, 4:], pr_pair[:, 4:]) >= iou_thresh)


def evaluate_recall(gt_rels, gt_boxes, gt_classes, pred_rels, pred_boxes, pred_classes, rel_scores=None, cls_scores=None,
                    iou_thresh=0.5, phrdet=False):
    """Reference-compatible return values (sg_eval.py:141-221): per-prediction lists of matched ground-truth indices,
    (s, o, class_s, class_o, predicate) rows and [score_s, score_o, score_rel] rows."""
    pred_rels = np.asarray(pred_rels)
    if pred_rels.size == 0:
        return [[]], np.zeros((0, 5)), np.zeros(0)
    assert np.asarray(gt_rels).shape[0] != 0
    pred_classes = np.asarray(pred_classes)
    assert pred_rels[:, :2].max() < pred_classes.shape[0] and np.all(pred_rels[:, 2] >= 0)
    m = match_matrix(gt_rels, gt_boxes, gt_classes, pred_rels, pred_boxes, pred_classes, iou_thresh, phrdet)
    pred_to_gt = [np.nonzero(m[:, j])[0].tolist() for j in range(m.shape[1])]
    pred_5ples = np.column_stack((pred_rels[:, :2], pred_classes[pred_rels[:, 0]], pred_classes[pred_rels[:, 1]], pred_rels[:, 2]))
    scores = None
    if rel_scores is not None and cls_scores is not None:
        cls_scores = np.asarray(cls_scores)
        scores = np.column_stack((cls_scores[pred_rels[:, 0]], cls_scores[pred_rels[:, 1]], np.asarray(rel_scores)))
    return pred_to_gt, pred_5ples, scores


def recall_at(match: np.ndarray, ks: Sequence[int]) -> Dict[int, float]:
    """Recall@K from the [n_gt, n_pred] match matrix (predictions in descending score order): a ground-truth relation
    counts for K if its first matching prediction has rank < K."""
    n_gt = match.shape[0]
    first = np.where(match.any(1), match.argmax(1), np.iinfo(np.int64).max)
    return {k: float((first < k).sum()) / float(n_gt) for k in ks}


class BasicSceneGraphEvaluator:
    def __init__(self, mode, multiple_preds=False):
        self.mode = mode
        self.multiple_preds = multiple_preds
        self.result_dict = {self.mode + "_recall": {20: [], 50: [], 100: []}}

    @classmethod
    def all_modes(cls, **kwargs):
        return {m: cls(mode=m, **kwargs) for m in MODES}

    @classmethod
    def vrd_modes(cls, **kwargs):
        return {m: cls(mode=m, multiple_preds=True, **kwargs) for m in ("preddet", "phrdet")}

    def evaluate_scene_graph_entry(self, gt_entry, pred_scores, viz_dict=None, iou_thresh=0.5):
        return evaluate_from_dict(gt_entry, pred_scores, self.mode, self.result_dict, viz_dict=viz_dict, iou_thresh=iou_thresh,
                                  multiple_preds=self.multiple_preds)

    def save(self, fn):
        np.save(fn, self.result_dict)

    def print_stats(self):
        mode = "recall without constraint" if self.multiple_preds else "recall with constraint"
        print("======================" + self.mode + "  " + mode + "============================")
        out = {}
        for k, v in self.result_dict[self.mode + "_recall"].items():
            out["R@%i" % k] = np.mean(v)
            print("R@%i: %f" % (k, out["R@%i" % k]))
        return out


def evaluate_from_dict(gt_entry, pred_entry, mode, result_dict, multiple_preds=False, viz_dict=None, iou_thresh=0.5, **kwargs):
    """sg_eval.py:74-139: the box-predicting modes (sgdet / phrdet), the ground-truth-box modes (predcls / sgcls) and preddet."""
    gt_rels = np.asarray(gt_entry["gt_relations"])
    gt_boxes = np.asarray(gt_entry["gt_boxes"]).astype(float)
    gt_classes = np.asarray(gt_entry["gt_classes"])
    pred_rel_inds = np.asarray(pred_entry["pred_rel_inds"])
    rel_scores = np.asarray(pred_entry["rel_scores"])
    if mode == "predcls":
        pred_boxes, pred_classes, obj_scores = gt_boxes, gt_classes, np.ones(gt_classes.shape[0])
    elif mode == "sgcls":
        pred_boxes, pred_classes, obj_scores = gt_boxes, np.asarray(pred_entry["pred_classes"]), np.asarray(pred_entry["obj_scores"])
    elif mode.startswith("sgdet") or mode == "phrdet":
        pred_boxes = np.asarray(pred_entry["pred_boxes"]).astype(float)
        pred_classes, obj_scores = np.asarray(pred_entry["pred_classes"]), np.asarray(pred_entry["obj_scores"])
    elif mode == "preddet":
        return _preddet(gt_rels, pred_rel_inds, rel_scores, result_dict)
    else:
        raise ValueError("invalid mode")
    if multiple_preds:
        pred_rels, predicate_scores = pred_rel_inds, rel_scores
    else:
        pred_rels = np.column_stack((pred_rel_inds, rel_scores.argmax(1)))
        predicate_scores = rel_scores.max(1)
    pred_to_gt, pred_5ples, scores = evaluate_recall(gt_rels, gt_boxes, gt_classes, pred_rels, pred_boxes, pred_classes,
                                                     predicate_scores, obj_scores, iou_thresh=iou_thresh, phrdet=mode == "phrdet")
    n_gt = float(gt_rels.shape[0])
    for k in result_dict[mode + "_recall"]:
        matched = set()
        for lst in pred_to_gt[:k]:
            matched.update(lst)
        result_dict[mode + "_recall"][k].append(float(len(matched)) / n_gt)
    return pred_to_gt, pred_5ples, scores


def _preddet(gt_rels, pred_rel_inds, rel_scores, result_dict):
    """Predicate detection (sg_eval.py:107-131): only the predicted (subject, object) pairs that appear in the ground truth are
    kept — the first prediction per ground-truth pair — and their per-predicate scores [pairs, P] are ranked jointly;
    recall@k = fraction of ground-truth triplets among the k best (pair, predicate) entries.  `pred_rel_inds` [n, 2],
    `rel_scores` [n, P].  Returns (None, None, None) like the reference."""
    key = "preddet_recall"
    pairs_equal = (pred_rel_inds[:, None, :2] == gt_rels[None, :, :2]).all(-1)  # [pred, gt]
    if pairs_equal.size == 0:
        for k in result_dict[key]:
            result_dict[key][k].append(0.0)
        return None, None, None
    first = pairs_equal.argmax(0)  # first matching prediction of every ground-truth pair (0 when none: the reference's argmax too)
    kept_pairs, kept_scores = pred_rel_inds[first, :2], rel_scores[first]
    order = np.argsort(-kept_scores.ravel())  # the reference's argsort_desc (quicksort: tie order unspecified)
    row, pred = np.unravel_index(order, kept_scores.shape)
    ranked = np.column_stack((kept_pairs[row], pred))  # (subject, object, predicate), best first
    hits = (ranked[:, None, :] == gt_rels[None, :, :]).all(-1)  # [ranked, gt]
    for k in result_dict[key]:
        result_dict[key][k].append(float(hits[:k].any(0).sum()) / float(gt_rels.shape[0]))
    return None, None, None


def calculate_mR_from_evaluator_list(evaluator_list, mode, multiple_preds=False):
    """Mean recall over predicates (sg_eval.py:330-372): predicates whose R@100 is NaN (never in the ground truth) add zero
    but still count in the denominator, as in the reference."""
    all_rel_results = {}
    for (_pred_id, pred_name, evaluator_rel) in evaluator_list:
        print("\n")
        print("relationship: ", pred_name)
        all_rel_results[pred_name] = evaluator_rel[mode].print_stats()
    sums = {20: 0.0, 50: 0.0, 100: 0.0}
    for value in all_rel_results.values():
        if math.isnan(value["R@100"]):
            continue
        for k in sums:
            sums[k] += value["R@%i" % k]
    n = len(evaluator_list)
    mean_recall = {"mR@%i" % k: sums[k] / n for k in (20, 50, 100)}
    all_rel_results["mean_recall"] = mean_recall
    print("\n")
    print("======================" + mode + "  " + ("mean recall without constraint" if multiple_preds else "mean recall with constraint")
          + "============================")
    for k in (20, 50, 100):
        print("mR@%i: " % k, mean_recall["mR@%i" % k])
    return mean_recall
