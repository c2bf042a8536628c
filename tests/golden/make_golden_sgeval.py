"""Golden vectors for egtr_b200/evaluation.py, produced by the UNMODIFIED reference evaluator
(/root/reference/lib/evaluation/sg_eval.py) on seeded random scenes.  The reference's compiled helper
`lib.fpn.box_intersections_cpu.bbox` (Cython, not built here; `np.float` no longer exists) is replaced by a literal
loop-for-loop Python restatement of bbox.pyx:14-61 installed as a stub module before the import.

    python tests/golden/make_golden_sgeval.py      # needs /root/reference; writes tests/golden/sgeval.npz
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"


def _bbox_overlaps_loops(boxes, query_boxes):  # bbox.pyx:14-61, loop for loop
    boxes = np.ascontiguousarray(boxes, dtype=float)
    query_boxes = np.ascontiguousarray(query_boxes, dtype=float)
    N, K = boxes.shape[0], query_boxes.shape[0]
    overlaps = np.zeros((N, K), dtype=float)
    for k in range(K):
        box_area = (query_boxes[k, 2] - query_boxes[k, 0] + 1) * (query_boxes[k, 3] - query_boxes[k, 1] + 1)
        for n in range(N):
            iw = min(boxes[n, 2], query_boxes[k, 2]) - max(boxes[n, 0], query_boxes[k, 0]) + 1
            if iw > 0:
                ih = min(boxes[n, 3], query_boxes[k, 3]) - max(boxes[n, 1], query_boxes[k, 1]) + 1
                if ih > 0:
                    ua = float((boxes[n, 2] - boxes[n, 0] + 1) * (boxes[n, 3] - boxes[n, 1] + 1) + box_area - iw * ih)
                    overlaps[n, k] = iw * ih / ua
    return overlaps


def scene(rng, n_gt_boxes, n_gt_rels, n_pred_boxes, n_pred_rels, n_cls, n_pred_cls, multiple):
    def boxes(n):
        xy = rng.uniform(0, 400, (n, 2))
        wh = rng.uniform(20, 200, (n, 2))
        return np.concatenate([xy, xy + wh], 1).round(1)
    gt_boxes = boxes(n_gt_boxes)
    gt_classes = rng.integers(0, n_cls, n_gt_boxes)
    so = np.array([(s, o) for s in range(n_gt_boxes) for o in range(n_gt_boxes) if s != o])
    n_gt_rels = min(n_gt_rels, len(so))
    gt_rels = np.column_stack([so[rng.choice(len(so), n_gt_rels, replace=False)], rng.integers(0, n_pred_cls, n_gt_rels)])
    # predictions: jittered copies of the ground truth (so that matches exist) plus random boxes
    idx = rng.integers(0, n_gt_boxes, n_pred_boxes)
    pred_boxes = np.where(rng.random((n_pred_boxes, 1)) < 0.6, gt_boxes[idx] + rng.normal(0, 8, (n_pred_boxes, 4)), boxes(n_pred_boxes))
    pred_classes = np.where(rng.random(n_pred_boxes) < 0.7, gt_classes[idx], rng.integers(0, n_cls, n_pred_boxes))
    obj_scores = rng.random(n_pred_boxes)
    pso = np.array([(s, o) for s in range(n_pred_boxes) for o in range(n_pred_boxes) if s != o])
    n_pred_rels = min(n_pred_rels, len(pso))
    sel = pso[rng.choice(len(pso), n_pred_rels, replace=False)]
    if multiple:
        pred_rel_inds = np.column_stack([sel, rng.integers(0, n_pred_cls, n_pred_rels)])
        rel_scores = np.sort(rng.random(n_pred_rels))[::-1].copy()
    else:
        pred_rel_inds = sel
        rel_scores = rng.random((n_pred_rels, n_pred_cls))
    return dict(gt_relations=gt_rels, gt_boxes=gt_boxes, gt_classes=gt_classes), dict(
        pred_rel_inds=pred_rel_inds, rel_scores=rel_scores, pred_boxes=pred_boxes, pred_classes=pred_classes, obj_scores=obj_scores)


def main():
    sys.path.insert(0, REF)
    stub = types.ModuleType("lib.fpn.box_intersections_cpu.bbox")
    stub.bbox_overlaps = _bbox_overlaps_loops
    for name in ("lib.fpn", "lib.fpn.box_intersections_cpu"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["lib.fpn.box_intersections_cpu.bbox"] = stub
    from lib.evaluation import sg_eval as ref  # the unmodified reference evaluator

    rng = np.random.default_rng(2024)
    out = {}
    cases = []
    for ci, (mode, multiple, iou) in enumerate([("sgdet", False, 0.5), ("sgdet", True, 0.5), ("phrdet", True, 0.5), ("sgdet", False, 0.3),
                                                ("sgcls", False, 0.5), ("predcls", True, 0.5)]):
        ev = ref.BasicSceneGraphEvaluator(mode, multiple_preds=multiple)
        for si in range(4):
            n_gt = int(rng.integers(4, 12))
            gt, pred = scene(rng, n_gt, int(rng.integers(3, 15)), n_gt if mode in ("sgcls", "predcls") else int(rng.integers(8, 30)),
                             int(rng.integers(30, 150)), 12, 6, multiple)
            if mode in ("sgcls", "predcls"):
                pred["pred_boxes"] = gt["gt_boxes"].copy()
            p2g, five, scores = ev.evaluate_scene_graph_entry(gt, pred, iou_thresh=iou)
            tag = f"c{ci}s{si}_"
            for k, v in {**gt, **pred}.items():
                out[tag + k] = np.asarray(v)
            out[tag + "p2g_len"] = np.array([len(x) for x in p2g])
            out[tag + "p2g_flat"] = np.array([g for x in p2g for g in x], dtype=np.int64)
            out[tag + "five"] = five
        for k, v in ev.result_dict[mode + "_recall"].items():
            out[f"c{ci}_recall{k}"] = np.array(v)
        cases.append((mode, int(multiple), iou))
    out["cases_mode"] = np.array([c[0] for c in cases])
    out["cases_multiple"] = np.array([c[1] for c in cases])
    out["cases_iou"] = np.array([c[2] for c in cases])
    # bbox_overlaps itself on random boxes (incl. disjoint and touching pairs)
    b = rng.uniform(0, 100, (40, 2))
    b = np.concatenate([b, b + rng.uniform(0, 60, (40, 2))], 1).round(0)
    out["iou_boxes"] = b
    out["iou_ref"] = _bbox_overlaps_loops(b[:25], b[15:])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sgeval.npz"), **out)
    print("wrote sgeval.npz with", len(cases), "cases")


if __name__ == "__main__":
    main()
