"""Golden values for egtr_b200/oi_evaluation.py, produced by the UNMODIFIED reference
(/root/reference/lib/evaluation/oi_eval.py::OIEvaluator.__call__ + eval_rel_results, ap_eval_rel.py) on seeded random scenes fed
through the same entry format `evaluate_batch` builds (train_egtr.py:154-173).  pycocotools is not installed: it is stubbed (only
the relation part is pinned — the COCO box AP the reference delegates to pycocotools is "parity unpinned"); the compiled
`lib.fpn.box_intersections_cpu.bbox` is replaced by the literal restatement used for sgeval.npz.  The scenes are regenerated from
their seed by the test, so only the reference's outputs are stored.

    python tests/golden/make_golden_oieval.py      # needs /root/reference; writes tests/golden/oieval.json
"""
import json
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden_sgeval import _bbox_overlaps_loops  # noqa: E402

REF = "/root/reference"
N_PRED_CLS, N_CLS = 9, 14


def scenes(seed, n_images):
    """(gt_entry, pred_entry) pairs in evaluate_batch's Open-Images form: all N*N (subject, object) pairs with [N*N, P] scores."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_images):
        n_gt, n_pred = int(rng.integers(3, 9)), int(rng.integers(6, 14))
        xy = rng.uniform(0, 300, (n_gt, 2))
        gt_boxes = np.concatenate([xy, xy + rng.uniform(30, 200, (n_gt, 2))], 1).round(1).astype(np.float32)
        gt_classes = rng.integers(0, N_CLS, n_gt)
        so = np.array([(s, o) for s in range(n_gt) for o in range(n_gt) if s != o])
        n_rel = int(min(len(so), rng.integers(1, 10)))  # the reference needs at least one relation per image (KeyError otherwise)
        gt_rels = np.column_stack([so[rng.choice(len(so), n_rel, replace=False)], rng.integers(0, N_PRED_CLS, n_rel)]).reshape(-1, 3)
        idx = rng.integers(0, n_gt, n_pred)
        jit = gt_boxes[idx] + rng.normal(0, 10, (n_pred, 4)).astype(np.float32)
        rnd_xy = rng.uniform(0, 300, (n_pred, 2))
        rnd = np.concatenate([rnd_xy, rnd_xy + rng.uniform(30, 200, (n_pred, 2))], 1).astype(np.float32)
        pred_boxes = np.where(rng.random((n_pred, 1)) < 0.7, jit, rnd).astype(np.float32)
        pred_classes = np.where(rng.random(n_pred) < 0.8, gt_classes[idx], rng.integers(0, N_CLS, n_pred))
        obj_scores = rng.random(n_pred).astype(np.float32)
        pairs = np.array([(s, o) for s in range(n_pred) for o in range(n_pred)])
        pred_scores = (rng.random((len(pairs), N_PRED_CLS)) ** 3).astype(np.float32)
        # make the ground-truth predicates likely winners of the pairs that reproduce them
        for s, o, p in gt_rels:
            for a in np.where(idx == s)[0]:
                for b in np.where(idx == o)[0]:
                    if rng.random() < 0.7:
                        pred_scores[a * n_pred + b, p] = 0.5 + 0.5 * rng.random()
        out.append((dict(gt_boxes=gt_boxes, gt_classes=gt_classes, gt_relations=gt_rels),
                    dict(pred_boxes=pred_boxes, pred_classes=pred_classes, obj_scores=obj_scores, sbj_obj_inds=pairs, pred_scores=pred_scores)))
    return out


def main():
    sys.path.insert(0, REF)
    stub = types.ModuleType("lib.fpn.box_intersections_cpu.bbox")
    stub.bbox_overlaps = _bbox_overlaps_loops
    for name in ("lib.fpn", "lib.fpn.box_intersections_cpu", "pycocotools", "pycocotools.coco", "pycocotools.cocoeval"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["lib.fpn.box_intersections_cpu.bbox"] = stub
    sys.modules["pycocotools.coco"].COCO = object
    sys.modules["pycocotools.cocoeval"].COCOeval = object
    from lib.evaluation import oi_eval as ref  # the unmodified reference evaluator

    golden = {}
    for seed, n_images in ((7, 6), (8, 12), (9, 3)):
        ev = ref.OIEvaluator([f"p{i}" for i in range(N_PRED_CLS)], [f"c{i}" for i in range(N_CLS)])
        for gt, pred in scenes(seed, n_images):
            ev(gt, pred)
        res = ref.eval_rel_results(ev.all_result, ev.predicate_cls_list)
        golden[str(seed)] = dict(n_images=n_images, **{k: float(v) for k, v in res.items()})
        print(seed, golden[str(seed)])
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oieval.json"), "w") as f:
        json.dump(golden, f, indent=1)


if __name__ == "__main__":
    main()
