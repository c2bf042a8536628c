"""Golden vectors for the `preddet` mode of egtr_b200/evaluation.py, produced by the UNMODIFIED reference evaluator
(/root/reference/lib/evaluation/sg_eval.py:107-131) on seeded random scenes (same import stub as make_golden_sgeval.py).

    python tests/golden/make_golden_preddet.py      # needs /root/reference; writes tests/golden/sgeval_preddet.npz
"""
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden_sgeval import REF, _bbox_overlaps_loops, scene  # noqa: E402


def main():
    sys.path.insert(0, REF)
    stub = types.ModuleType("lib.fpn.box_intersections_cpu.bbox")
    stub.bbox_overlaps = _bbox_overlaps_loops
    for name in ("lib.fpn", "lib.fpn.box_intersections_cpu"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["lib.fpn.box_intersections_cpu.bbox"] = stub
    from lib.evaluation import sg_eval as ref  # the unmodified reference evaluator

    rng = np.random.default_rng(77)
    out = {}
    ev = ref.BasicSceneGraphEvaluator.vrd_modes()["preddet"]
    n_scenes = 6
    for si in range(n_scenes):
        n_gt = int(rng.integers(4, 10))
        # predictions are made on the ground-truth boxes (predicate detection), in the per-pair form: inds [n,2], scores [n,P]
        gt, pred = scene(rng, n_gt, int(rng.integers(3, 15)), n_gt, int(rng.integers(10, n_gt * (n_gt - 1) + 1)), 12, 6, False)
        if si == 4:  # no predicted pair at all
            pred["pred_rel_inds"] = pred["pred_rel_inds"][:0]
            pred["rel_scores"] = pred["rel_scores"][:0]
        ev.evaluate_scene_graph_entry(gt, pred)
        for k, v in {**gt, **pred}.items():
            out[f"s{si}_{k}"] = np.asarray(v)
    for k, v in ev.result_dict["preddet_recall"].items():
        out[f"recall{k}"] = np.array(v)
    out["n_scenes"] = np.array(n_scenes)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sgeval_preddet.npz"), **out)
    print("wrote sgeval_preddet.npz", {k: v for k, v in ev.result_dict["preddet_recall"].items()})


if __name__ == "__main__":
    main()
