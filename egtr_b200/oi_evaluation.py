"""Open-Images relation metrics and COCO-style box AP (the rest of `lib/evaluation`).

Same interfaces as the reference's `OIEvaluator` (`/root/reference/lib/evaluation/oi_eval.py:437-483`; fed by
`train_egtr.py:154-173`) and `CocoEvaluator` (`lib/evaluation/coco_eval.py:24-170`; fed by `evaluate_egtr.py:86-103`), so the
evaluation loop keeps working unchanged on the triplets / boxes this library emits.  Host code, numpy only:

* relation part (`eval_rel_results`, oi_eval.py:77-279 + `ap_eval_rel.py`): per image the top-100 (pair, predicate) candidates
  out of `score_s * score_o * top-2 predicate scores`, recall@K against the ground-truth triplets, and the per-predicate
  VOC-style average precision of relation (`min(IoU_s, IoU_o)`) and phrase (IoU of the enclosing boxes) detection, weighted by
  class frequency: `score = 0.4 wmAP_rel + 0.4 wmAP_phr + 0.2 R@50`.  Vectorised per image and per predicate class (the
  reference walks detections one by one through torch tensors); pinned against the unmodified reference by
  `tests/golden/make_golden_oieval.py`.  Kept quirks: the IoU of `ap_eval_rel.bbox_iou` adds one pixel to the intersection
  extents but not to the areas; a detection whose best ground truth was already taken is a false positive.
* box AP (`eval_entites_detection`, `CocoEvaluator`): the reference delegates to pycocotools' `COCOeval`, which is not installed
  here — **parity unpinned**; `CocoBoxEval` restates the published bbox algorithm of pycocotools 2.0 (`cocoeval.py`: greedy
  matching per IoU threshold in score order, ignore regions by area range, 101-point interpolated precision) and is checked
  against hand-computed cases only.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, List, Optional, Sequence

import numpy as np

from .evaluation import match_matrix  # noqa: F401  (re-exported for callers that build their own loops)


# ------------------------------------------------------------------------------------------------ relation metrics
def _iou_plus1_inter(box: np.ndarray, boxes: np.ndarray) -> np.ndarray:
    """`ap_eval_rel.bbox_iou` (ap_eval_rel.py:41-66) of one box against [M, 4]: +1 on the intersection extents only, float32."""
    box = box.astype(np.float32)
    boxes = boxes.astype(np.float32)
    lt = np.maximum(box[None, :2], boxes[:, :2])
    rb = np.minimum(box[None, 2:], boxes[:, 2:])
    wh = np.clip(rb - lt + np.float32(1), 0, None)
    inter = wh[:, 0] * wh[:, 1]
    a1 = (box[2] - box[0]) * (box[3] - box[1])
    a2 = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    return inter / (a1 + a2 - inter)


def _boxes_union(b1: np.ndarray, b2: np.ndarray) -> np.ndarray:
    return np.concatenate([np.minimum(b1[:, :2], b2[:, :2]), np.maximum(b1[:, 2:], b2[:, 2:])], 1) if len(b1) else np.zeros((0, 4), b1.dtype)


def _voc_ap(rec: np.ndarray, prec: np.ndarray) -> float:
    """Area under the monotone precision envelope (ap_eval_rel.py:148-165)."""
    mrec = np.concatenate(([0.0], rec, [1.0]))
    mpre = np.concatenate(([0.0], prec, [0.0]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return float(np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1]))


def _pair_iou_inclusive(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    from .evaluation import bbox_overlaps
    return bbox_overlaps(a, b)


def top_relation_candidates(res: dict, topk: int = 100, prd_k: int = 2):
    """oi_eval.py:87-160: the image's top-`topk` (pair, predicate) candidates -> subject boxes, object boxes, (s, p, o) labels, scores."""
    if res.get("pred_scores") is None or len(res["pred_scores"]) == 0:
        z4 = np.zeros((0, 4), np.float32)
        return z4, z4, np.zeros((0, 3), np.int64), np.zeros(0, np.float32)
    prd = np.asarray(res["pred_scores"])
    order = np.argsort(-prd, axis=1)
    prd_sorted = -np.sort(-prd, axis=1)
    so = np.asarray(res["sbj_scores"]) * np.asarray(res["obj_scores"])
    spo = so[:, None] * prd_sorted[:, :prd_k]
    flat = np.argsort(-spo.ravel())[:topk]
    pi, ki = np.unravel_index(flat, spo.shape)
    score = spo[pi, ki]
    keep = score > 0.00001
    pi, ki, score = pi[keep], ki[keep], score[keep]
    labels = np.stack([np.asarray(res["sbj_labels"])[pi], order[pi, ki], np.asarray(res["obj_labels"])[pi]], 1)
    return np.asarray(res["sbj_boxes"])[pi], np.asarray(res["obj_boxes"])[pi], labels, score


def eval_rel_results(all_results: Sequence[dict], predicate_cls_list: Sequence, topk: int = 100) -> Dict[str, float]:
    """Weighted relation / phrase mAP, micro recall@50 and the Open-Images score (oi_eval.py:77-279)."""
    n_cls = len(predicate_cls_list)
    ks = (1, 5, 10, 20, 50, 100)
    hit = {k: 0 for k in ks}
    n_gt_total = 0
    dets: List[dict] = []
    for im, res in enumerate(all_results):
        bs, bo, lab, sc = top_relation_candidates(res, topk)
        g_bs, g_bo = np.asarray(res["gt_sbj_boxes"], np.float64).reshape(-1, 4), np.asarray(res["gt_obj_boxes"], np.float64).reshape(-1, 4)
        g_lab = np.stack([np.asarray(res["gt_sbj_labels"]), np.asarray(res["gt_prd_labels"]), np.asarray(res["gt_obj_labels"])], 1).reshape(-1, 3)
        # recall: ground truth i is recalled at K when one of the first K candidates has its triplet and overlaps both boxes
        if len(g_lab) and len(lab):
            same = (g_lab[:, None, :] == lab[None, :, :]).all(-1)
            m = same & (_pair_iou_inclusive(g_bs, bs) >= 0.5) & (_pair_iou_inclusive(g_bo, bo) >= 0.5)
            first = np.where(m.any(1), m.argmax(1), np.iinfo(np.int64).max)
            for k in ks:
                hit[k] += int((first < k).sum())
        n_gt_total += len(g_lab)
        dets.append(dict(image=im, bs=bs, bo=bo, lab=lab, score=sc, g_bs=g_bs, g_bo=g_bo, g_lab=g_lab))
    recalls = {k: hit[k] / (n_gt_total + 1e-12) for k in ks}

    def class_ap(c: int, rel: bool):
        conf, img, b_s, b_o, l_s, l_o = [], [], [], [], [], []
        gts, npos = {}, 0
        for d in dets:
            sel = np.where(d["lab"][:, 1] == c)[0] if len(d["lab"]) else np.zeros(0, np.int64)
            if len(sel):
                conf.append(d["score"][sel]); img.append(np.full(len(sel), d["image"])); b_s.append(d["bs"][sel]); b_o.append(d["bo"][sel])
                l_s.append(d["lab"][sel, 0]); l_o.append(d["lab"][sel, 2])
            gsel = np.where(d["g_lab"][:, 1] == c)[0] if len(d["g_lab"]) else np.zeros(0, np.int64)
            gts[d["image"]] = dict(bs=d["g_bs"][gsel], bo=d["g_bo"][gsel], ls=d["g_lab"][gsel, 0] if len(gsel) else np.zeros(0, np.int64),
                                   lo=d["g_lab"][gsel, 2] if len(gsel) else np.zeros(0, np.int64), seen=np.zeros(len(gsel), bool))
            npos += len(gsel)
        if not conf:
            return 0.0, npos
        conf, img = np.concatenate(conf), np.concatenate(img)
        b_s, b_o, l_s, l_o = np.concatenate(b_s), np.concatenate(b_o), np.concatenate(l_s), np.concatenate(l_o)
        order = np.argsort(-conf)
        tp = np.zeros(len(order))
        for rank, d in enumerate(order):  # greedy, in confidence order (ap_eval_rel.py:202-252)
            g = gts[int(img[d])]
            if len(g["ls"]) == 0:
                continue
            valid = (g["ls"] == l_s[d]) & (g["lo"] == l_o[d])
            if not valid.any():
                continue
            if rel:
                ov = np.minimum(_iou_plus1_inter(b_s[d], g["bs"]), _iou_plus1_inter(b_o[d], g["bo"]))
            else:
                ov = _iou_plus1_inter(_boxes_union(b_s[d][None], b_o[d][None])[0], _boxes_union(g["bs"], g["bo"]))
            ov = ov * valid
            j = int(np.argmax(ov))
            if ov[j] > 0.5 and not g["seen"][j]:
                tp[rank] = 1.0
                g["seen"][j] = True
        ctp = np.cumsum(tp)
        cfp = np.cumsum(1.0 - tp)
        rec = ctp / (float(npos) + 1e-12)
        prec = ctp / np.maximum(ctp + cfp, np.finfo(np.float64).eps)
        return _voc_ap(rec, prec), npos

    out = {}
    for name, rel in (("rel", True), ("phr", False)):
        aps, npos = zip(*[class_ap(c, rel) for c in range(n_cls)])
        total = float(sum(npos))
        out[f"w_{name}_mAP"] = float(sum(ap * n / total for ap, n in zip(aps, npos))) if total > 0 else 0.0
        out[f"{name}_mAP"] = float(np.mean(aps))
    r50 = recalls[50]
    return {"w_rel_mAP": out["w_rel_mAP"], "w_phr_mAP": out["w_phr_mAP"], "microR@50": r50,
            "score": out["w_rel_mAP"] * 0.4 + out["w_phr_mAP"] * 0.4 + r50 * 0.2,
            "rel_mAP": out["rel_mAP"], "phr_mAP": out["phr_mAP"], **{f"R@{k}": recalls[k] for k in ks}}


# ------------------------------------------------------------------------------------------------ COCO box AP
class CocoBoxEval:
    """Bounding-box evaluation with pycocotools' `COCOeval` semantics (bbox only): `stats` = [AP, AP50, AP75, APs, APm, APl,
    AR1, AR10, AR100, ARs, ARm, ARl].  Ground truth / detections are COCO-style dicts (`bbox` xywh, `category_id`, `image_id`,
    optional `iscrowd`, `area`; detections carry `score`)."""

    iou_thrs = np.linspace(0.5, 0.95, 10)
    rec_thrs = np.linspace(0.0, 1.0, 101)
    max_dets = (1, 10, 100)
    area_rng = ((0, 1e5 ** 2), (0, 32 ** 2), (32 ** 2, 96 ** 2), (96 ** 2, 1e5 ** 2))

    def __init__(self, annotations: Sequence[dict], img_ids: Optional[Sequence] = None, cat_ids: Optional[Sequence] = None):
        self.gts = defaultdict(list)
        for a in annotations:
            a = dict(a)
            a.setdefault("iscrowd", 0)
            a.setdefault("area", a["bbox"][2] * a["bbox"][3])
            self.gts[(a["image_id"], a["category_id"])].append(a)
        self.img_ids = sorted(set(img_ids) if img_ids is not None else {a["image_id"] for a in annotations})
        self.cat_ids = sorted(set(cat_ids) if cat_ids is not None else {a["category_id"] for a in annotations})
        self.dts = defaultdict(list)
        self.stats = np.zeros(12)
        self.eval = None

    def add_detections(self, dets: Sequence[dict]):
        for d in dets:
            d = dict(d)
            d.setdefault("area", d["bbox"][2] * d["bbox"][3])
            self.dts[(d["image_id"], d["category_id"])].append(d)

    @staticmethod
    def _iou(dt: np.ndarray, gt: np.ndarray, crowd: np.ndarray) -> np.ndarray:
        """maskApi bbIou: xywh boxes, no +1; for a crowd ground truth the union is the detection's own area."""
        if len(dt) == 0 or len(gt) == 0:
            return np.zeros((len(dt), len(gt)))
        dx1, dy1, dx2, dy2 = dt[:, 0, None], dt[:, 1, None], dt[:, 0, None] + dt[:, 2, None], dt[:, 1, None] + dt[:, 3, None]
        gx1, gy1, gx2, gy2 = gt[None, :, 0], gt[None, :, 1], gt[None, :, 0] + gt[None, :, 2], gt[None, :, 1] + gt[None, :, 3]
        iw = np.clip(np.minimum(dx2, gx2) - np.maximum(dx1, gx1), 0, None)
        ih = np.clip(np.minimum(dy2, gy2) - np.maximum(dy1, gy1), 0, None)
        inter = iw * ih
        da, ga = (dt[:, 2] * dt[:, 3])[:, None], (gt[:, 2] * gt[:, 3])[None, :]
        union = np.where(crowd[None, :], da, da + ga - inter)
        return np.where(union > 0, inter / np.where(union > 0, union, 1), 0.0)

    def _evaluate_img(self, img, cat, rng, max_det):
        gt, dt = self.gts.get((img, cat), []), self.dts.get((img, cat), [])
        if not gt and not dt:
            return None
        g_ig = np.array([bool(g["iscrowd"]) or g["area"] < rng[0] or g["area"] > rng[1] for g in gt], bool)
        g_ord = np.argsort(g_ig, kind="mergesort")
        gt = [gt[i] for i in g_ord]
        g_ig = g_ig[g_ord]
        d_ord = np.argsort([-d["score"] for d in dt], kind="mergesort")[:max_det]
        dt = [dt[i] for i in d_ord]
        crowd = np.array([bool(g["iscrowd"]) for g in gt], bool)
        ious = self._iou(np.array([d["bbox"] for d in dt], float).reshape(-1, 4), np.array([g["bbox"] for g in gt], float).reshape(-1, 4), crowd)
        T, G, D = len(self.iou_thrs), len(gt), len(dt)
        gtm = np.zeros((T, G), bool)
        dtm = np.zeros((T, D), bool)
        dt_ig = np.zeros((T, D), bool)
        for ti, t in enumerate(self.iou_thrs):
            for di in range(D):
                best, m = min(t, 1 - 1e-10), -1
                for gi in range(G):
                    if gtm[ti, gi] and not crowd[gi]:
                        continue
                    if m > -1 and not g_ig[m] and g_ig[gi]:
                        break  # regular ground truths come first: stop at the ignore ones once a regular match exists
                    if ious[di, gi] < best:
                        continue
                    best, m = ious[di, gi], gi
                if m == -1:
                    continue
                dt_ig[ti, di] = g_ig[m]
                dtm[ti, di] = True
                gtm[ti, m] = True
        d_area_out = np.array([d["area"] < rng[0] or d["area"] > rng[1] for d in dt], bool)
        dt_ig |= (~dtm) & d_area_out[None, :]
        return dict(scores=np.array([d["score"] for d in dt]), dtm=dtm, dt_ig=dt_ig, n_gt=int((~g_ig).sum()))

    def evaluate_and_accumulate(self):
        T, R, K, A, M = len(self.iou_thrs), len(self.rec_thrs), len(self.cat_ids), len(self.area_rng), len(self.max_dets)
        precision = -np.ones((T, R, K, A, M))
        recall = -np.ones((T, K, A, M))
        for ki, cat in enumerate(self.cat_ids):
            for ai, rng in enumerate(self.area_rng):
                per_img = [self._evaluate_img(img, cat, rng, self.max_dets[-1]) for img in self.img_ids]
                per_img = [e for e in per_img if e is not None]
                if not per_img:
                    continue
                for mi, md in enumerate(self.max_dets):
                    scores = np.concatenate([e["scores"][:md] for e in per_img])
                    order = np.argsort(-scores, kind="mergesort")
                    dtm = np.concatenate([e["dtm"][:, :md] for e in per_img], 1)[:, order]
                    dig = np.concatenate([e["dt_ig"][:, :md] for e in per_img], 1)[:, order]
                    npig = sum(e["n_gt"] for e in per_img)
                    if npig == 0:
                        continue
                    tps = np.cumsum(dtm & ~dig, 1).astype(float)
                    fps = np.cumsum(~dtm & ~dig, 1).astype(float)
                    for ti in range(T):
                        tp, fp = tps[ti], fps[ti]
                        nd = len(tp)
                        rc = tp / npig
                        pr = tp / (fp + tp + np.spacing(1))
                        recall[ti, ki, ai, mi] = rc[-1] if nd else 0
                        pr = np.maximum.accumulate(pr[::-1])[::-1] if nd else pr
                        inds = np.searchsorted(rc, self.rec_thrs, side="left")
                        q = np.zeros(R)
                        ok = inds < nd
                        q[ok] = pr[inds[ok]]
                        precision[ti, :, ki, ai, mi] = q
        self.eval = dict(precision=precision, recall=recall)
        return self.eval

    def summarize(self):
        if self.eval is None:
            self.evaluate_and_accumulate()
        p, r = self.eval["precision"], self.eval["recall"]

        def mean(x):
            x = x[x > -1]
            return float(x.mean()) if x.size else -1.0

        t50, t75 = 0, 5
        self.stats = np.array([
            mean(p[:, :, :, 0, 2]), mean(p[t50, :, :, 0, 2]), mean(p[t75, :, :, 0, 2]), mean(p[:, :, :, 1, 2]), mean(p[:, :, :, 2, 2]),
            mean(p[:, :, :, 3, 2]), mean(r[:, :, 0, 0]), mean(r[:, :, 0, 1]), mean(r[:, :, 0, 2]), mean(r[:, :, 1, 2]), mean(r[:, :, 2, 2]),
            mean(r[:, :, 3, 2])])
        return self.stats


def _xyxy_to_xywh(b):
    b = np.asarray(b, float)
    return [float(b[0]), float(b[1]), float(b[2] - b[0]), float(b[3] - b[1])]


class CocoEvaluator:
    """`lib/evaluation/coco_eval.py:24-170` for iou type "bbox": `coco_gt` is a COCO-format dataset dict (`annotations`, `images`,
    `categories`) or any object with a `.dataset` of that form (e.g. a pycocotools `COCO`).  `update` takes the reference's
    `{image_id: {"boxes" xyxy, "scores", "labels"}}` (labels are shifted by +1 exactly as coco_eval.py:44-45 does)."""

    def __init__(self, coco_gt, iou_types=("bbox",)):
        if list(iou_types) != ["bbox"]:
            raise NotImplementedError("only the 'bbox' iou type is evaluated on this path")
        ds = coco_gt if isinstance(coco_gt, dict) else coco_gt.dataset
        self.iou_types = list(iou_types)
        self._eval = CocoBoxEval(ds["annotations"], [im["id"] for im in ds.get("images", [])] or None,
                                 [c["id"] for c in ds.get("categories", [])] or None)
        self.coco_eval = {"bbox": self._eval}
        self.img_ids: List = []

    def update(self, predictions: dict):
        self.img_ids.extend(np.unique(list(predictions.keys())).tolist())
        dets = []
        for image_id, p in predictions.items():
            if len(p) == 0:
                continue
            boxes, scores, labels = (np.asarray(getattr(p[k], "cpu", lambda: p[k])()) for k in ("boxes", "scores", "labels"))
            dets += [dict(image_id=image_id, category_id=int(l) + 1, bbox=_xyxy_to_xywh(b), score=float(s)) for b, s, l in zip(boxes, scores, labels)]
        self._eval.add_detections(dets)

    def synchronize_between_processes(self):
        """One process per GPU evaluates its own images; gather the detections of all ranks (torch.distributed, if initialised)."""
        try:
            import torch.distributed as dist
        except ImportError:
            return
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        mine = [d for ds in self._eval.dts.values() for d in ds]
        everyone = [None] * dist.get_world_size()
        dist.all_gather_object(everyone, mine)
        self._eval.dts = defaultdict(list)
        for part in everyone:
            self._eval.add_detections(part)

    def accumulate(self):
        self._eval.img_ids = sorted(set(self.img_ids)) or self._eval.img_ids
        self._eval.evaluate_and_accumulate()

    def summarize(self):
        stats = self._eval.summarize()
        names = ["AP", "AP50", "AP75", "APs", "APm", "APl", "AR@1", "AR@10", "AR@100", "ARs", "ARm", "ARl"]
        print("IoU metric: bbox")
        for n, v in zip(names, stats):
            print(f" {n:<7} = {v:0.3f}")


class OIEvaluator:
    """`lib/evaluation/oi_eval.py:437-483`: accumulates per-image entries of `evaluate_batch` and aggregates the Open-Images metrics."""

    def __init__(self, predicate_cls_list, ind_to_classes):
        self.predicate_cls_list = list(predicate_cls_list)
        self.ind_to_classes = list(ind_to_classes)
        self.all_result: List[dict] = []

    def __call__(self, gt_entry, pred_entry):
        gt_boxes, gt_class = np.asarray(gt_entry["gt_boxes"]), np.asarray(gt_entry["gt_classes"])
        rels = np.asarray(gt_entry["gt_relations"]).reshape(-1, 3)
        pb, pc, ps = np.asarray(pred_entry["pred_boxes"]), np.asarray(pred_entry["pred_classes"]), np.asarray(pred_entry["obj_scores"])
        so = np.asarray(pred_entry["sbj_obj_inds"]).reshape(-1, 2)
        self.all_result.append(dict(
            gt_boxes=gt_boxes, gt_class=gt_class,
            gt_sbj_boxes=gt_boxes[rels[:, 0]], gt_obj_boxes=gt_boxes[rels[:, 1]], gt_sbj_labels=gt_class[rels[:, 0]],
            gt_obj_labels=gt_class[rels[:, 1]], gt_prd_labels=rels[:, 2],
            pred_boxes=pb, pred_class=pc, pred_cls_scores=ps,
            sbj_boxes=pb[so[:, 0]], obj_boxes=pb[so[:, 1]], sbj_labels=pc[so[:, 0]], obj_labels=pc[so[:, 1]],
            sbj_scores=ps[so[:, 0]], obj_scores=ps[so[:, 1]], pred_scores=np.asarray(pred_entry["pred_scores"])))

    def detection_metrics(self) -> Dict[str, float]:
        """`eval_entites_detection` (oi_eval.py:283-392): COCO-style box AP of the predicted entities."""
        anns = []
        for image_id, r in enumerate(self.all_result):
            for cls, box in zip(r["gt_class"].tolist(), r["gt_boxes"].tolist()):
                anns.append(dict(area=(box[3] - box[1] + 1) * (box[2] - box[0] + 1), bbox=[box[0], box[1], box[2] - box[0] + 1, box[3] - box[1] + 1],
                                 category_id=cls, id=len(anns), image_id=image_id, iscrowd=0))
        cats = [i for i, name in enumerate(self.ind_to_classes) if name != "__background__"]
        ev = CocoBoxEval(anns, list(range(len(self.all_result))), cats)
        dets = []
        for image_id, r in enumerate(self.all_result):
            for b, s, l in zip(r["pred_boxes"], r["pred_cls_scores"], r["pred_class"]):
                dets.append(dict(image_id=image_id, category_id=int(l), score=float(s), bbox=[float(b[0]), float(b[1]), float(b[2] - b[0] + 1), float(b[3] - b[1] + 1)]))
        ev.add_detections(dets)
        stats = ev.summarize()
        return {f"bbox/{n}": float(v) for n, v in zip(["AP", "AP50", "AP75", "APs", "APm", "APl"], stats[:6])}

    def aggregate_metrics(self) -> Dict[str, float]:
        out = {}
        out.update(self.detection_metrics())
        rel = eval_rel_results(self.all_result, self.predicate_cls_list)
        out.update({k: rel[k] for k in ("w_rel_mAP", "w_phr_mAP", "microR@50", "score")})
        return out
