"""NumPy restatement of the reference's PR-curve / mAP code (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/utils/measurement.py: per-image matching :217-292,
PR points :297-321, ``__call__`` :328-338, ``get_map`` :393-447,
``create_score_mat`` :16-150.  decode / NMS / IoU come from oracle/tools.py.

The O(D^2) prefix loop of the reference (:302-319) is restated as a
first-occurrence flag + cumulative sum, which yields the same integers.
Reference behaviour on ill-posed inputs (a class with no ground truth ->
ZeroDivisionError, a class with ground truth but no detections reusing the
previous class's counters) is NOT reproduced: inputs must give every class at
least one ground truth and one detection.

Tie rule for the two confidence sorts (``np.argsort(...)[::-1]``): descending,
equal keys later-position-first (see oracle/tools.py).

Parity pinning: checked against the unmodified reference executed in the build
container and tests/golden/map.npz.
"""
import numpy as np

from . import tools


def _detect(per_image_preds, class_num, conf_threshold, nms_mode, nms_threshold, version, nms_sigma=0.5):
    det = tools.decode(*per_image_preds, class_num=class_num, threshold=conf_threshold, version=version)
    if nms_mode > 0 and len(det) > 0:
        if nms_mode == 1:
            det = tools.nms(det, class_num, nms_threshold)
        elif nms_mode == 3:
            det = tools.nms(det, class_num, nms_threshold, 2)
        elif nms_mode == 2:
            det = tools.soft_nms(det, class_num, nms_threshold, conf_threshold, nms_sigma)
    return det.reshape(-1, 7)


def match_image(gt, det, class_num, iou_threshold):
    """Per class: (conf, best_gt_local, tp_flag, n_gt) for the detections of one image."""
    out = []
    gcls = gt[:, 5].astype("int") if len(gt) else np.zeros(0, int)
    dcls = det[:, 5].astype("int") if len(det) else np.zeros(0, int)
    for k in range(class_num):
        g, d = gt[gcls == k], det[dcls == k]
        conf = d[:, 4] * d[:, 6]
        if len(d) and len(g):
            m = tools.pair_iou(g[:, None, :5], d[None, :, :5])
            best, arg = m.max(axis=0), m.argmax(axis=0)
            flag = (best >= iou_threshold)
        else:
            arg = np.zeros(len(d), int)
            flag = np.zeros(len(d), bool)
        out.append((conf, arg, flag, len(g)))
    return out


class PRfunc:
    def __init__(self, y_trues, *y_preds, class_names=[], conf_threshold=0.05, nms_mode=1,
                 nms_threshold=0.5, nms_sigma=0.5, iou_threshold=0.5, precision_mode=2,
                 max_per_img=100, version=3):
        C = len(class_names)
        self.class_num, self.class_names = C, class_names
        gts = [0] * C
        dets = [[] for _ in range(C)]
        for i, y_true in enumerate(y_trues):
            gt = tools.decode(y_true, class_num=C, version=version).reshape(-1, 7)
            det = _detect([p[i] for p in y_preds], C, conf_threshold, nms_mode, nms_threshold, version, nms_sigma)
            for k, (conf, arg, flag, n_gt) in enumerate(match_image(gt, det, C, iou_threshold)):
                if len(conf):
                    gid = arg + gts[k] if n_gt > 0 else np.zeros(len(conf))
                    trip = np.stack((conf, gid, flag.astype("float32")), axis=1)
                    if max_per_img is not None and len(trip) > max_per_img:
                        trip = trip[tools.visit_order(trip[:, 0])][:max_per_img]
                    dets[k].append(trip)
                gts[k] += n_gt
        self.precisions, self.recalls = [], []
        for k in range(C):
            d = np.vstack(dets[k]) if dets[k] else np.empty((0, 3))
            d = d[tools.visit_order(d[:, 0])]
            flag = d[:, 2].astype(bool)
            ids = d[:, 1]
            first = np.zeros(len(d), bool)
            seen = set()
            for j in range(len(d)):
                if flag[j] and ids[j] not in seen:
                    seen.add(ids[j])
                    first[j] = True
            tp = np.cumsum(first)
            tpp = np.cumsum(flag)
            n = np.arange(1, len(d) + 1)
            if precision_mode == 0:
                pc = tpp / n
            elif precision_mode == 1:
                pc = tp / (tp + (n - tpp))
            else:
                pc = tp / n
            rc = tp / gts[k]
            self.precisions.append(np.append(pc, 0))
            self.recalls.append(np.append(rc, rc[-1]))

    def __call__(self, recall, class_idx=0):
        pc, rc = self.precisions[class_idx], self.recalls[class_idx]
        k = (rc > recall).sum()
        return 0 if k == 0 else pc[-k:].max()

    def get_ap(self, mode="voc2012"):
        aps = []
        for k in range(self.class_num):
            pc, rc = self.precisions[k], self.recalls[k]
            if mode in ("area", "smootharea"):
                if mode == "smootharea":
                    pc = pc.copy()
                    top = 0
                    for i in range(len(pc) - 1, -1, -1):
                        if pc[i] > top:
                            top = pc[i]
                        else:
                            pc[i] = top
                ap = 0
                for i in range(len(pc) - 1):
                    ap += (rc[i + 1] - rc[i]) * ((pc[i + 1] - pc[i]) / 2 + pc[i])
            else:
                pts = [0, 0.14, 0.29, 0.43, 0.57, 0.71, 1] if mode == "voc2012" else [i / 10 for i in range(11)]
                ap = 0
                for r in pts:
                    ap += self(r, k)
                ap = ap / len(pts)
            aps.append(ap)
        aps.append(sum(aps) / len(aps))
        return np.array(aps, dtype=np.float64)


def score_table(y_trues, *y_preds, class_names=[], conf_threshold=0.5, nms_mode=0, nms_threshold=0.5,
                nms_sigma=0.5, iou_threshold=0.5, precision_mode=2, version=3):
    """create_score_mat (:16-150) as a dict of per-class arrays:
    precision, recall, F1-score, gts, dets."""
    C = len(class_names)
    denom = np.zeros((C, 2))
    tps = np.zeros((C, 2))
    det_counts = np.zeros(C, dtype="int")
    for i, y_true in enumerate(y_trues):
        gt = tools.decode(y_true, class_num=C, version=version).reshape(-1, 7)
        det = _detect([p[i] for p in y_preds], C, conf_threshold, nms_mode, nms_threshold, version, nms_sigma)
        for k, (conf, arg, flag, n_gt) in enumerate(match_image(gt, det, C, iou_threshold)):
            denom[k] += (len(conf), n_gt)
            det_counts[k] += len(conf)
            if n_gt > 0 and len(conf) > 0:
                tpp, tp = int(flag.sum()), len(set(arg[flag]))
                if precision_mode == 1:
                    denom[k, 0] -= tpp - tp
                if precision_mode > 0:
                    tpp = tp
                tps[k] += (tpp, tp)
    with np.errstate(invalid="ignore", divide="ignore"):
        pr = np.true_divide(tps, denom)
        f1 = (2 * pr[:, 0] * pr[:, 1]) / (pr[:, 0] + pr[:, 1])
    return {"precision": pr[:, 0], "recall": pr[:, 1], "F1-score": f1,
            "gts": denom[:, 1].astype("int"), "dets": det_counts}
