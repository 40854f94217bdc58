"""Mirror of /root/reference/utils/measurement.py: ``create_score_mat`` (:16-150), ``PRfunc``
(:198-447) and the deprecated ``PR_func`` alias (:450-455), same constructor keywords,
``__call__`` / ``get_map`` semantics and DataFrame layouts.

The per-image Python pipeline of the reference (decode ground truth, decode predictions, NMS,
per-class IoU matching, per-image top-k) runs batched on the GPU: yb_decode -> yb_nms ->
yb_map_match -> yb_map_accumulate, chunk of images by chunk, without leaving the device.  The
O(D^2) prefix loop (:302-319) is one device sort by (class, confidence) plus a first-occurrence
scan (yb_pr_curve).  Only the per-class integer prefix counts come back; precision / recall /
AP arithmetic on those few numbers follows the reference line by line on the host.

Well-posed inputs only (as SURVEY.md 8a R10 documents, the reference itself fails otherwise):
every class needs at least one ground truth; a class without detections gets the single sentinel
point (precision 0, recall 0).

With ``process_group`` the images are this rank's shard (rank order = image order): per-class
ground-truth counts are all-gathered to offset ``gt_id`` and the triples are all-gathered before
the final sort, so every rank ends with the same curves.
"""
import math
import warnings

import numpy as np
import pandas as pd
import torch

from .. import dist as dist_util
from .. import engine
from .._native import YoloB200Error
from .tools import cal_iou, decode, nms, soft_nms  # noqa: F401  (same module globals as the reference)

_CHUNK_BYTES = 2 << 30  # device bytes of head tensors per chunk of images


def _n_images(a):
    return a.shape[0]


def _to_dev(a, sl, dev, keep_f64):
    if torch.is_tensor(a):
        t = a[sl]
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a[sl])))
    if t.dtype == torch.float64 and keep_f64:
        pass
    elif t.dtype != torch.float32:
        t = t.to(torch.float32) if t.is_cuda else t.float()
    return t.to(dev, non_blocking=True).contiguous()


def _nms_args(nms_mode, nms_threshold):
    if nms_mode == 0:
        return math.inf, 1          # nothing is suppressed: pure grouping by class
    if nms_mode in (1, 2):
        return float(nms_threshold), 1
    if nms_mode == 3:
        return float(nms_threshold), 2
    raise ValueError(f"Invalid nms_mode: {nms_mode}")


class _Accumulated:
    __slots__ = ("conf", "gid", "flag", "cls", "class_counts", "gts", "score")


def _accumulate(y_trues, y_preds, class_num, conf_threshold, nms_mode, nms_threshold, iou_threshold,
                max_per_img, version, process_group=None, nms_sigma=0.5):
    if not torch.cuda.is_available():
        raise YoloB200Error("no CUDA device: tf2_yolo_b200 has no CPU fallback")
    if len(y_preds) == 0:
        raise ValueError("at least one prediction array is needed")
    dev = torch.device("cuda", torch.cuda.current_device())
    n_img = _n_images(y_trues)
    thr, iou_mode = _nms_args(nms_mode, nms_threshold)
    per_img = sum(int(np.prod(p.shape[1:])) * 4 for p in y_preds) + int(np.prod(y_trues.shape[1:])) * 8
    chunk = max(1, min(n_img, _CHUNK_BYTES // max(per_img, 1)))
    gt_is_f64 = (y_trues.dtype == torch.float64) if torch.is_tensor(y_trues) else (np.asarray(y_trues[:0]).dtype == np.float64)

    C = class_num
    score = torch.zeros(3 * C, dtype=torch.int64, device=dev)
    gt_base = torch.zeros(C, dtype=torch.int64, device=dev)
    gts_total = None
    if process_group is not None:
        # ground-truth counts of every rank first: gt_id = local argmax + counts of earlier ranks/images
        local = torch.zeros(C, dtype=torch.int64, device=dev)
        for s in range(0, n_img, chunk):
            yt = _to_dev(y_trues, slice(s, s + chunk), dev, True)
            rows, off = engine.decode_batch_exact([yt], C, 0.5, version)
            if rows.shape[0]:
                local += torch.bincount(rows[:, 5].long(), minlength=C)[:C]
        before, gts_total = dist_util.rank_offsets(local, process_group)
        gt_base = gt_base + before
    parts = []
    counts_total = np.zeros(C, dtype=np.int64)
    for s in range(0, n_img, chunk):
        sl = slice(s, min(n_img, s + chunk))
        yt = _to_dev(y_trues, sl, dev, True)
        preds = [_to_dev(p, sl, dev, False) for p in y_preds]
        gt_rows, gt_off = engine.decode_batch_exact([yt], C, 0.5, version)
        det_rows, det_off = engine.decode_batch_exact(preds, C, conf_threshold, version)
        res = engine.nms_batch(det_rows, det_off, C, thr, iou_mode, want_seg_offsets=True,
                               soft=(conf_threshold, nms_sigma) if nms_mode == 2 else None)
        n_keep = int(res["out_offsets"][-1].item())
        dets = res["out_rows"][:n_keep].contiguous()
        best_iou, best_gt, counts = engine.map_match(gt_rows, gt_off, dets, res["out_offsets"], C)
        conf, gid, flag, cls, coff = engine.map_accumulate(
            dets, res["seg_offsets"], best_iou, best_gt, counts, C, iou_threshold, max_per_img, gt_base, score)
        coff_h = coff.cpu().numpy()
        n_out = int(coff_h[-1])
        parts.append((conf[:n_out], gid[:n_out], flag[:n_out], cls[:n_out]))
        counts_total += np.diff(coff_h)
        gt_base = gt_base + counts.sum(dim=0, dtype=torch.int64)
    out = _Accumulated()
    out.conf = torch.cat([p[0] for p in parts])
    out.gid = torch.cat([p[1] for p in parts])
    out.flag = torch.cat([p[2] for p in parts])
    out.cls = torch.cat([p[3] for p in parts])
    out.class_counts = counts_total
    out.score = score
    if process_group is None:
        out.gts = gt_base.cpu().numpy()
    else:
        out.gts = gts_total.cpu().numpy()
        out.conf = dist_util.gather_varlen(out.conf, process_group)
        out.gid = dist_util.gather_varlen(out.gid, process_group)
        out.flag = dist_util.gather_varlen(out.flag, process_group)
        out.cls = dist_util.gather_varlen(out.cls, process_group)
        out.class_counts = dist_util.allreduce_sum(torch.from_numpy(counts_total).to(dev), process_group).cpu().numpy()
        dist_util.allreduce_sum(out.score, process_group)
    return out


def create_score_mat(y_trues, *y_preds,
                     class_names=[],
                     conf_threshold=0.5,
                     nms_mode=0,
                     nms_threshold=0.5,
                     nms_sigma=0.5,
                     iou_threshold=0.5,
                     precision_mode=2,
                     version=3,
                     process_group=None):
    """Score matrix table: precision, recall, F1-score, gts and dets per class."""
    class_num = len(class_names)
    acc = _accumulate(y_trues, y_preds, class_num, conf_threshold, nms_mode, nms_threshold,
                      iou_threshold, None, version, process_group, nms_sigma)
    sc = acc.score.cpu().numpy().reshape(3, class_num)
    pp, tpp, tp = sc[0].astype(np.float64), sc[1].astype(np.float64), sc[2].astype(np.float64)
    denom_array = np.zeros((class_num, 2))
    tp_array = np.zeros((class_num, 2))
    denom_array[:, 0] = pp
    denom_array[:, 1] = acc.gts
    det_counts = sc[0].astype("int")
    if precision_mode == 1:
        denom_array[:, 0] -= (tpp - tp)
    if precision_mode > 0:
        tpp = tp
    tp_array[:, 0] = tpp
    tp_array[:, 1] = tp
    with np.errstate(invalid="ignore", divide="ignore"):
        score_table = np.true_divide(tp_array, denom_array)
    score_table = pd.DataFrame(score_table)
    score_table.columns = ["precision", "recall"]

    precision = score_table["precision"]
    recall = score_table["recall"]
    f1_score = (2*precision*recall)/(precision + recall)
    score_table["F1-score"] = f1_score
    score_table["gts"] = denom_array[:, 1].astype("int")
    score_table["dets"] = det_counts

    score_table.index = class_names

    return score_table


class PRfunc(object):
    """Precision-recall function; call it with a recall value to get a precision value."""

    def __init__(self,
                 y_trues, *y_preds,
                 class_names=[],
                 conf_threshold=0.05,
                 nms_mode=1,
                 nms_threshold=0.5,
                 nms_sigma=0.5,
                 iou_threshold=0.5,
                 precision_mode=2,
                 max_per_img=100,
                 version=3,
                 process_group=None):
        class_num = len(class_names)
        self.class_num = class_num
        self.class_names = class_names

        acc = _accumulate(y_trues, y_preds, class_num, conf_threshold, nms_mode, nms_threshold,
                          iou_threshold, max_per_img, version, process_group, nms_sigma)
        gts = [int(g) for g in acc.gts]
        dev = acc.conf.device
        table = np.concatenate([[0], np.cumsum(acc.gts)]).astype(np.int64)
        order, tp_cum, tpp_cum = engine.pr_curve(acc.conf, acc.cls, acc.gid, acc.flag,
                                                 torch.from_numpy(table).to(dev), int(table[-1]))
        tp_cum = tp_cum.cpu().numpy()
        tpp_cum = tpp_cum.cpu().numpy()
        starts = np.concatenate([[0], np.cumsum(acc.class_counts)]).astype(np.int64)

        precisions, recalls = [], []
        for class_i in range(class_num):
            num_gts = gts[class_i]
            if num_gts == 0:
                raise ZeroDivisionError(f"class {class_i} has no ground truth (the reference fails here too)")
            a, b = int(starts[class_i]), int(starts[class_i + 1])
            num_tp = tp_cum[a + 1:b + 1] - tp_cum[a]
            num_tpp = tpp_cum[a + 1:b + 1] - tpp_cum[a]
            num_dets = np.arange(1, b - a + 1, dtype=np.int64)
            num_fp = num_dets - num_tpp
            if precision_mode == 0:
                precision = num_tpp/num_dets
            elif precision_mode == 1:
                precision = num_tp/(num_tp + num_fp)
            elif precision_mode == 2:
                precision = num_tp/num_dets
            recall = num_tp/num_gts
            last = recall[-1] if b > a else 0.0
            precisions.append(np.append(precision, 0))
            recalls.append(np.append(recall, last))

        self.precisions = precisions
        self.recalls = recalls

    def __call__(self, recall, class_idx=0):
        if class_idx >= self.class_num:
            raise IndexError("Class index out of range")
        precisions = self.precisions[class_idx]
        recalls = self.recalls[class_idx]
        pc_idx = (recalls > recall).sum()
        if pc_idx == 0:
            precision = 0
        else:
            precision = precisions[-pc_idx:].max()
        return precision

    def plot_pr_curve(self, class_idx=-1, smooth=False, figsize=None, return_fig=False):
        """Plot PR curve (presentation only; needs matplotlib)."""
        import matplotlib.pyplot as plt
        if class_idx >= self.class_num:
            raise IndexError("Class index out of range")
        sel = slice(class_idx, class_idx + 1) if class_idx >= 0 else slice(None)
        fig = plt.figure(figsize=figsize)
        for precision, recall in zip(self.precisions[sel], self.recalls[sel]):
            if smooth:
                precision = np.maximum.accumulate(precision[::-1])[::-1]
            plt.plot(recall, precision)
        plt.legend(self.class_names[sel])
        plt.title("PR curve")
        plt.xlabel("recall")
        plt.ylabel("precision")
        plt.xlim(-0.05, 1.05)
        plt.ylim(-0.05, 1.05)
        if return_fig:
            return fig
        plt.show()

    def get_map(self, mode="voc2012"):
        """mAP table: "voc2007" (11 points), "voc2012" (7 points), "area", "smootharea"."""
        aps = [0 for _ in range(self.class_num)]

        if mode == "area" or mode == "smootharea":
            for class_i in range(self.class_num):
                precisions = self.precisions[class_i]
                if mode == "smootharea":
                    precisions = np.maximum.accumulate(precisions[::-1])[::-1]
                recalls = self.recalls[class_i]
                # sequential accumulation (cumsum) == the reference's `aps[class_i] += delta*value` loop
                delta = recalls[1:] - recalls[:-1]
                value = (precisions[1:] - precisions[:-1])/2 + precisions[:-1]
                if len(delta):
                    aps[class_i] += np.cumsum(delta*value)[-1]
        else:
            if mode == "voc2012":
                recall_list = [0, 0.14, 0.29, 0.43, 0.57, 0.71, 1]
            elif mode == "voc2007":
                recall_list = [i/10 for i in range(0, 11)]
            else:
                raise UnboundLocalError("recall_list")  # the reference's failure for an unknown mode
            for class_i in range(self.class_num):
                for rc in recall_list:
                    aps[class_i] += self(rc, class_i)
            aps = [ap/len(recall_list) for ap in aps]
        aps.append(sum(aps)/len(aps))

        ap_table = pd.DataFrame(aps)
        ap_table.columns = ["ap"]
        ap_table.index = list(self.class_names) + ["mAP"]

        return ap_table


class PR_func(PRfunc):
    def __init__(self, *args, **kwargs):
        warnings.warn(
            "`PR_func` is deprecated and renamed to `PRfunc`.",
            Warning)
        super().__init__(*args, **kwargs)
