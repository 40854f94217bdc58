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
    __slots__ = ("conf", "gid", "flag", "cls", "class_counts", "gts", "score", "owned")


class _Phase1:
    """Per-rank accumulation over chunks of images (utils/measurement.py:210-292, batched): decode
    ground truth and predictions, NMS, per-class IoU matching, per-image top-k; the records
    (confidence, ground-truth id, true-positive flag, class) of every chunk are appended to
    rank-level arrays ON THE DEVICE.  After the first chunk (which measures the row density) no
    chunk makes the host wait: row capacities come from the measured density and every
    "did it fit" question is answered once, in ``finish``."""

    def __init__(self, class_num, conf_threshold, nms_mode, nms_threshold, iou_threshold, max_per_img, version,
                 nms_sigma, dev, gt_base):
        self.C, self.conf_thr, self.nms_mode, self.iou_thr = class_num, conf_threshold, nms_mode, iou_threshold
        self.thr, self.iou_mode = _nms_args(nms_mode, nms_threshold)
        self.max_per_img, self.version, self.sigma, self.dev = max_per_img, version, nms_sigma, dev
        self.score = torch.zeros(3 * class_num, dtype=torch.int64, device=dev)
        self.gt_base = gt_base.clone()
        self.class_counts = torch.zeros(class_num, dtype=torch.int64, device=dev)
        self.total = torch.zeros(1, dtype=torch.int64, device=dev)
        self.counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self.dst = None
        self.coffs = []            # per chunk: class offsets of its records (device, C+1)
        self.checks = []           # (device scalar, capacity, what): verified in finish()
        self.det_density = None    # decode rows per image of the first chunk
        self.n_img = 0

    def _grow(self, need):
        cap = 0 if self.dst is None else self.dst[0].shape[0]
        if need <= cap:
            return
        new_cap = max(need, 2 * cap, 1 << 16)
        dev = self.dev
        new = (torch.empty(new_cap, dtype=torch.float64, device=dev), torch.empty(new_cap, dtype=torch.int64, device=dev),
               torch.empty(new_cap, dtype=torch.uint8, device=dev), torch.empty(new_cap, dtype=torch.int32, device=dev))
        if self.dst is not None:   # the prefix written so far (all of it: capacity never shrinks)
            for d, o in zip(new, self.dst):
                d[:cap].copy_(o)
        self.dst = new

    def add(self, yt, preds):
        """One chunk: ``yt`` (n, gh, gw, 5+C) and ``preds`` (list of head tensors), on the device."""
        C, dev = self.C, self.dev
        n = yt.shape[0]
        cells = yt.shape[1] * yt.shape[2]
        # ground truth: at most one box per cell and class bit; labels are one-hot -> cells bound the rows
        gt_cap = n * cells
        gt_rows, gt_off = engine.decode_batch([yt], C, 0.5, self.version, capacity=gt_cap)
        self.checks.append((gt_off[-1:], gt_cap, "ground-truth rows"))
        if self.det_density is None:       # first chunk: exact (one sync), measures the density
            det_rows, det_off = engine.decode_batch_exact(preds, C, self.conf_thr, self.version)
            self.det_density = max(det_rows.shape[0] / max(n, 1), 1.0)
            if det_rows.shape[0] == 0:
                det_rows = torch.empty((1, 7), dtype=torch.float64, device=dev)
        else:
            det_cap = int(2.0 * self.det_density * n) + 4096
            det_rows, det_off = engine.decode_batch(preds, C, self.conf_thr, self.version, capacity=det_cap)
            self.checks.append((det_off[-1:], det_cap, "decode rows"))
        res = engine.nms_batch(det_rows, det_off, C, self.thr, self.iou_mode, want_seg_offsets=True,
                               soft=(self.conf_thr, self.sigma) if self.nms_mode == 2 else None)
        dets = res["out_rows"]              # capacity-sized; every consumer reads the counts on the device
        best_iou, best_gt, counts = engine.map_match(gt_rows, gt_off, dets, res["out_offsets"], C)
        conf, gid, flag, cls, coff = engine.map_accumulate(
            dets, res["seg_offsets"], best_iou, best_gt, counts, C, self.iou_thr, self.max_per_img, self.gt_base,
            self.score)
        # records of this chunk: at most max_per_img per (image, class), never more than the survivors
        bound = dets.shape[0] if not self.max_per_img else min(dets.shape[0], n * C * int(self.max_per_img))
        self.n_img += n
        self.upper = getattr(self, "upper", 0) + bound
        self._grow(self.upper)
        engine.map_append((conf, gid, flag, cls), coff[C:C + 1], self.dst, self.total, self.counter)
        self.coffs.append(coff)
        self.class_counts += coff[1:] - coff[:-1]
        self.gt_base = self.gt_base + counts.sum(dim=0, dtype=torch.int64)

    def finish(self):
        """One sync: capacities respected? -> (conf, gid, flag, cls) of this rank, class counts."""
        if self.checks:
            got = torch.cat([c[0] for c in self.checks]).cpu().numpy()
            for v, (_, cap, what) in zip(got, self.checks):
                if v > cap:
                    raise _CapacityExceeded(f"{what}: {int(v)} > capacity {cap}")
        n = int(self.total.item())
        if self.dst is None:
            self._grow(1)
        return tuple(t[:n] for t in self.dst)

    def class_segments(self):
        """(n_chunks, C+1) int64 ndarray: where the records of class c of chunk k sit in this rank's
        arrays - [seg[k, c], seg[k, c+1]) (every chunk's block is class-major)."""
        if not self.coffs:
            return np.zeros((0, self.C + 1), dtype=np.int64)
        co = torch.stack(self.coffs).cpu().numpy()
        base = np.concatenate([[0], np.cumsum(co[:, -1])[:-1]])
        return co + base[:, None]


class _CapacityExceeded(RuntimeError):
    pass


def _chunks_of(y_trues, y_preds, dev):
    """Device chunks of whole arrays (host or device), sized to _CHUNK_BYTES of head tensors."""
    n_img = _n_images(y_trues)
    per_img = sum(int(np.prod(p.shape[1:])) * 4 for p in y_preds) + int(np.prod(y_trues.shape[1:])) * 8
    chunk = max(1, min(n_img, _CHUNK_BYTES // max(per_img, 1)))
    for s in range(0, n_img, chunk):
        sl = slice(s, min(n_img, s + chunk))
        yield _to_dev(y_trues, sl, dev, True), [_to_dev(p, sl, dev, False) for p in y_preds]


def _local_gt_counts(chunks, C, version, dev):
    local = torch.zeros(C, dtype=torch.int64, device=dev)
    for yt, _ in chunks:
        rows, off = engine.decode_batch_exact([yt], C, 0.5, version)
        if rows.shape[0]:
            local += torch.bincount(rows[:, 5].long(), minlength=C)[:C]
    return local


def _tick(timings, key, t0):
    """Phase timer (only when the caller asked for timings: it synchronises the device)."""
    if timings is None:
        return t0
    import time
    torch.cuda.synchronize()
    now = time.perf_counter()
    timings[key] = timings.get(key, 0.0) + (now - t0)
    return now


def _accumulate(y_trues, y_preds, class_num, conf_threshold, nms_mode, nms_threshold, iou_threshold,
                max_per_img, version, process_group=None, nms_sigma=0.5, chunk_source=None, partition=False,
                timings=None):
    """Phase 1 on this rank's images, then the exchange of SURVEY.md 8(e): per-class ground-truth
    counts are all-gathered (gt ids need the counts of earlier ranks: rank order = image order) and
    the records travel to the rank that OWNS their class (class % world: all-to-all, not an
    all-gather of everything), so the sort + scan of phase 2 runs on 1/world of the records per rank.
    ``chunk_source``: a callable returning an iterator of (y_true chunk, [pred chunks]) on the
    device, for inputs that never exist as whole arrays (it is called twice when sharded)."""
    if not torch.cuda.is_available():
        raise YoloB200Error("no CUDA device: tf2_yolo_b200 has no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device())
    if chunk_source is None:
        if len(y_preds) == 0:
            raise ValueError("at least one prediction array is needed")
        chunk_source = lambda: _chunks_of(y_trues, y_preds, dev)   # noqa: E731
    C = class_num
    import time
    t0 = time.perf_counter() if timings is not None else 0.0
    world, rank = dist_util.world(process_group) if process_group is not None else (1, 0)
    gt_base = torch.zeros(C, dtype=torch.int64, device=dev)
    gts_total = None
    if process_group is not None:
        local = _local_gt_counts(chunk_source(), C, version, dev)
        before, gts_total = dist_util.rank_offsets(local, process_group)
        gt_base = gt_base + before
        t0 = _tick(timings, "gt_count_pass_and_allgather_s", t0)
    for attempt in range(2):
        ph = _Phase1(C, conf_threshold, nms_mode, nms_threshold, iou_threshold, max_per_img, version, nms_sigma, dev,
                     gt_base)
        if attempt == 1:
            ph.det_density = float("inf")      # capacities from exact counts (one sync per chunk)
        try:
            for yt, preds in chunk_source():
                if attempt == 1:
                    ph.det_density = None
                ph.add(yt, preds)
            conf, gid, flag, cls = ph.finish()
            break
        except _CapacityExceeded:
            if attempt == 1:
                raise
    t0 = _tick(timings, "phase1_decode_nms_match_s", t0)
    out = _Accumulated()
    out.score = ph.score
    out.owned = None
    class_counts = ph.class_counts
    if process_group is None:
        out.gts = ph.gt_base.cpu().numpy()
    else:
        out.gts = gts_total.cpu().numpy()
        dist_util.allreduce_sum(out.score, process_group)
        if partition:
            conf, gid, flag, cls, class_counts = dist_util.route_records_by_class(
                (conf, gid, flag, cls), ph.class_segments(), C, process_group)
            out.owned = [c for c in range(C) if c % world == rank]
        else:
            conf = dist_util.gather_varlen(conf, process_group)
            gid = dist_util.gather_varlen(gid, process_group)
            flag = dist_util.gather_varlen(flag, process_group)
            cls = dist_util.gather_varlen(cls, process_group)
            class_counts = dist_util.allreduce_sum(class_counts, process_group)
    out.conf, out.gid, out.flag, out.cls = conf, gid, flag, cls
    out.class_counts = class_counts.cpu().numpy()
    _tick(timings, "exchange_s", t0)
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
                 process_group=None,
                 partition_classes=False,
                 chunk_source=None,
                 timings=None):
        """Keywords as the reference (utils/measurement.py:198-208).  Extensions for sharded / very
        large evaluations: ``process_group`` (the arrays are this rank's images, rank order = image
        order), ``partition_classes`` (phase 2 per class owner, class % world: ``precisions[c]`` /
        ``recalls[c]`` exist on the owner only, ``get_map`` is the same on every rank),
        ``chunk_source`` (callable -> iterator of device chunks ``(y_true, [preds])`` instead of
        whole arrays; pass ``None`` for ``y_trues``), ``timings`` (a dict that receives the seconds per
        phase; asking for it synchronises the device between phases)."""
        class_num = len(class_names)
        self.class_num = class_num
        self.class_names = class_names
        self._group = process_group if partition_classes else None

        acc = _accumulate(y_trues, y_preds, class_num, conf_threshold, nms_mode, nms_threshold,
                          iou_threshold, max_per_img, version, process_group, nms_sigma,
                          chunk_source=chunk_source, partition=partition_classes and process_group is not None,
                          timings=timings)
        import time
        t0 = time.perf_counter() if timings is not None else 0.0
        self.owned = acc.owned          # None: every class lives here
        gts = [int(g) for g in acc.gts]
        dev = acc.conf.device
        table = np.concatenate([[0], np.cumsum(acc.gts)]).astype(np.int64)
        order, tp_cum, tpp_cum = engine.pr_curve(acc.conf, acc.cls, acc.gid, acc.flag,
                                                 torch.from_numpy(table).to(dev), int(table[-1]))
        t0 = _tick(timings, "phase2_sort_scan_s", t0)
        if precision_mode not in (0, 1, 2):
            raise UnboundLocalError("precision")      # the reference's failure for an unknown mode
        starts = np.concatenate([[0], np.cumsum(acc.class_counts)]).astype(np.int64)
        for class_i in range(class_num):
            if (self.owned is None or class_i in self.owned) and gts[class_i] == 0:
                raise ZeroDivisionError(f"class {class_i} has no ground truth (the reference fails here too)")
        # precision / recall of every prefix on the device (int64 / int64 true divisions as in
        # measurement.py:311-319), one copy back per array
        prec_d, rec_d = engine.pr_points(tp_cum, tpp_cum, torch.from_numpy(starts).to(dev),
                                         torch.from_numpy(np.asarray(acc.gts, dtype=np.int64)).to(dev), precision_mode)
        prec_h, rec_h = prec_d.cpu().numpy(), rec_d.cpu().numpy()
        precisions, recalls = [], []
        for class_i in range(class_num):
            if self.owned is not None and class_i not in self.owned:
                precisions.append(None)
                recalls.append(None)
                continue
            a, b = int(starts[class_i]), int(starts[class_i + 1])
            last = rec_h[b - 1] if b > a else 0.0
            precisions.append(np.append(prec_h[a:b], 0))
            recalls.append(np.append(rec_h[a:b], last))

        self.precisions = precisions
        self.recalls = recalls
        _tick(timings, "host_curves_s", t0)

    def __call__(self, recall, class_idx=0):
        if class_idx >= self.class_num:
            raise IndexError("Class index out of range")
        precisions = self.precisions[class_idx]
        recalls = self.recalls[class_idx]
        if precisions is None:
            raise KeyError(f"class {class_idx} is owned by another rank (partition_classes=True)")
        # the reference's `(recalls > recall).sum()` / `precisions[-pc_idx:].max()` (measurement.py:333-337):
        # recall is a running count over a fixed total, i.e. non-decreasing, so the count is a binary
        # search and the maximum a lookup in the suffix maxima - same numbers, without two passes over
        # millions of points per query (get_map asks 7 or 11 times per class)
        fast = self._fast_lookup(class_idx, precisions, recalls)
        if fast is None:
            pc_idx = (recalls > recall).sum()
            return 0 if pc_idx == 0 else precisions[-pc_idx:].max()
        pc_idx = len(recalls) - int(np.searchsorted(recalls, recall, side="right"))
        if pc_idx == 0:
            return 0
        return fast[len(precisions) - pc_idx]

    def _fast_lookup(self, class_idx, precisions, recalls):
        """Suffix maxima of the class's precisions, or None when its recalls are not sorted."""
        cache = self.__dict__.setdefault("_suffix_max", {})
        if class_idx not in cache:
            ok = (len(recalls) == len(precisions) and len(recalls) > 0 and not np.isnan(precisions).any()
                  and bool(np.all(recalls[1:] >= recalls[:-1])))
            cache[class_idx] = np.maximum.accumulate(precisions[::-1])[::-1] if ok else None
        return cache[class_idx]

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
        mine = range(self.class_num) if self.owned is None else self.owned

        if mode == "area" or mode == "smootharea":
            for class_i in mine:
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
            for class_i in mine:
                for rc in recall_list:
                    aps[class_i] += self(rc, class_i)
            aps = [ap/len(recall_list) for ap in aps]
        if self.owned is not None:      # every class has exactly one owner: the sum is a gather
            t = torch.tensor([float(a) for a in aps], dtype=torch.float64, device="cuda")
            dist_util.allreduce_sum(t, self._group)
            # numpy scalars, like the unsharded path: Python 3.12's sum() is compensated for exact
            # floats only, so the final mean would otherwise differ from the reference's by an ulp
            aps = [np.float64(a) for a in t.cpu().numpy()]
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
