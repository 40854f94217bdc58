"""NumPy restatement of the reference's decode / pairwise IoU / NMS
(TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Follows /root/reference/utils/tools.py:
    decode      :370-438
    pair_iou    :630-684  (cal_iou, mode 1 = IoU, mode 2 = DIoU)
    nms         :687-733
    soft_nms    :736-786

Float semantics that bind bit-exactness are kept: the threshold test runs in
the INPUT dtype (fp32 heads -> fp32 product and fp32-rounded threshold), the
row arithmetic is float64 with one rounding per operation in the reference's
order, and NMS suppresses on ``>=``.

Documented tie rule (reference: ``np.argsort(conf)[::-1]``, whose order among
equal keys is implementation-defined): descending confidence, equal
confidences visited HIGHER ORIGINAL INDEX FIRST (what a stable sort reversed
gives, and what numpy does for short inputs).

Parity pinning: checked against the unmodified reference functions executed in
the build container (oracle/refexec.py) and tests/golden/decode_nms_*.npz.
"""
import numpy as np

EPS = 1e-07  # utils/tools.py:26


def decode(*grids, class_num=1, threshold=0.5, version=1):
    """Rows [x, y, w, h, c, class, p] (float64), one per (cell, box, class) whose
    joint confidence c*p reaches ``threshold``; scan order = argument order, then
    row-major (y, x, box, class).  No hits -> shape (0,), like the reference."""
    chunks = []
    for g in grids:
        g = np.asarray(g)
        gh, gw = g.shape[:2]
        if version == 1:
            nb = (g.shape[-1] - class_num) // 5
            geo = g[..., :-class_num].reshape(gh, gw, nb, 5)
            prob = g[..., -class_num:][:, :, None, :]
        elif version in (2, 3, 4):
            nb = g.shape[-1] // (5 + class_num)
            cell = g.reshape(gh, gw, nb, 5 + class_num)
            geo, prob = cell[..., :5], cell[..., 5:]
        else:
            raise ValueError(f"Invalid version: {version}")
        joint = geo[..., 4:5] * prob                       # input dtype
        yy, xx, bb, kk = np.nonzero(joint >= threshold)    # python float is "weak"
        if len(yy) == 0:
            continue
        sel = geo[yy, xx, bb]                              # (K,5) input dtype
        rows = np.empty((len(yy), 7), dtype=np.float64)
        rows[:, 0] = (xx + sel[:, 0]) / gw                 # int64 + fp32 -> fp64
        rows[:, 1] = (yy + sel[:, 1]) / gh
        rows[:, 2:5] = sel[:, 2:5]
        rows[:, 5] = kk
        rows[:, 6] = prob[yy, xx, 0 if version == 1 else bb, kk]
        chunks.append(rows)
    if not chunks:
        return np.array([], dtype=np.float64)
    return np.concatenate(chunks, axis=0)


def pair_iou(a, b, mode=1):
    """Broadcast IoU (mode 1) or DIoU (mode 2) of boxes [...,(x,y,w,h,..)]; ``a`` is
    the reference's xywh_true, ``b`` its xywh_pred.  tools.py:649-682."""
    a = np.asarray(a)
    b = np.asarray(b)
    ac, asz = a[..., 0:2], a[..., 2:4]
    bc, bsz = b[..., 0:2], b[..., 2:4]
    ah, bh = asz / 2.0, bsz / 2.0
    alo, ahi = ac - ah, ac + ah
    blo, bhi = bc - bh, bc + bh
    ov = np.maximum(np.minimum(bhi, ahi) - np.maximum(blo, alo), 0.0)
    inter = ov[..., 0] * ov[..., 1]
    union = bsz[..., 0] * bsz[..., 1] + asz[..., 0] * asz[..., 1] - inter
    iou = inter / (union + EPS)
    if mode == 1:
        return iou
    hull = np.maximum(bhi, ahi) - np.minimum(blo, alo)
    with np.errstate(invalid="ignore", divide="ignore"):
        diag2 = np.power(hull[..., 0], 2) + np.power(hull[..., 1], 2)
        dist2 = np.power(ac[..., 0] - bc[..., 0], 2) + np.power(ac[..., 1] - bc[..., 1], 2)
        return iou - dist2 / diag2                         # no epsilon on diag2 (:672-682)


def visit_order(conf):
    """Descending confidence, ties -> higher index first (documented rule)."""
    return np.argsort(conf, kind="stable")[::-1]


def nms_keep(rows, class_num=1, nms_threshold=0.45, iou_mode=1):
    """Boolean keep mask over ``rows`` (K,7).  Rows whose class id is outside
    [0, class_num) belong to no class and are dropped (tools.py:702-705)."""
    rows = np.asarray(rows, dtype=np.float64).reshape(-1, 7)
    keep = np.zeros(len(rows), dtype=bool)
    cls = rows[:, 5].astype("int")
    for k in range(class_num):
        idx = np.nonzero(cls == k)[0]
        if len(idx) == 0:
            continue
        sub = rows[idx]
        m = pair_iou(sub[:, None, :5], sub[None, :, :5], mode=iou_mode)
        order = visit_order(sub[:, 4] * sub[:, 6])
        seen = np.zeros(len(idx), dtype=bool)
        dead = np.zeros(len(idx), dtype=bool)
        for i in order:
            seen[i] = True
            if not dead[i]:
                dead |= (m[i] >= nms_threshold) & ~seen     # NaN compares False
        keep[idx[~dead]] = True
    return keep


def nms(rows, class_num=1, nms_threshold=0.45, iou_mode=1):
    """Survivors grouped by class 0..C-1, original order inside a class."""
    rows = np.asarray(rows, dtype=np.float64).reshape(-1, 7)
    keep = nms_keep(rows, class_num, nms_threshold, iou_mode)
    cls = rows[:, 5].astype("int")
    parts = [rows[keep & (cls == k)] for k in range(class_num)]
    return np.vstack(parts) if parts else rows[:0]


def soft_nms(rows, class_num=1, nms_threshold=0.45, conf_threshold=0.5, sigma=0.5):
    """Gaussian soft-NMS, tools.py:736-786 (every visited box decays the
    confidence of each unvisited overlapping box, even if itself deleted)."""
    rows = np.asarray(rows, dtype=np.float64).reshape(-1, 7)
    cls = rows[:, 5].astype("int")
    parts = []
    for k in range(class_num):
        sub = rows[cls == k]
        n = len(sub)
        m = pair_iou(sub[:, None, :5], sub[None, :, :5], mode=1)
        conf = sub[:, 4] * sub[:, 6]
        order = visit_order(conf)          # fixed before any decay, like the reference
        seen = np.zeros(n, dtype=bool)
        dead = np.zeros(n, dtype=bool)
        for i in order:
            seen[i] = True
            hit = np.nonzero((m[i] >= nms_threshold) & ~seen)[0]
            for j in hit:
                conf[j] *= np.exp(-1 * (m[i][j] ** 2) / sigma)
                if conf[j] < conf_threshold:
                    dead[j] = True
        parts.append(sub[~dead])
    return np.vstack(parts) if parts else rows[:0]
