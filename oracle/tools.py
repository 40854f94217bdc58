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


def encode_labels(boxes, box_offsets, img_size, grid_shape, class_num):
    """Box lists -> label grid (n_img, gh, gw, 5+C) float64: the ``_encode_to_array`` closure of
    ``YoloDataSequence.__getitem__`` (utils/tools.py:179-209), boxes given as pixel corners
    ``[x1, y1, x2, y2, class index]`` (what the closure reads from imgaug's BoundingBox and the
    ``labels`` list), ``img_size = img.shape[:2]`` of the resized image.

    Kept from the reference: Python-float arithmetic (``//`` and ``%`` are Python's floored
    division), boxes applied in list order so a later box overwrites x, y, w, h of a cell while
    the class bits of earlier boxes in that cell stay set, a centre at or beyond the last
    column / row is skipped (:199), negative cell indices wrap like NumPy indexing."""
    boxes = np.asarray(boxes, dtype=np.float64).reshape(-1, 5)
    off = np.asarray(box_offsets, dtype=np.int64)
    gh, gw = int(grid_shape[0]), int(grid_shape[1])
    img_height, img_width = int(img_size[0]), int(img_size[1])
    label_data = np.zeros((len(off) - 1, gh, gw, 5 + class_num))
    grid_height = img_height / gh
    grid_width = img_width / gw
    for pos in range(len(off) - 1):
        for b in range(off[pos], off[pos + 1]):
            x1, y1, x2, y2 = (float(v) for v in boxes[b, :4])
            label = int(boxes[b, 4])
            box_x = x1 + (x2 - x1) / 2
            box_y = y1 + (y2 - y1) / 2
            box_w = x2 - x1
            box_h = y2 - y1
            x_i = int(box_x // grid_width)
            y_i = int(box_y // grid_height)
            if x_i < gw and y_i < gh:
                label_data[pos, y_i, x_i, 0] = box_x % grid_width / grid_width
                label_data[pos, y_i, x_i, 1] = box_y % grid_height / grid_height
                label_data[pos, y_i, x_i, 2] = box_w / img_width
                label_data[pos, y_i, x_i, 3] = box_h / img_height
                label_data[pos, y_i, x_i, 4] = 1
                label_data[pos, y_i, x_i, 5 + label] = 1
    return label_data


def down2xlabel(label_data):
    """2x label downsample, utils/tools.py:342-367: a 2x2 block whose largest obj flag equals 1
    keeps its largest-area entry (np.argmax: first maximum in row-major order, areas in the input
    dtype) with the xy offset re-expressed in the coarser cell; every other block stays zero.
    Output float64."""
    lab = np.asarray(label_data)
    n, gh, gw, ch = lab.shape
    if gh % 2 or gw % 2:
        raise IndexError("down2xlabel: odd grid (the reference indexes out of bounds)")
    blk = lab.reshape(n, gh // 2, 2, gw // 2, 2, ch).transpose(0, 1, 3, 2, 4, 5).reshape(n, gh // 2, gw // 2, 4, ch)
    has = blk[..., 4].max(axis=-1) == 1
    pick = (blk[..., 2] * blk[..., 3]).argmax(axis=-1)
    sel = np.take_along_axis(blk, pick[..., None, None], axis=3)[..., 0, :].astype(np.float64)
    out = np.zeros((n, gh // 2, gw // 2, ch))
    sel[..., 0] = (sel[..., 0] + pick % 2) / 2
    sel[..., 1] = (sel[..., 1] + pick // 2) / 2
    out[has] = sel[has]
    return out


def encode_label_pyramid(boxes, box_offsets, img_size, grid_shape, class_num, n_levels=1, dtype=np.float64):
    """encode_labels on the finest grid followed by (n_levels-1) down2xlabel passes, returned
    coarse grid first like ``_Yolov4DataSequence.__getitem__`` (yolov4/__init__.py:47-53);
    ``dtype=np.float32`` adds the cast Keras applies to y_true before the loss sees it."""
    lab = encode_labels(boxes, box_offsets, img_size, grid_shape, class_num)
    out = [lab]
    for _ in range(n_levels - 1):
        lab = down2xlabel(lab)
        out.insert(0, lab)
    return [o.astype(dtype) for o in out]
