"""Seeded synthetic anchor-grid inputs for the BASELINE.json configs.

There is no dataset and no network, so every test and benchmark runs on
synthetic head outputs and labels of the shapes the reference produces
(SURVEY.md section 8d).  Labels follow the reference's grid encoding
(utils/tools.py:179-209: cell-relative xy offsets, image-normalised wh,
obj flag, one-hot class) and coarser scales are derived with the same rule
as ``down2xlabel`` (utils/tools.py:342-367).

Everything here is host-side numpy; it is input generation, not the hot path.
"""
import numpy as np

# Default anchors of the reference models (w, h normalised to the image).
ANCHORS_V4 = np.array([  # yolov4/models/__init__.py:15-23
    [0.75493421, 0.65953947], [0.31578947, 0.39967105], [0.23355263, 0.18092105],
    [0.11842105, 0.24013158], [0.12500000, 0.09046053], [0.05921053, 0.12335526],
    [0.06578947, 0.04605263], [0.03125000, 0.05921053], [0.01973684, 0.02631579]])
ANCHORS_V3 = np.array([  # yolov3/__init__.py:101-109
    [0.89663461, 0.78365384], [0.37500000, 0.47596153], [0.27884615, 0.21634615],
    [0.14182692, 0.28605769], [0.14903846, 0.10817307], [0.07211538, 0.14663461],
    [0.07932692, 0.05528846], [0.03846153, 0.07211538], [0.02403846, 0.03125000]])
ANCHORS_V2 = np.array([  # yolov2/__init__.py:70-74
    [0.75157846, 0.70525231], [0.60637077, 0.27136769], [0.25680231, 0.42110308],
    [0.14418923, 0.15865615], [0.04405615, 0.05210654]])

CONFIGS = {
    # name: version, finest->coarsest listed coarse first like the reference
    "v2-416": dict(version=2, grids=[13], bbox_num=5, class_num=20, batch=8, anchors=ANCHORS_V2),
    "v3-416": dict(version=3, grids=[13, 26, 52], bbox_num=3, class_num=80, batch=64, anchors=ANCHORS_V3),
    "v4-608": dict(version=4, grids=[19, 38, 76], bbox_num=3, class_num=80, batch=128, anchors=ANCHORS_V4),
}


def halve_labels(label):
    """2x label-grid downsample, same rule as utils/tools.py:342-367: a 2x2 block
    containing an object keeps its largest-area entry (first on ties, row-major),
    with the xy offset re-expressed in the coarser cell."""
    n, gh, gw, ch = label.shape
    blk = label[:, :gh // 2 * 2, :gw // 2 * 2].reshape(n, gh // 2, 2, gw // 2, 2, ch)
    blk = blk.transpose(0, 1, 3, 2, 4, 5).reshape(n, gh // 2, gw // 2, 4, ch)
    has = blk[..., 4].max(axis=-1) == 1
    pick = (blk[..., 2] * blk[..., 3]).argmax(axis=-1)
    sel = np.take_along_axis(blk, pick[..., None, None], axis=3)[..., 0, :]
    out = np.zeros((n, gh // 2, gw // 2, ch), dtype=np.float64)
    shift = np.stack([pick % 2, pick // 2], axis=-1)
    xy = (sel[..., :2] + shift) / 2
    out[..., :2] = np.where(has[..., None], xy, 0)
    out[..., 2:] = np.where(has[..., None], sel[..., 2:], 0)
    return out


def make_labels(rng, n_img, grids, class_num, anchors, mean_boxes=8.0, dtype=np.float32):
    """Labels for every scale, coarse grid first (reference order,
    yolov4/__init__.py:518-528).  Finest grid: max(1, Poisson(mean_boxes)) boxes
    per image in distinct cells; wh = random anchor * exp(N(0, .25^2))."""
    ch = 5 + class_num
    s = grids[-1]
    fine = np.zeros((n_img, s, s, ch), dtype=np.float64)
    counts = np.maximum(1, rng.poisson(mean_boxes, n_img))
    for i in range(n_img):
        k = min(int(counts[i]), s * s)
        cells = rng.choice(s * s, size=k, replace=False)
        a = anchors[rng.integers(0, len(anchors), k)]
        wh = np.clip(a * np.exp(rng.normal(0, 0.25, (k, 2))), 1e-3, 1.0)
        cy, cx = cells // s, cells % s
        fine[i, cy, cx, 0:2] = rng.uniform(0, 1, (k, 2))
        fine[i, cy, cx, 2:4] = wh
        fine[i, cy, cx, 4] = 1
        fine[i, cy, cx, 5 + rng.integers(0, class_num, k)] = 1
    out = [fine]
    for g in reversed(grids[:-1]):
        nxt = halve_labels(out[0])
        assert nxt.shape[1] == g, (nxt.shape, g)
        out.insert(0, nxt)
    return [o.astype(dtype) for o in out]


def boxes_from_labels(fine_label, img_size):
    """Box lists (pixels of an img_size = (H, W) image) of a finest-grid label tensor: the inverse
    of the reference's grid encoding (utils/tools.py:185-209), i.e. what its file reader would
    have parsed.  Returns (boxes (n, 5) float64 [x1, y1, x2, y2, class], offsets (n_img+1,) int64),
    boxes in row-major cell order per image."""
    lab = np.asarray(fine_label, dtype=np.float64)
    n_img, gh, gw, _ = lab.shape
    H, W = float(img_size[0]), float(img_size[1])
    img, cy, cx = np.nonzero(lab[..., 4] == 1)
    sel = lab[img, cy, cx]
    bx, by = (cx + sel[:, 0]) * (W / gw), (cy + sel[:, 1]) * (H / gh)
    bw, bh = sel[:, 2] * W, sel[:, 3] * H
    boxes = np.column_stack([bx - bw / 2, by - bh / 2, bx + bw / 2, by + bh / 2,
                             np.argmax(sel[:, 5:], axis=1).astype(np.float64)])
    offsets = np.concatenate([[0], np.cumsum(np.bincount(img, minlength=n_img))]).astype(np.int64)
    return np.ascontiguousarray(boxes), offsets


def make_head_outputs(rng, y_trues, grids, bbox_num, class_num, anchors,
                      det_per_gt=30, stray_frac=0.01, dtype=np.float32):
    """Activated head outputs per scale, (N, S, S, B*(5+C)), coarse first.

    Background: xy ~ U(.02,.98); wh = anchor*exp(N(0,.35^2)) in (1e-3, 1];
    objectness ~ U(.02,.30); class scores ~ U(.02,.20) with one dominant class
    ~ U(.60,.98).  Detections: around every label box ``det_per_gt`` predicted
    boxes (random scale/anchor, neighbouring cells) get a jittered copy of the
    label geometry, objectness ~ U(.55,.98) and the label class ~ U(.90,.98),
    so decode at thr .5 yields a few hundred rows per image in overlapping
    clusters (NMS has work to do); ``stray_frac`` of the boxes are confident
    false positives.  Everything stays inside the loss's clip band.
    """
    ch = 5 + class_num
    n_img = y_trues[0].shape[0]
    n_sc = len(grids)
    preds = []
    for si, s in enumerate(grids):
        a = anchors[si * bbox_num:(si + 1) * bbox_num] if len(anchors) >= (si + 1) * bbox_num else anchors[:bbox_num]
        p = np.empty((n_img, s, s, bbox_num, ch), dtype=np.float32)
        p[..., 0:2] = rng.uniform(0.02, 0.98, (n_img, s, s, bbox_num, 2))
        p[..., 2:4] = np.clip(a[None, None, None] * np.exp(
            rng.normal(0, 0.35, (n_img, s, s, bbox_num, 2))), 1e-3, 1.0)
        p[..., 4] = rng.uniform(0.02, 0.30, (n_img, s, s, bbox_num))
        p[..., 5:] = rng.uniform(0.02, 0.20, (n_img, s, s, bbox_num, class_num))
        dom = rng.integers(0, class_num, (n_img, s, s, bbox_num))
        np.put_along_axis(p[..., 5:], dom[..., None],
                          rng.uniform(0.60, 0.98, (n_img, s, s, bbox_num, 1)).astype(np.float32), axis=-1)
        stray = rng.uniform(0, 1, (n_img, s, s, bbox_num)) < stray_frac
        p[..., 4][stray] = rng.uniform(0.55, 0.98, int(stray.sum()))
        preds.append(p)

    fine = y_trues[-1]
    sf = grids[-1]
    img, cy, cx = np.nonzero(fine[..., 4] == 1)
    for i, y, x in zip(img, cy, cx):
        lab = fine[i, y, x]
        bx, by = (x + lab[0]) / sf, (y + lab[1]) / sf
        cls = int(np.argmax(lab[5:]))
        for _ in range(det_per_gt):
            si = int(rng.integers(0, n_sc))
            s = grids[si]
            jx = float(np.clip(bx + rng.normal(0, 0.15) * lab[2], 0, 1 - 1e-6))
            jy = float(np.clip(by + rng.normal(0, 0.15) * lab[3], 0, 1 - 1e-6))
            gx, gy = int(jx * s), int(jy * s)
            b = int(rng.integers(0, bbox_num))
            q = preds[si][i, gy, gx, b]
            q[0] = np.clip(jx * s - gx, 0.02, 0.98)
            q[1] = np.clip(jy * s - gy, 0.02, 0.98)
            q[2:4] = np.clip(lab[2:4] * np.exp(rng.normal(0, 0.12, 2)), 1e-3, 1.0)
            q[4] = rng.uniform(0.55, 0.98)
            q[5:] = rng.uniform(0.02, 0.20, class_num)
            q[5 + cls] = rng.uniform(0.90, 0.98)
    return [p.reshape(n_img, s, s, bbox_num * ch).astype(dtype, copy=False)
            for p, s in zip(preds, grids)]


def make_config(name, batch=None, seed=0, rank=0, **overrides):
    """Inputs for one BASELINE config.  Seed rule: seed*1000 + rank."""
    cfg = dict(CONFIGS[name])
    cfg.update(overrides)
    if batch is not None:
        cfg["batch"] = batch
    rng = np.random.default_rng(seed * 1000 + rank)
    y_trues = make_labels(rng, cfg["batch"], cfg["grids"], cfg["class_num"], cfg["anchors"])
    y_preds = make_head_outputs(rng, y_trues, cfg["grids"], cfg["bbox_num"],
                                cfg["class_num"], cfg["anchors"])
    cfg.update(y_trues=y_trues, y_preds=y_preds, name=name)
    return cfg


def make_dense_candidates(rng, n_rows, class_num, n_clusters_per_class=20):
    """Config-4 style NMS stress rows (K,7) float64: [x, y, w, h, c, class, p];
    boxes are drawn around per-class cluster centres so suppression is real."""
    cls = rng.integers(0, class_num, n_rows)
    centres = rng.uniform(0.1, 0.9, (class_num, n_clusters_per_class, 2))
    which = rng.integers(0, n_clusters_per_class, n_rows)
    xy = centres[cls, which] + rng.normal(0, 0.03, (n_rows, 2))
    wh = np.exp(rng.uniform(np.log(0.02), np.log(0.3), (n_rows, 2)))
    c = rng.uniform(0.5, 1.0, n_rows)
    p = rng.uniform(0.5, 1.0, n_rows)
    return np.column_stack([xy, wh, c, cls.astype(np.float64), p]).astype(np.float64)


def make_kmeans_boxes(rng, n, k=9):
    """Config-5 style (n,2) float64 (w,h): k-component log-normal mixture in (0,1]."""
    comp = rng.integers(0, k, n)
    mu = np.log(np.linspace(0.03, 0.7, k))
    w = np.exp(mu[comp] + rng.normal(0, 0.25, n))
    h = np.exp(mu[comp] + rng.normal(0, 0.35, n))
    return np.clip(np.column_stack([w, h]), 1e-3, 1.0)
