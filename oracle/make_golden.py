"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE SOURCE (container only).

    python -m oracle.make_golden

The reference ships no tests or fixtures, so these are the
known-answer vectors that pin both the oracle restatements (CPU tests) and the
CUDA path (GPU tests).  Inputs are small, seeded and stored alongside the
reference's outputs so the fixtures are self-contained on the GPU box, where
/root/reference does not exist.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refexec  # noqa: E402
from tf2_yolo_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def small_grid_case(rng, version, n_img, grids, B, C, anchors):
    y_trues = synth.make_labels(rng, n_img, grids, C, anchors, mean_boxes=3.0)
    if version == 1:
        s = grids[0]
        yp = rng.uniform(0.02, 0.98, (n_img, s, s, B * 5 + C)).astype(np.float32)
        # make responsible boxes plausible: copy label geometry with jitter into box 0
        obj = y_trues[0][..., 4] == 1
        yp[obj, 0:4] = np.clip(y_trues[0][obj, 0:4] * np.exp(rng.normal(0, 0.2, (int(obj.sum()), 4))), 0.02, 0.98)
        return y_trues, [yp]
    y_preds = synth.make_head_outputs(rng, y_trues, grids, B, C, anchors, det_per_gt=6, stray_frac=0.02)
    return y_trues, y_preds


def gen_losses():
    cases = []
    rng = np.random.default_rng(101)
    anc4 = synth.ANCHORS_V4
    # (name, version, grids, B, C, anchors, list of kwargs variants)
    plan = [
        ("v4", 4, [4, 8], 3, 6, anc4[3:9], [
            dict(loss_weight=[1, 5, 1]),
            dict(loss_weight=[2, 3, 0.5], binary_weight=np.array([0.4]), wh_reg_weight=0.05,
                 ignore_thresh=0.5, truth_thresh=0.7, label_smooth=0.05, focal_loss_gamma=1.5),
            dict(loss_weight=[1, 1, 1], focal_loss_gamma=0, anchors=None),
        ]),
        ("v3", 3, [4, 8], 3, 6, synth.ANCHORS_V3[3:9], [
            dict(loss_weight=[1, 1, 5, 1]),
            dict(loss_weight=[1, 2, 3, 4], use_focal_loss=True, focal_loss_gamma=2, use_scale=False,
                 binary_weight=0.5, ignore_thresh=0.4),
            dict(loss_weight=[1, 1, 1, 1], use_focal_loss=True, focal_loss_gamma=1.5),
        ]),
        ("v2", 2, [6], 5, 4, synth.ANCHORS_V2, [
            dict(loss_weight=[1, 1, 5, 1]),
            dict(loss_weight=[2, 1, 1, 3], binary_weight=0.3, ignore_thresh=0.3),
        ]),
        ("v1", 1, [7], 2, 5, None, [
            dict(loss_weight=[5, 5, 1, 1], binary_weight=0.5),
            dict(loss_weight=[1, 1, 1, 1]),
        ]),
    ]
    for name, ver, grids, B, C, anchors, variants in plan:
        y_trues, y_preds = small_grid_case(rng, ver, 3, grids, B, C, anchors if anchors is not None else anc4)
        for vi, kw in enumerate(variants):
            for si, s in enumerate(grids):
                k = dict(kw)
                if ver != 1:
                    if "anchors" not in k:
                        k["anchors"] = np.asarray(anchors[si * B:(si + 1) * B]) if len(anchors) >= (si + 1) * B else np.asarray(anchors[:B])
                    elif k["anchors"] is None and ver == 2:
                        continue
                loss, grad = refexec.reference_loss(ver, y_trues[si], y_preds[si], grid_shape=(s, s),
                                                    bbox_num=B, class_num=C, **k)
                js = {kk: (np.asarray(vv).tolist() if vv is not None else None) for kk, vv in k.items()}
                js["binary_weight_is_array"] = isinstance(k.get("binary_weight"), np.ndarray)
                cases.append(dict(name=f"{name}_var{vi}_s{s}", version=ver, grid=s, B=B, C=C,
                                  kwargs=json.dumps(js), y_true=y_trues[si], y_pred=y_preds[si],
                                  loss=np.asarray(loss, dtype=np.float64).reshape(-1), grad=grad))
    pack = {}
    names = []
    for c in cases:
        n = c["name"]
        names.append(n)
        pack[n + "/meta"] = np.array(json.dumps(dict(version=c["version"], grid=c["grid"], B=c["B"], C=c["C"],
                                                     kwargs=json.loads(c["kwargs"]))))
        pack[n + "/y_true"] = c["y_true"]
        pack[n + "/y_pred"] = c["y_pred"]
        pack[n + "/loss"] = c["loss"]
        pack[n + "/grad"] = c["grad"].astype(np.float64)
    pack["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "loss.npz"), **pack)
    print("loss.npz:", len(names), "cases")


def gen_decode_nms():
    tools, _, _ = refexec.load_numpy_half()
    rng = np.random.default_rng(202)
    pack = {}
    # v3/v4-style: 3 images, 2 scales, C=5
    grids, B, C = [5, 10], 3, 5
    y_trues = synth.make_labels(rng, 3, grids, C, synth.ANCHORS_V4[3:9], mean_boxes=4.0)
    y_preds = synth.make_head_outputs(rng, y_trues, grids, B, C, synth.ANCHORS_V4[3:9], det_per_gt=25,
                                      stray_frac=0.05)
    for si in range(2):
        pack[f"m/pred{si}"] = y_preds[si]
    pack["m/y_true"] = y_trues[-1].astype(np.float64)
    for thr in (0.5, 0.3):
        for i in range(3):
            rows = tools.decode(*[p[i] for p in y_preds], class_num=C, threshold=thr, version=4)
            pack[f"m/rows_t{thr}_i{i}"] = rows
            if len(rows):
                pack[f"m/nms1_t{thr}_i{i}"] = tools.nms(rows, C, 0.45, 1)
                pack[f"m/nms2_t{thr}_i{i}"] = tools.nms(rows, C, 0.45, 2)
                pack[f"m/soft_t{thr}_i{i}"] = tools.soft_nms(rows.copy(), C, 0.45, thr, 0.5)
    for i in range(3):
        pack[f"m/gt_i{i}"] = tools.decode(y_trues[-1][i].astype(np.float64), class_num=C, version=4)
    # v2 single scale fp32 and v1 layout
    c2 = synth.make_config("v2-416", batch=2, seed=9)
    pack["v2/pred"] = c2["y_preds"][0]
    for i in range(2):
        rows = tools.decode(c2["y_preds"][0][i], class_num=20, threshold=0.4, version=2)
        pack[f"v2/rows_i{i}"] = rows
        pack[f"v2/nms1_i{i}"] = tools.nms(rows, 20, 0.45, 1)
    g1 = rng.uniform(0, 1, (2, 7, 7, 2 * 5 + 6)).astype(np.float32)
    pack["v1/pred"] = g1
    for i in range(2):
        pack[f"v1/rows_i{i}"] = tools.decode(g1[i], class_num=6, threshold=0.5, version=1)
    # dense candidates, ties and degenerate boxes
    rows = synth.make_dense_candidates(rng, 1500, 4)
    rows[100:110] = rows[90:100]            # exact duplicates -> confidence ties, IoU == 1
    rows[200:204, 2:4] = 0.0                # zero-size boxes (DIoU 0/0 with duplicates below)
    rows[204:208] = rows[200:204]
    pack["dense/rows"] = rows
    with np.errstate(invalid="ignore", divide="ignore"):
        pack["dense/nms1"] = tools.nms(rows, 4, 0.45, 1)
        pack["dense/nms2"] = tools.nms(rows, 4, 0.45, 2)
    a = rng.uniform(0.1, 0.9, (7, 5))
    b = rng.uniform(0.1, 0.9, (9, 5))
    pack["iou/a"], pack["iou/b"] = a, b
    pack["iou/m1"] = tools.cal_iou(a[:, None], b[None], mode=1)
    pack["iou/m2"] = tools.cal_iou(a[:, None], b[None], mode=2)
    np.savez_compressed(os.path.join(OUT, "decode_nms.npz"), **pack)
    print("decode_nms.npz:", len(pack), "arrays")


def gen_kmeans():
    _, km, _ = refexec.load_numpy_half()
    rng = np.random.default_rng(303)
    data = synth.make_kmeans_boxes(rng, 4000, k=5)
    pack = {"data": data}
    for name, fn, k, stop in (("iou", km.iou_dist, 5, 1e-5), ("euclid", km.euclidean_dist, 3, 1e-4)):
        np.random.seed(12)
        pack[f"{name}/centers"] = km.kmeans(data, k, fn, stop, verbose=False)
        pack[f"{name}/k"] = np.array(k)
        pack[f"{name}/stop"] = np.array(stop)
        c0 = rng.uniform(0.05, 0.6, (k, 2))
        d = fn(c0[:, None, :], data[None, :, :])
        pack[f"{name}/c0"] = c0
        pack[f"{name}/assign0"] = np.argmin(d, axis=0).astype(np.int32)
    # empty-cluster re-draw path: a centre far from every box
    np.random.seed(5)
    tiny = data[:50]
    pack["empty/data"] = tiny
    pack["empty/centers"] = km.kmeans(tiny, 9, km.iou_dist, 1e-5, verbose=False)
    np.savez_compressed(os.path.join(OUT, "kmeans.npz"), **pack)
    print("kmeans.npz ok")


def gen_map():
    _, _, meas = refexec.load_numpy_half()
    rng = np.random.default_rng(404)
    grids, B, C = [4, 8], 3, 3
    n_img = 12
    y_trues = synth.make_labels(rng, n_img, grids, C, synth.ANCHORS_V4[3:9], mean_boxes=4.0)
    y_preds = synth.make_head_outputs(rng, y_trues, grids, B, C, synth.ANCHORS_V4[3:9], det_per_gt=12,
                                      stray_frac=0.08)
    names = [f"c{i}" for i in range(C)]
    pack = {"y_true": y_trues[-1].astype(np.float64), "pred0": y_preds[0], "pred1": y_preds[1],
            "class_num": np.array(C)}
    variants = {
        "a": dict(conf_threshold=0.05, nms_mode=1, nms_threshold=0.5, iou_threshold=0.5, precision_mode=2, max_per_img=100),
        "b": dict(conf_threshold=0.3, nms_mode=3, nms_threshold=0.45, iou_threshold=0.4, precision_mode=1, max_per_img=3),
        "c": dict(conf_threshold=0.3, nms_mode=0, iou_threshold=0.5, precision_mode=0, max_per_img=None),
    }
    for vn, kw in variants.items():
        pr = meas.PRfunc(y_trues[-1].astype(np.float64), *y_preds, class_names=names, version=4, **kw)
        pack[f"{vn}/kwargs"] = np.array(json.dumps(kw))
        for k in range(C):
            pack[f"{vn}/precisions{k}"] = pr.precisions[k]
            pack[f"{vn}/recalls{k}"] = pr.recalls[k]
        for mode in ("voc2007", "voc2012", "area", "smootharea"):
            pack[f"{vn}/ap_{mode}"] = pr.get_map(mode)["ap"].to_numpy(dtype=np.float64)
        pack[f"{vn}/call"] = np.array([[pr(r, k) for r in (0.0, 0.3, 0.6, 0.9)] for k in range(C)], dtype=np.float64)
    for vn, kw in {"s0": dict(conf_threshold=0.5, nms_mode=1, precision_mode=2),
                   "s1": dict(conf_threshold=0.3, nms_mode=3, nms_threshold=0.45, precision_mode=1),
                   "s2": dict(conf_threshold=0.3, nms_mode=0, precision_mode=0)}.items():
        tab = meas.create_score_mat(y_trues[-1].astype(np.float64), *y_preds, class_names=names, version=4, **kw)
        pack[f"{vn}/kwargs"] = np.array(json.dumps(kw))
        pack[f"{vn}/table"] = tab.to_numpy(dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "map.npz"), **pack)
    print("map.npz ok")


def gen_metrics():
    rng = np.random.default_rng(505)
    pack = {}
    names = []
    plan = [("v4", 4, [4, 8], 3, 6, synth.ANCHORS_V4[3:9]), ("v3", 3, [4, 8], 3, 6, synth.ANCHORS_V3[3:9]),
            ("v2", 2, [6], 5, 4, synth.ANCHORS_V2), ("v1", 1, [7], 2, 5, None)]
    for name, ver, grids, B, C, anchors in plan:
        y_trues, y_preds = small_grid_case(rng, ver, 4, grids, B, C, anchors if anchors is not None else synth.ANCHORS_V4)
        if ver == 1:   # make some class predictions right
            obj = y_trues[0][..., 4] == 1
            y_preds[0][obj, -C:] = 0.5 * y_preds[0][obj, -C:] + 0.5 * y_trues[0][obj, 5:]
        for si, s in enumerate(grids):
            for thr in (0.5, 0.25):
                n = f"{name}_s{s}_t{thr}"
                names.append(n)
                pack[n + "/meta"] = np.array(json.dumps(dict(version=ver, grid=s, B=B, C=C, thr=thr)))
                pack[n + "/y_true"] = y_trues[si]
                pack[n + "/y_pred"] = y_preds[si]
                pack[n + "/metrics"] = refexec.reference_metrics(ver, y_trues[si], y_preds[si], (s, s), B, C, thr)
    pack["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **pack)
    print("metrics.npz:", len(names), "cases")


def gen_labels():
    """down2xlabel / get_class_weight fixtures (utils/tools.py:342-367, :592-627)."""
    tools, _, _ = refexec.load_numpy_half()
    rng = np.random.default_rng(606)
    lab = synth.make_labels(rng, 3, [12], 5, synth.ANCHORS_V4, mean_boxes=14.0, dtype=np.float64)[0]
    pack = {"lab": lab, "down64": tools.down2xlabel(lab), "down32": tools.down2xlabel(lab.astype(np.float32))}
    for m in ("alpha", "log", "effective", "binary"):
        pack["cw_" + m] = tools.get_class_weight(lab[..., 5:], method=m)
    np.savez_compressed(os.path.join(OUT, "labels.npz"), **pack)
    print("labels.npz ok")


def encode_case(rng, n_img, size, grid, C, mean_boxes, degenerate=True):
    """Box lists in pixels of a (H, W) image: random boxes, several per cell, some centred outside."""
    H, W = size
    counts = rng.poisson(mean_boxes, n_img)
    counts[rng.integers(0, n_img)] = 0                     # an image without boxes
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    nb = int(off[-1])
    cx, cy = rng.uniform(0, W, nb), rng.uniform(0, H, nb)
    w, h = rng.uniform(1, 0.5 * W, nb), rng.uniform(1, 0.5 * H, nb)
    boxes = np.column_stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2,
                             rng.integers(0, C, nb).astype(np.float64)])
    if degenerate and nb >= 12:
        gh, gw = grid
        boxes[1, :4] = boxes[0, :4] + [0.25, 0.25, 0.5, 0.5]          # same cell, later box wins, both classes stay
        boxes[1, 4] = (boxes[0, 4] + 1) % C
        boxes[2, :4] = [W + 3.0, 10.0, W + 9.0, 20.0]                # centre beyond the last column: skipped
        boxes[3, :4] = [10.0, H, 20.0, H + 2.0]                      # centre beyond the last row: skipped
        boxes[4, :4] = [-9.0, 5.0, -1.0, 15.0]                       # negative centre: column index wraps
        boxes[5, :4] = [W / gw * 3, H / gh * 2, W / gw * 3, H / gh * 2]   # zero size, exactly on a cell corner
        boxes[6, :4] = [30.0, 40.0, 20.0, 60.0]                      # x2 < x1: negative width
        boxes[7, :4] = [W / gw * 2 - 4.0, 8.0, W / gw * 2 + 4.0, 12.0]    # centre exactly on a column boundary
        boxes[8, :4] = boxes[5, :4] + [W / gw, 0.0, W / gw, 0.0]     # zero-area neighbour in the same 2x2 block
    return boxes, off


def gen_encode():
    """Label-grid encoding fixtures: the UNMODIFIED YoloDataSequence.__getitem__ (utils/tools.py:176-339)
    reading labelme files, then down2xlabel (utils/tools.py:342-367) per extra level."""
    tools, _, _ = refexec.load_numpy_half()
    rng = np.random.default_rng(707)
    pack = {}
    names = []
    for name, n_img, size, grid, C, mean, levels in (
            ("small", 6, (100, 90), (8, 12), 5, 9.0, 3),
            ("dense", 3, (64, 64), (8, 8), 3, 60.0, 4),
            ("v4", 2, (608, 608), (76, 76), 80, 12.0, 3),
            ("v2", 3, (416, 416), (13, 13), 20, 5.0, 1)):
        boxes, off = encode_case(rng, n_img, size, grid, C, mean)
        lab = refexec.reference_encode_labels(boxes, off, size, grid, C)
        names.append(name)
        pack[name + "/boxes"], pack[name + "/offsets"] = boxes, off
        pack[name + "/meta"] = np.array(json.dumps(dict(size=size, grid=grid, C=C, levels=levels)))
        pack[name + "/level0"] = lab
        for l in range(1, levels):
            lab = tools.down2xlabel(lab)
            pack[name + f"/level{l}"] = lab
    pack["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "encode.npz"), **pack)
    print("encode.npz:", names)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1:
        for name in sys.argv[1:]:
            globals()["gen_" + name]()
        sys.exit(0)
    gen_losses()
    gen_decode_nms()
    gen_kmeans()
    gen_map()
    gen_metrics()
    gen_labels()
    gen_encode()
