"""GPU parity (bit-exact): decode rows, NMS / DIoU-NMS keep sets and output order."""
import numpy as np
import pytest
import torch

from oracle import tools as ot
from tf2_yolo_b200 import engine, synth
from tf2_yolo_b200.utils import tools

pytestmark = pytest.mark.gpu


def test_reference_fixtures(golden):
    z = golden("decode_nms")
    preds = [z["m/pred0"], z["m/pred1"]]
    for thr in (0.5, 0.3):
        for i in range(3):
            rows = tools.decode(*[p[i] for p in preds], class_num=5, threshold=thr, version=4)
            assert np.array_equal(rows, z[f"m/rows_t{thr}_i{i}"])
            assert np.array_equal(tools.nms(rows, 5, 0.45, 1), z[f"m/nms1_t{thr}_i{i}"])
            assert np.array_equal(tools.nms(rows, 5, 0.45, 2), z[f"m/nms2_t{thr}_i{i}"])
    for i in range(3):   # float64 labels decode bit-exactly too
        assert np.array_equal(tools.decode(z["m/y_true"][i], class_num=5, version=4), z[f"m/gt_i{i}"])
    for i in range(2):
        rows = tools.decode(z["v2/pred"][i], class_num=20, threshold=0.4, version=2)
        assert np.array_equal(rows, z[f"v2/rows_i{i}"])
        assert np.array_equal(tools.nms(rows, 20, 0.45, 1), z[f"v2/nms1_i{i}"])
        assert np.array_equal(tools.decode(z["v1/pred"][i], class_num=6, threshold=0.5, version=1), z[f"v1/rows_i{i}"])
    rows = z["dense/rows"]      # ties, duplicates, zero-size boxes, segments > 32 (CTA path)
    assert np.array_equal(tools.nms(rows, 4, 0.45, 1), z["dense/nms1"])
    assert np.array_equal(tools.nms(rows, 4, 0.45, 2), z["dense/nms2"])
    a, b = z["iou/a"], z["iou/b"]
    assert np.array_equal(tools.cal_iou(a[:, None], b[None], mode=1), z["iou/m1"])
    assert np.array_equal(tools.cal_iou(a[:, None], b[None], mode=2), z["iou/m2"])
    assert np.array_equal(tools.cal_iou(a[:7], b[:7], mode=2), np.diagonal(z["iou/m2"][:, :7]))


def test_batched_pipeline_vs_oracle():
    """Whole batch, three scales, decode -> DIoU-NMS chained on the device without a host sync."""
    cfg = synth.make_config("v4-608", batch=6, seed=31)
    C = cfg["class_num"]
    dev_preds = [torch.from_numpy(p).cuda() for p in cfg["y_preds"]]
    rows, offsets = engine.decode_batch(dev_preds, C, 0.5, 4, capacity=6 * 2048)
    res = engine.nms_batch(rows, offsets, C, 0.45, 2, want_seg_offsets=True)
    offs = offsets.cpu().numpy()
    rows_h = rows.cpu().numpy()
    keep_h = res["keep"].cpu().numpy().astype(bool)
    out_h = res["out_rows"].cpu().numpy()
    oo = res["out_offsets"].cpu().numpy()
    seg = res["seg_offsets"].cpu().numpy()
    assert offs[-1] <= rows.shape[0]
    for i in range(6):
        ref_rows = ot.decode(*[p[i] for p in cfg["y_preds"]], class_num=C, threshold=0.5, version=4)
        got = rows_h[offs[i]:offs[i + 1]]
        assert np.array_equal(got, ref_rows.reshape(-1, 7)), i
        ref_keep = ot.nms_keep(ref_rows, C, 0.45, 2)
        assert np.array_equal(keep_h[offs[i]:offs[i + 1]], ref_keep), i
        assert np.array_equal(out_h[oo[i]:oo[i + 1]], ot.nms(ref_rows, C, 0.45, 2)), i
        # per-(image, class) extents
        cls = out_h[oo[i]:oo[i + 1], 5].astype(int)
        assert np.array_equal(np.diff(seg[i * C:(i + 1) * C + 1]), np.bincount(cls, minlength=C))


def test_decode_capacity_overflow_and_empty():
    cfg = synth.make_config("v3-416", batch=2, seed=2)
    C = 80
    dev_preds = [torch.from_numpy(p).cuda() for p in cfg["y_preds"]]
    rows, offsets = engine.decode_batch(dev_preds, C, 0.5, 3, capacity=16)
    total = int(offsets[-1])
    assert total > 16                     # true count reported although the buffer is small
    full, off2 = engine.decode_batch_exact(dev_preds, C, 0.5, 3, capacity=16)
    assert full.shape[0] == total and torch.equal(off2, offsets)
    assert torch.equal(full[:16], rows[:16])
    # nothing passes the threshold
    zero = [torch.zeros_like(p) for p in dev_preds]
    r, o = engine.decode_batch_exact(zero, C, 0.5, 3)
    assert r.shape == (0, 7) and int(o[-1]) == 0
    assert tools.decode(np.zeros((13, 13, 255), np.float32), class_num=80, version=3).shape == (0,)
    res = engine.nms_batch(r, o, C, 0.45, 1)
    assert int(res["out_offsets"][-1]) == 0


def test_threshold_runs_in_input_dtype():
    """fp32 heads: c*p rounded to fp32 and compared with the fp32-rounded threshold (numpy semantics)."""
    rng = np.random.default_rng(0)
    g = rng.uniform(0.3, 1.0, (8, 8, 3 * 9)).astype(np.float32)
    for thr in (0.5, 0.45, 0.6, 0.3):
        a = tools.decode(g, class_num=4, threshold=thr, version=3)
        assert np.array_equal(a, ot.decode(g, class_num=4, threshold=thr, version=3))
        a64 = tools.decode(g.astype(np.float64), class_num=4, threshold=thr, version=3)
        assert np.array_equal(a64, ot.decode(g.astype(np.float64), class_num=4, threshold=thr, version=3))


@pytest.mark.parametrize("mode", [1, 2])
def test_dense_scene_nms(mode):
    """Config-4 style: thousands of candidates per class (CTA path, shared memory and >2048 global path)."""
    rng = np.random.default_rng(44)
    rows = synth.make_dense_candidates(rng, 20000, 4)        # ~5000 per class -> global-scratch path
    rows2 = synth.make_dense_candidates(rng, 6000, 12)       # ~500 per class -> shared-memory path
    for r, C in ((rows, 4), (rows2, 12)):
        with np.errstate(invalid="ignore", divide="ignore"):
            ref = ot.nms_keep(r, C, 0.45, mode)
        dev = torch.from_numpy(r).cuda()
        off = torch.tensor([0, len(r)], dtype=torch.int64, device="cuda")
        res = engine.nms_batch(dev, off, C, 0.45, mode)
        assert np.array_equal(res["keep"].cpu().numpy().astype(bool), ref)


def test_nms_edge_cases():
    rng = np.random.default_rng(9)
    rows = synth.make_dense_candidates(rng, 300, 3)
    rows[10, 5] = 7.0          # class id outside [0, C): belongs to no class, dropped
    rows[11, 5] = -1.0
    rows[20:25, 4] = rows[19, 4]
    rows[20:25, 6] = rows[19, 6]   # equal confidences: higher original index first
    rows[30:33, :4] = rows[29, :4]
    ref = ot.nms(rows, 3, 0.45, 1)
    assert np.array_equal(tools.nms(rows, 3, 0.45, 1), ref)
    # IoU exactly at the threshold suppresses (>=)
    two = np.array([[0.5, 0.5, 0.2, 0.2, 0.9, 0, 0.9], [0.5, 0.5, 0.2, 0.2, 0.8, 0, 0.9]])
    iou = ot.pair_iou(two[0], two[1])
    assert len(tools.nms(two, 1, float(iou), 1)) == 1
    assert len(tools.nms(two, 1, float(np.nextafter(iou, 2)), 1)) == 2
    # ragged batch: images with zero rows in between
    parts = [rows[:0], rows[:100], rows[:0], rows[100:300]]
    offs = np.cumsum([0] + [len(p) for p in parts])
    res = engine.nms_batch(torch.from_numpy(rows).cuda(), torch.from_numpy(offs).cuda(), 3, 0.45, 2)
    keep = res["keep"].cpu().numpy().astype(bool)
    for i, p in enumerate(parts):
        if len(p):
            assert np.array_equal(keep[offs[i]:offs[i + 1]], ot.nms_keep(p, 3, 0.45, 2))


def test_soft_nms(golden):
    """Gaussian soft-NMS (nms_mode=2): keep sets equal the reference fixtures and the oracle."""
    z = golden("decode_nms")
    for thr in (0.5, 0.3):
        for i in range(3):
            rows = z[f"m/rows_t{thr}_i{i}"]
            assert np.array_equal(tools.soft_nms(rows, 5, 0.45, thr, 0.5), z[f"m/soft_t{thr}_i{i}"])
    rng = np.random.default_rng(12)
    rows = synth.make_dense_candidates(rng, 4000, 6)      # warp path and CTA path
    for sigma, cthr in ((0.5, 0.5), (0.3, 0.6), (1.0, 0.3)):
        assert np.array_equal(tools.soft_nms(rows, 6, 0.45, cthr, sigma), ot.soft_nms(rows, 6, 0.45, cthr, sigma))
    rows = synth.make_dense_candidates(rng, 9000, 3)      # > 2048 per class: global-scratch path
    assert np.array_equal(tools.soft_nms(rows, 3, 0.5, 0.5, 0.5), ot.soft_nms(rows, 3, 0.5, 0.5, 0.5))


def test_label_helpers(golden):
    """down2xlabel and get_class_weight vs the reference fixtures (bit-exact / 1e-15)."""
    z = golden("labels")
    lab = z["lab"]
    assert np.array_equal(tools.down2xlabel(lab), z["down64"])
    assert np.array_equal(tools.down2xlabel(lab.astype(np.float32)), z["down32"])
    assert np.array_equal(tools.down2xlabel(lab), synth.halve_labels(lab))
    with pytest.raises(IndexError):
        tools.down2xlabel(lab[:, :11])
    for m in ("alpha", "log", "effective", "binary"):
        assert np.allclose(tools.get_class_weight(lab[..., 5:], method=m), z["cw_" + m], rtol=1e-14, atol=0), m
    big = synth.make_labels(np.random.default_rng(2), 16, [76], 80, synth.ANCHORS_V4)[0]   # fp32 labels
    ref = big.astype(np.float64).reshape(-1, 85)[:, 5:].sum(axis=0)
    got = engine.column_sums(torch.from_numpy(big[..., 5:].reshape(-1, 80).copy()).cuda()).cpu().numpy()
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("mode", [1, 2])
def test_nms_nonfinite_boxes_and_nonpositive_threshold(mode):
    """NaN / Inf coordinates behave as in NumPy (np.maximum / np.minimum propagate NaN: such a box
    neither suppresses nor is suppressed); thresholds <= 0 take the exact expression for every pair."""
    rng = np.random.default_rng(77)
    rows = synth.make_dense_candidates(rng, 400, 2)
    rows[5, 0] = np.nan
    rows[17, 3] = np.nan
    rows[30, 2] = np.inf            # infinitely wide box
    rows[31, 0] = np.inf            # centre at infinity: inf - inf inside the corners
    rows[40:43, 2:4] = 0.0          # zero-size boxes ...
    rows[41, :2] = rows[40, :2]     # ... two of them identical (0/0 in the DIoU term)
    rows[50, 4] = np.nan            # NaN confidences are visited first, like np.argsort()[::-1];
    rows[300, 6] = np.nan           # two of them: higher index first
    rows[[60, 61], 4] = np.inf      # and infinite ones right after
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        for thr in (0.45, 0.0, -0.25):
            ref = ot.nms_keep(rows, 2, thr, mode)
            res = engine.nms_batch(torch.from_numpy(rows).cuda(),
                                   torch.tensor([0, len(rows)], dtype=torch.int64, device="cuda"), 2, thr, mode)
            got = res["keep"].cpu().numpy().astype(bool)
            assert np.array_equal(got, ref), (mode, thr, np.nonzero(got != ref)[0][:10])


def test_nms_class_counts_and_ragged_images():
    """One class with thousands of rows (CTA + global-scratch paths), hundreds of classes with a few
    rows each (chunked per-image scans), empty first / last images, 33..128-box segments (rank-counting
    order) and >128 (bitonic)."""
    rng = np.random.default_rng(78)
    # (a) class_num = 1, the reference's default
    rows = synth.make_dense_candidates(rng, 3000, 1)
    parts = [rows[:0], rows[:40], rows[40:140], rows[140:3000], rows[:0]]
    offs = np.cumsum([0] + [len(p) for p in parts])
    res = engine.nms_batch(torch.from_numpy(rows).cuda(), torch.from_numpy(offs).cuda(), 1, 0.45, 1, want_seg_offsets=True)
    keep = res["keep"].cpu().numpy().astype(bool)
    out_rows, out_off = res["out_rows"].cpu().numpy(), res["out_offsets"].cpu().numpy()
    for i, p in enumerate(parts):
        want = ot.nms(p, 1, 0.45, 1) if len(p) else p
        assert np.array_equal(keep[offs[i]:offs[i + 1]], ot.nms_keep(p, 1, 0.45, 1) if len(p) else keep[:0]), i
        assert np.array_equal(out_rows[out_off[i]:out_off[i + 1]], want), i
    assert np.array_equal(res["seg_offsets"].cpu().numpy(), out_off)     # C = 1: segments are images
    # (b) 700 classes (more than one 256-thread chunk in the per-image scans), two images
    rows = synth.make_dense_candidates(rng, 5000, 700)
    offs = np.array([0, 2100, 5000])
    res = engine.nms_batch(torch.from_numpy(rows).cuda(), torch.from_numpy(offs).cuda(), 700, 0.3, 2, want_seg_offsets=True)
    out_rows, out_off = res["out_rows"].cpu().numpy(), res["out_offsets"].cpu().numpy()
    seg = res["seg_offsets"].cpu().numpy()
    for i in range(2):
        p = rows[offs[i]:offs[i + 1]]
        want = ot.nms(p, 700, 0.3, 2)
        assert np.array_equal(out_rows[out_off[i]:out_off[i + 1]], want), i
        per_class = np.bincount(want[:, 5].astype(int), minlength=700)
        assert np.array_equal(np.diff(seg[i * 700:(i + 1) * 700 + 1]), per_class), i


def test_decode_scan_sizes():
    """Per-cell count scan: single-block path (<= 64k cells) and the chained look-back path with a
    ragged last block, against the oracle's decode."""
    rng = np.random.default_rng(79)
    for n_img, S, thr in ((3, 13, 0.3), (420, 13, 0.6), (37, 52, 0.7)):     # 507 / 70,980 / 100,048 cells
        B, C = 3, 7
        yp = rng.uniform(0, 1, (n_img, S, S, B * (5 + C))).astype(np.float32)
        rows, offs = engine.decode_batch_exact([torch.from_numpy(yp).cuda()], C, thr, 3)
        rows_h, offs_h = rows.cpu().numpy(), offs.cpu().numpy()
        for i in (0, n_img // 2, n_img - 1):
            ref = ot.decode(yp[i], class_num=C, threshold=thr, version=3).reshape(-1, 7)
            assert np.array_equal(rows_h[offs_h[i]:offs_h[i + 1]], ref), (n_img, S, i)
        assert offs_h[-1] == len(rows_h)


@pytest.mark.parametrize("mode", [1, 2])
def test_pairwise_iou_matrix(mode):
    """(na, nb) IoU / DIoU matrix bit for bit: ragged tile sizes, zero-size and non-finite boxes."""
    rng = np.random.default_rng(80)
    a = synth.make_dense_candidates(rng, 131, 1)[:, :5]
    b = synth.make_dense_candidates(rng, 517, 1)[:, :5]
    a[3, 2:4] = 0.0
    b[5] = a[3]                      # identical zero-size boxes: 0/0 in the DIoU term
    a[7, 0] = np.nan
    b[9, 3] = np.inf
    b[11, 1] = -np.inf
    a[13, 2] = np.inf                # inf - inf inside the overlap
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        ref = ot.pair_iou(a[:, None, :], b[None, :, :], mode=mode)
    got = engine.pairwise_iou(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), mode).cpu().numpy()
    assert got.shape == ref.shape
    assert np.array_equal(got, ref, equal_nan=True), np.argwhere(~((got == ref) | (np.isnan(got) & np.isnan(ref))))[:5]
    one = engine.pairwise_iou(torch.from_numpy(a[:1]).cuda(), torch.from_numpy(b[:1]).cuda(), mode).cpu().numpy()
    assert np.array_equal(one, ref[:1, :1], equal_nan=True)
