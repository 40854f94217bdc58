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
