"""GPU parity: k-means assignment (bit-exact), Lloyd loop vs the reference fixtures, mAP matching."""
import numpy as np
import pytest
import torch

from oracle import kmeans as okm
from oracle import measurement as om
from oracle import tools as ot
from tf2_yolo_b200 import engine, synth
from tf2_yolo_b200._native import YB_DIST_EUCLID, YB_DIST_IOU
from tf2_yolo_b200.utils import kmeans as km

pytestmark = pytest.mark.gpu


def test_assignments_bit_exact_vs_reference_fixture(golden):
    z = golden("kmeans")
    data = torch.from_numpy(z["data"]).cuda()
    for name, kind in (("iou", YB_DIST_IOU), ("euclid", YB_DIST_EUCLID)):
        c0 = torch.from_numpy(z[f"{name}/c0"]).cuda()
        assign, sums, counts = engine.kmeans_assign(data, c0, kind, want_assign=True)
        ref = z[f"{name}/assign0"]
        assert np.array_equal(assign.cpu().numpy(), ref), name
        k = c0.shape[0]
        assert np.array_equal(counts.cpu().numpy(), np.bincount(ref, minlength=k))
        for c in range(k):
            want = z["data"][ref == c].sum(axis=0)
            assert np.allclose(sums[c].cpu().numpy(), want, rtol=1e-13, atol=0)


def test_lloyd_loop_matches_reference_fixture(golden):
    z = golden("kmeans")
    for name, fn in (("iou", km.iou_dist), ("euclid", km.euclidean_dist)):
        np.random.seed(12)
        c = km.kmeans(z["data"], int(z[f"{name}/k"]), fn, float(z[f"{name}/stop"]), verbose=False)
        assert c.dtype == np.float32 and np.array_equal(c, z[f"{name}/centers"]), name
    np.random.seed(5)   # empty clusters re-drawn from the host RNG stream in the reference's order
    c = km.kmeans(z["empty/data"], 9, km.iou_dist, 1e-5, verbose=False)
    assert np.array_equal(c, z["empty/centers"])


def test_large_assign_vs_oracle_and_ties():
    rng = np.random.default_rng(50)
    data = synth.make_kmeans_boxes(rng, 300_001, k=9)       # odd count: ragged last tile
    centers = rng.uniform(0.02, 0.8, (9, 2))
    centers[4] = centers[2]                                 # duplicate centre: first minimum wins
    data[:9] = centers                                      # zero distances
    data[10:20] = 1e-4                                      # area ratio < 2^-20: per-point exact fallback
    data[20:30] = [0.999, 1.0]                              # larger than every centroid
    ref = okm.assign(data, centers, okm.iou_dist)
    a, sums, counts = engine.kmeans_assign(torch.from_numpy(data).cuda(), torch.from_numpy(centers).cuda(),
                                           YB_DIST_IOU, want_assign=True)
    assert np.array_equal(a.cpu().numpy(), ref)
    assert counts.cpu().numpy()[4] == 0
    a2, s2, c2 = engine.kmeans_assign(torch.from_numpy(data).cuda(), torch.from_numpy(centers).cuda(), YB_DIST_IOU)
    assert a2 is None and torch.equal(s2, sums) and torch.equal(c2, counts)   # deterministic reduction
    mm = engine.minmax(torch.from_numpy(data).cuda()).cpu().numpy()
    assert mm[0] == data.min() and mm[1] == data.max()
    # well-separated centroids (the bracketed two-division fast path), incl. boxes on a centroid area
    c2 = np.array([[0.03, 0.05], [0.6, 0.5], [0.1, 0.12], [0.3, 0.2], [0.9, 0.95], [0.05, 0.08], [0.2, 0.1],
                   [0.15, 0.3], [0.4, 0.45]])
    data[100:109] = c2
    data[110:119] = c2[:, ::-1]                             # same area, swapped sides
    ref2 = okm.assign(data, c2, okm.iou_dist)
    a3, s3, n3 = engine.kmeans_assign(torch.from_numpy(data).cuda(), torch.from_numpy(c2).cuda(), YB_DIST_IOU, True)
    assert np.array_equal(a3.cpu().numpy(), ref2)
    assert np.array_equal(n3.cpu().numpy(), np.bincount(ref2, minlength=9))
    for c in range(9):
        assert np.allclose(s3[c].cpu().numpy(), data[ref2 == c].sum(axis=0), rtol=1e-12, atol=0)
    for d in (1, 3, 4):
        x = rng.uniform(0, 1, (5000, d))
        c = rng.uniform(0, 1, (6, d))
        a, _, _ = engine.kmeans_assign(torch.from_numpy(x).cuda(), torch.from_numpy(c).cuda(), YB_DIST_EUCLID, True)
        assert np.array_equal(a.cpu().numpy(), okm.assign(x, c, okm.euclidean_dist))


def test_map_match_vs_oracle(golden):
    z = golden("map")
    C = int(z["class_num"])
    y_true = z["y_true"]
    preds = [z["pred0"], z["pred1"]]
    n_img = y_true.shape[0]
    gt_rows, gt_off = engine.decode_batch_exact([torch.from_numpy(y_true).cuda()], C, 0.5, 4)
    det_rows, det_off = engine.decode_batch_exact([torch.from_numpy(p).cuda() for p in preds], C, 0.3, 4)
    res = engine.nms_batch(det_rows, det_off, C, 0.5, 1)
    n_keep = int(res["out_offsets"][-1])
    dets = res["out_rows"][:n_keep].contiguous()
    best_iou, best_gt, counts = engine.map_match(gt_rows, gt_off, dets, res["out_offsets"], C)
    best_iou, best_gt, counts = best_iou.cpu().numpy(), best_gt.cpu().numpy(), counts.cpu().numpy()
    oo = res["out_offsets"].cpu().numpy()
    dets_h = dets.cpu().numpy()
    for i in range(n_img):
        gt = ot.decode(y_true[i], class_num=C, version=4).reshape(-1, 7)
        det = ot.nms(ot.decode(*[p[i] for p in preds], class_num=C, threshold=0.3, version=4), C, 0.5, 1)
        assert np.array_equal(det, dets_h[oo[i]:oo[i + 1]])
        pos = oo[i]
        for k, (conf, arg, flag, n_gt) in enumerate(om.match_image(gt, det, C, 0.5)):
            assert counts[i, k] == n_gt
            n = len(conf)
            if n_gt > 0:
                assert np.array_equal(best_gt[pos:pos + n], arg)
                assert np.array_equal(best_iou[pos:pos + n] >= 0.5, flag)
            else:
                assert (best_gt[pos:pos + n] == -1).all()
            pos += n


def test_kmeans_many_clusters_and_extreme_areas():
    """k > 9 (counters in shared memory instead of packed registers), k = 1, areas outside the cell
    table (below 2^-15 and above 2), non-finite boxes: assignments equal np.argmin bit for bit."""
    rng = np.random.default_rng(51)
    data = synth.make_kmeans_boxes(rng, 100_003, k=9)
    data[:50] = rng.uniform(1e-4, 4e-3, (50, 2))            # areas down to 1e-8
    data[50:60] = rng.uniform(1.5, 3.0, (10, 2))            # areas above 2
    data[60] = [np.nan, 0.5]
    data[61] = [np.inf, 0.5]
    data[62] = [0.0, 0.5]
    for k in (1, 2, 12, 16):
        centers = rng.uniform(0.02, 0.9, (k, 2))
        with np.errstate(invalid="ignore", divide="ignore"):
            ref = okm.assign(data, centers, okm.iou_dist)
        a, sums, counts = engine.kmeans_assign(torch.from_numpy(data).cuda(), torch.from_numpy(centers).cuda(),
                                               YB_DIST_IOU, want_assign=True)
        assert np.array_equal(a.cpu().numpy(), ref), k
        assert np.array_equal(counts.cpu().numpy(), np.bincount(ref, minlength=k)), k
    # centroids 2^40 apart: the certainty window is clamped to ratios >= 2^-20, the rest is exact
    centers = np.array([[1e-7, 1e-6], [0.5, 0.5], [1e3, 1e3]])
    d2 = np.concatenate([data[:1000], rng.uniform(1e-7, 1e-3, (200, 2)), rng.uniform(10, 1e3, (200, 2))])
    with np.errstate(invalid="ignore", divide="ignore"):
        ref = okm.assign(d2, centers, okm.iou_dist)
    a, _, _ = engine.kmeans_assign(torch.from_numpy(d2).cuda(), torch.from_numpy(centers).cuda(), YB_DIST_IOU, True)
    assert np.array_equal(a.cpu().numpy(), ref)
