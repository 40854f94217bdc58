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


def test_device_lloyd_loop_is_the_reference_loop(golden, capsys):
    """The loop runs on the device (update, loss and stop test in the assignment launch) and the
    host looks at it every `check_every` iterations: centres, iteration count and the printed losses
    equal the reference's whatever the batch size, including an empty-cluster hand-back."""
    z = golden("kmeans")
    rng = np.random.default_rng(77)
    data = synth.make_kmeans_boxes(rng, 20000, k=6)
    for fn, ofn, k, stop in ((km.iou_dist, okm.iou_dist, 6, 1e-5), (km.euclidean_dist, okm.euclidean_dist, 4, 1e-4)):
        np.random.seed(3)
        trace = []
        ref = okm.kmeans(data, k, ofn, stop, trace=trace)
        lines = [f"epoch {e + 1:2d}: loss = {t[2]:.4f}" for e, t in enumerate(trace)]
        for every in (1, 3, 8, 64):
            np.random.seed(3)
            got = km.kmeans(data, k, fn, stop, verbose=True, check_every=every)
            out = capsys.readouterr().out.strip().split("\n")
            assert np.array_equal(got, ref), (fn.__name__, every)
            assert out == lines, (fn.__name__, every)
    # iteration cap (kmeans.py:97: epoch > max_iternum)
    for cap in (1, 2, 5):
        np.random.seed(3)
        ref = okm.kmeans(data, 6, okm.iou_dist, 0.0, max_iternum=cap)
        np.random.seed(3)
        assert np.array_equal(km.kmeans(data, 6, km.iou_dist, 0.0, max_iternum=cap, verbose=False), ref), cap
    # empty clusters in several iterations (9 clusters over 50 boxes), every batch size
    for every in (1, 2, 8):
        np.random.seed(5)
        c = km.kmeans(z["empty/data"], 9, km.iou_dist, 1e-5, verbose=False, check_every=every)
        assert np.array_equal(c, z["empty/centers"]), every
    # assignments refer to the returned centres (they were one update behind)
    np.random.seed(3)
    c, a = km.kmeans(data, 6, km.iou_dist, 1e-5, verbose=False, return_assignments=True)
    np.random.seed(3)
    tr = []
    okm.kmeans(data, 6, okm.iou_dist, 1e-5, trace=tr)
    assert np.array_equal(a.cpu().numpy(), okm.assign(data, tr[-1][0], okm.iou_dist))


def test_reference_distance_functions_and_data_sized_helpers():
    """dist_func may be the reference's own function object (recognised by behaviour); the distance
    helpers take data-sized arrays (utils/kmeans.py:43-45 accepts both)."""
    rng = np.random.default_rng(8)
    data = synth.make_kmeans_boxes(rng, 70000, k=5)

    def ref_iou_dist(center_boxes, data_boxes):          # utils/kmeans.py:9-33, as a user would pass it
        ca = center_boxes[..., 0] * center_boxes[..., 1]
        da = data_boxes[..., 0] * data_boxes[..., 1]
        return 1 - np.minimum(ca, da) / np.maximum(ca, da)
    np.random.seed(1)
    a = km.kmeans(data, 5, ref_iou_dist, 1e-5, verbose=False)
    np.random.seed(1)
    assert np.array_equal(a, km.kmeans(data, 5, km.iou_dist, 1e-5, verbose=False))
    with pytest.raises(Exception):
        km.kmeans(data, 5, lambda c, d: np.abs(c - d).sum(-1), 1e-5, verbose=False)
    c = rng.uniform(0.05, 0.7, (5, 1, 2))
    assert np.array_equal(km.iou_dist(c, data[None]), okm.iou_dist(c, data[None]))           # (5, 70000) on the GPU
    assert np.array_equal(km.iou(c, data[None]), okm.area_ratio(c, data[None]))
    assert np.array_equal(km.euclidean_dist(c, data[None]), okm.euclidean_dist(c, data[None]))
    assert np.array_equal(km.iou_dist(data, data[::-1]), okm.iou_dist(data, data[::-1]))     # element by element
    x3 = rng.uniform(0, 1, (40000, 3))
    assert np.array_equal(km.euclidean_dist(x3, x3[::-1]), okm.euclidean_dist(x3, x3[::-1]))


def test_map_match_non_finite_boxes():
    """NaN / Inf / zero-size boxes: np.maximum / np.minimum propagate NaN into the IoU, np.max
    then returns NaN (never >= threshold) and np.argmax the FIRST NaN (measurement.py:270-277)."""
    rng = np.random.default_rng(9)
    C, n_img = 3, 4
    gts, dets, g_off, d_off = [], [], [0], [0]
    for i in range(n_img):
        g = np.column_stack([rng.uniform(0.2, 0.8, (6, 2)), rng.uniform(0.1, 0.4, (6, 2)), np.ones(6),
                             rng.integers(0, C, 6), np.ones(6)])
        d = np.column_stack([rng.uniform(0.2, 0.8, (12, 2)), rng.uniform(0.1, 0.4, (12, 2)), rng.uniform(0.3, 1, 12),
                             rng.integers(0, C, 12), rng.uniform(0.3, 1, 12)])
        gts.append(g)
        dets.append(d)
    gts[0][1, 2] = np.nan                 # NaN ground-truth width
    gts[0][3, 0] = np.inf                 # infinite centre
    gts[1][0, 2:4] = 0.0                  # zero-size ground truth
    dets[1][2, 2:4] = 0.0                 # zero-size detection
    dets[2][5, 1] = np.nan                # NaN detection
    dets[3][0, 3] = np.inf
    for g, d in zip(gts, dets):
        g_off.append(g_off[-1] + len(g))
        d_off.append(d_off[-1] + len(d))
    gt_rows, det_rows = np.concatenate(gts), np.concatenate(dets)
    dev = torch.device("cuda")
    bi, bg, counts = engine.map_match(torch.from_numpy(gt_rows).to(dev), torch.tensor(g_off, device=dev),
                                      torch.from_numpy(det_rows).to(dev), torch.tensor(d_off, device=dev), C)
    bi, bg = bi.cpu().numpy(), bg.cpu().numpy()
    with np.errstate(invalid="ignore", divide="ignore"):
        for i in range(n_img):
            for j in range(len(dets[i])):
                d = dets[i][j]
                same = gts[i][gts[i][:, 5].astype(int) == int(d[5])]
                k = d_off[i] + j
                if len(same) == 0:
                    assert bg[k] == -1
                    continue
                iou = ot.pair_iou(same[:, None, :5], d[None, None, :5], 1)[:, 0]
                want, arg = np.max(iou), int(np.argmax(iou))
                assert bg[k] == arg, (i, j)
                assert (np.isnan(want) and np.isnan(bi[k])) or bi[k] == want, (i, j)


def test_lloyd_step_with_the_in_kernel_exchange_single_rank():
    """yb_kmeans_lloyd_step_peers with a world of one (the rank's own mailbox): the exchange code of
    the last CTA runs - stores, release, acquire, rank-ordered sum - and the loop equals the plain
    device loop bit for bit, iteration after iteration (both parities of the mailbox)."""
    rng = np.random.default_rng(12)
    data = torch.from_numpy(synth.make_kmeans_boxes(rng, 200_000, k=7)).cuda()
    c0 = np.sort(rng.uniform(0.02, 0.8, (7, 2)), axis=0)
    a = engine.KMeansLloyd(data, torch.from_numpy(c0).cuda(), YB_DIST_IOU, 1e-6, 1000)
    b = engine.KMeansLloyd(data, torch.from_numpy(c0).cuda(), YB_DIST_IOU, 1e-6, 1000, sharded=True, peer_group="self")
    for it in range(9):
        a.step()
        b.step()
        sa, sb = a.read_state(), b.read_state()
        assert sa[0] == sb[0] and sa[1] == sb[1] == it + 1
        assert torch.equal(a.centers, b.centers), it
        assert np.array_equal(sa[2][:it + 1], sb[2][:it + 1])
    b.close()
