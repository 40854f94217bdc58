"""GPU parity at BASELINE.json's FULL sizes, through properties that do not need the oracle to run
the whole workload (it would take hours): per-image independence checked against the oracle on a
few sampled images, batch linearity (the sharding contract of SURVEY 8e), idempotence and the
greedy-NMS invariants, k-means conservation laws and sampled assignments.

config 3: YOLOv4-608, batch 128         config 4: 100,000 candidates / image, 80 classes
config 5: 50,000,000 boxes, k = 9       (1024 images / 8 GPUs = 128 per GPU; 16 here keeps the
                                         host-side generation of the inputs within seconds)
"""
import numpy as np
import pytest
import torch

from oracle import kmeans as okm
from oracle import losses as ol
from oracle import tools as ot
from tf2_yolo_b200 import engine, synth
from tf2_yolo_b200._native import YB_DIST_IOU
from tf2_yolo_b200.yolov4.losses import wrap_yolo_loss

pytestmark = pytest.mark.gpu

LOSS_RTOL, GRAD_RTOL = 1e-5, 1e-4


@pytest.fixture(scope="module")
def v4_full():
    cfg = synth.make_config("v4-608", batch=128, seed=3)
    B, C = cfg["bbox_num"], cfg["class_num"]
    fns, specs = [], []
    for si, S in enumerate(cfg["grids"]):
        kw = dict(anchors=cfg["anchors"][si * B:(si + 1) * B], loss_weight=[1, 5, 1])
        fns.append(wrap_yolo_loss((S, S), B, C, **kw))
        specs.append(ol.GridLossSpec(version=4, grid_shape=(S, S), bbox_num=B, class_num=C, **kw))
    yt = [torch.from_numpy(a).cuda() for a in cfg["y_trues"]]
    yp = [torch.from_numpy(a).cuda() for a in cfg["y_preds"]]
    return cfg, fns, specs, yt, yp


def test_config3_loss_sampled_images_and_batch_linearity(v4_full):
    cfg, fns, specs, yt, yp = v4_full
    params = [f.params for f in fns]
    loss, dpreds, terms = engine.loss_fwd_bwd(params, yt, yp, want_terms=True)
    torch.cuda.synchronize()
    loss = loss.cpu().numpy().astype(np.float64)
    # (1) cells are independent and every sum is (1/N) * sum over cells: the gradient of image i
    #     in the batch of 128 is the oracle's gradient of image i alone, divided by 128
    for i in (0, 57, 127):
        for si in range(3):
            _, g_ref, _ = ol.loss_and_grad(specs[si], cfg["y_trues"][si][i:i + 1], cfg["y_preds"][si][i:i + 1])
            g = dpreds[si][i].cpu().numpy().astype(np.float64) * 128.0
            scale = np.abs(g_ref).max()
            assert np.all(np.abs(g - g_ref[0]) <= GRAD_RTOL * np.abs(g_ref[0]) + 1e-6 * scale), (i, si)
    # (2) batch linearity = the multi-GPU sharding contract: four shards of 32 images, each divided by
    #     the GLOBAL batch, sum to the full-batch loss; gradients of a shard are those of the full batch
    parts = np.zeros(3)
    for r in range(4):
        sl = slice(32 * r, 32 * r + 32)
        l_r, d_r, _ = engine.loss_fwd_bwd(params, [t[sl] for t in yt], [p[sl] for p in yp], global_batch=128)
        parts += l_r.cpu().numpy().astype(np.float64)
        for si in range(3):
            assert torch.equal(d_r[si], dpreds[si][sl]), (r, si)
    assert np.all(np.abs(parts - loss) <= 4e-6 * np.abs(loss)), (parts, loss)
    # (3) the per-scale loss equals the mean over images of the oracle's single-image losses on a sample,
    #     scaled: checked through the terms (box, conf, prob, reg) of one image alone
    l1, _, _ = engine.loss_fwd_bwd(params, [t[5:6] for t in yt], [p[5:6] for p in yp])
    for si in range(3):
        l_ref, _, _ = ol.loss_and_grad(specs[si], cfg["y_trues"][si][5:6], cfg["y_preds"][si][5:6])
        assert abs(float(l1[si]) - l_ref) <= LOSS_RTOL * abs(l_ref)
    # (4) same bits on a second run (order-independent accumulation)
    loss2, dpreds2, _ = engine.loss_fwd_bwd(params, yt, yp)
    assert np.array_equal(loss2.cpu().numpy().astype(np.float64), loss)
    assert all(torch.equal(a, b) for a, b in zip(dpreds, dpreds2))
    assert all(bool(torch.isfinite(d).all()) for d in dpreds)


def test_config3_decode_nms_full_batch(v4_full):
    cfg, fns, specs, yt, yp = v4_full
    C = cfg["class_num"]
    rows, offs = engine.decode_batch_exact(yp, C, 0.5, 4)
    res = engine.nms_batch(rows, offs, C, 0.45, 2, want_seg_offsets=True)
    offs_h = offs.cpu().numpy()
    rows_h = rows.cpu().numpy()
    keep_h = res["keep"].cpu().numpy().astype(bool)
    out_off = res["out_offsets"].cpu().numpy()
    out_rows = res["out_rows"].cpu().numpy()
    for i in (0, 31, 64, 127):
        ref = ot.decode(*[p[i] for p in cfg["y_preds"]], class_num=C, threshold=0.5, version=4).reshape(-1, 7)
        assert np.array_equal(rows_h[offs_h[i]:offs_h[i + 1]], ref), i
        assert np.array_equal(keep_h[offs_h[i]:offs_h[i + 1]], ot.nms_keep(ref, C, 0.45, 2)), i
        assert np.array_equal(out_rows[out_off[i]:out_off[i + 1]], ot.nms(ref, C, 0.45, 2)), i
    # idempotence over the whole batch: the survivors survive a second pass, in the same order
    n_out = int(out_off[-1])
    res2 = engine.nms_batch(res["out_rows"][:n_out].contiguous(), res["out_offsets"], C, 0.45, 2)
    assert bool(res2["keep"][:n_out].all())
    assert torch.equal(res2["out_rows"][:n_out], res["out_rows"][:n_out])
    # the fused loss+decode entry point yields the same rows
    params = [f.params for f in fns]
    cap = int(offs_h[-1]) + 16
    _, _, _, rows_f, offs_f = engine.loss_decode_fused(params, yt, yp, 0.5, capacity=cap)
    assert torch.equal(offs_f, offs) and torch.equal(rows_f[:offs_h[-1]], rows)


@pytest.mark.parametrize("mode", [1, 2])
def test_config4_dense_nms_invariants(mode):
    n_img, per_img, C, thr = 16, 100_000, 80, 0.45
    rng = np.random.default_rng(404)
    rows = np.concatenate([synth.make_dense_candidates(rng, per_img, C) for _ in range(n_img)])
    offs_h = np.arange(0, (n_img + 1) * per_img, per_img, dtype=np.int64)
    dev = torch.from_numpy(rows).cuda()
    offs = torch.from_numpy(offs_h).cuda()
    res = engine.nms_batch(dev, offs, C, thr, mode, want_seg_offsets=True)
    keep = res["keep"].cpu().numpy().astype(bool)
    seg = res["seg_offsets"].cpu().numpy()
    out_off = res["out_offsets"].cpu().numpy()
    n_out = int(out_off[-1])
    assert n_out == int(keep.sum())
    cls = rows[:, 5].astype(np.int64)
    img = np.repeat(np.arange(n_img), per_img)
    # per-(image, class) survivor counts match the segment table
    want = np.bincount(img[keep] * C + cls[keep], minlength=n_img * C)
    assert np.array_equal(np.diff(seg), want)
    # oracle on sampled (image, class) segments (~1250 boxes each: seconds in NumPy)
    for i, k in ((0, 0), (7, 41), (15, 79)):
        idx = np.nonzero((img == i) & (cls == k))[0]
        sub = rows[idx].copy()
        sub[:, 5] = 0.0
        with np.errstate(invalid="ignore", divide="ignore"):
            assert np.array_equal(keep[idx], ot.nms_keep(sub, 1, thr, mode)), (i, k)
    # greedy-NMS invariants on other segments: survivors are mutually below the threshold, and every
    # removed box has a survivor with confidence >= its own that overlaps it at or above the threshold
    for i, k in ((3, 5), (11, 60)):
        idx = np.nonzero((img == i) & (cls == k))[0]
        sub, kp = rows[idx], keep[idx]
        with np.errstate(invalid="ignore", divide="ignore"):
            m = ot.pair_iou(sub[:, None, :5], sub[None, :, :5], mode=mode)
        conf = sub[:, 4] * sub[:, 6]
        mk = m[np.ix_(kp, kp)].copy()
        np.fill_diagonal(mk, -1.0)
        assert not (mk >= thr).any()
        hit = (m[np.ix_(~kp, kp)] >= thr) & (conf[kp][None, :] >= conf[~kp][:, None])
        assert hit.any(axis=1).all()
    # idempotence over all 1.6 M rows
    res2 = engine.nms_batch(res["out_rows"][:n_out].contiguous(), res["out_offsets"], C, thr, mode)
    assert bool(res2["keep"][:n_out].all())
    # output order: class-major per image, original order inside a class
    out_rows = res["out_rows"][:n_out].cpu().numpy()
    i = 9
    kept_i = rows[offs_h[i]:offs_h[i + 1]][keep[offs_h[i]:offs_h[i + 1]]]
    order = np.argsort(kept_i[:, 5].astype(np.int64), kind="stable")
    assert np.array_equal(out_rows[out_off[i]:out_off[i + 1]], kept_i[order])


def test_config5_kmeans_50m_conservation_and_sampled_assignments():
    n, k = 50_000_000, 9
    rng = np.random.default_rng(505)
    data_h = synth.make_kmeans_boxes(rng, n, k)
    centers = np.sort(rng.uniform(0.02, 0.8, (k, 2)), axis=0)
    data = torch.from_numpy(data_h).cuda()
    cen = torch.from_numpy(centers).cuda()
    assign, sums, counts = engine.kmeans_assign(data, cen, YB_DIST_IOU, want_assign=True)
    a = assign.cpu().numpy()
    counts_h, sums_h = counts.cpu().numpy(), sums.cpu().numpy()
    # conservation: every box counted once, and the cluster sums add up to the column sums
    assert counts_h.sum() == n and np.array_equal(counts_h, np.bincount(a, minlength=k))
    total = data.sum(dim=0).cpu().numpy()
    assert np.allclose(sums_h.sum(axis=0), total, rtol=1e-12, atol=0)
    # assignments are per box: the oracle on sampled slices (first, middle, ragged tail)
    for lo, hi in ((0, 200_000), (25_000_000, 25_200_000), (n - 100_001, n)):
        assert np.array_equal(a[lo:hi], okm.assign(data_h[lo:hi], centers, okm.iou_dist)), (lo, hi)
    # per-cluster sums against NumPy on the device assignments (summation order differs: 1e-12)
    for c in range(k):
        assert np.allclose(sums_h[c], data_h[a == c].sum(axis=0), rtol=1e-12, atol=0), c
    # sharding contract (8e): partial sums of two halves add up; counts exactly
    h = n // 2
    _, s0, c0 = engine.kmeans_assign(data[:h], cen, YB_DIST_IOU)
    _, s1, c1 = engine.kmeans_assign(data[h:], cen, YB_DIST_IOU)
    assert torch.equal(c0 + c1, counts)
    assert np.allclose((s0 + s1).cpu().numpy(), sums_h, rtol=1e-12, atol=0)
    # deterministic: same bits on a second pass
    _, s2, c2 = engine.kmeans_assign(data, cen, YB_DIST_IOU)
    assert torch.equal(s2, sums) and torch.equal(c2, counts)


def test_config3_two_launch_step_equals_the_chain_at_batch_128():
    """BASELINE config 3 at its stated size: the two-launch step (loss + counting pass, decode + NMS
    with one CTA per image) returns the loss, the gradients and the survivors of the separate calls."""
    import torch
    from tf2_yolo_b200 import engine, synth
    from tf2_yolo_b200.grid_loss import fused_losses
    from tf2_yolo_b200.yolov4.losses import wrap_yolo_loss
    cfg = synth.make_config("v4-608", batch=128, seed=2)
    B, C = 3, 80
    fns = [wrap_yolo_loss((S, S), B, C, anchors=cfg["anchors"][si * B:(si + 1) * B], loss_weight=[1, 5, 1])
           for si, S in enumerate(cfg["grids"])]
    yts = [torch.from_numpy(a).cuda() for a in cfg["y_trues"]]
    yps = [torch.from_numpy(a).cuda() for a in cfg["y_preds"]]
    loss0, d0, _ = fused_losses(fns, yts, yps)
    rows, offs = engine.decode_batch_exact(yps, C, 0.5, 4)
    g = engine.nms_batch(rows, offs, C, 0.45, 2)
    n = int(g["out_offsets"][-1].item())
    loss, d, _, r = engine.loss_decode_nms_fused([f.params for f in fns], yts, yps, 0.5, 0.45, 2, rows_per_img_cap=1024)
    assert int(r["n_overflow"].item()) == 0
    assert torch.equal(loss, loss0) and all(torch.equal(a, b) for a, b in zip(d, d0))
    assert torch.equal(r["out_offsets"], g["out_offsets"]) and torch.equal(r["out_rows"][:n], g["out_rows"][:n])


def test_config5_ap_table_of_500_images_equals_the_oracle_checked_fixture():
    """PR curves / mAP (utils/measurement.py:198-447) over 500 v4-608 images: the 81-row AP table
    must equal tests/golden/map500_ap_oracle_checked.npz bit for bit.  That table was produced by
    this path on a B200 and then compared with the CPU oracle over the same 500 images
    (benchmarks/map_subsample_check.py, 1925 s of oracle time: profiles/r2/map500_oracle_check.json) -
    identical, max |diff| 0."""
    import os
    from tf2_yolo_b200.utils import measurement as meas
    ref = np.load(os.path.join(os.path.dirname(__file__), "golden", "map500_ap_oracle_checked.npz"))
    n = int(ref["images"])
    cfg = synth.make_config("v4-608", batch=n, seed=50)
    assert sum(float(p.sum()) for p in cfg["y_preds"]) == float(ref["checksum"])   # the same synthetic images
    pr = meas.PRfunc(cfg["y_trues"][-1].astype(np.float64), *cfg["y_preds"], class_names=[str(i) for i in range(80)],
                     conf_threshold=0.05, nms_mode=1, nms_threshold=0.5, max_per_img=100, version=4)
    got = pr.get_map()["ap"].values.astype(np.float64)
    assert np.array_equal(got, ref["ap"])
