"""GPU parity: in-training metrics accumulated inside the fused loss pass."""
import importlib
import json

import numpy as np
import pytest
import torch

from oracle import losses as ol
from oracle import metrics as omet
from tf2_yolo_b200 import synth
from tf2_yolo_b200.grid_loss import fused_losses

pytestmark = pytest.mark.gpu
PKG = {1: "yolov1_5", 2: "yolov2", 3: "yolov3", 4: "yolov4"}


def test_reference_fixtures(golden):
    z = golden("metrics")
    for name in z["names"]:
        name = str(name)
        meta = json.loads(str(z[name + "/meta"]))
        ver, S, B, C, thr = meta["version"], meta["grid"], meta["B"], meta["C"], meta["thr"]
        mod = importlib.import_module(f"tf2_yolo_b200.{PKG[ver]}.metrics")
        yt, yp = z[name + "/y_true"], z[name + "/y_pred"]
        fns = [mod.wrap_obj_acc((S, S), B, C), mod.wrap_mean_iou((S, S), B, C),
               mod.wrap_class_acc((S, S), C) if ver == 1 else mod.wrap_class_acc((S, S), B, C),
               mod.wrap_recall((S, S), B, C, iou_threshold=thr)]
        got = np.array([float(f(yt, yp)) for f in fns])
        assert np.allclose(got, z[name + "/metrics"], rtol=2e-6, atol=1e-7), (name, got, z[name + "/metrics"])


def test_metrics_ride_along_with_the_loss():
    """Same launch as loss + gradient: loss/grad bits unchanged, metrics equal the oracle."""
    cfg = synth.make_config("v4-608", batch=4, seed=41)
    B, C = cfg["bbox_num"], cfg["class_num"]
    from tf2_yolo_b200.yolov4.losses import wrap_yolo_loss
    fns = [wrap_yolo_loss((S, S), B, C, anchors=cfg["anchors"][si * B:(si + 1) * B], loss_weight=[1, 5, 1])
           for si, S in enumerate(cfg["grids"])]
    yts = [torch.from_numpy(a).cuda() for a in cfg["y_trues"]]
    yps = [torch.from_numpy(a).cuda() for a in cfg["y_preds"]]
    loss0, d0, _ = fused_losses(fns, yts, yps)
    loss1, d1, _, met = fused_losses(fns, yts, yps, want_metrics=True, recall_iou_threshold=0.4)
    assert torch.equal(loss0, loss1) and all(torch.equal(a, b) for a, b in zip(d0, d1))
    met = met.cpu().numpy()
    for si, S in enumerate(cfg["grids"]):
        m = omet.grid_metrics(4, cfg["y_trues"][si], cfg["y_preds"][si], (S, S), B, C, 0.4)
        want = np.array([m["obj_acc"], m["mean_iou"], m["class_acc"], m["recall"]])
        assert np.allclose(met[si, :4], want, rtol=2e-6, atol=1e-7), (S, met[si], want)
        assert np.allclose(met[si, 4:10], m["raw"], rtol=2e-6)


def _oracle4(ver, yt, yp, S, B, C, thr):
    m = omet.grid_metrics(ver, yt, yp, (S, S), B, C, thr)
    return np.array([m["obj_acc"], m["mean_iou"], m["class_acc"], m["recall"]])


def test_back_to_back_batches_of_equal_shape_are_not_confused():
    """Two different batches of identical shape, back to back: NumPy inputs (their device copies die
    on return and the allocator reuses the addresses) and freshly allocated CUDA tensors.  Every
    metric must describe ITS batch (round-1 bug: a cache keyed on data_ptr returned batch 1's)."""
    import gc
    mod = importlib.import_module("tf2_yolo_b200.yolov4.metrics")
    S, B, C, thr = 19, 3, 80, 0.5
    anc = synth.ANCHORS_V4
    fns = [mod.wrap_obj_acc((S, S), B, C), mod.wrap_mean_iou((S, S), B, C), mod.wrap_class_acc((S, S), B, C),
           mod.wrap_recall((S, S), B, C, iou_threshold=thr)]
    batches = []
    for seed in (71, 72):
        cfg = synth.make_config("v4-608", batch=2, seed=seed)
        batches.append((cfg["y_trues"][0], cfg["y_preds"][0]))
    want = [_oracle4(4, yt, yp, S, B, C, thr) for yt, yp in batches]
    assert not np.allclose(want[0], want[1])
    # NumPy in, one metric at a time, batches interleaved
    for f_i, f in enumerate(fns):
        for b_i, (yt, yp) in enumerate(batches):
            assert np.isclose(float(f(yt, yp)), want[b_i][f_i], rtol=2e-6, atol=1e-7), (f_i, b_i)
    # fresh CUDA tensors of the same shape: free batch 1 before creating batch 2 (same addresses)
    got = []
    for yt, yp in batches:
        t, p = torch.from_numpy(yt).cuda(), torch.from_numpy(yp).cuda()
        got.append(np.array([float(f(t, p)) for f in fns]))
        del t, p
        gc.collect()
    for b_i in range(2):
        assert np.allclose(got[b_i], want[b_i], rtol=2e-6, atol=1e-7), b_i
    # an in-place update of the same tensor object is seen as well
    t, p = torch.from_numpy(batches[0][0]).cuda(), torch.from_numpy(batches[0][1]).cuda()
    a = float(fns[1](t, p))
    p.copy_(torch.from_numpy(batches[1][1]))
    t.copy_(torch.from_numpy(batches[1][0]))
    b = float(fns[1](t, p))
    assert np.isclose(a, want[0][1], rtol=2e-6) and np.isclose(b, want[1][1], rtol=2e-6)


def test_four_metrics_of_one_pair_share_one_launch():
    from tf2_yolo_b200 import grid_metrics as gm
    cfg = synth.make_config("v3-416", batch=2, seed=5)
    t, p = torch.from_numpy(cfg["y_trues"][1]).cuda(), torch.from_numpy(cfg["y_preds"][1]).cuda()
    m1 = gm.grid_metrics(3, t, p, (26, 26), 3, 80)
    m2 = gm.grid_metrics(3, t, p, (26, 26), 3, 80)
    assert m2 is m1                                           # same tensors, same version: remembered
    m3 = gm.grid_metrics(3, t, p, (26, 26), 3, 80, iou_threshold=0.3)
    assert m3 is not m1
    m4 = gm.grid_metrics(3, cfg["y_trues"][1], cfg["y_preds"][1], (26, 26), 3, 80, iou_threshold=0.3)
    assert m4 is not m3 and torch.equal(m4, m3)               # host arrays are never cached
