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
