#!/usr/bin/env python
"""Short driver for ncu captures of the non-loss kernels (k-means assignment, dense-scene NMS,
PR matching):  python benchmarks/profile_targets.py [kmeans] [nms] [iou]
Each target runs warm-ups and a few launches; use with `ncu -k regex:<kernel> -s <skip> -c <n>`."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tf2_yolo_b200 import engine, synth  # noqa: E402
from tf2_yolo_b200._native import YB_DIST_IOU  # noqa: E402

what = sys.argv[1:] or ["kmeans", "nms"]
if "kmeans" in what:
    rng = np.random.default_rng(4)
    n = int(os.environ.get("YB_PROF_KM_N", 50_000_000))
    data = torch.from_numpy(synth.make_kmeans_boxes(rng, n, 9)).cuda()
    centers = torch.from_numpy(np.sort(rng.uniform(0.02, 0.8, (9, 2)), axis=0)).cuda()
    for _ in range(4):
        engine.kmeans_assign(data, centers, YB_DIST_IOU)
    loop = engine.KMeansLloyd(data, centers.clone(), YB_DIST_IOU, 0.0, 1 << 40)   # the device Lloyd loop
    for _ in range(4):
        loop.step()
    torch.cuda.synchronize()
    del data
if "nms" in what:
    rng = np.random.default_rng(4)
    n_img, per_img, C = int(os.environ.get("YB_PROF_NMS_IMG", 8)), 100_000, 80
    rows = np.concatenate([synth.make_dense_candidates(rng, per_img, C) for _ in range(n_img)])
    offs = torch.arange(0, (n_img + 1) * per_img, per_img, dtype=torch.int64, device="cuda")
    dev = torch.from_numpy(rows).cuda()
    for mode in (1, 2):
        for _ in range(2):
            engine.nms_batch(dev, offs, C, 0.45, mode)
    torch.cuda.synchronize()
if "iou" in what:
    rng = np.random.default_rng(5)
    a = torch.from_numpy(synth.make_dense_candidates(rng, 8192, 1)).cuda()
    b = torch.from_numpy(synth.make_dense_candidates(rng, 8192, 1)).cuda()
    for mode in (1, 2):
        for _ in range(3):
            engine.pairwise_iou(a, b, mode)       # 67 M pair IoUs per launch, fp64, (n, n) matrix out
    torch.cuda.synchronize()
if "step" in what:   # the two-launch train-and-evaluate step of bench.py (v4-608)
    from tf2_yolo_b200.yolov4.losses import wrap_yolo_loss
    batch = int(os.environ.get("YB_PROF_BATCH", 128))
    cfg = synth.make_config("v4-608", batch=batch, seed=2)
    B, C = 3, 80
    fns = [wrap_yolo_loss((S, S), B, C, anchors=cfg["anchors"][si * B:(si + 1) * B], loss_weight=[1, 5, 1])
           for si, S in enumerate(cfg["grids"])]
    yts = [torch.from_numpy(a).cuda() for a in cfg["y_trues"]]
    yps = [torch.from_numpy(a).cuda() for a in cfg["y_preds"]]
    dps = [torch.empty_like(a) for a in yps]
    for _ in range(4):
        engine.loss_decode_nms_fused([f.params for f in fns], yts, yps, 0.5, 0.45, 2, rows_per_img_cap=1024, dpreds=dps)
    torch.cuda.synchronize()
print("ok")
