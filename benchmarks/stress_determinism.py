#!/usr/bin/env python
"""Repeat the atomics-heavy paths and compare every run with the first one bit for bit (a cheap
race detector: compute-sanitizer is closed on this pool).

    python benchmarks/stress_determinism.py [repeats]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tf2_yolo_b200 import engine, synth  # noqa: E402
from tf2_yolo_b200._native import YB_DIST_IOU  # noqa: E402
from tf2_yolo_b200.yolov4.losses import wrap_yolo_loss  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rng = np.random.default_rng(123)
bad = 0


def check(name, fn):
    global bad
    first = None
    for i in range(reps):
        out = [t.clone() for t in fn()]
        torch.cuda.synchronize()
        if first is None:
            first = out
        elif not all(torch.equal(a, b) for a, b in zip(first, out)):
            bad += 1
            print(f"MISMATCH {name} at repeat {i}")
            return
    print(f"ok {name}: {reps} identical runs")


# dense scene: big segments (shared-memory and global-scratch paths), both modes, soft-NMS
rows = np.concatenate([synth.make_dense_candidates(rng, 100_000, 80) for _ in range(4)] +
                      [synth.make_dense_candidates(rng, 20_000, 4)])
offs = torch.tensor([0, 100_000, 200_000, 300_000, 400_000, 420_000], dtype=torch.int64, device="cuda")
d = torch.from_numpy(rows).cuda()
for mode in (1, 2):
    def run(mode=mode):
        r = engine.nms_batch(d, offs, 80, 0.45, mode, want_seg_offsets=True)
        n = int(r["out_offsets"][-1])
        return r["keep"], r["out_rows"][:n], r["out_offsets"], r["seg_offsets"]
    check(f"dense nms mode {mode}", run)
check("dense soft-nms", lambda: (lambda r: (r["keep"], r["out_offsets"]))(
    engine.nms_batch(d[:420_000], offs, 80, 0.45, 1, soft=(0.5, 0.5))))

# sparse scene through the fused loss + decode + NMS step
cfg = synth.make_config("v4-608", batch=32, seed=9)
B, C = cfg["bbox_num"], cfg["class_num"]
fns = [wrap_yolo_loss((S, S), B, C, anchors=cfg["anchors"][i * B:(i + 1) * B], loss_weight=[1, 5, 1])
       for i, S in enumerate(cfg["grids"])]
yt = [torch.from_numpy(a).cuda() for a in cfg["y_trues"]]
yp = [torch.from_numpy(a).cuda() for a in cfg["y_preds"]]
params = [f.params for f in fns]


def step():
    loss, dp, _, rws, ro = engine.loss_decode_fused(params, yt, yp, 0.5, capacity=4096 * 32)
    r = engine.nms_batch(rws, ro, C, 0.45, 2)
    n = int(r["out_offsets"][-1])
    return [loss, *dp, rws[:int(ro[-1])], ro, r["keep"][:int(ro[-1])], r["out_rows"][:n], r["out_offsets"]]


check("fused loss + decode + nms step", step)

# k-means assignment (shared-memory accumulators, last-CTA reduction)
data = torch.from_numpy(synth.make_kmeans_boxes(rng, 5_000_001, 9)).cuda()
cen = torch.from_numpy(np.sort(rng.uniform(0.02, 0.8, (9, 2)), axis=0)).cuda()
check("kmeans assign", lambda: engine.kmeans_assign(data, cen, YB_DIST_IOU, want_assign=True))
sys.exit(1 if bad else 0)
