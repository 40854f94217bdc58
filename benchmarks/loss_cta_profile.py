#!/usr/bin/env python
"""Per-CTA timeline of the fused loss kernel (debug build: make EXTRA=-DYB_LOSS_PROFILE).
Prints when CTAs start, finish streaming, and how long the final reduction takes."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tf2_yolo_b200 import _native as N, engine, synth  # noqa: E402
from tf2_yolo_b200.yolov4.losses import wrap_yolo_loss  # noqa: E402

fused = "--fused" in sys.argv
batch = 128
cfg = synth.make_config("v4-608", batch=batch, seed=3)
B, Cn = cfg["bbox_num"], cfg["class_num"]
fns = [wrap_yolo_loss((S, S), B, Cn, anchors=cfg["anchors"][si * B:(si + 1) * B], loss_weight=[1, 5, 1])
       for si, S in enumerate(cfg["grids"])]
params = [f.params for f in fns]
yt = [torch.from_numpy(a).cuda() for a in cfg["y_trues"]]
yp = [torch.from_numpy(a).cuda() for a in cfg["y_preds"]]
dp = [torch.empty_like(a) for a in yp]
rows = torch.empty((4096 * batch, 7), dtype=torch.float64, device="cuda")


def run():
    if fused:
        engine.loss_decode_fused(params, yt, yp, 0.5, dpreds=dp, rows=rows, split_hook=lambda: None)
    else:
        engine.loss_fwd_bwd(params, yt, yp, dpreds=dp)


for _ in range(5):
    run()
torch.cuda.synchronize()
buf = (C.c_ulonglong * (1024 * 6))()
N.lib.yb_debug_loss_profile.restype = C.c_int
assert N.lib.yb_debug_loss_profile(buf) == 0
t = np.frombuffer(buf, dtype=np.uint64).reshape(1024, 6)[:592].astype(np.float64)
t0 = t[:, 0].min()
t = (t - t0) / 1e3
names = ["start", "consumers done", "producer loop done", "stores drained", "cta exit", "final reduce done"]
for k, nm in enumerate(names):
    col = t[:, k]
    col = col[col > -1e6]
    if k == 5:
        col = col[col > 0]
    print(f"{nm:20s} min {col.min():8.1f}  p50 {np.median(col):8.1f}  p90 {np.percentile(col, 90):8.1f}  max {col.max():8.1f} us   n={len(col)}")
end = t[:, 4]
order = np.argsort(end)
print("earliest exits (cta, us):", [(int(i), round(float(end[i]), 1)) for i in order[:6]])
print("latest exits   (cta, us):", [(int(i), round(float(end[i]), 1)) for i in order[-6:]])
sm = np.arange(592) % 148
per_sm = np.array([end[sm == i].max() for i in range(148)])
print("per-(cta mod 148) max exit: min %.1f p50 %.1f max %.1f" % (per_sm.min(), np.median(per_sm), per_sm.max()))
