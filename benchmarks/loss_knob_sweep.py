"""Sweep the loss kernel's ring geometry (YB_LOSS_STAGES / CTAS_PER_SM / TILE_CELLS / WARPS, read once per
process) for the two-launch v4-608 step: one subprocess per setting, CUDA-graph replays of
engine.TrainEvalStep, prints ms per step.   python benchmarks/loss_knob_sweep.py [batch]"""
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, %r)
from tf2_yolo_b200 import engine, synth
from tf2_yolo_b200.yolov4.losses import wrap_yolo_loss
batch = int(sys.argv[1])
cfg = synth.make_config("v4-608", batch=batch, seed=2, rank=0)
B, C = cfg["bbox_num"], cfg["class_num"]
fns = [wrap_yolo_loss((S, S), B, C, anchors=cfg["anchors"][si * B:(si + 1) * B], loss_weight=[1, 5, 1],
                      wh_reg_weight=0.01, ignore_thresh=0.6) for si, S in enumerate(cfg["grids"])]
dev = torch.device("cuda", 0)
yp = [torch.from_numpy(a).to(dev) for a in cfg["y_preds"]]
yt = [torch.from_numpy(a).to(dev) for a in cfg["y_trues"]]
st = engine.TrainEvalStep([f.params for f in fns], yt, yp, 0.5, 0.45, 2, rows_per_img_cap=1024)
for _ in range(5):
    st.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ms = []
for _ in range(15):
    e0.record()
    for _ in range(20):
        st.run()
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1) / 20)
print(json.dumps({"ms_per_step": float(np.median(ms)), "loss": st.loss.cpu().tolist()}))
''' % ROOT

batch = sys.argv[1] if len(sys.argv) > 1 else "128"
settings = [dict()] + [dict(YB_LOSS_STAGES=s, YB_LOSS_CTAS_PER_SM=c, YB_LOSS_TILE_CELLS=t)
                       for s, c, t in [(2, 4, 16), (2, 5, 16), (3, 3, 20), (3, 4, 12), (2, 3, 20), (2, 3, 28),
                                       (3, 2, 28), (2, 6, 12), (4, 2, 20), (3, 3, 16)]]
for env in settings:
    e = dict(os.environ)
    e.update({k: str(v) for k, v in env.items()})
    r = subprocess.run([sys.executable, "-c", CHILD, batch], env=e, capture_output=True, text=True, timeout=300)
    out = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr.strip()[-300:]
    print(json.dumps(env), out, flush=True)
