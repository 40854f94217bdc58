#!/usr/bin/env python
"""Fixed cost vs streaming cost of the fused loss kernel: time it at several batch sizes with the
launch already queued behind a spin kernel (so host-side launch latency is outside the interval)
and fit  t = a + bytes / bw.

    python benchmarks/loss_scaling.py [v4-608|v3-416] [--fused]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tf2_yolo_b200 import engine, synth  # noqa: E402
from tf2_yolo_b200.grid_loss import fused_losses  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "v4-608"
    fused = "--fused" in sys.argv
    version = 4 if name.startswith("v4") else 3
    import importlib
    wrap = importlib.import_module(f"tf2_yolo_b200.yolov{version}.losses").wrap_yolo_loss
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    pts = []
    for batch in (8, 16, 32, 64, 128, 192):
        cfg = synth.make_config(name, batch=batch, seed=3)
        B, C = cfg["bbox_num"], cfg["class_num"]
        fns = [wrap((S, S), B, C, anchors=cfg["anchors"][si * B:(si + 1) * B],
                    loss_weight=[1, 5, 1] if version == 4 else [1, 1, 5, 1]) for si, S in enumerate(cfg["grids"])]
        yt = [torch.from_numpy(a).cuda() for a in cfg["y_trues"]]
        yp = [torch.from_numpy(a).cuda() for a in cfg["y_preds"]]
        dp = [torch.empty_like(a) for a in yp]
        rows = torch.empty((4096 * batch, 7), dtype=torch.float64, device="cuda")
        params = [f.params for f in fns]
        nbytes = sum(4 * batch * s * s * (2 * B * (5 + C) + 5 + C) for s in cfg["grids"])

        def run(hook=None):
            if fused:
                engine.loss_decode_fused(params, yt, yp, 0.5, dpreds=dp, rows=rows, split_hook=hook or (lambda: None))
            else:
                fused_losses(fns, yt, yp, dpreds=dp)
                if hook:
                    hook()
        for _ in range(3):
            run()
        ts = []
        for _ in range(7):
            flush.zero_()
            torch.cuda.synchronize()
            torch.cuda._sleep(600_000)           # ~0.3 ms: the CPU queues the launch meanwhile
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run(e1.record)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        t = float(np.median(ts))
        pts.append((nbytes, t))
        print(f"{name} batch {batch:4d}: {t:8.1f} us  {nbytes / t / 1e3:8.1f} GB/s  (min {min(ts):.1f})")
    x = np.array([p[0] for p in pts], dtype=np.float64)
    y = np.array([p[1] for p in pts], dtype=np.float64)
    b, a = np.polyfit(x, y, 1)
    print(json.dumps({"config": name, "fused_decode": fused, "fixed_us": a, "asymptotic_GBps": 1e-3 / b}))


if __name__ == "__main__":
    main()
