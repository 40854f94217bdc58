#!/usr/bin/env python
"""Where the two-launch step's time goes: device time of each launch alone (CUDA events around
back-to-back calls) and the host's cost of issuing a step (Python + ctypes), batch 128 v4-608."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tf2_yolo_b200 import _native as N  # noqa: E402
from tf2_yolo_b200 import engine, synth  # noqa: E402
from tf2_yolo_b200.yolov4.losses import wrap_yolo_loss  # noqa: E402

batch = int(os.environ.get("YB_BATCH", 128))
cfg = synth.make_config("v4-608", batch=batch, seed=2)
B, Cn = 3, 80
fns = [wrap_yolo_loss((S, S), B, Cn, anchors=cfg["anchors"][si * B:(si + 1) * B], loss_weight=[1, 5, 1])
       for si, S in enumerate(cfg["grids"])]
params = [f.params for f in fns]
yts = [torch.from_numpy(a).cuda() for a in cfg["y_trues"]]
yps = [torch.from_numpy(a).cuda() for a in cfg["y_preds"]]
dps = [torch.empty_like(a) for a in yps]
out = dict(out_rows=torch.empty((1024 * batch, 7), dtype=torch.float64, device="cuda"),
           out_offsets=torch.empty(batch + 1, dtype=torch.int64, device="cuda"),
           n_overflow=torch.zeros(1, dtype=torch.int32, device="cuda"))


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / reps
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, host * 1e6     # device us, host-issue us


res = {}
res["step_two_launches"] = timed(lambda: engine.loss_decode_nms_fused(params, yts, yps, 0.5, 0.45, 2, dpreds=dps, out=out))
res["loss_only"] = timed(lambda: engine.loss_fwd_bwd(params, yts, yps, dpreds=dps))
# the per-image kernel alone: buckets filled once, then the finish call back to back
dparams, _ = engine.make_decode_params(yps, Cn, 0.5, 4)
fws_bytes = N.lib.yb_decode_nms_workspace_bytes(C.byref(dparams), batch, 1024)
fws = engine.workspaces.get("decode_nms", fws_bytes, yps[0].device)
ptrs = (C.c_void_p * 3)(*[t.data_ptr() for t in yps])
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def finish():
    # the kernel consumes its ticket / look-back words: clear them (the memset is part of a real step too)
    N.check(N.lib.yb_decode_nms(ptrs, batch, C.byref(dparams), 0.45, 2, 1024, C.c_void_p(out["out_rows"].data_ptr()),
                                out["out_rows"].shape[0], C.c_void_p(out["out_offsets"].data_ptr()),
                                C.c_void_p(out["n_overflow"].data_ptr()), C.c_void_p(fws.data_ptr()), fws_bytes, st))


res["decode_count_plus_per_image_kernel"] = timed(finish)
res["decode_count_only"] = timed(lambda: engine.decode_batch(yps, Cn, 0.5, 4, rows=out["out_rows"]))
res["chain_decode_nms"] = timed(lambda: engine.nms_batch(*engine.decode_batch(yps, Cn, 0.5, 4, rows=out["out_rows"]), Cn, 0.45, 2))
print(json.dumps({k: {"device_us": round(v[0], 1), "host_issue_us": round(v[1], 1)} for k, v in res.items()}, indent=1))
