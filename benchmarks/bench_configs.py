#!/usr/bin/env python
"""Kernel-level measurements of the other BASELINE.json configs (not the driver's bench contract):

    python benchmarks/bench_configs.py [v2 v3 v4 nms kmeans map cpu] [--json out.json]

CUDA-event timings after warm-up, inputs resident in HBM; every number is printed with the
algorithmic bytes / work it is measured against.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tf2_yolo_b200 import engine, synth  # noqa: E402
from tf2_yolo_b200._native import YB_DIST_IOU  # noqa: E402
from tf2_yolo_b200.grid_loss import fused_losses  # noqa: E402

PEAK = 6549.8
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])


def timed(fn, warm=3, reps=10, flush=None, burst=1):
    """Median / min CUDA-event time of one call.  burst > 1 queues that many calls between the two
    events (inputs larger than L2 only), so the host's launch latency is hidden behind the GPU work
    as it is in a real loop; burst == 1 with a flush buffer is for working sets that fit L2."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(burst):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / burst)
    return float(np.median(ts)), float(np.min(ts))


def wrap(version):
    import importlib
    pkg = {2: "yolov2", 3: "yolov3", 4: "yolov4"}[version]
    return importlib.import_module(f"tf2_yolo_b200.{pkg}.losses").wrap_yolo_loss


def grid_config(name, version, out):
    cfg = synth.make_config(name, seed=version)
    B, C, batch = cfg["bbox_num"], cfg["class_num"], cfg["batch"]
    fns = []
    for si, S in enumerate(cfg["grids"]):
        kw = dict(anchors=cfg["anchors"][si * B:(si + 1) * B])
        kw["loss_weight"] = [1, 5, 1] if version == 4 else [1, 1, 5, 1]
        fns.append(wrap(version)((S, S), B, C, **kw))
    yt = [torch.from_numpy(a).cuda() for a in cfg["y_trues"]]
    yp = [torch.from_numpy(a).cuda() for a in cfg["y_preds"]]
    dp = [torch.empty_like(a) for a in yp]
    ch = 5 + C
    loss_bytes = sum(4 * batch * s * s * (2 * B * ch + ch) for s in cfg["grids"])
    dec_bytes = sum(4 * batch * s * s * B * ch for s in cfg["grids"])
    small = loss_bytes < 126e6
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if small else None
    rows = torch.empty((4096 * batch, 7), dtype=torch.float64, device="cuda")
    burst = 1 if small else 8
    t_loss, t_loss_min = timed(lambda: fused_losses(fns, yt, yp, dpreds=dp), flush=flush, burst=burst)
    t_dec, _ = timed(lambda: engine.decode_batch(yp, C, 0.5, version, rows=rows), flush=flush, burst=burst)
    _, offs = engine.decode_batch(yp, C, 0.5, version, rows=rows)
    mode = 2 if version == 4 else 1
    t_nms, _ = timed(lambda: engine.nms_batch(rows, offs, C, 0.45, mode))
    # the train-and-evaluate step in two launches, and inference decode + NMS in two (count, per-image kernel)
    params = [f.params for f in fns]
    fout = dict(out_rows=torch.empty((2048 * batch, 7), dtype=torch.float64, device="cuda"),
                out_offsets=torch.empty(batch + 1, dtype=torch.int64, device="cuda"),
                n_overflow=torch.zeros(1, dtype=torch.int32, device="cuda"))
    step = lambda: engine.loss_decode_nms_fused(params, yt, yp, 0.5, 0.45, mode, rows_per_img_cap=2048, dpreds=dp, out=fout)  # noqa: E731
    t_step, _ = timed(step, flush=flush, burst=burst)
    assert int(fout["n_overflow"].item()) == 0
    t_inf, _ = timed(lambda: engine.decode_nms_batch(yp, C, 0.5, version, 0.45, mode, rows_per_img_cap=2048, out=fout),
                     flush=flush, burst=burst)
    # the same step over static buffers replayed from a CUDA graph, workspaces zeroed once (no memsets):
    # what a training loop with fixed shapes runs (engine.TrainEvalStep)
    t_graph = None
    try:
        st = engine.TrainEvalStep(params, yt, yp, 0.5, 0.45, mode, rows_per_img_cap=2048, dpreds=dp)
        t_graph, _ = timed(st.run, flush=flush, burst=burst)
        assert torch.equal(st.out_offsets, fout["out_offsets"]) and int(st.n_overflow.item()) == 0
    except Exception as e:  # noqa: BLE001
        t_graph = f"graph capture failed: {type(e).__name__}: {e}"
        torch.cuda.synchronize()
    out[name] = {
        "step_two_launches_ms": t_step, "step_graph_replay_ms": t_graph, "decode_nms_two_launches_ms": t_inf,
        "step_images_per_s": batch / (t_step * 1e-3), "step_frac_of_measured_hbm": loss_bytes / t_step / 1e6 / PEAK,
        "graph_step_images_per_s": batch / (t_graph * 1e-3) if isinstance(t_graph, float) else None,
        "graph_step_frac_of_measured_hbm": loss_bytes / t_graph / 1e6 / PEAK if isinstance(t_graph, float) else None,
        "batch": batch, "loss_ms": t_loss, "loss_GBps": loss_bytes / t_loss / 1e6,
        "loss_frac_of_measured_hbm": loss_bytes / t_loss / 1e6 / PEAK, "loss_bytes": loss_bytes,
        "decode_ms": t_dec, "decode_GBps": dec_bytes / t_dec / 1e6, "decode_frac": dec_bytes / t_dec / 1e6 / PEAK,
        "nms_ms": t_nms, "rows_per_image": int(offs[-1]) / batch,
        "images_per_s_loss_decode_nms": batch / ((t_loss + t_dec + t_nms) * 1e-3),
        "l2": "flushed between iterations (working set < L2), one call per timing" if small else "inputs larger than L2, 8 calls queued per timing",
    }
    print(name, json.dumps(out[name]))


def nms_stress(out, n_img=16, per_img=100_000, C=80):
    rng = np.random.default_rng(4)
    rows = np.concatenate([synth.make_dense_candidates(rng, per_img, C) for _ in range(n_img)])
    offs = torch.arange(0, (n_img + 1) * per_img, per_img, dtype=torch.int64, device="cuda")
    dev = torch.from_numpy(rows).cuda()
    for mode in (1, 2):
        t, tmin = timed(lambda: engine.nms_batch(dev, offs, C, 0.45, mode), warm=1, reps=3)
        res = engine.nms_batch(dev, offs, C, 0.45, mode)
        pairs = n_img * C * (per_img / C) ** 2 / 2
        out[f"nms_dense_mode{mode}"] = {"images": n_img, "candidates_per_image": per_img, "ms": t,
                                        "images_per_s": n_img / (t * 1e-3),
                                        "upper_bound_pair_iou_per_s": pairs / (t * 1e-3),
                                        "kept_per_image": int(res["out_offsets"][-1]) / n_img}
        print(f"nms_dense_mode{mode}", json.dumps(out[f"nms_dense_mode{mode}"]))


def kmeans_bench(out, n=50_000_000, k=9):
    rng = np.random.default_rng(4)
    data = torch.from_numpy(synth.make_kmeans_boxes(rng, n, k)).cuda()
    centers = torch.from_numpy(np.sort(rng.uniform(0.02, 0.8, (k, 2)), axis=0)).cuda()
    t, tmin = timed(lambda: engine.kmeans_assign(data, centers, YB_DIST_IOU), burst=8)
    t2, _ = timed(lambda: engine.kmeans_assign(data, centers, YB_DIST_IOU, want_assign=True), burst=8)
    # one iteration of the device Lloyd loop (assignment + update + stop test in one launch)
    loop = engine.KMeansLloyd(data, centers.clone(), YB_DIST_IOU, 0.0, 1 << 40)
    t3, _ = timed(loop.step, burst=8)
    t4, _ = timed(lambda: loop.step_many(8), burst=1)
    t4 /= 8
    out["kmeans_50M"] = {"boxes": n, "k": k, "ms_per_iteration": t, "GBps": 16 * n / t / 1e6,
                         "frac_of_measured_hbm": 16 * n / t / 1e6 / PEAK, "ms_with_assignments": t2,
                         "ms_per_lloyd_iteration_device_loop": t3, "ms_per_lloyd_iteration_graph_batches": t4,
                         "lloyd_frac_of_measured_hbm": 16 * n / t3 / 1e6 / PEAK}
    from tf2_yolo_b200.utils import kmeans as km
    import contextlib
    import io
    km.kmeans(data[:100_000], k, km.iou_dist, 1e-5, verbose=False)     # warm
    np.random.seed(4)
    buf = io.StringIO()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(buf):
        c = km.kmeans(data, k, km.iou_dist, 1e-5, verbose=True)
    dt = time.perf_counter() - t0
    iters = len([ln for ln in buf.getvalue().split("\n") if ln.startswith("epoch")])
    out["kmeans_50M"].update(full_run_s=dt, full_run_iterations=iters, full_run_ms_per_iteration=1e3 * dt / max(iters, 1))
    print("kmeans_50M", json.dumps(out["kmeans_50M"]), c[:3].tolist())


def map_bench(out, n_img=512):
    from tf2_yolo_b200.utils import measurement as meas
    cfg = synth.make_config("v4-608", batch=n_img, seed=5)
    names = [str(i) for i in range(80)]
    yt = cfg["y_trues"][-1]
    t0 = time.perf_counter()
    pr = meas.PRfunc(yt, *cfg["y_preds"], class_names=names, conf_threshold=0.05, version=4)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # the same with the tensors already on the device (what a predict() on the GPU hands over)
    d_yt = torch.from_numpy(yt).cuda()
    d_yp = [torch.from_numpy(p).cuda() for p in cfg["y_preds"]]
    meas.PRfunc(d_yt, *d_yp, class_names=names, conf_threshold=0.05, version=4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pr2 = meas.PRfunc(d_yt, *d_yp, class_names=names, conf_threshold=0.05, version=4)
    torch.cuda.synchronize()
    dt2 = time.perf_counter() - t0
    same = bool(np.array_equal(pr.get_map()["ap"].values, pr2.get_map()["ap"].values))
    out["prfunc_v4_608"] = {"images": n_img, "seconds_host_inputs": dt, "images_per_s": n_img / dt,
                            "seconds_device_inputs": dt2, "images_per_s_device_inputs": n_img / dt2,
                            "same_table": same, "mAP_voc2012": float(pr.get_map()["ap"].iloc[-1])}
    print("prfunc", json.dumps(out["prfunc_v4_608"]))


def cpu_baselines(out):
    """The CPU side of BASELINE.md section 3 for configs 4 and 5: the oracle port of the reference's
    NumPy code (effectively one core: Python loops around NumPy) on bounded samples, with the law
    used to extrapolate stated next to each number.  Test infrastructure used as a timed baseline
    only (like bench.py's cpu_baseline)."""
    from oracle import kmeans as okm
    from oracle import measurement as om
    from oracle import tools as ot
    cores = os.cpu_count()
    rng = np.random.default_rng(4)
    rows = synth.make_dense_candidates(rng, 100_000, 80)
    for mode in (1, 2):
        t0 = time.perf_counter()
        with np.errstate(invalid="ignore", divide="ignore"):
            ot.nms(rows, 80, 0.45, mode)
        dt = time.perf_counter() - t0
        out[f"cpu_nms_dense_mode{mode}"] = {"images": 1, "seconds": dt, "images_per_s": 1 / dt, "cores_used": 1,
                                            "host_cores": cores, "law": "linear in images"}
        print(f"cpu_nms_dense_mode{mode}", json.dumps(out[f"cpu_nms_dense_mode{mode}"]))
    n = 5_000_000
    data = synth.make_kmeans_boxes(rng, n, 9)
    centers = np.sort(rng.uniform(0.02, 0.8, (9, 2)), axis=0)
    t0 = time.perf_counter()
    okm.lloyd_step(data, centers, okm.iou_dist, data.min(), data.max())
    dt = time.perf_counter() - t0
    out["cpu_kmeans"] = {"boxes": n, "seconds_per_iteration": dt, "extrapolated_ms_per_iteration_50M": dt * 10 * 1e3,
                         "cores_used": 1, "host_cores": cores, "law": "linear in boxes (x10 for 50 M)"}
    print("cpu_kmeans", json.dumps(out["cpu_kmeans"]))
    n_img = 4
    cfg = synth.make_config("v4-608", batch=n_img, seed=5)
    names = [str(i) for i in range(80)]
    t0 = time.perf_counter()
    with np.errstate(invalid="ignore", divide="ignore"):
        om.PRfunc(cfg["y_trues"][-1], *cfg["y_preds"], class_names=names, conf_threshold=0.05, version=4)
    dt = time.perf_counter() - t0
    out["cpu_prfunc"] = {"images": n_img, "seconds": dt, "images_per_s": n_img / dt, "cores_used": 1,
                         "host_cores": cores, "law": "phase 1 (decode + NMS + match) linear in images"}
    print("cpu_prfunc", json.dumps(out["cpu_prfunc"]))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="*", default=["v2", "v3", "v4", "nms", "kmeans", "map"])
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    out = {"peak_hbm_GBps": PEAK}
    if "v2" in a.what:
        grid_config("v2-416", 2, out)
    if "v3" in a.what:
        grid_config("v3-416", 3, out)
    if "v4" in a.what:
        grid_config("v4-608", 4, out)
    if "nms" in a.what:
        nms_stress(out)
    if "kmeans" in a.what:
        kmeans_bench(out)
    if "map" in a.what:
        map_bench(out)
    if "cpu" in a.what:
        cpu_baselines(out)
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)
