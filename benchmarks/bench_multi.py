#!/usr/bin/env python
"""Multi-GPU runs of BASELINE.json configs 4 and 5 (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 benchmarks/bench_multi.py [nms] [kmeans] [map] [--json out.json]

    nms    : dense-scene NMS stress, images sharded across ranks, no collective
    kmeans : anchor k-means over 50 M boxes sharded across ranks; per Lloyd iteration one
             all-reduce of k*(d+1) doubles; the centres are compared with the unsharded run
    map    : PRfunc over images sharded across ranks (all-gather of ground-truth counts and of the
             (conf, gt_id, flag, class) records); the mAP table is compared with the unsharded run

Times are CUDA events on each rank, max over ranks.  Works with N = 1 too (no process group).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tf2_yolo_b200 import dist as ydist  # noqa: E402
from tf2_yolo_b200 import engine, synth  # noqa: E402
from tf2_yolo_b200._native import YB_DIST_IOU  # noqa: E402
from tf2_yolo_b200.utils import kmeans as km  # noqa: E402
from tf2_yolo_b200.utils import measurement as meas  # noqa: E402


def max_over_ranks(x, world, dev):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def barrier(world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run_nms(out, rank, world, dev, group, imgs_per_gpu, per_img=100_000, C=80):
    rng = np.random.default_rng(4000 + rank)
    rows = np.concatenate([synth.make_dense_candidates(rng, per_img, C) for _ in range(imgs_per_gpu)])
    offs = torch.arange(0, (imgs_per_gpu + 1) * per_img, per_img, dtype=torch.int64, device=dev)
    d = torch.from_numpy(rows).to(dev)
    for mode in (1, 2):
        for _ in range(2):
            res = engine.nms_batch(d, offs, C, 0.45, mode)
        barrier(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            res = engine.nms_batch(d, offs, C, 0.45, mode)
        e1.record()
        barrier(world)
        ms = max_over_ranks(e0.elapsed_time(e1) / reps, world, dev)
        kept = torch.tensor([float(res["out_offsets"][-1])], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(kept)
        out[f"nms_dense_mode{mode}"] = {
            "n_gpus": world, "images_per_gpu": imgs_per_gpu, "candidates_per_image": per_img, "classes": C,
            "ms": ms, "images_per_s": imgs_per_gpu * world / (ms * 1e-3),
            "kept_per_image": float(kept) / (imgs_per_gpu * world), "collective": "none"}


def run_kmeans(out, rank, world, dev, group, n=50_000_000, k=9):
    rng = np.random.default_rng(4)
    full = synth.make_kmeans_boxes(rng, n, k)           # same on every rank (seeded)
    a, b = ydist.shard_range(n, rank, world)
    shard = torch.from_numpy(full[a:b]).to(dev)
    # one sharded Lloyd iteration, timed on the device: assignment pass + the all-reduce
    centers = torch.from_numpy(np.sort(rng.uniform(0.02, 0.8, (k, 2)), axis=0)).to(dev)
    def one_iter():
        _, sums, counts = engine.kmeans_assign(shard, centers, YB_DIST_IOU)
        return ydist.allreduce_kmeans(sums, counts, group)
    for _ in range(3):
        one_iter()
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        sums, counts = one_iter()
    e1.record()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1) / reps, world, dev)
    assert int(counts.sum()) == n
    # full run: the reference's loop with its RNG stream (same seed on every rank)
    np.random.seed(4)
    t0 = time.perf_counter()
    c_sharded = km.kmeans(shard, k, km.iou_dist, 1e-5, verbose=False, process_group=group)
    torch.cuda.synchronize()
    t_full = time.perf_counter() - t0
    same = None
    if rank == 0:
        np.random.seed(4)
        t0 = time.perf_counter()
        c_single = km.kmeans(torch.from_numpy(full).to(dev), k, km.iou_dist, 1e-5, verbose=False)
        t_single = time.perf_counter() - t0
        # fp64 partial sums are reduced in a different order across ranks: equal to ~1e-15 relative, so
        # the float32 centres are normally bit-identical; report both
        same = bool(np.array_equal(c_single, c_sharded))
        out["kmeans_50M"] = {
            "n_gpus": world, "boxes": n, "k": k, "ms_per_iteration": ms,
            "boxes_per_s": n / (ms * 1e-3), "GBps_aggregate": 16 * n / ms / 1e6,
            "collective": f"all-reduce of {k * 3} doubles per iteration",
            "full_run_s_sharded": t_full, "full_run_s_single_gpu": t_single,
            "centres_equal_single_gpu_run": same,
            "max_rel_centre_diff": float(np.max(np.abs(c_single - c_sharded) / np.abs(c_single))),
            "centres": c_sharded.tolist()}
    del shard


def run_map(out, rank, world, dev, group, n_img):
    cfg = synth.make_config("v4-608", batch=n_img, seed=5)       # same on every rank
    names = [str(i) for i in range(80)]
    a, b = ydist.shard_range(n_img, rank, world)
    yt = cfg["y_trues"][-1][a:b]
    yps = [p[a:b] for p in cfg["y_preds"]]
    barrier(world)
    t0 = time.perf_counter()
    pr = meas.PRfunc(yt, *yps, class_names=names, conf_threshold=0.05, version=4, process_group=group)
    tab = pr.get_map()
    barrier(world)
    dt = max_over_ranks(time.perf_counter() - t0, world, dev)
    if rank == 0:
        t0 = time.perf_counter()
        pr1 = meas.PRfunc(cfg["y_trues"][-1], *cfg["y_preds"], class_names=names, conf_threshold=0.05, version=4)
        tab1 = pr1.get_map()
        dt1 = time.perf_counter() - t0
        out["prfunc_v4_608"] = {
            "n_gpus": world, "images": n_img, "seconds_sharded_host_inputs": dt, "images_per_s": n_img / dt,
            "seconds_single_gpu": dt1, "mAP_voc2012": float(tab["ap"].iloc[-1]),
            "ap_table_equals_single_gpu_run": bool(np.array_equal(tab["ap"].values, tab1["ap"].values)),
            "collective": "all-gather of per-class GT counts + variable-length (conf, gt_id, flag, class) records"}
    barrier(world)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="*", default=["nms", "kmeans", "map"])
    ap.add_argument("--json", default=None)
    ap.add_argument("--nms-images-per-gpu", type=int, default=32)
    ap.add_argument("--map-images", type=int, default=1024)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    out = {"n_gpus": world}
    if "nms" in a.what:
        run_nms(out, rank, world, dev, group, a.nms_images_per_gpu)
    if "kmeans" in a.what:
        run_kmeans(out, rank, world, dev, group)
    if "map" in a.what:
        run_map(out, rank, world, dev, group, a.map_images)
    if rank == 0:
        print(json.dumps(out))
        if a.json:
            json.dump(out, open(a.json, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
