#!/usr/bin/env python
"""Multi-GPU runs of BASELINE.json configs 4 and 5 (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 benchmarks/bench_multi.py [nms] [kmeans] [map] [--json out.json]

    nms    : dense-scene NMS stress, images sharded across ranks, no collective
    kmeans : anchor k-means over 50 M boxes sharded across ranks; per Lloyd iteration one
             all-reduce of k*(d+1) doubles; the centres are compared with the unsharded run
    map    : PRfunc over images sharded across ranks (all-gather of ground-truth counts; the records
             gathered everywhere or routed to their class owner); the mAP table is compared with the
             unsharded run
    map5   : config 5 at its stated size (50k images streamed in device chunks), class-partitioned

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
    # one sharded Lloyd iteration of the device loop, timed on the device: assignment pass, the
    # all-reduce of k*(d+1) doubles on the same stream, the update kernel - no host sync inside
    c0 = np.sort(rng.uniform(0.02, 0.8, (k, 2)), axis=0)
    def fresh_loop():
        return engine.KMeansLloyd(shard, torch.from_numpy(c0).to(dev), YB_DIST_IOU, 0.0, 1 << 40, sharded=world > 1)
    def one_iter(loop):
        loop.step()
        if world > 1:
            ydist.allreduce_sum(loop.packed, group)
            loop.update()
    loop = fresh_loop()
    for _ in range(5):
        one_iter(loop)
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 40
    e0.record()
    for _ in range(reps):
        one_iter(loop)
    e1.record()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1) / reps, world, dev)
    # the same iterations replayed from a CUDA graph (the host's launch overhead out of the picture)
    ms_graph = None
    try:
        loop = fresh_loop()
        for _ in range(3):
            one_iter(loop)
        barrier(world)
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for _ in range(8):
                    one_iter(loop)
        torch.cuda.current_stream().wait_stream(side)
        for _ in range(2):
            g.replay()
        barrier(world)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        barrier(world)
        ms_graph = max_over_ranks(e0.elapsed_time(e1) / 40, world, dev)
    except Exception as e:  # noqa: BLE001
        ms_graph = f"graph capture failed: {type(e).__name__}: {e}"
        torch.cuda.synchronize()
    # the all-reduce inside the assignment launch, over NVLink peer memory (one launch per iteration)
    ms_peer, ms_peer_graph, t_full_peer, peer_same = None, None, None, None
    if world > 1:
        loop_p = engine.KMeansLloyd(shard, torch.from_numpy(c0).to(dev), YB_DIST_IOU, 0.0, 1 << 40, sharded=True,
                                    peer_group=group)
        for _ in range(5):
            loop_p.step()
        barrier(world)
        e0.record()
        for _ in range(reps):
            loop_p.step()
        e1.record()
        barrier(world)
        ms_peer = max_over_ranks(e0.elapsed_time(e1) / reps, world, dev)
        # the same from CUDA-graph batches of 8 iterations (what utils.kmeans queues between two looks)
        for _ in range(2):
            loop_p.step_many(8)
        barrier(world)
        e0.record()
        for _ in range(5):
            loop_p.step_many(8)
        e1.record()
        barrier(world)
        ms_peer_graph = max_over_ranks(e0.elapsed_time(e1) / 40, world, dev)
        peer_same = int(loop_p.read_state()[0])      # 0 = running (no exchange timed out)
        barrier(world)
        loop_p.close()
        np.random.seed(4)
        barrier(world)
        t0 = time.perf_counter()
        c_peer = km.kmeans(shard, k, km.iou_dist, 1e-5, verbose=False, process_group=group, exchange="peer")
        torch.cuda.synchronize()
        t_full_peer = time.perf_counter() - t0
    # full run: the reference's loop with its RNG stream
    np.random.seed(4)
    barrier(world)
    t0 = time.perf_counter()
    trace = []
    c_sharded = km.kmeans(shard, k, km.iou_dist, 1e-5, verbose=False, process_group=group)
    torch.cuda.synchronize()
    t_full = time.perf_counter() - t0
    if rank == 0:
        np.random.seed(4)
        dev_full = torch.from_numpy(full).to(dev)
        km.kmeans(dev_full[:1000], k, km.iou_dist, 1e-5, verbose=False)    # warm
        np.random.seed(4)
        t0 = time.perf_counter()
        import io, contextlib
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            c_single = km.kmeans(dev_full, k, km.iou_dist, 1e-5, verbose=True)
        t_single = time.perf_counter() - t0
        n_iter = len([ln for ln in buf.getvalue().split("\n") if ln.startswith("epoch")])
        # fp64 partial sums are reduced in a different order across ranks: equal to ~1e-15 relative, so
        # the float32 centres are normally bit-identical; report both
        out["kmeans_50M"] = {
            "n_gpus": world, "boxes": n, "k": k, "ms_per_iteration": ms, "ms_per_iteration_cuda_graph": ms_graph,
            "boxes_per_s": n / (ms * 1e-3), "GBps_aggregate": 16 * n / ms / 1e6,
            "ms_per_iteration_peer_exchange_in_kernel": ms_peer,
            "ms_per_iteration_peer_exchange_cuda_graph": ms_peer_graph, "peer_exchange_status": peer_same,
            "full_run_s_sharded_peer_exchange": t_full_peer,
            "centres_peer_equal_nccl": bool(np.array_equal(c_peer, c_sharded)) if world > 1 else None,
            "collective": f"all-reduce of {k * 3} doubles per iteration, same stream, no host sync",
            "full_run_s_sharded": t_full, "full_run_s_single_gpu": t_single, "iterations": n_iter,
            "single_gpu_ms_per_iteration_whole_run": 1e3 * t_single / max(n_iter, 1),
            "centres_equal_single_gpu_run": bool(np.array_equal(c_single, c_sharded)),
            "max_rel_centre_diff": float(np.max(np.abs(c_single - c_sharded) / np.abs(c_single))),
            "centres": c_sharded.tolist()}
        del dev_full
    del shard
    barrier(world)


def run_map(out, rank, world, dev, group, n_img):
    cfg = synth.make_config("v4-608", batch=n_img, seed=5)       # same on every rank
    names = [str(i) for i in range(80)]
    a, b = ydist.shard_range(n_img, rank, world)
    yt = cfg["y_trues"][-1][a:b]
    yps = [p[a:b] for p in cfg["y_preds"]]
    res = {}
    for part in (False, True):
        barrier(world)
        t0 = time.perf_counter()
        pr = meas.PRfunc(yt, *yps, class_names=names, conf_threshold=0.05, version=4, process_group=group,
                         partition_classes=part)
        tab = pr.get_map()
        barrier(world)
        res[part] = (max_over_ranks(time.perf_counter() - t0, world, dev), tab)
    if rank == 0:
        t0 = time.perf_counter()
        pr1 = meas.PRfunc(cfg["y_trues"][-1], *cfg["y_preds"], class_names=names, conf_threshold=0.05, version=4)
        tab1 = pr1.get_map()
        dt1 = time.perf_counter() - t0
        out["prfunc_v4_608"] = {
            "n_gpus": world, "images": n_img, "seconds_sharded_gather_all": res[False][0],
            "seconds_sharded_class_partitioned": res[True][0], "images_per_s": n_img / res[True][0],
            "seconds_single_gpu": dt1, "mAP_voc2012": float(tab1["ap"].iloc[-1]),
            "ap_table_equals_single_gpu_run": bool(np.array_equal(res[False][1]["ap"].values, tab1["ap"].values)),
            "ap_table_class_partitioned_equals_single_gpu_run":
                bool(np.array_equal(res[True][1]["ap"].values, tab1["ap"].values)),
            "collective": "all-gather of per-class GT counts; records all-to-all to the class owner (class % world)"}
    barrier(world)


def run_map_config5(out, rank, world, dev, group, n_img_total, pool=256, check=128, oracle_check=False):
    """BASELINE config 5 at its stated size: PR curves / mAP over n_img_total v4-608 images sharded
    across the ranks.  The images never exist as whole arrays (50k images are 387 GB of head
    outputs): every rank streams device chunks of its slice of ONE global image sequence, drawn
    with repetition from a pool of `pool` synthetic images (same pool on every rank).  Rank 0 then
    evaluates the whole sequence alone and the two AP tables are compared; with `oracle_check` the
    first `check` images are also evaluated by the CPU oracle."""
    from oracle import measurement as om
    cfg = synth.make_config("v4-608", batch=pool, seed=50)
    names = [str(i) for i in range(80)]
    a, b = ydist.shard_range(n_img_total, rank, world)
    d_true = torch.from_numpy(cfg["y_trues"][-1].astype(np.float64)).to(dev)
    d_preds = [torch.from_numpy(p).to(dev) for p in cfg["y_preds"]]
    chunk = 250
    order = np.random.default_rng(7).integers(0, pool, n_img_total)
    order[:check] = np.arange(check)           # the oracle-checked subsample comes first

    def source_of(lo, hi):
        def source():
            for s in range(lo, hi, chunk):
                idx = torch.from_numpy(order[s:min(hi, s + chunk)]).to(dev)
                yield d_true[idx], [p[idx] for p in d_preds]
        return source
    kw = dict(class_names=names, conf_threshold=0.05, nms_mode=1, nms_threshold=0.5, max_per_img=100, version=4)
    # warm-up on a slice of this rank's images, through the same collectives (NCCL sets up its
    # all-to-all connections on first use: seconds at 8 ranks), then the timed run
    meas.PRfunc(None, process_group=group, partition_classes=world > 1,
                chunk_source=source_of(a, min(b, a + 2 * chunk)), **kw).get_map()
    torch.cuda.synchronize()
    barrier(world)
    t0 = time.perf_counter()
    pr = meas.PRfunc(None, process_group=group, partition_classes=world > 1, chunk_source=source_of(a, b), **kw)
    t1 = time.perf_counter()
    tab = pr.get_map()
    barrier(world)
    dt = max_over_ranks(time.perf_counter() - t0, world, dev)
    dt_curves = max_over_ranks(t1 - t0, world, dev)
    split = {}      # the same once more with the device synchronised between the phases
    meas.PRfunc(None, process_group=group, partition_classes=world > 1, chunk_source=source_of(a, b), timings=split, **kw)
    barrier(world)
    n_rec = sum(len(p) - 1 for p in pr.precisions if p is not None)
    n_rec = int(max_over_ranks(float(n_rec), world, dev))
    if rank == 0:
        res = {
            "n_gpus": world, "images": n_img_total, "images_per_rank": b - a, "pool_images": pool,
            "seconds_total": dt, "seconds_until_curves": dt_curves, "images_per_s": n_img_total / dt,
            "records_in_largest_rank": n_rec, "mAP_voc2012": float(tab["ap"].iloc[-1]),
            "phase_split_rank0_s (second run, device synchronised between phases)": split,
            "collective": "all-gather of per-class GT counts; records all-to-all to the class owner (class % world)"}
        if world > 1:
            t0 = time.perf_counter()
            one = meas.PRfunc(None, chunk_source=source_of(0, n_img_total), **kw).get_map()
            res["seconds_single_gpu_same_sequence"] = time.perf_counter() - t0
            res["ap_table_equals_single_gpu_run"] = bool(np.array_equal(one["ap"].values, tab["ap"].values))
        if oracle_check:
            sub_t = cfg["y_trues"][-1][:check].astype(np.float64)
            sub_p = [p[:check] for p in cfg["y_preds"]]
            t0 = time.perf_counter()
            got = meas.PRfunc(sub_t, *sub_p, **kw).get_map()["ap"].values
            res["subsample_seconds_gpu"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            ref = om.PRfunc(sub_t, *sub_p, **kw).get_ap("voc2012")
            res["subsample_seconds_cpu_oracle"] = time.perf_counter() - t0
            res["subsample_images"] = check
            res["subsample_ap_table_equals_cpu_oracle"] = bool(np.allclose(got, ref, rtol=0, atol=1e-15))
        out["prfunc_config5"] = res
    barrier(world)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="*", default=["nms", "kmeans", "map"])
    ap.add_argument("--json", default=None)
    ap.add_argument("--nms-images-per-gpu", type=int, default=32)
    ap.add_argument("--map-images", type=int, default=1024)
    ap.add_argument("--map5-images", type=int, default=50_000)
    ap.add_argument("--kmeans-boxes", type=int, default=50_000_000)
    ap.add_argument("--oracle-check", action="store_true", help="map5: also run the CPU oracle on a 128-image subsample")
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
        run_kmeans(out, rank, world, dev, group, n=a.kmeans_boxes)
    if "map" in a.what:
        run_map(out, rank, world, dev, group, a.map_images)
    if "map5" in a.what:
        run_map_config5(out, rank, world, dev, group, a.map5_images, oracle_check=a.oracle_check)
    if rank == 0:
        print(json.dumps(out))
        if a.json:
            json.dump(out, open(a.json, "w"), indent=1)
    if world > 1:
        import threading
        dog = threading.Timer(60.0, os._exit, (0,))   # a stuck teardown must not hold the launcher
        dog.daemon = True
        dog.start()
        dist.destroy_process_group()
        dog.cancel()


if __name__ == "__main__":
    main()
