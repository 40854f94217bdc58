#!/usr/bin/env python
"""Host-to-device copy ceiling of the box: plain pinned cudaMemcpyAsync of 1 GB per rank, every
rank at once (one process per GPU), nothing else running.  This is the bound of bench.py's `e2e`
number (the step's head outputs must cross PCIe every step); run it at N = 1, 2, 4, 8:

    python benchmarks/h2d_ceiling.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29520 benchmarks/h2d_ceiling.py [--json out.json]

Prints GB/s per GPU (slowest rank), the sum over the ranks and each rank's own rate.
"""
import argparse
import json
import os

import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--bytes", type=int, default=1 << 30)
ap.add_argument("--reps", type=int, default=8)
ap.add_argument("--json", default=None)
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
host = torch.empty(a.bytes, dtype=torch.uint8).pin_memory()
host.fill_(rank + 1)
devbuf = torch.empty(a.bytes, dtype=torch.uint8, device=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for direction in ("h2d", "d2h"):
    src, dst = (host, devbuf) if direction == "h2d" else (devbuf, host)
    dst.copy_(src, non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        dst.copy_(src, non_blocking=True)      # one cudaMemcpyAsync per copy (pinned <-> device)
    e1.record()
    barrier()
    gbs = a.bytes * a.reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    t = torch.tensor([gbs], dtype=torch.float64, device=dev)
    every = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(every, t)
    else:
        every = [t]
    rates = [float(x) for x in every]
    if rank == 0:
        out = {"direction": direction, "n_gpus": world, "bytes_per_copy": a.bytes, "reps": a.reps,
               "GBps_per_gpu_slowest": min(rates), "GBps_sum": sum(rates), "GBps_each": [round(r, 2) for r in rates]}
        print(json.dumps(out))
        if a.json:
            with open(a.json, "a") as f:
                f.write(json.dumps(out) + "\n")
if world > 1:
    dist.destroy_process_group()
