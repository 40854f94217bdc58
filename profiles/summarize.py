#!/usr/bin/env python
"""Turn ncu outputs (gpurun_out/) into the small text summaries kept under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_X.csv  > profiles/launches_X.txt
    python profiles/summarize.py kernel   gpurun_out/X.ncu-rep       > profiles/X.txt
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [(r["Kernel Name"], float(r["Metric Value"]), r["Grid Size"], r["Block Size"]) for r in rows]
    loss = [i for i, n in enumerate(names) if "loss_fwd_bwd" in n[0]]
    # one device-resident timed step = from one loss launch to the next, taken inside the longest
    # run of identical steps (round 2: the graph replays of the timed region)
    start, end = (loss[4], loss[5]) if len(loss) > 5 else (loss[0], len(names))
    gaps = [(loss[k + 1] - loss[k], k) for k in range(len(loss) - 1)]
    for k in range(1, len(gaps) - 1):
        if gaps[k - 1][0] == gaps[k][0] == gaps[k + 1][0] and \
                [n[0] for n in names[loss[k]:loss[k + 1]]] == [n[0] for n in names[loss[k - 1]:loss[k]]] and \
                "encode_labels" not in names[loss[k] - 1][0]:
            start, end = loss[k], loss[k + 1]
            break
    step = names[start:end]
    total = sum(v for _, v, _, _ in step)
    print(f"# one step of `bench.py --steps 2 --warmup 3` under ncu (gpu__time_duration, cold cache, serialised)")
    print(f"# {len(step)} launches, {total / 1000:.1f} us in kernels")
    for n, v, g, b in step:
        print(f"{v / 1000:9.1f} us  {100 * v / total:5.1f}%  grid {g:>13s} block {b:>12s}  {n[:100]}")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_col = hdr.index("Kernel Name")
    for d in data:
        print(f"## {d[name_col][:110]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:80s} {d[i]:>18s} {units[i]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
