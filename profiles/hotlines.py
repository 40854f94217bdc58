#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an ncu report (needs -lineinfo + --import-source on):
    python profiles/hotlines.py REPORT.ncu-rep KERNEL_REGEX [launch_skip] [top_n]
Prints the lines with the most executed warp instructions and stall samples."""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{pat}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, lines = "", None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        print("#", r[1][:120])
    elif r[0] == "Line No":
        hdr = {n: j for j, n in enumerate(r)}
    elif hdr and r[2] == "-" and r[0].isdigit():
        g = lambda k: float(r[hdr[k]] or 0)
        stalls = {k[6:]: g(k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
        lines.append((g("Instructions Executed"), g("# Samples"), g("Avg. Threads Executed"), fname, r[0], r[1].strip()[:90], stalls))
ti = sum(l[0] for l in lines) or 1
ts = sum(l[1] for l in lines) or 1
print(f"# total warp instructions {ti:.0f}, samples {ts:.0f}")
agg = {}
for l in lines:
    for k, v in l[6].items():
        agg[k] = agg.get(k, 0) + v
print("# stall mix:", ", ".join(f"{k} {100 * v / ts:.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for l in sorted(lines, key=lambda l: -l[1])[:top]:
    top_st = ",".join(f"{k}:{100 * v / max(l[1], 1):.0f}" for k, v in sorted(l[6].items(), key=lambda kv: -kv[1])[:3])
    print(f"{100 * l[1] / ts:5.1f}% samp {100 * l[0] / ti:5.1f}% inst thr/warp {l[2]:4.0f} {l[3]}:{l[4]:>4s} {l[5]}  [{top_st}]")
