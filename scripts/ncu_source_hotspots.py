#!/usr/bin/env python
"""Per-CUDA-source-line hot spots of one kernel launch of an .ncu-rep (needs -lineinfo + --import-source on):
   python scripts/ncu_source_hotspots.py <report.ncu-rep> <launch index> [top N]
Aggregates warp-stall samples and executed warp instructions per source line (ncu --page source --print-source cuda,sass)."""
import csv
import subprocess
import sys

rep, k = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(k),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hdr_i]
ci = {n: i for i, n in enumerate(h)}
samp, inst = ci["# Samples"], ci["Instructions Executed"]
stall_cols = [(n, i) for n, i in ci.items() if n.startswith("stall_") and "Not Issued" not in n]
data = []
cur_file = ""
for r in rows[hdr_i + 1:]:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) <= max(samp, inst) or r[2] != "-":  # keep the per-line summary rows (Address == "-"), skip SASS rows
        continue
    try:
        s, n = float(r[samp] or 0), float(r[inst] or 0)
    except ValueError:
        continue
    if s > 0 or n > 0:
        st = sorted(((float(r[i] or 0), nm) for nm, i in stall_cols), reverse=True)[:2]
        data.append((s, n, cur_file, r[0], r[1].strip()[:100], st))
tot, totn = sum(d[0] for d in data), sum(d[1] for d in data)
print(f"launch {k}: {tot:.0f} stall samples, {totn:.0f} warp instructions")
for s, n, f, l, src, st in sorted(data, reverse=True)[:top]:
    print(f"{100 * s / tot:5.1f}% smp {100 * n / totn:5.1f}% inst  {f}:{l}: {src}   [{st[0][1]} {st[0][0]:.0f}, {st[1][1]} {st[1][0]:.0f}]")
