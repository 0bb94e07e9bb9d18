#!/usr/bin/env python
"""Turns the raw evidence of a gpurun measurement call into the tables committed under profiles/:

  python scripts/summarize_profiles.py launches gpurun_out/r02_launches_step.csv profiles/r02_launches_step_summary.md
      ncu --csv launch list (gpu__time_duration + dram bytes per launch) -> per-kernel table (launches, ms, share, DRAM
      GB, achieved GB/s) + the per-launch GEMM DRAM traffic json bench.py reports as roofline.traffic
  python scripts/summarize_profiles.py gemm gpurun_out/r02_gemm_launches.json profiles/r02_gemm_shapes.md
      bench.py --dump-gemm-profile (CUDA events around every GEMM launch of one step) -> per-shape-class table with each
      class's own roofline (max(flops / peak, minimal bytes / peak bandwidth))
"""
import csv
import json
import os
import re
import subprocess
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return j["bf16_tflops_sustained"], j["hbm_gbs"]
    return 1400.0, 6650.0


def commit():
    try:
        return subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    except OSError:
        return "unknown"


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:70]


def launches(src, dst):
    rows = OrderedDict()
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        d = rows.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "byte": 1.0,
                                                                         "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1.0)
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for d in rows.values():
        a = agg[short(d["name"])]
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tot_ns = sum(a[1] for a in agg.values())
    tot_b = sum(a[2] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list of ONE optimisation step (R50-L1-H1024, batch 256, 1x B200) at commit {commit()}\n\n")
        f.write("Command: `ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                "--clock-control none --csv python scripts/profile_step.py` (raw list next to this file). Per-launch times under ncu "
                "are serialised / cold-cache: compare SHARES.\n\n")
        f.write(f"Sum of kernel time {tot_ns / 1e6:.2f} ms over {len(rows)} launches; DRAM traffic {tot_b / 1e9:.1f} GB per step.\n\n")
        f.write("| kernel | launches | ms | share | DRAM GB | GB/s |\n|---|---:|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / tot_ns:.1f}% | {a[2] / 1e9:.2f} | "
                    f"{(a[2] / a[1] if a[1] else 0):.0f} |\n")
    g = [d for d in rows.values() if "gemm_tc_kernel" in d["name"]]
    if g:
        out = {"commit": commit(), "launches": len(g),
               "dram_bytes_per_launch": sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in g) / len(g),
               "source": os.path.basename(src)}
        tag = os.path.basename(dst).split("_")[0]  # r02s_launches_step_summary.md -> r02s_gemm_dram_traffic.json
        with open(os.path.join(os.path.dirname(dst), f"{tag}_gemm_dram_traffic.json"), "w") as f:
            json.dump(out, f, indent=1)
    print(f"{len(rows)} launches, {tot_ns / 1e6:.2f} ms, {tot_b / 1e9:.1f} GB -> {dst}")


def min_bytes(M, N, K, mode, a_mn, b_mn, extra=0):
    return _min_bytes(M, N, K, mode, a_mn, b_mn) + extra


def _min_bytes(M, N, K, mode, a_mn, b_mn):
    if mode in (1, 3, 5):
        return 2 * (M * K // (9 if mode != 5 else 16) + N * K) + 2 * M * N
    if mode in (2, 6):
        return 2 * (K * M + K * N // (9 if mode != 6 else 16)) + 4 * M * N
    if mode == 4:
        return 2 * (K * N + K * 64) + 4 * M * N
    return 2 * (M * K + N * K) + (4 if (a_mn and b_mn) else 2) * M * N


def gemm(src, dst):
    peak_tf, peak_bw = peaks()
    data = json.load(open(src))
    cls = OrderedDict()
    for d in data:
        key = (d["M"], d["N"], d["K"], d["conv_mode"], d["a_mn"], d["b_mn"], d.get("extra_bytes", 0))
        c = cls.setdefault(key, [0, 0.0])
        c[0] += 1
        c[1] += d["ms"]
    tot = sum(c[1] for c in cls.values())
    tmin_all = 0.0
    rows = []
    for (M, N, K, mode, a, b, extra), (n, ms) in cls.items():
        fl = 2.0 * M * N * K
        by = min_bytes(M, N, K, mode, a, b, extra)
        tmin = max(fl / (peak_tf * 1e12), by / (peak_bw * 1e9)) * 1e3  # ms per launch
        tmin_all += tmin * n
        rows.append((ms - tmin * n, M, N, K, mode + (0.5 if extra else 0), a, b, n, ms, ms / n * 1e3, fl * n / ms / 1e9, by * n / ms / 1e6, tmin * n / ms))
    with open(dst, "w") as f:
        f.write(f"# tcgen05 GEMM launches of one step, grouped by shape (CUDA events around every launch) at commit {commit()}\n\n")
        f.write(f"Total {tot:.2f} ms over {len(data)} launches; per-launch roofline (max(flops / {peak_tf:.0f} TFLOP/s, minimal bytes / "
                f"{peak_bw:.0f} GB/s)) sums to {tmin_all:.2f} ms -> {tmin_all / tot:.3f}. mode: 0 plain, 1 implicit 3x3 fprop/dgrad "
                "(64->64 shapes run the halo variant), 2 implicit 3x3 wgrad, 4 halo-reuse wgrad, 5/6 stem fprop/wgrad, +.5: the epilogue "
                "also reads a residual tile (and ReLU bit mask), counted in the minimal bytes; a/b = operand "
                "MN-major flags. Sorted by time lost against the class's own roofline.\n\n")
        f.write("| M | N | K | mode | a | b | launches | ms total | us/launch | TFLOP/s | min-bytes GB/s | frac of own roofline | ms lost |\n"
                "|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for lost, M, N, K, mode, a, b, n, ms, us, tf, gbs, frac in sorted(rows, reverse=True):
            f.write(f"| {M} | {N} | {K} | {mode} | {a} | {b} | {n} | {ms:.3f} | {us:.1f} | {tf:.0f} | {gbs:.0f} | {frac:.2f} | {lost:.3f} |\n")
    print(f"{len(data)} GEMM launches, {tot:.2f} ms, roofline {tmin_all / tot:.3f} -> {dst}")


if __name__ == "__main__":
    {"launches": launches, "gemm": gemm}[sys.argv[1]](sys.argv[2], sys.argv[3])
