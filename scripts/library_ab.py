#!/usr/bin/env python
"""Per-shape-class A/B of gemm_tc_kernel against the vendor libraries on the same B200 (measurement tooling):

    python scripts/library_ab.py profiles/r02f_gemm_launches.json > profiles/r02_library_ab.md

For every GEMM class of one optimisation step: this library's time inside the step (CUDA events of bench.py's profile,
epilogue work included: BN statistics, bias, residual, ...) next to the library call that computes the same product with
NO epilogue -- torch.matmul (cuBLASLt, bf16) for plain classes, F.conv2d / conv2d_input / conv2d_weight (cuDNN, bf16,
channels_last, cudnn.benchmark) for the convolution classes.  Median of 9 launches with an L2 flush in between.
The library numbers are an optimistic bound for eager PyTorch (which adds separate BN / bias / residual kernels).
"""
import json
import sys
from collections import OrderedDict

import torch
import torch.nn.functional as F

dev = "cuda"
torch.backends.cudnn.benchmark = True
flush = None


def timeit(fn, reps=9):
    global flush
    if flush is None:
        flush = torch.empty(64 * 1024 * 1024, device=dev)
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def bf(*s):
    return (torch.randn(*s, device=dev) * 0.1).bfloat16()


def conv_case(M, N, K, mode):
    """(library callable, description) for an implicit-conv class; B = 256 images."""
    if mode in (1, 3):  # fprop or (stride-1) dgrad: 3x3 conv C -> N over M = B*H*W outputs
        C = K // 9
        side = int(round((M // 256) ** 0.5))
        if K % 9 or side * side * 256 != M or C % 64:
            raise ValueError("strided / parity-class conv class: no single library call computes the same product")
        x = bf(256, C, side, side).contiguous(memory_format=torch.channels_last)
        w = bf(N, C, 3, 3).contiguous(memory_format=torch.channels_last)
        return (lambda: F.conv2d(x, w, padding=1)), f"cuDNN conv2d 3x3 {C}->{N} @ {side}x{side}"
    if mode == 2:       # wgrad of a 3x3 conv: M = Cout, N = 9*C, K = positions
        C = N // 9
        side = int(round((K // 256) ** 0.5))
        if N % 9 or side * side * 256 != K:
            raise ValueError("strided conv class")
        x = bf(256, C, side, side).contiguous(memory_format=torch.channels_last)
        dy = bf(256, M, side, side).contiguous(memory_format=torch.channels_last)
        return (lambda: torch.nn.grad.conv2d_weight(x, (M, C, 3, 3), dy, padding=1)), f"cuDNN wgrad 3x3 {C}->{M} @ {side}x{side}"
    if mode == 4:
        x = bf(256, 64, 56, 56).contiguous(memory_format=torch.channels_last)
        dy = bf(256, 64, 56, 56).contiguous(memory_format=torch.channels_last)
        return (lambda: torch.nn.grad.conv2d_weight(x, (64, 64, 3, 3), dy, padding=1)), "cuDNN wgrad 3x3 64->64 @ 56x56"
    if mode == 5:
        x = bf(256, 3, 224, 224).contiguous(memory_format=torch.channels_last)
        w = bf(64, 3, 7, 7).contiguous(memory_format=torch.channels_last)
        return (lambda: F.conv2d(x, w, stride=2, padding=3)), "cuDNN conv2d 7x7/2 3->64 @ 224x224"
    if mode == 6:
        x = bf(256, 3, 224, 224).contiguous(memory_format=torch.channels_last)
        dy = bf(256, 64, 112, 112).contiguous(memory_format=torch.channels_last)
        return (lambda: torch.nn.grad.conv2d_weight(x, (64, 3, 7, 7), dy, stride=2, padding=3)), "cuDNN wgrad 7x7/2 3->64"
    raise ValueError(mode)


def main():
    data = json.load(open(sys.argv[1]))
    cls = OrderedDict()
    for d in data:
        k = (d["M"], d["N"], d["K"], d["conv_mode"], d["a_mn"], d["b_mn"])
        c = cls.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += d["ms"]
    rows = []
    for (M, N, K, mode, a_mn, b_mn), (n, ms) in sorted(cls.items(), key=lambda kv: -kv[1][1]):
        ours = ms / n * 1e3
        try:
            if mode == 0:
                A = bf(K, M).t() if a_mn else bf(M, K)
                B = bf(K, N) if b_mn else bf(N, K).t()
                fn, what = (lambda A=A, B=B: torch.matmul(A, B)), "cuBLASLt bf16 matmul" + (" (fp32 out in ours)" if a_mn and b_mn else "")
            else:
                fn, what = conv_case(M, N, K, mode)
            lib = timeit(fn)
        except Exception as e:  # noqa: BLE001
            lib, what = float("nan"), f"n/a ({type(e).__name__})"
        rows.append((M, N, K, mode, a_mn, b_mn, n, ours, lib, what))
        torch.cuda.empty_cache()
    tot_ours = sum(r[7] * r[6] for r in rows) / 1e3
    tot_lib = sum(r[8] * r[6] for r in rows if r[8] == r[8]) / 1e3
    print(f"# gemm_tc_kernel vs cuBLASLt / cuDNN per shape class, one optimisation step at batch 256 (torch {torch.__version__})\n")
    print(f"Sum over classes: ours {tot_ours:.2f} ms (with its fused epilogues, inside the step), libraries {tot_lib:.2f} ms (bare "
          "products, stand-alone, L2 flushed).  `ratio` = ours / library: > 1 means the library call alone is faster than our "
          "launch including its epilogue work.\n")
    print("| M | N | K | mode | a | b | launches | ours us | library us | ratio | library call |")
    print("|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|")
    for M, N, K, mode, a, b, n, ours, lib, what in rows:
        print(f"| {M} | {N} | {K} | {mode} | {a} | {b} | {n} | {ours:.1f} | {lib:.1f} | {ours / lib:.2f} | {what} |")
    lose = [r for r in rows if r[8] == r[8] and r[7] > 1.1 * r[8]]
    print(f"\nClasses where the bare library call is > 10 % faster than our launch: {len(lose)} of {len(rows)}, "
          f"{sum((r[7] - r[8]) * r[6] for r in lose) / 1e3:.2f} ms per step in total.")


if __name__ == "__main__":
    main()
