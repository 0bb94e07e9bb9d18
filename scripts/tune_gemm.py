#!/usr/bin/env python
"""Per-shape-class tuning sweep of vtx_gemm's free parameters (tile_n, split_k) on a B200 (measurement tooling).

    python scripts/tune_gemm.py profiles/r02a_gemm_launches.json [--top 40] > gpurun_out/tune_gemm.txt

For every distinct (M, N, K, conv_mode, a_mn, b_mn) class of one optimisation step it rebuilds operands of the right
shapes, times the library's own choice (tile_n = 0, the engine's split_k) and the alternatives with CUDA events (median
of 7 launches, a 256 MB buffer written between launches to flush L2) and prints one line per class, sorted by the time
the best alternative would save per step.  Nothing here changes the product: the winners are folded into the heuristics
of vtx_gemm / Engine._wgrad by hand.
"""
import argparse
import json
import sys
from collections import OrderedDict

import torch

sys.path.insert(0, ".")
from virtex_b200 import ops  # noqa: E402

dev = "cuda"


def bf(*s):
    return (torch.randn(*s, device=dev) * 0.1).bfloat16()


def conv_geometry(M, N, K, mode):
    """(B, H, W, C) of the implicit-conv classes of ResNet-50 at batch 256 (M or K = B*H*W)."""
    if mode in (1, 3):
        C = K // 9
        pos = M
    elif mode == 2:
        C = N // 9
        pos = K
    elif mode == 4:
        C, pos = 64, K
    elif mode == 5:
        return 256, 112, 112, 64
    elif mode == 6:
        return 256, 112, 112, 64
    hw = pos // 256
    side = int(round(hw ** 0.5))
    return 256, side, side, C


def build(M, N, K, mode, a_mn, b_mn):
    kw = {}
    if mode == 0:
        A = bf(K, M) if a_mn else bf(M, K)
        B = bf(K, N) if b_mn else bf(N, K)
        f32 = bool(a_mn and b_mn)
        D = torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
        kw = dict(a_mn=a_mn, b_mn=b_mn, atomic=f32, out_f32=f32)
    else:
        Bn, H, W, C = conv_geometry(M, N, K, mode)
        if mode == 1:
            A, B = bf(Bn, H, W, C), bf(N, K)
            D = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            kw = dict(lda=C, conv=(Bn, H, W, C), conv_mode=1)
        elif mode == 2:
            A, B = bf(Bn, H, W, M), bf(Bn, H, W, C)
            D = torch.zeros(M, N, device=dev)
            kw = dict(lda=M, ldb=C, atomic=True, out_f32=True, conv=(Bn, H, W, C), conv_mode=2)
        elif mode == 4:
            A, B = bf(Bn, H, W, 64), bf(Bn, H, W, 64)
            D = torch.zeros(576, 64, device=dev)
            kw = dict(lda=64, ldb=64, ldd=64, atomic=True, out_f32=True, conv=(Bn, H, W, 64), conv_mode=4)
        elif mode == 5:
            A, B = bf(Bn, H + 3, W + 3, 16), bf(64, 256)
            D = torch.empty(M, 64, device=dev, dtype=torch.bfloat16)
            kw = dict(lda=64, ldb=256, conv=(Bn, H, W, 64), conv_mode=5)
        else:
            A, B = bf(Bn, H, W, 64), bf(Bn, H + 3, W + 3, 16)
            D = torch.zeros(64, 256, device=dev)
            kw = dict(lda=64, ldb=64, atomic=True, out_f32=True, conv=(Bn, H, W, 64), conv_mode=6)
    return A, B, D, kw


flush = None


def timeit(fn, reps=7):
    global flush
    if flush is None:
        flush = torch.empty(64 * 1024 * 1024, device=dev)
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("profile")
    ap.add_argument("--top", type=int, default=60)
    args = ap.parse_args()
    cls = OrderedDict()
    for d in json.load(open(args.profile)):
        k = (d["M"], d["N"], d["K"], d["conv_mode"], d["a_mn"], d["b_mn"])
        c = cls.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += d["ms"]
    order = sorted(cls.items(), key=lambda kv: -kv[1][1])[:args.top]
    rows = []
    for (M, N, K, mode, a_mn, b_mn), (n, ms) in order:
        try:
            A, B, D, kw = build(M, N, K, mode, a_mn, b_mn)
        except Exception as e:  # noqa: BLE001
            print(f"# skip {(M, N, K, mode)}: {e}")
            continue
        wgrad = kw.get("atomic", False)
        if not wgrad and M % 12544 == 0 and not kw.get("b_mn") and mode in (0, 1, 5):
            kw["stats"] = torch.zeros(2, N, device=dev)  # conv fprop: BN statistics in the epilogue, as in the step
        if not wgrad and mode == 0 and kw.get("b_mn") and M % 12544 == 0 and N == 4 * K:
            # conv1 dgrad of a bottleneck: + shortcut gradient under the ReLU bit mask (engine.backbone_backward)
            kw["residual"] = bf(M, N)
            kw["residual_mask"] = torch.randint(0, 256, (M, N // 8), device=dev, dtype=torch.uint8)
        if mode == 0 and wgrad:
            tiles = ((M + 127) // 128) * ((N + 255) // 256)
            sk0 = ops.split_k_for(tiles, (K + 63) // 64)
        elif mode in (2, 6):
            tiles = ((M + 127) // 128) * ((N + 255) // 256)
            sk0 = ops.split_k_for(tiles, K // 64)
        else:
            sk0 = 1
        gran = 64 if (kw.get("b_mn") or mode in (2, 4, 6)) else 16
        tns = [0] + [t for t in (64, 128, 192, 256) if t % gran == 0 and t <= max(64, ((N + gran - 1) // gran) * gran)]
        sks = sorted({sk0, max(1, sk0 // 2), sk0 * 2, sk0 * 4}) if wgrad and mode != 4 else [sk0]
        if mode in (3, 4, 5, 6):
            tns = [0]
        res = {}
        for tn in tns:
            for sk in sks:
                def run(tn=tn, sk=sk):
                    ops.gemm(A, B, D, M, N, K, tile_n=tn, split_k=sk, **kw)
                try:
                    run()
                    torch.cuda.synchronize()
                    res[(tn, sk)] = timeit(run)
                except Exception:  # noqa: BLE001 -- an unsupported (tile_n, split_k) combination
                    continue
        base = res.get((0, sk0))
        if base is None:
            continue
        best = min(res, key=res.get)
        rows.append(((base - res[best]) * n, (M, N, K, mode, a_mn, b_mn), n, base, best, res[best], res))
        del A, B, D
    rows.sort(key=lambda r: -r[0])
    print("saved_us_per_step | class (M,N,K,mode,a,b) | launches | default us | best (tile_n, split_k) us | all")
    for saved, k, n, base, best, tb, res in rows:
        allr = " ".join(f"{t}/{s}:{v:.1f}" for (t, s), v in sorted(res.items()))
        print(f"{saved:8.1f} | {k} | {n} | {base:.1f} | {best} {tb:.1f} | {allr}")
    print(f"# total potential saving {sum(r[0] for r in rows) / 1e3:.3f} ms per step")


if __name__ == "__main__":
    main()
