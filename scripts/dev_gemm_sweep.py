"""Sweep tile_n / staging-buffer count for the HBM-bound GEMM shapes (developer tool, run under gpurun)."""
import os, sys
import torch
sys.path.insert(0, ".")
from virtex_b200 import ops

def bf(*s):
    return (torch.randn(*s, device="cuda") * 0.5).bfloat16()

def timeit(fn, iters=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

shapes = [(802816, 256, 64), (802816, 64, 256), (200704, 512, 128), (200704, 128, 512), (50176, 1024, 256),
          (50176, 256, 1024), (12544, 2048, 512), (12544, 512, 2048), (7680, 1024, 1024), (7680, 3072, 1024)]
for (M, N, K) in shapes:
    A, B, D = bf(M, K), bf(N, K), torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    res = []
    for tn in (64, 128, 256):
        if tn > N: continue
        for nb in ("1", "2"):
            os.environ["VTX_GEMM_NBUF"] = nb
            try:
                us = timeit(lambda: ops.gemm(A, B, D, M, N, K, tile_n=tn))
            except Exception as e:
                us = float("nan")
            res.append((us, tn, nb))
    os.environ.pop("VTX_GEMM_NBUF", None)
    auto = timeit(lambda: ops.gemm(A, B, D, M, N, K))
    tref = timeit(lambda: torch.matmul(A, B.t()))
    best = min(res)
    print(f"{M}x{N}x{K}: auto {auto:6.1f} us | cublas {tref:6.1f} | best {best[0]:6.1f} (tile_n {best[1]}, nbuf {best[2]}) | " +
          " ".join(f"{tn}/{nb}:{us:.0f}" for us, tn, nb in res), flush=True)
