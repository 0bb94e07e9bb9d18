"""Attribution experiments for the HBM-bound GEMM shapes (developer tool, run under gpurun)."""
import os, subprocess, sys
if len(sys.argv) > 1:
    import torch
    sys.path.insert(0, ".")
    from virtex_b200 import ops
    M, N, K, tn = [int(x) for x in sys.argv[1:5]]
    A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16(); B = (torch.randn(N, K, device="cuda") * 0.5).bfloat16()
    D = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3): ops.gemm(A, B, D, M, N, K, tile_n=tn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(20): ops.gemm(A, B, D, M, N, K, tile_n=tn)
    e1.record(); torch.cuda.synchronize()
    err = ((D[:4096].float() - A[:4096].float() @ B.float().t()).norm() / (A[:4096].float() @ B.float().t()).norm()).item()
    print(f"M{M} N{N} K{K} tile_n={tn} dbg={os.environ.get('VTX_GEMM_DBG','0'):>3s}: {e0.elapsed_time(e1)/20*1e3:7.1f} us  relerr {err:.1e}", flush=True)
else:
    runs = [(802816, 256, 64, 256, 0), (802816, 256, 64, 256, 16), (802816, 256, 64, 256, 17), (802816, 256, 64, 128, 0),
            (802816, 256, 64, 64, 0), (802816, 64, 64, 64, 0), (802816, 64, 64, 64, 16), (802816, 64, 64, 64, 17),
            (802816, 128, 64, 128, 16), (200704, 256, 64, 256, 16)]
    for (M, N, K, tn, dbg) in runs:
        env = dict(os.environ, VTX_GEMM_DBG=str(dbg))
        subprocess.run([sys.executable, __file__, str(M), str(N), str(K), str(tn)], env=env)
