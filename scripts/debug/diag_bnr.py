"""Diagnostic: where does the fused-BN-reduce epilogue differ from the plain one (sentinel-prefilled outputs)?"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from virtex_b200 import ops

BF16 = torch.bfloat16


def pack_mask(keep):
    M, C = keep.shape
    w = (1 << torch.arange(8, device=keep.device)).to(torch.int32)
    return (keep.view(M, C // 8, 8).to(torch.int32) * w).sum(-1).to(torch.uint8).contiguous()


def describe(tag, X, Y):
    d = (X != Y) | torch.isnan(X.float())
    n = int(d.sum())
    msg = f"{tag}: {n} differ"
    if n:
        rows = d.any(1).nonzero().flatten()
        cols = d.any(0).nonzero().flatten()
        tiles = sorted(set((rows // 128).tolist()))
        sent = int(((X.float() == 777.0) & d).sum())
        nan = int((torch.isnan(X.float()) & d).sum())
        msg += f"; rows {int(rows.min())}..{int(rows.max())} ({len(rows)} rows, tiles {tiles[:6]}..{tiles[-3:]} n={len(tiles)}); cols {int(cols.min())}..{int(cols.max())} ({len(cols)}); still-sentinel {sent}, nan {nan}"
    print(msg, flush=True)


def main():
    g = torch.Generator().manual_seed(1)
    for (M, N, K, pf) in ((20000, 256, 64, "1"), (20000, 256, 64, "0"), (19000, 256, 64, "0"), (20000, 512, 64, "1"), (20000, 512, 64, "0"),
                          (7680, 1024, 1024, "1"), (7680, 1024, 1024, "0")):
        os.environ["VTX_BNR_PREFETCH"] = pf
        A = (torch.randn(M, K, generator=g) * 0.5).bfloat16().cuda()
        B = (torch.randn(K, N, generator=g) * 0.2).bfloat16().cuda()
        y = (torch.randn(M, N, generator=g) * 1.5).bfloat16().cuda()
        bnp = torch.stack([torch.randn(N, generator=g) * 0.5, torch.rand(N, generator=g) + 0.5, torch.rand(N, generator=g) + 0.5,
                           torch.randn(N, generator=g) * 0.3]).contiguous().cuda()
        R = torch.randn(M, N, generator=g).bfloat16().cuda()
        rkeep = (torch.rand(M, N, generator=g) > 0.5).cuda()
        bkeep = (torch.rand(M, N, generator=g) > 0.45).cuda()
        rbits, bbits = pack_mask(rkeep), pack_mask(bkeep)
        print(f"--- M={M} N={N} K={K} prefetch={pf}", flush=True)

        def run(**kw):
            D = torch.full((M, N), 777.0, dtype=BF16, device="cuda")
            ops.gemm(A, B, D, M, N, K, b_mn=1, **kw)
            torch.cuda.synchronize()
            return D

        sums = torch.zeros(2, N, device="cuda")
        P_res_mask = run(residual=R, residual_mask=rbits)
        P_res = run(residual=R)
        P = run()
        describe("bnr(bits)+res+mask vs plain", run(residual=R, residual_mask=rbits, bnr=(y, bnp, sums, bbits)), P_res_mask)
        describe("bnr(bits)+res      vs plain", run(residual=R, bnr=(y, bnp, sums, bbits)), P_res)
        describe("bnr(from y)+res    vs plain", run(residual=R, bnr=(y, bnp, sums, None)), P_res)
        describe("bnr(bits)          vs plain", run(bnr=(y, bnp, sums, bbits)), P)
        describe("bnr(from y)        vs plain", run(bnr=(y, bnp, sums, None)), P)
        describe("stats              vs plain", run(stats=sums), P)


main()
