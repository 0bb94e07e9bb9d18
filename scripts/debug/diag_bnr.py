"""Diagnostic: is the masked-residual GEMM epilogue deterministic, and where does the fused-BN-reduce variant differ?"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from virtex_b200 import ops

BF16 = torch.bfloat16


def pack_mask(keep):
    M, C = keep.shape
    w = (1 << torch.arange(8, device=keep.device)).to(torch.int32)
    return (keep.view(M, C // 8, 8).to(torch.int32) * w).sum(-1).to(torch.uint8).contiguous()


def describe(tag, X, Y, ref):
    d = (X != Y)
    n = int(d.sum())
    print(f"{tag}: {n} differing elements of {X.numel()}")
    if n:
        idx = d.nonzero()[:2000]
        rows, cols = idx[:, 0], idx[:, 1]
        print("   rows: min", int(rows.min()), "max", int(rows.max()), " distinct tiles(128 rows):", len(set((rows // 128).tolist())),
              " row%128 sample:", sorted(set((rows % 128).tolist()))[:20])
        print("   cols: distinct", len(set(cols.tolist())), "sample", sorted(set(cols.tolist()))[:24])
        r0, c0 = int(rows[0]), int(cols[0])
        print("   first:", (r0, c0), "X", float(X[r0, c0]), "Y", float(Y[r0, c0]), "ref", float(ref[r0, c0]))
        ex = (X.float() - ref).abs()[d].max().item()
        ey = (Y.float() - ref).abs()[d].max().item()
        print(f"   max |X-ref| on differing {ex:.4f}   max |Y-ref| {ey:.4f}")


def main():
    g = torch.Generator().manual_seed(20000 + 256 + 64)
    for (M, N, K) in ((20000, 256, 64), (60000, 256, 64), (20000, 256, 128)):
        A = (torch.randn(M, K, generator=g) * 0.5).bfloat16().cuda()
        B = (torch.randn(K, N, generator=g) * 0.2).bfloat16().cuda()
        y = (torch.randn(M, N, generator=g) * 1.5).bfloat16().cuda()
        mean = torch.randn(N, generator=g) * 0.5
        bnp = torch.stack([mean, torch.rand(N, generator=g) + 0.5, torch.rand(N, generator=g) + 0.5,
                           torch.randn(N, generator=g) * 0.3]).contiguous().cuda()
        R = torch.randn(M, N, generator=g).bfloat16().cuda()
        rkeep = (torch.rand(M, N, generator=g) > 0.5).cuda()
        bkeep = (torch.rand(M, N, generator=g) > 0.45).cuda()
        rbits, bbits = pack_mask(rkeep), pack_mask(bkeep)
        ref = A.float() @ B.float() + R.float() * rkeep.float()
        outs = []
        for i in range(3):
            D0 = torch.empty(M, N, dtype=BF16, device="cuda")
            ops.gemm(A, B, D0, M, N, K, b_mn=1, residual=R, residual_mask=rbits)
            outs.append(D0)
        torch.cuda.synchronize()
        print(f"--- M={M} N={N} K={K}")
        describe("plain run0 vs run1", outs[0], outs[1], ref)
        describe("plain run0 vs run2", outs[0], outs[2], ref)
        sums_all = []
        for i in range(3):
            D = torch.empty(M, N, dtype=BF16, device="cuda")
            sums = torch.zeros(2, N, device="cuda")
            ops.gemm(A, B, D, M, N, K, b_mn=1, residual=R, residual_mask=rbits, bnr=(y, bnp, sums, bbits))
            torch.cuda.synchronize()
            describe(f"bnr run{i} vs plain run0", D, outs[0], ref)
            sums_all.append(sums)
            dz = D.double() * bkeep.double()
            xh = (y.double() - bnp[0].double()) * bnp[1].double()
            e0 = ((sums[0].double() - dz.sum(0)).norm() / dz.sum(0).norm()).item()
            e1 = ((sums[1].double() - (dz * xh).sum(0)).norm() / (dz * xh).sum(0).norm()).item()
            print(f"   sums rel err {e0:.2e} {e1:.2e}")


main()
