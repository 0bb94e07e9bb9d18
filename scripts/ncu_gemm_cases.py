"""A few representative GEMM launches for `ncu --set full` captures (run under gpurun)."""
import sys
import torch
sys.path.insert(0, ".")
from virtex_b200 import ops

dev = "cuda"
torch.manual_seed(0)


def bf(*s):
    return (torch.randn(*s, device=dev) * 0.5).bfloat16()


cases = sys.argv[1:] or ["l1conv3", "ffn1", "l1dgrad"]
for _ in range(2):
    if "l1conv3" in cases:  # layer1 1x1 conv 64->256 fprop with BN statistics: HBM bound
        M, N, K = 802816, 256, 64
        A, B, D = bf(M, K), bf(N, K), torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        st = torch.zeros(2, N, device=dev)
        ops.gemm(A, B, D, M, N, K, stats=st)
    if "l1dgrad" in cases:  # layer1 conv1 dgrad + residual
        M, N, K = 802816, 256, 64
        A, B, D = bf(M, K), bf(K, N), torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        R = bf(M, N)
        ops.gemm(A, B, D, M, N, K, b_mn=1, residual=R)
    if "ffn1" in cases:  # FFN linear1 with bias: tensor bound
        M, N, K = 7680, 4096, 1024
        A, B, D = bf(M, K), bf(N, K), torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        bias = torch.randn(N, device=dev)
        ops.gemm(A, B, D, M, N, K, bias=bias)
    if "l1conv" in cases:  # layer1 3x3 implicit conv 64->64 with BN statistics
        x, w = bf(256, 56, 56, 64), bf(64, 576)
        D = torch.empty(256 * 3136, 64, device=dev, dtype=torch.bfloat16)
        st = torch.zeros(2, 64, device=dev)
        ops.gemm(x, w, D, 256 * 3136, 64, 576, lda=64, stats=st, conv=(256, 56, 56, 64), conv_mode=1)
    if "l1wgrad" in cases:  # layer1 3x3 halo-reuse wgrad (conv_mode 4)
        x, dy = bf(256, 56, 56, 64), bf(256, 56, 56, 64)
        D = torch.zeros(576, 64, device=dev)
        ops.gemm(dy, x, D, 576, 64, 256 * 3136, lda=64, ldb=64, ldd=64, atomic=True, out_f32=True,
                 conv=(256, 56, 56, 64), conv_mode=4)
    if "l2wgrad" in cases:  # layer2 3x3 implicit wgrad (conv_mode 2), 128 -> 128 channels at 28 x 28
        x, dy = bf(256, 28, 28, 128), bf(256, 28, 28, 128)
        D = torch.zeros(128, 1152, device=dev)
        ops.gemm(dy, x, D, 128, 1152, 256 * 784, lda=128, ldb=128, atomic=True, out_f32=True,
                 split_k=ops.split_k_for(5, 256 * 784 // 64), conv=(256, 28, 28, 128), conv_mode=2)
    if "l3conv" in cases:  # layer3 3x3 implicit fprop 256 -> 256 at 14 x 14 with BN statistics: tensor bound
        x, w = bf(256, 14, 14, 256), bf(256, 2304)
        D = torch.empty(256 * 196, 256, device=dev, dtype=torch.bfloat16)
        st = torch.zeros(2, 256, device=dev)
        ops.gemm(x, w, D, 256 * 196, 256, 2304, lda=256, stats=st, conv=(256, 14, 14, 256), conv_mode=1)
    if "l1dgrad_bnr" in cases:  # layer1 conv1 dgrad + masked shortcut gradient + fused bn3-backward sums (kBnr = 2)
        M, N, K = 802816, 256, 64
        A, B, D = bf(M, K), bf(K, N), torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        R, y = bf(M, N), bf(M, N)
        bits = torch.randint(0, 256, (M, N // 8), device=dev, dtype=torch.uint8)
        bnp = torch.rand(4, N, device=dev) + 0.5
        sums = torch.zeros(2, N, device=dev)
        ops.gemm(A, B, D, M, N, K, b_mn=1, residual=R, residual_mask=bits, bnr=(y, bnp, sums, bits))
    if "l1conv3_dgrad_bnr" in cases:  # layer1 conv3 dgrad 256 -> 64 + fused bn2-backward sums (kBnr = 1, mask from y)
        M, N, K = 802816, 64, 256
        A, B, D = bf(M, K), bf(K, N), torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        y, bnp, sums = bf(M, N), torch.rand(4, N, device=dev) + 0.5, torch.zeros(2, N, device=dev)
        ops.gemm(A, B, D, M, N, K, b_mn=1, bnr=(y, bnp, sums, None))
    if "vocab_pair" in cases:  # vocabulary projection dgrad 7680 x 1024 x 10000: CTA pairs (cta_group::2)
        M, N, K = 7680, 1024, 10000
        A, B, D = bf(M, K), bf(K, N), torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        ops.gemm(A, B, D, M, N, K, b_mn=1)
    if "ffn2_pair" in cases:  # FFN linear2 7680 x 1024 x 4096 with bias: CTA pairs
        M, N, K = 7680, 1024, 4096
        A, B, D = bf(M, K), bf(N, K), torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        ops.gemm(A, B, D, M, N, K, bias=torch.randn(N, device=dev))
    torch.cuda.synchronize()
print("done")
