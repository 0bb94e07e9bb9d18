"""CPU checks of the index maps behind the space-to-depth stem path (csrc/stem_s2d.cu, vtx_gemm conv_mode 5 / 6):
the space-to-depth formulation IS the 7x7 / stride-2 / pad-3 convolution, and a line-by-line transliteration of the
packing kernel's index arithmetic reproduces the layout definition.  (The kernels themselves need a GPU: see
tests/test_gpu_parity.py::test_stem_space_to_depth_conv.)"""
import torch

from tests.test_gpu_parity import _s2d_ref, _stem_wpack_ref


def _views(S, Ho, Wo):
    """A operand of k-block a: 4 pixels x 16 channels = 64 contiguous elements per output position."""
    N, Hs, Ws, _ = S.shape
    flat = S.reshape(N, Hs, Ws * 16)
    for a in range(4):
        for ow in range(Wo):
            yield a, ow, flat[:, a:a + Ho, ow * 16:ow * 16 + 64]


def test_space_to_depth_formulation_equals_the_stem_convolution():
    torch.manual_seed(0)
    x, w = torch.randn(2, 3, 16, 24, dtype=torch.float64), torch.randn(5, 3, 7, 7, dtype=torch.float64)
    S, wp = _s2d_ref(x), _stem_wpack_ref(w)
    Ho, Wo = 8, 12
    assert S.shape == (2, Ho + 3, Wo + 3, 16) and wp.shape == (5, 256)
    out = torch.zeros(2, Ho, Wo, 5, dtype=torch.float64)
    for a, ow, A in _views(S, Ho, Wo):
        out[:, :, ow, :] += A @ wp[:, a * 64:(a + 1) * 64].t()
    ref = torch.nn.functional.conv2d(x, w, stride=2, padding=3).permute(0, 2, 3, 1)
    assert torch.allclose(out, ref, atol=1e-12)
    # weight gradient through the same views, folded back with the unpack map of stem_w_unpack_add_kernel
    dy = torch.randn(2, Ho, Wo, 5, dtype=torch.float64)
    dwp = torch.zeros(5, 256, dtype=torch.float64)
    for a, ow, A in _views(S, Ho, Wo):
        dwp[:, a * 64:(a + 1) * 64] += dy[:, :, ow, :].reshape(-1, 5).t() @ A.reshape(-1, 64)
    g = torch.zeros(5, 3, 7, 7, dtype=torch.float64)
    for kh in range(7):
        for kw in range(7):
            for c in range(3):
                g[:, c, kh, kw] = dwp[:, (kh >> 1) * 64 + (kw >> 1) * 16 + ((kh & 1) * 2 + (kw & 1)) * 3 + c]
    gref = torch.nn.grad.conv2d_weight(x, (5, 3, 7, 7), dy.permute(0, 3, 1, 2), stride=2, padding=3)
    assert torch.allclose(g, gref, atol=1e-10)


def test_s2d_kernel_index_arithmetic():
    """stem_s2d_kernel, statement by statement (shared-memory tile of 3 x 2 rows with 4 zero columns on each side)."""
    torch.manual_seed(1)
    for (N, H, W) in [(2, 8, 12), (1, 16, 16)]:
        img = torch.randn(N, 3, H, W)
        Hs, Ws, Wp = H // 2 + 3, W // 2 + 3, W + 8
        quads = Wp // 4
        S = torch.full((N, Hs, Ws, 16), 99.0)
        for blk in range(N * Hs):
            n, i = blk // Hs, blk % Hs
            tile = torch.full((6 * Wp,), float("nan"))
            for e in range(6 * quads):
                cr, qd = e // quads, e % quads
                c, r = cr >> 1, cr & 1
                h, w0 = 2 * i + r - 3, qd * 4 - 4
                v = torch.zeros(4)
                if 0 <= h < H and w0 >= 0 and w0 + 3 < W:
                    v = img[n, c, h, w0:w0 + 4]
                tile[cr * Wp + qd * 4: cr * Wp + qd * 4 + 4] = v
            for e in range(Ws * 2):
                j, half = e >> 1, e & 1
                for t in range(8):
                    ch = half * 8 + t
                    f = 0.0
                    if ch < 12:
                        rq, c = ch // 3, ch % 3
                        f = tile[(c * 2 + (rq >> 1)) * Wp + 2 * j + (rq & 1) + 1]
                    S[n, i, j, ch] = f
        assert torch.equal(S, _s2d_ref(img))
