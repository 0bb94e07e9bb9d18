/* virtex_b200_x.h -- EXPERIMENTAL entry points of libvirtex_b200_x.so.
 *
 * Kernels in this library were written without hardware access and are NOT on the default path: the engine uses
 * them only when the environment variable VTX_EXPERIMENTAL names the feature (see virtex_b200/experimental.py), and
 * their GPU tests are skipped unless it is set.  Once validated on a B200 they move into virtex_b200.h.
 * Conventions (error codes, streams, no allocation, no synchronisation) are those of virtex_b200.h.
 */
#ifndef VIRTEX_B200_X_H
#define VIRTEX_B200_X_H
#include "virtex_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* vtx_gemm with two more implicit-conv modes (everything else identical to vtx_gemm):
 *   conv_mode 5: stem fprop.  A = S, the space-to-depth view [conv_n, conv_h + 3, conv_w + 3, 16] bf16 written by
 *                vtx_x_stem_s2d (conv_h x conv_w = OUTPUT size, conv_c = 64 = 4 pixels x 16 channels);
 *                B = packed weights [N = 64, K = 256] (vtx_x_stem_w_pack); D [conv_n*conv_h*conv_w, 64] NHWC.
 *                Replaces torchvision resnet.py:197 conv1 forward (with BN statistics through `stats`).
 *   conv_mode 6: stem wgrad.  A = dy [conv_n, conv_h, conv_w, M = 64], B = S; D [64, 256] fp32 += (atomic). */
int vtx_gemm_x(const VtxGemm* g, void* stream);

/* image fp32 NCHW [N, 3, H, W] -> S bf16 [N, H/2 + 3, W/2 + 3, 16]:
 * S[n, i, j, (r*2+q)*3 + c] = img[n, c, 2i + r - 3, 2j + q - 3], zero outside the image and in channels 12..15 */
int vtx_x_stem_s2d(const float* img, void* S, int N, int H, int W, void* stream);
/* conv1.weight fp32 [O, 3, 7, 7] -> bf16 [O, 256], k = a*64 + b*16 + (r*2+q)*3 + c for tap (kh, kw) = (2a+r, 2b+q) */
int vtx_x_stem_w_pack(const float* w, void* wp, int O, void* stream);
/* grad fp32 [O, 3, 7, 7] += dwp fp32 [O, 256] (same index map) */
int vtx_x_stem_w_unpack_add(const float* dwp, float* grad, int O, void* stream);

/* Feature `head_x`: this library also exports vtx_ln_bwd and vtx_embed_bwd with the signatures of virtex_b200.h,
 * implemented by register-accumulating kernels (csrc/head.cu, -DVTX_HEAD_X); virtex_b200/ops.py routes those two
 * entry points here when the feature is enabled. */

#ifdef __cplusplus
}
#endif
#endif
