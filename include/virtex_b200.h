/*
 * virtex_b200 -- C ABI of the B200-native (sm_100a) kernels behind the VirTex bicaptioning pretraining step.
 *
 * The reference (kdexd/virtex) has no FFI of its own: its hot path is `VirTexModel.forward` + autograd
 * (virtex/models/captioning.py:71-143) executed by torch / torchvision library calls.
 * Each entry point below replaces one family of those library calls; the file:line it stands in for is cited.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless stated otherwise
 *   - `stream` is a cudaStream_t passed as void*
 *   - returns 0 on success, a negative VTX_E* code on failure; vtx_last_error() gives a message (thread local)
 *   - no allocation, no synchronisation, no global state beyond cached device properties
 *   - activations are bf16 (NHWC for the backbone, [tokens, features] row-major for the head); statistics,
 *     master parameters, gradients and the decoder residual stream are fp32
 */
#ifndef VIRTEX_B200_H_
#define VIRTEX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VTX_OK 0
#define VTX_EINVAL (-1)
#define VTX_ECUDA (-2)
#define VTX_EUNSUPPORTED (-3)

const char* vtx_last_error(void);
int vtx_version(void);
/* Number of SMs of the current device (cached). */
int vtx_num_sms(void);

/* ------------------------------------------------------------------------------------------------------------------
 * tcgen05 GEMM:  D[M,N] = epilogue( sum_k A[m,k] * B[n,k] )       bf16 x bf16 -> fp32 accumulate in TMEM
 * Replaces every cuBLASLt / cuDNN GEMM-shaped call on the path: nn.Linear fwd/dgrad/wgrad
 * (virtex/modules/textual_heads.py:168-170,199,245,277; torch/nn/modules/transformer.py:1158-1199) and the 1x1 /
 * im2col'd convolutions of torchvision Bottleneck (torchvision/models/resnet.py:146-158).
 *   a_mn = 0: A is stored [M, K] row major (K contiguous, leading dim lda)     ("K-major")
 *   a_mn = 1: A is stored [K, M] row major (M contiguous, leading dim lda)     ("MN-major", used by wgrad)
 *   b_mn = 0: B is stored [N, K] row major;   b_mn = 1: B is stored [K, N] row major.
 * Epilogue order: acc -> (stats: per-column sum / sum of squares of acc, fp32 atomics) -> *alpha -> +bias[n]
 *                 -> +residual[m,n] (bf16) -> activation -> store (bf16 or fp32; or fp32 atomic accumulate).
 * split_k > 1 requires atomic = 1 (fp32 output, caller zero-initialises).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct VtxGemm {
  const void* A;
  const void* B;
  void* D;
  const float* bias;    /* [N] or NULL */
  const void* residual; /* bf16 [M, ldr] or NULL */
  float* stats;         /* [2, N]: sum, sumsq  or NULL */
  int64_t lda, ldb, ldd, ldr;
  int32_t M, N, K;
  int32_t a_mn, b_mn;
  int32_t out_f32; /* 0: bf16 output, 1: fp32 output */
  int32_t atomic;  /* 1: D += result with fp32 atomics (needs out_f32) */
  int32_t act;     /* 0 none, 1 relu, 2 gelu(erf) */
  int32_t split_k; /* >= 1 */
  int32_t tile_n;  /* 0 = auto; else multiple of 16 (64 if b_mn) and <= 256 */
  float alpha;
  /* implicit 3x3 / stride 1 / pad 1 convolution over an NHWC bf16 tensor (conv_c > 0):
     A is the activation [conv_n, conv_h, conv_w, conv_c]; M = n*h*w, K = 9*conv_c, B = weights [N, (kh,kw,c)].
     conv_wgrad = 1 swaps roles for the weight gradient (see gemm_tc.cu). */
  int32_t conv_n, conv_h, conv_w, conv_c;
  int32_t conv_mode; /* 0 = plain GEMM, 1 = implicit fprop/dgrad gather on A, 2 = wgrad gather on B */
} VtxGemm;

int vtx_gemm(const VtxGemm* g, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIRTEX_B200_H_ */
