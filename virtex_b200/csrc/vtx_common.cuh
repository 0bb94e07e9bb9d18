// Shared host/device helpers for the virtex_b200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string.h>

// Programmatic dependent launch (validated on B200 in round 2: -0.4 ms/step).  Every kernel of the library signals at
// its first instruction that dependents may be scheduled; the GEMM (the only kernel launched with the
// programmatic-serialisation attribute) runs its prologue -- barrier init, TMEM allocation, tensor-map prefetch --
// while the previous kernel drains, then waits for it to complete and flush before its first global access.
// Kernels launched without the attribute keep plain stream order, so nothing else changes semantics.
#define VTX_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define VTX_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")

namespace vtx {
// printf-style error recording; returns `code` so call sites can `return set_error(...)`.
int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);

__device__ __forceinline__ float bf2f(__nv_bfloat16 x) { return __bfloat162float(x); }
__device__ __forceinline__ __nv_bfloat16 f2bf(float x) { return __float2bfloat16_rn(x); }

// 8 x bf16 <-> 8 x float through one 16-byte vector
struct alignas(16) bf16x8 { __nv_bfloat162 h[4]; };
__device__ __forceinline__ void unpack8(const bf16x8& u, float* f) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __bfloat1622float2(u.h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 u;
#pragma unroll
  for (int j = 0; j < 4; ++j) u.h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
  return u;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Counter-based RNG for dropout: keep-decision for element `idx` of dropout site `site` at step seed `seed`.
// (murmur3-style finaliser over a 64-bit counter; masks are recomputed in backward, never stored.)
__device__ __forceinline__ uint32_t hash_u32(uint64_t seed, uint32_t site, uint64_t idx) {
  uint64_t x = seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(site + 1)) ^ (idx * 0xD6E8FEB86659FD93ull);
  x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull;
  x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull;
  x ^= x >> 32;
  return (uint32_t)x;
}
// returns scale to multiply by: 0 if dropped, 1/(1-p) if kept.  p == 0 -> always 1.
__device__ __forceinline__ float dropout_scale(float p, float inv_keep, uint64_t seed, uint32_t site, uint64_t idx) {
  if (p <= 0.f) return 1.f;
  const uint32_t h = hash_u32(seed, site, idx);
  const float u = (float)(h >> 8) * (1.0f / 16777216.0f);
  return u < p ? 0.f : inv_keep;
}
}  // namespace vtx
