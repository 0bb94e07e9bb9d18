// Memory/latency-bound kernels of the TransformerDecoder textual head (everything that is not a GEMM):
// fused embedding (gather + LayerNorm(1e-8) + dropout + pad mask), residual-add + dropout + LayerNorm, multi-head
// attention cores for T<=32 queries / S<=64 keys with the causal + key-padding mask generated from caption_lengths
// (never materialised), exact-erf GELU + dropout, token cross-entropy (fwd + dlogits in place), bias-gradient column
// sums and argmax.  Reference semantics: virtex/modules/embedding.py:46-74, torch/nn/modules/transformer.py:1144-1199,
// torch/nn/functional.py:6244-6690, virtex/models/captioning.py:111-114.
// The decoder residual stream is fp32; GEMM operands/outputs are bf16 (the bf16-autocast placement of the reference).
#include "vtx_common.cuh"
#include "../../include/virtex_b200.h"

namespace vtx {

constexpr int kWarpsPerBlock = 4;

// ------------------------------------------------------------------------------------------------ LayerNorm helpers
// one warp per row; two-pass statistics (mean, then centred variance) in fp32
__device__ __forceinline__ void warp_row_stats(const float* row, int H, int lane, float eps, float& mean, float& rstd) {
  float s = 0.f;
  for (int i = lane * 4; i < H; i += 128) {
    const float4 v = *reinterpret_cast<const float4*>(row + i);
    s += v.x + v.y + v.z + v.w;
  }
  mean = warp_sum(s) / H;
  float q = 0.f;
  for (int i = lane * 4; i < H; i += 128) {
    const float4 v = *reinterpret_cast<const float4*>(row + i);
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    q += a * a + b * b + c * c + d * d;
  }
  rstd = rsqrtf(warp_sum(q) / H + eps);
}

// ------------------------------------------------------------------------------------------------ embedding
// z = words[tok] + positions[t];  out = LN_eps(z) -> dropout -> * [tok != pad]
__global__ void embed_fwd_kernel(const long long* __restrict__ tokens, const float* __restrict__ words,
                                 const float* __restrict__ positions, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float* __restrict__ z, float* __restrict__ stats,
                                 float* __restrict__ out, __nv_bfloat16* __restrict__ out_bf, int M, int T, int H,
                                 int pad, float eps, float p, const uint64_t* seed_ptr, uint32_t site) {
  const uint64_t seed = seed_ptr ? *seed_ptr : 0ull;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= M) return;
  const long long tok = tokens[row];
  const int t = row % T;
  float* zr = z + (long long)row * H;
  for (int i = lane * 4; i < H; i += 128) {
    const float4 a = *reinterpret_cast<const float4*>(words + tok * H + i);
    const float4 b = *reinterpret_cast<const float4*>(positions + (long long)t * H + i);
    *reinterpret_cast<float4*>(zr + i) = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
  __syncwarp();
  float mean, rstd;
  warp_row_stats(zr, H, lane, eps, mean, rstd);
  if (lane == 0) {
    stats[2 * row] = mean;
    stats[2 * row + 1] = rstd;
  }
  const float keep = (tok != pad) ? 1.f : 0.f;
  const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  for (int i = lane * 4; i < H; i += 128) {
    const float4 v = *reinterpret_cast<const float4*>(zr + i);
    const float4 g = *reinterpret_cast<const float4*>(gamma + i);
    const float4 b = *reinterpret_cast<const float4*>(beta + i);
    float o[4] = {(v.x - mean) * rstd * g.x + b.x, (v.y - mean) * rstd * g.y + b.y, (v.z - mean) * rstd * g.z + b.z,
                  (v.w - mean) * rstd * g.w + b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] *= keep * dropout_scale(p, inv_keep, seed, site, (uint64_t)row * H + i + j);
    *reinterpret_cast<float4*>(out + (long long)row * H + i) = make_float4(o[0], o[1], o[2], o[3]);
    __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[3]);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(out_bf + (long long)row * H + i) = u;
  }
}

// upstream g = (dy_a + dy_b) * [tok != pad] * dropmask -> LN backward -> scatter-add into d_words[tok], d_positions[t]
__global__ void embed_bwd_kernel(const float* __restrict__ dy_a, const __nv_bfloat16* __restrict__ dy_b,
                                 const long long* __restrict__ tokens, const float* __restrict__ z,
                                 const float* __restrict__ stats, const float* __restrict__ gamma,
                                 float* __restrict__ d_words, float* __restrict__ d_pos, float* __restrict__ d_gamma,
                                 float* __restrict__ d_beta, int M, int T, int H, int pad, float p, const uint64_t* seed_ptr,
                                 uint32_t site) {
  const uint64_t seed = seed_ptr ? *seed_ptr : 0ull;
  extern __shared__ float acc[];  // [2][H] : dgamma, dbeta partials of this CTA
  for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  for (int row = blockIdx.x * kWarpsPerBlock + warp; row < M; row += gridDim.x * kWarpsPerBlock) {
    const long long tok = tokens[row];
    if (tok == pad) continue;  // zero upstream gradient: contributes nothing anywhere
    const int t = row % T;
    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
    const float* zr = z + (long long)row * H;
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < H; i += 32) {
      float g = dy_a ? dy_a[(long long)row * H + i] : 0.f;
      if (dy_b) g += bf2f(dy_b[(long long)row * H + i]);
      g *= dropout_scale(p, inv_keep, seed, site, (uint64_t)row * H + i);
      const float xh = (zr[i] - mean) * rstd;
      atomicAdd(&acc[i], g * xh);
      atomicAdd(&acc[H + i], g);
      const float dxh = g * gamma[i];
      s1 += dxh;
      s2 += dxh * xh;
    }
    s1 = warp_sum(s1) / H;
    s2 = warp_sum(s2) / H;
    for (int i = lane; i < H; i += 32) {
      float g = dy_a ? dy_a[(long long)row * H + i] : 0.f;
      if (dy_b) g += bf2f(dy_b[(long long)row * H + i]);
      g *= dropout_scale(p, inv_keep, seed, site, (uint64_t)row * H + i);
      const float xh = (zr[i] - mean) * rstd;
      const float dz = rstd * (g * gamma[i] - s1 - xh * s2);
      atomicAdd(d_words + tok * H + i, dz);
      atomicAdd(d_pos + (long long)t * H + i, dz);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    atomicAdd(d_gamma + i, acc[i]);
    atomicAdd(d_beta + i, acc[H + i]);
  }
}

// ------------------------------------------------------------------------------------------------ add + dropout + LN
// z = res + dropout(branch);  out = LN(z)*gamma + beta     (ln == 0: out = z, the pre-norm residual update)
__global__ void add_ln_fwd_kernel(const float* __restrict__ res, const __nv_bfloat16* __restrict__ branch,
                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                  float* __restrict__ z, float* __restrict__ stats, float* __restrict__ out,
                                  __nv_bfloat16* __restrict__ out_bf, int M, int H, float eps, float p, const uint64_t* seed_ptr,
                                  uint32_t site, int ln) {
  const uint64_t seed = seed_ptr ? *seed_ptr : 0ull;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= M) return;
  const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  float* zr = z + (long long)row * H;
  for (int i = lane * 4; i < H; i += 128) {
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (res) r = *reinterpret_cast<const float4*>(res + (long long)row * H + i);
    if (branch) {
      const uint2 u = *reinterpret_cast<const uint2*>(branch + (long long)row * H + i);
      const float2 b0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
      const float2 b1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
      const uint64_t e = (uint64_t)row * H + i;
      r.x += b0.x * dropout_scale(p, inv_keep, seed, site, e);
      r.y += b0.y * dropout_scale(p, inv_keep, seed, site, e + 1);
      r.z += b1.x * dropout_scale(p, inv_keep, seed, site, e + 2);
      r.w += b1.y * dropout_scale(p, inv_keep, seed, site, e + 3);
    }
    *reinterpret_cast<float4*>(zr + i) = r;
  }
  __syncwarp();
  float mean = 0.f, rstd = 1.f;
  if (ln) {
    warp_row_stats(zr, H, lane, eps, mean, rstd);
    if (lane == 0) {
      stats[2 * row] = mean;
      stats[2 * row + 1] = rstd;
    }
  }
  for (int i = lane * 4; i < H; i += 128) {
    const float4 v = *reinterpret_cast<const float4*>(zr + i);
    float o[4] = {v.x, v.y, v.z, v.w};
    if (ln) {
      const float4 g = *reinterpret_cast<const float4*>(gamma + i);
      const float4 b = *reinterpret_cast<const float4*>(beta + i);
      o[0] = (v.x - mean) * rstd * g.x + b.x;
      o[1] = (v.y - mean) * rstd * g.y + b.y;
      o[2] = (v.z - mean) * rstd * g.z + b.z;
      o[3] = (v.w - mean) * rstd * g.w + b.w;
    }
    if (out) *reinterpret_cast<float4*>(out + (long long)row * H + i) = make_float4(o[0], o[1], o[2], o[3]);
    if (out_bf) {
      __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[3]);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0);
      u.y = *reinterpret_cast<uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(out_bf + (long long)row * H + i) = u;
    }
  }
}

// g = dy_a + dy_b;  LN backward -> dz;  d_res = dz (+ d_skip);  d_branch = dz * dropmask (bf16);  dgamma/dbeta +=
// ln == 0: dz = g (plain residual split).
__global__ void ln_bwd_kernel(const float* __restrict__ dy_a, const __nv_bfloat16* __restrict__ dy_b,
                              const float* __restrict__ z, const float* __restrict__ stats,
                              const float* __restrict__ gamma, const float* __restrict__ d_skip,
                              float* __restrict__ d_res, __nv_bfloat16* __restrict__ d_branch,
                              float* __restrict__ d_gamma, float* __restrict__ d_beta, int M, int H, float p,
                              const uint64_t* seed_ptr, uint32_t site, int ln) {
  const uint64_t seed = seed_ptr ? *seed_ptr : 0ull;
  extern __shared__ float acc[];  // [warps][2][H] per-warp partial dgamma / dbeta (no atomics in the row loop)
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float* my = acc + (size_t)warp * 2 * H;
  if (ln) {
    for (int i = lane; i < 2 * H; i += 32) my[i] = 0.f;
    __syncwarp();
  }
  const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  for (int row = blockIdx.x * kWarpsPerBlock + warp; row < M; row += gridDim.x * kWarpsPerBlock) {
    const long long base = (long long)row * H;
    float mean = 0.f, rstd = 1.f, s1 = 0.f, s2 = 0.f;
    if (ln) {
      mean = stats[2 * row];
      rstd = stats[2 * row + 1];
      for (int i = lane * 4; i < H; i += 128) {
        float g[4] = {0.f, 0.f, 0.f, 0.f};
        if (dy_a) {
          const float4 t = *reinterpret_cast<const float4*>(dy_a + base + i);
          g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
        }
        if (dy_b) {
          const uint2 u = *reinterpret_cast<const uint2*>(dy_b + base + i);
          const float2 b0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
          const float2 b1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
          g[0] += b0.x; g[1] += b0.y; g[2] += b1.x; g[3] += b1.y;
        }
        const float4 zz = *reinterpret_cast<const float4*>(z + base + i);
        const float4 gm = *reinterpret_cast<const float4*>(gamma + i);
        const float xh[4] = {(zz.x - mean) * rstd, (zz.y - mean) * rstd, (zz.z - mean) * rstd, (zz.w - mean) * rstd};
        const float gw[4] = {gm.x, gm.y, gm.z, gm.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          my[i + j] += g[j] * xh[j];
          my[H + i + j] += g[j];
          const float dxh = g[j] * gw[j];
          s1 += dxh;
          s2 += dxh * xh[j];
        }
      }
      s1 = warp_sum(s1) / H;
      s2 = warp_sum(s2) / H;
    }
    for (int i = lane * 4; i < H; i += 128) {
      float g[4] = {0.f, 0.f, 0.f, 0.f};
      if (dy_a) {
        const float4 t = *reinterpret_cast<const float4*>(dy_a + base + i);
        g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
      }
      if (dy_b) {
        const uint2 u = *reinterpret_cast<const uint2*>(dy_b + base + i);
        const float2 b0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
        const float2 b1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
        g[0] += b0.x; g[1] += b0.y; g[2] += b1.x; g[3] += b1.y;
      }
      float dz[4] = {g[0], g[1], g[2], g[3]};
      if (ln) {
        const float4 zz = *reinterpret_cast<const float4*>(z + base + i);
        const float4 gm = *reinterpret_cast<const float4*>(gamma + i);
        dz[0] = rstd * (g[0] * gm.x - s1 - (zz.x - mean) * rstd * s2);
        dz[1] = rstd * (g[1] * gm.y - s1 - (zz.y - mean) * rstd * s2);
        dz[2] = rstd * (g[2] * gm.z - s1 - (zz.z - mean) * rstd * s2);
        dz[3] = rstd * (g[3] * gm.w - s1 - (zz.w - mean) * rstd * s2);
      }
      if (d_branch) {
        float t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] = dz[j] * dropout_scale(p, inv_keep, seed, site, (uint64_t)base + i + j);
        __nv_bfloat162 h0 = __floats2bfloat162_rn(t[0], t[1]), h1 = __floats2bfloat162_rn(t[2], t[3]);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&h0);
        u.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(d_branch + base + i) = u;
      }
      if (d_res) {
        float4 o = make_float4(dz[0], dz[1], dz[2], dz[3]);
        if (d_skip) {
          const float4 k = *reinterpret_cast<const float4*>(d_skip + base + i);
          o.x += k.x; o.y += k.y; o.z += k.z; o.w += k.w;
        }
        *reinterpret_cast<float4*>(d_res + base + i) = o;
      }
    }
  }
  if (ln) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kWarpsPerBlock; ++w) t += acc[(size_t)w * 2 * H + i];
      atomicAdd((i < H ? d_gamma + i : d_beta + (i - H)), t);
    }
  }
}

// ------------------------------------------------------------------------------------------------ attention
// One warp per (batch b, head h); head_dim = 64; Tq <= 32 queries (lane = query), Tk <= 64 keys.
// causal != 0: key j allowed for query i iff j <= i and j < lengths[b]  (self-attention of the captioning head)
// causal == 0: all Tk keys allowed (cross-attention over the visual grid).
constexpr int kD = 64;
constexpr int kPS = 65;  // padded row stride (floats) of per-query smem rows -> conflict free for lane = row access

struct AttnArgs {
  const __nv_bfloat16 *q, *k, *v;
  long long ldq, ldk, ldv;
  int B, heads, Tq, Tk;
  const long long* lengths;
  int causal;
  float scale, p;
  const uint64_t* seed_ptr;
  uint32_t site;
};

__device__ __forceinline__ void load_rows_f32(float* dst, int dst_stride, const __nv_bfloat16* src, long long ld,
                                              int rows, int lane) {
  // rows x 64 bf16 -> fp32 smem; each lane moves 8 elements (16 B) per step
  for (int e = lane; e < rows * 8; e += 32) {
    const int r = e >> 3, c = (e & 7) * 8;
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(src + (long long)r * ld + c), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[r * dst_stride + c + j] = f[j];
  }
}

__global__ void __launch_bounds__(128) attn_fwd_kernel(const AttnArgs a, __nv_bfloat16* __restrict__ out, long long ldo,
                                                        float* __restrict__ lse) {
  const uint64_t seed = a.seed_ptr ? *a.seed_ptr : 0ull;
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x * kWarpsPerBlock + warp;
  if (unit >= a.B * a.heads) return;
  const int b = unit / a.heads, h = unit % a.heads;
  float* sK = sm + (size_t)warp * (2 * 64 * kD + 32 * kPS);
  float* sV = sK + 64 * kD;
  float* sS = sV + 64 * kD;  // [32][kPS]
  load_rows_f32(sK, kD, a.k + (long long)b * a.Tk * a.ldk + h * kD, a.ldk, a.Tk, lane);
  load_rows_f32(sV, kD, a.v + (long long)b * a.Tk * a.ldv + h * kD, a.ldv, a.Tk, lane);
  __syncwarp();
  const int i = lane;
  const bool active = i < a.Tq;
  const int len = a.causal ? (int)a.lengths[b] : a.Tk;
  float q[kD];
  if (active) {
    const __nv_bfloat16* qp = a.q + ((long long)b * a.Tq + i) * a.ldq + h * kD;
#pragma unroll
    for (int c = 0; c < kD; c += 8) unpack8(*reinterpret_cast<const bf16x8*>(qp + c), q + c);
  } else {
#pragma unroll
    for (int c = 0; c < kD; ++c) q[c] = 0.f;
  }
  float mx = -INFINITY;
  for (int j = 0; j < a.Tk; ++j) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kD; c += 4) {
      const float4 kk = *reinterpret_cast<const float4*>(sK + j * kD + c);
      s += q[c] * kk.x + q[c + 1] * kk.y + q[c + 2] * kk.z + q[c + 3] * kk.w;
    }
    s *= a.scale;
    const bool ok = a.causal ? (j <= i && j < len) : true;
    s = ok ? s : -INFINITY;
    sS[i * kPS + j] = s;
    mx = fmaxf(mx, s);
  }
  float o[kD];
#pragma unroll
  for (int c = 0; c < kD; ++c) o[c] = 0.f;
  float l = 0.f;
  const float inv_keep = a.p > 0.f ? 1.f / (1.f - a.p) : 1.f;
  for (int j = 0; j < a.Tk; ++j) {
    const float s = sS[i * kPS + j];
    float pr = (s == -INFINITY) ? 0.f : __expf(s - mx);
    l += pr;
    pr *= dropout_scale(a.p, inv_keep, seed, a.site, ((uint64_t)unit * 32 + i) * 64 + j);
#pragma unroll
    for (int c = 0; c < kD; c += 4) {
      const float4 vv = *reinterpret_cast<const float4*>(sV + j * kD + c);
      o[c] += pr * vv.x; o[c + 1] += pr * vv.y; o[c + 2] += pr * vv.z; o[c + 3] += pr * vv.w;
    }
  }
  if (active) {
    const float inv = 1.f / l;
    __nv_bfloat16* op = out + ((long long)b * a.Tq + i) * ldo + h * kD;
#pragma unroll
    for (int c = 0; c < kD; c += 8) {
      float t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = o[c + j] * inv;
      *reinterpret_cast<bf16x8*>(op + c) = pack8(t);
    }
    if (lse) lse[(long long)unit * 32 + i] = mx + __logf(l);
  }
}

// backward: dq [.., ldq-like layout given by ldgq], dk, dv
__global__ void __launch_bounds__(64) attn_bwd_kernel(const AttnArgs a, const __nv_bfloat16* __restrict__ dout,
                                                       long long ldo, const float* __restrict__ lse,
                                                       __nv_bfloat16* __restrict__ dq, long long lddq,
                                                       __nv_bfloat16* __restrict__ dk, long long lddk,
                                                       __nv_bfloat16* __restrict__ dv, long long lddv) {
  const uint64_t seed = a.seed_ptr ? *a.seed_ptr : 0ull;
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x * 2 + warp;
  if (unit >= a.B * a.heads) return;
  const int b = unit / a.heads, h = unit % a.heads;
  constexpr int per_warp = 2 * 64 * kD + 4 * 32 * kPS;
  float* sK = sm + (size_t)warp * per_warp;
  float* sV = sK + 64 * kD;
  float* sQ = sV + 64 * kD;   // [32][kPS]
  float* sdO = sQ + 32 * kPS;  // [32][kPS]
  float* sP = sdO + 32 * kPS;  // [32][kPS]  dropped probabilities Pd
  float* sdS = sP + 32 * kPS;  // [32][kPS]
  load_rows_f32(sK, kD, a.k + (long long)b * a.Tk * a.ldk + h * kD, a.ldk, a.Tk, lane);
  load_rows_f32(sV, kD, a.v + (long long)b * a.Tk * a.ldv + h * kD, a.ldv, a.Tk, lane);
  load_rows_f32(sQ, kPS, a.q + (long long)b * a.Tq * a.ldq + h * kD, a.ldq, a.Tq, lane);
  load_rows_f32(sdO, kPS, dout + (long long)b * a.Tq * ldo + h * kD, ldo, a.Tq, lane);
  __syncwarp();
  const int i = lane;
  const bool active = i < a.Tq;
  const int len = a.causal ? (int)a.lengths[b] : a.Tk;
  const float inv_keep = a.p > 0.f ? 1.f / (1.f - a.p) : 1.f;
  // ---- phase A (lane = query): P, dP, D_i, dS, dq
  const float L = active ? lse[(long long)unit * 32 + i] : 0.f;
  float Di = 0.f;
  for (int j = 0; j < a.Tk; ++j) {
    float s = 0.f, dp = 0.f;
    if (active) {
#pragma unroll 16
      for (int c = 0; c < kD; ++c) {
        s += sQ[i * kPS + c] * sK[j * kD + c];
        dp += sdO[i * kPS + c] * sV[j * kD + c];
      }
    }
    s *= a.scale;
    const bool ok = active && (a.causal ? (j <= i && j < len) : true);
    const float pr = ok ? __expf(s - L) : 0.f;
    const float mk = dropout_scale(a.p, inv_keep, seed, a.site, ((uint64_t)unit * 32 + i) * 64 + j);
    const float dpr = dp * mk;  // dP = dPd * mask
    Di += pr * dpr;
    sP[i * kPS + j] = pr * mk;   // Pd
    sdS[i * kPS + j] = dpr;      // dP for now (turned into dS by the second pass)
  }
  float dqv[kD];
#pragma unroll
  for (int c = 0; c < kD; ++c) dqv[c] = 0.f;
  for (int j = 0; j < a.Tk; ++j) {
    // recompute s -> P (cheap: 64 FMAs) to form dS exactly, including dropped entries
    float s = 0.f;
    if (active) {
#pragma unroll 16
      for (int c = 0; c < kD; ++c) s += sQ[i * kPS + c] * sK[j * kD + c];
    }
    s *= a.scale;
    const bool ok = active && (a.causal ? (j <= i && j < len) : true);
    const float pr = ok ? __expf(s - L) : 0.f;
    const float ds = pr * (sdS[i * kPS + j] - Di) * a.scale;
    sdS[i * kPS + j] = ds;
#pragma unroll
    for (int c = 0; c < kD; c += 4) {
      const float4 kk = *reinterpret_cast<const float4*>(sK + j * kD + c);
      dqv[c] += ds * kk.x; dqv[c + 1] += ds * kk.y; dqv[c + 2] += ds * kk.z; dqv[c + 3] += ds * kk.w;
    }
  }
  if (active) {
    __nv_bfloat16* p = dq + ((long long)b * a.Tq + i) * lddq + h * kD;
#pragma unroll
    for (int c = 0; c < kD; c += 8) *reinterpret_cast<bf16x8*>(p + c) = pack8(dqv + c);
  }
  __syncwarp();
  // ---- phase B (lane = key): dV_j = sum_i Pd_ij dO_i ; dK_j = sum_i dS_ij Q_i
  for (int j0 = 0; j0 < a.Tk; j0 += 32) {
    const int j = j0 + lane;
    float dvv[kD], dkv[kD];
#pragma unroll
    for (int c = 0; c < kD; ++c) dvv[c] = dkv[c] = 0.f;
    if (j < a.Tk) {
      for (int ii = 0; ii < a.Tq; ++ii) {
        const float pd = sP[ii * kPS + j];
        const float ds = sdS[ii * kPS + j];
#pragma unroll
        for (int c = 0; c < kD; ++c) {
          dvv[c] += pd * sdO[ii * kPS + c];
          dkv[c] += ds * sQ[ii * kPS + c];
        }
      }
      __nv_bfloat16* pk = dk + ((long long)b * a.Tk + j) * lddk + h * kD;
      __nv_bfloat16* pv = dv + ((long long)b * a.Tk + j) * lddv + h * kD;
#pragma unroll
      for (int c = 0; c < kD; c += 8) {
        *reinterpret_cast<bf16x8*>(pk + c) = pack8(dkv + c);
        *reinterpret_cast<bf16x8*>(pv + c) = pack8(dvv + c);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ GELU + dropout
__global__ void gelu_dropout_fwd_kernel(const __nv_bfloat16* __restrict__ u, __nv_bfloat16* __restrict__ h,
                                        long long n8, float p, const uint64_t* seed_ptr, uint32_t site) {
  const uint64_t seed = seed_ptr ? *seed_ptr : 0ull;
  const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(u + i * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // the reference evaluates GELU on the bf16 tensor and rounds the result to bf16 before dropout
      const float g = bf2f(f2bf(0.5f * f[j] * (1.0f + erff(f[j] * 0.70710678118654752f))));
      f[j] = g * dropout_scale(p, inv_keep, seed, site, (uint64_t)i * 8 + j);
    }
    *reinterpret_cast<bf16x8*>(h + i * 8) = pack8(f);
  }
}
// du = dh * dropmask * gelu'(u)     (in place over dh allowed)
__global__ void gelu_dropout_bwd_kernel(const __nv_bfloat16* __restrict__ dh, const __nv_bfloat16* __restrict__ u,
                                        __nv_bfloat16* __restrict__ du, long long n8, float p, const uint64_t* seed_ptr,
                                        uint32_t site) {
  const uint64_t seed = seed_ptr ? *seed_ptr : 0ull;
  const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float d[8], x[8];
    unpack8(*reinterpret_cast<const bf16x8*>(dh + i * 8), d);
    unpack8(*reinterpret_cast<const bf16x8*>(u + i * 8), x);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float cdf = 0.5f * (1.0f + erff(x[j] * 0.70710678118654752f));
      const float pdf = 0.3989422804014327f * __expf(-0.5f * x[j] * x[j]);
      d[j] = d[j] * dropout_scale(p, inv_keep, seed, site, (uint64_t)i * 8 + j) * (cdf + x[j] * pdf);
    }
    *reinterpret_cast<bf16x8*>(du + i * 8) = pack8(d);
  }
}

// ------------------------------------------------------------------------------------------------ cross entropy
// counts[0] = number of targets tokens[b, t>=1] != pad
__global__ void count_valid_kernel(const long long* __restrict__ tokens, int B, int T, int pad, float* __restrict__ count) {
  float c = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * T; i += gridDim.x * blockDim.x)
    if ((i % T) >= 1 && tokens[i] != pad) c += 1.f;
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0 && c != 0.f) atomicAdd(count, c);
}

// One CTA per row (b,t) of bf16 logits [B*T, ldl].  Target = tokens[b,t+1] for t < T-1 (else ignored); ignored when
// == pad.  loss += nll / count;  if write_grad: logits row overwritten by dlogits = (softmax - onehot)/count (or 0).
__global__ void ce_kernel(__nv_bfloat16* __restrict__ logits, long long ldl, const long long* __restrict__ tokens, int T,
                          int V, int pad, const float* __restrict__ count, float* __restrict__ loss, int write_grad) {
  __shared__ float red[32];
  __shared__ float bcast;
  const int row = blockIdx.x;
  const int t = row % T;
  __nv_bfloat16* z = logits + (long long)row * ldl;
  const long long target = (t < T - 1) ? tokens[row + 1] : (long long)pad;
  const bool valid = target != pad;
  const int nv = V / 8;
  if (!valid) {
    if (write_grad)
      for (int i = threadIdx.x; i < nv; i += blockDim.x) *reinterpret_cast<uint4*>(z + i * 8) = make_uint4(0, 0, 0, 0);
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(z + i * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) mx = fmaxf(mx, f[j]);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (warp == 0) {
    float v = lane < nwarps ? red[lane] : -INFINITY;
    v = warp_max(v);
    if (lane == 0) bcast = v;
  }
  __syncthreads();
  mx = bcast;
  float s = 0.f;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(z + i * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += __expf(f[j] - mx);
  }
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (warp == 0) {
    float v = lane < nwarps ? red[lane] : 0.f;
    v = warp_sum(v);
    if (lane == 0) bcast = v;
  }
  __syncthreads();
  s = bcast;
  const float inv_n = 1.f / fmaxf(*count, 1.f);
  if (threadIdx.x == 0) {
    const float zt = bf2f(z[target]);
    atomicAdd(loss, (mx + __logf(s) - zt) * inv_n);
  }
  if (write_grad) {
    const float inv_s = 1.f / s;
    __syncthreads();  // everyone (incl. thread 0's read of z[target]) is done with the original logits
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
      float f[8];
      unpack8(*reinterpret_cast<const bf16x8*>(z + i * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float pr = __expf(f[j] - mx) * inv_s;
        f[j] = (pr - ((long long)(i * 8 + j) == target ? 1.f : 0.f)) * inv_n;
      }
      *reinterpret_cast<bf16x8*>(z + i * 8) = pack8(f);
    }
  }
}

// out[n] += sum_m X[m,n]    X bf16 [M, ld]
__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ X, long long ld, int M, int N, float* __restrict__ out,
                              int rows_per_block) {
  const int g = blockIdx.y * blockDim.x + threadIdx.x;  // 8-column group
  if (g * 8 >= N) return;
  const int m0 = blockIdx.x * rows_per_block;
  const int m1 = min(M, m0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int m = m0; m < m1; ++m) {
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(X + (long long)m * ld + g * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += f[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (g * 8 + j < N) atomicAdd(out + g * 8 + j, acc[j]);
}

// first-index argmax of each fp32 row
__global__ void argmax_rows_kernel(const float* __restrict__ X, long long ld, int N, long long* __restrict__ out) {
  __shared__ float bv[32];
  __shared__ int bi[32];
  const float* x = X + (long long)blockIdx.x * ld;
  float best = -INFINITY;
  int idx = 0x7fffffff;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float v = x[i];
    if (v > best || (v == best && i < idx)) { best = v; idx = i; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { bv[warp] = best; bi[warp] = idx; }
  __syncthreads();
  if (warp == 0) {
    best = lane < (blockDim.x >> 5) ? bv[lane] : -INFINITY;
    idx = lane < (blockDim.x >> 5) ? bi[lane] : 0x7fffffff;
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
    }
    if (lane == 0) out[blockIdx.x] = idx;
  }
}

}  // namespace vtx

using namespace vtx;
#define STREAM reinterpret_cast<cudaStream_t>(stream)
#define REQ(cond, msg) \
  if (!(cond)) return set_error(VTX_EINVAL, "%s: %s", __func__, msg)

extern "C" int vtx_embed_fwd(const int64_t* tokens, const float* words, const float* positions, const float* gamma,
                             const float* beta, float* z, float* stats, float* out, void* out_bf, int M, int T, int H,
                             int pad, float eps, float p, const uint64_t* seed_ptr, uint32_t site, void* stream) {
  REQ(tokens && words && positions && gamma && beta && z && stats && out && out_bf && H % 128 == 0, "bad arguments");
  embed_fwd_kernel<<<(M + kWarpsPerBlock - 1) / kWarpsPerBlock, 32 * kWarpsPerBlock, 0, STREAM>>>(
      (const long long*)tokens, words, positions, gamma, beta, z, stats, out, (__nv_bfloat16*)out_bf, M, T, H, pad, eps,
      p, seed_ptr, site);
  return check_launch("embed_fwd");
}
extern "C" int vtx_embed_bwd(const float* dy_a, const void* dy_b, const int64_t* tokens, const float* z,
                             const float* stats, const float* gamma, float* d_words, float* d_pos, float* d_gamma,
                             float* d_beta, int M, int T, int H, int pad, float p, const uint64_t* seed_ptr, uint32_t site,
                             void* stream) {
  REQ(tokens && z && stats && gamma && d_words && d_pos && d_gamma && d_beta && (dy_a || dy_b), "bad arguments");
  int blocks = (M + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int cap = vtx_num_sms() * 2;
  if (blocks > cap) blocks = cap;
  embed_bwd_kernel<<<blocks, 32 * kWarpsPerBlock, 2 * H * sizeof(float), STREAM>>>(
      dy_a, (const __nv_bfloat16*)dy_b, (const long long*)tokens, z, stats, gamma, d_words, d_pos, d_gamma, d_beta, M, T,
      H, pad, p, seed_ptr, site);
  return check_launch("embed_bwd");
}
extern "C" int vtx_add_ln_fwd(const float* res, const void* branch, const float* gamma, const float* beta, float* z,
                              float* stats, float* out, void* out_bf, int M, int H, float eps, float p, const uint64_t* seed_ptr,
                              uint32_t site, int ln, void* stream) {
  REQ(z && (res || branch) && H % 128 == 0 && (!ln || (gamma && beta && stats)), "bad arguments");
  add_ln_fwd_kernel<<<(M + kWarpsPerBlock - 1) / kWarpsPerBlock, 32 * kWarpsPerBlock, 0, STREAM>>>(
      res, (const __nv_bfloat16*)branch, gamma, beta, z, stats, out, (__nv_bfloat16*)out_bf, M, H, eps, p, seed_ptr, site,
      ln);
  return check_launch("add_ln_fwd");
}
extern "C" int vtx_ln_bwd(const float* dy_a, const void* dy_b, const float* z, const float* stats, const float* gamma,
                          const float* d_skip, float* d_res, void* d_branch, float* d_gamma, float* d_beta, int M,
                          int H, float p, const uint64_t* seed_ptr, uint32_t site, int ln, void* stream) {
  REQ((dy_a || dy_b) && (!ln || (z && stats && gamma && d_gamma && d_beta)), "bad arguments");
  int blocks = (M + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int cap = vtx_num_sms() * 2;
  if (blocks > cap) blocks = cap;
  const size_t ln_smem = ln ? (size_t)kWarpsPerBlock * 2 * H * sizeof(float) : 0;
  static size_t ln_attr = 0;
  if (ln_smem > 48 * 1024 && ln_smem > ln_attr) {
    cudaFuncSetAttribute(ln_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ln_smem);
    ln_attr = ln_smem;
  }
  ln_bwd_kernel<<<blocks, 32 * kWarpsPerBlock, ln_smem, STREAM>>>(
      dy_a, (const __nv_bfloat16*)dy_b, z, stats, gamma, d_skip, d_res, (__nv_bfloat16*)d_branch, d_gamma, d_beta, M, H,
      p, seed_ptr, site, ln);
  return check_launch("ln_bwd");
}

static int fill_attn(AttnArgs* a, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                     int B, int heads, int Tq, int Tk, const int64_t* lengths, int causal, float p, const uint64_t* seed_ptr,
                     uint32_t site) {
  if (!q || !k || !v || Tq < 1 || Tq > 32 || Tk < 1 || Tk > 64 || (causal && !lengths))
    return set_error(VTX_EINVAL, "attention: unsupported shape (Tq<=32, Tk<=64, head_dim 64)");
  if (ldq % 8 || ldk % 8 || ldv % 8) return set_error(VTX_EINVAL, "attention: leading dims must be multiples of 8");
  a->q = (const __nv_bfloat16*)q; a->k = (const __nv_bfloat16*)k; a->v = (const __nv_bfloat16*)v;
  a->ldq = ldq; a->ldk = ldk; a->ldv = ldv;
  a->B = B; a->heads = heads; a->Tq = Tq; a->Tk = Tk;
  a->lengths = (const long long*)lengths; a->causal = causal;
  a->scale = 0.125f;  // 1/sqrt(64)
  a->p = p; a->seed_ptr = seed_ptr; a->site = site;
  return VTX_OK;
}

extern "C" int vtx_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                            void* out, int64_t ldo, float* lse, int B, int heads, int Tq, int Tk,
                            const int64_t* lengths, int causal, float p, const uint64_t* seed_ptr, uint32_t site, void* stream) {
  AttnArgs a;
  int rc = fill_attn(&a, q, ldq, k, ldk, v, ldv, B, heads, Tq, Tk, lengths, causal, p, seed_ptr, site);
  if (rc) return rc;
  REQ(out && ldo % 8 == 0, "bad output");
  const size_t smem = (size_t)kWarpsPerBlock * (2 * 64 * kD + 32 * kPS) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  const int units = B * heads;
  attn_fwd_kernel<<<(units + kWarpsPerBlock - 1) / kWarpsPerBlock, 32 * kWarpsPerBlock, smem, STREAM>>>(
      a, (__nv_bfloat16*)out, ldo, lse);
  return check_launch("attn_fwd");
}
extern "C" int vtx_attn_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                            const void* dout, int64_t ldo, const float* lse, void* dq, int64_t lddq, void* dk,
                            int64_t lddk, void* dv, int64_t lddv, int B, int heads, int Tq, int Tk,
                            const int64_t* lengths, int causal, float p, const uint64_t* seed_ptr, uint32_t site, void* stream) {
  AttnArgs a;
  int rc = fill_attn(&a, q, ldq, k, ldk, v, ldv, B, heads, Tq, Tk, lengths, causal, p, seed_ptr, site);
  if (rc) return rc;
  REQ(dout && lse && dq && dk && dv && ldo % 8 == 0 && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0, "bad arguments");
  const size_t smem = (size_t)2 * (2 * 64 * kD + 4 * 32 * kPS) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  const int units = B * heads;
  attn_bwd_kernel<<<(units + 1) / 2, 64, smem, STREAM>>>(a, (const __nv_bfloat16*)dout, ldo, lse, (__nv_bfloat16*)dq,
                                                         lddq, (__nv_bfloat16*)dk, lddk, (__nv_bfloat16*)dv, lddv);
  return check_launch("attn_bwd");
}
extern "C" int vtx_gelu_dropout_fwd(const void* u, void* h, int64_t n, float p, const uint64_t* seed_ptr, uint32_t site,
                                    void* stream) {
  REQ(u && h && n % 8 == 0, "bad arguments");
  const long long n8 = n / 8;
  long long blocks = (n8 + 255) / 256;
  const long long cap = (long long)vtx_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  gelu_dropout_fwd_kernel<<<(int)blocks, 256, 0, STREAM>>>((const __nv_bfloat16*)u, (__nv_bfloat16*)h, n8, p, seed_ptr, site);
  return check_launch("gelu_dropout_fwd");
}
extern "C" int vtx_gelu_dropout_bwd(const void* dh, const void* u, void* du, int64_t n, float p, const uint64_t* seed_ptr,
                                    uint32_t site, void* stream) {
  REQ(dh && u && du && n % 8 == 0, "bad arguments");
  const long long n8 = n / 8;
  long long blocks = (n8 + 255) / 256;
  const long long cap = (long long)vtx_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  gelu_dropout_bwd_kernel<<<(int)blocks, 256, 0, STREAM>>>((const __nv_bfloat16*)dh, (const __nv_bfloat16*)u,
                                                           (__nv_bfloat16*)du, n8, p, seed_ptr, site);
  return check_launch("gelu_dropout_bwd");
}
extern "C" int vtx_count_valid(const int64_t* tokens, int B, int T, int pad, float* count, void* stream) {
  REQ(tokens && count, "bad arguments");
  count_valid_kernel<<<32, 256, 0, STREAM>>>((const long long*)tokens, B, T, pad, count);
  return check_launch("count_valid");
}
extern "C" int vtx_cross_entropy(void* logits, int64_t ldl, const int64_t* tokens, int B, int T, int V, int pad,
                                 const float* count, float* loss, int write_grad, void* stream) {
  REQ(logits && tokens && count && loss && V % 8 == 0 && ldl % 8 == 0, "bad arguments");
  ce_kernel<<<B * T, 256, 0, STREAM>>>((__nv_bfloat16*)logits, ldl, (const long long*)tokens, T, V, pad, count, loss,
                                       write_grad);
  return check_launch("cross_entropy");
}
extern "C" int vtx_colsum(const void* X, int64_t ld, int M, int N, float* out, void* stream) {
  REQ(X && out && ld % 8 == 0, "bad arguments");
  const int groups = (N + 7) / 8;
  const int threads = 128;
  const int gy = (groups + threads - 1) / threads;
  int gx = (vtx_num_sms() * 4 + gy - 1) / gy;
  if (gx > M) gx = M;
  if (gx < 1) gx = 1;
  const int rows_per_block = (M + gx - 1) / gx;
  gx = (M + rows_per_block - 1) / rows_per_block;
  colsum_kernel<<<dim3(gx, gy), threads, 0, STREAM>>>((const __nv_bfloat16*)X, ld, M, N, out, rows_per_block);
  return check_launch("colsum");
}
extern "C" int vtx_argmax_rows(const float* X, int64_t ld, int M, int N, int64_t* out, void* stream) {
  REQ(X && out, "bad arguments");
  argmax_rows_kernel<<<M, 256, 0, STREAM>>>(X, ld, N, (long long*)out);
  return check_launch("argmax_rows");
}
