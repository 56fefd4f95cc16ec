// Operand builder of the gradient-penalty attention node (shared-QK L2-distance form, bf16, dim_head 64, key axis padded
// to a multiple of 64) and its first / second derivative - one bandwidth kernel each instead of a chain of row reductions,
// casts, zero fills, three concatenations and (backwards) strided slice copies and gradient accumulations:
//   qa [n, seq, h, 80] = [ q_i , 1 , 1 , 0 x 14 ]
//   ka [n, Lp,  h, 80] = [ k_j , hi_j , lo_j , 0 x 14 ]   k_0 = null key, k_j = q_{j-1}, k_j = 0 for the padding rows;
//                          hi + lo = -|k_j|^2 / 2 split into two bf16 numbers, hi = -1e30 on the padding rows
//   vf [n, Lp,  h, 64] = [ null value ; v ; 0 ]
// so that 2 s (qa ka^T) = -s |q_i - k_j|^2 + s |q_i|^2  (gigagan_pytorch.py:574-586 with the row-constant term dropped).
// First derivative (d(hi + lo)/dk_j = -k_j, taken through the hi column as the composed form does):
//   dq_i = dqa_i[:64] + dka_{i+1}[:64] - dka_{i+1}[64] q_i ,  dv_i = dvf_{i+1} ,  dnull = the j = 0 rows summed over images.
// Second derivative = the adjoint of that map (the same gather / scatter pattern with q replaced by the cotangent).
// 8 lanes per 64-wide row (one 16-byte chunk each); everything is a single pass over its operands.
#include "../../include/gigagan_sm100.h"
#include "gg_common.cuh"

#define ST ((cudaStream_t)stream)
#define AUG_D 64
#define AUG_W 80

__device__ __forceinline__ void unpack8(uint4 v, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int e = 0; e < 4; ++e) { float2 t = __bfloat1622float2(h[e]); f[2 * e] = t.x; f[2 * e + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 o;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
  return o;
}
__device__ __forceinline__ float sum8(float s) {
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  return s;
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// rows = n * Lp * heads, 8 lanes each.  mode 0: forward (sources q, v, null_kv).  mode 1: second derivative (sources wq, wv,
// wnull; the hi column becomes -(w_k . k), the ones columns of qa become 0, and g_q = -ghi * wq is written as well).
__global__ void attn_aug_build_kernel(const bf16* __restrict__ q, const bf16* __restrict__ v, const float* __restrict__ null_kv,
                                      const bf16* __restrict__ wq, const bf16* __restrict__ wv, const float* __restrict__ wnull,
                                      const bf16* __restrict__ dka_in, bf16* __restrict__ qa, bf16* __restrict__ ka,
                                      bf16* __restrict__ vf, bf16* __restrict__ gq, int n, int seq, int Lp, int heads, int mode) {
  const long rows = (long)n * Lp * heads;
  const int sub = threadIdx.x & 7;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  for (long row = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 3; row < rows; row += ((long)gridDim.x * blockDim.x) >> 3) {
    const int h = (int)(row % heads);
    const long bj = row / heads;
    const int j = (int)(bj % Lp);
    const long b = bj / Lp;
    float kf[8], vv[8], src[8];
    float extra_hi = 0.f, extra_lo = 0.f;
    const bool tok = j >= 1 && j <= seq;
    const long trow = tok ? ((b * seq + (j - 1)) * heads + h) : 0;          // token row of q / v / wq / wv
    // the key row (bf16 values as the product sees them)
    if (j == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) kf[e] = bf16_round(null_kv[h * AUG_D + sub * 8 + e]);
    } else if (tok) {
      unpack8(*reinterpret_cast<const uint4*>(q + trow * AUG_D + sub * 8), kf);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) kf[e] = 0.f;
    }
    if (mode == 0) {
      float ss = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) { src[e] = kf[e]; ss = fmaf(kf[e], kf[e], ss); }
      ss = sum8(ss);
      const float t = -0.5f * ss + (j > seq ? -1e30f : 0.f);
      extra_hi = bf16_round(t);
      extra_lo = t - extra_hi;
      if (j == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) vv[e] = null_kv[(heads + h) * AUG_D + sub * 8 + e];
      } else if (tok) {
        unpack8(*reinterpret_cast<const uint4*>(v + trow * AUG_D + sub * 8), vv);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) vv[e] = 0.f;
      }
    } else {
      // cotangent rows: w_k (of dq through the key role), w_v; the hi column carries -(w_k . k)
      if (j == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          src[e] = wnull ? wnull[h * AUG_D + sub * 8 + e] : 0.f;
          vv[e] = wnull ? wnull[(heads + h) * AUG_D + sub * 8 + e] : 0.f;
        }
      } else if (tok) {
        unpack8(*reinterpret_cast<const uint4*>(wq + trow * AUG_D + sub * 8), src);
        unpack8(*reinterpret_cast<const uint4*>(wv + trow * AUG_D + sub * 8), vv);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) { src[e] = 0.f; vv[e] = 0.f; }
      }
      float dot = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) dot = fmaf(src[e], kf[e], dot);
      extra_hi = -sum8(dot);
      extra_lo = 0.f;
    }
    bf16* karow = ka + row * AUG_W;
    *reinterpret_cast<uint4*>(karow + sub * 8) = pack8(src);
    if (sub == 0) {
      float ex[8] = {extra_hi, extra_lo, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      *reinterpret_cast<uint4*>(karow + 64) = pack8(ex);
    } else if (sub == 1) {
      *reinterpret_cast<uint4*>(karow + 72) = zero4;
    }
    *reinterpret_cast<uint4*>(vf + row * AUG_D + sub * 8) = pack8(vv);
    if (tok) {
      bf16* qarow = qa + trow * AUG_W;
      *reinterpret_cast<uint4*>(qarow + sub * 8) = pack8(src);                 // mode 0: q itself (src == kf); mode 1: wq
      if (sub == 0) {
        float one = mode == 0 ? 1.f : 0.f;
        float ex[8] = {one, one, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        *reinterpret_cast<uint4*>(qarow + 64) = pack8(ex);
      } else if (sub == 1) {
        *reinterpret_cast<uint4*>(qarow + 72) = zero4;
      }
      if (mode == 1) {
        const float ghi = __bfloat162float(dka_in[row * AUG_W + 64]);
        float g[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) g[e] = -ghi * src[e];
        *reinterpret_cast<uint4*>(gq + trow * AUG_D + sub * 8) = pack8(g);
      }
    }
  }
}

// first derivative: rows = n * seq * heads
__global__ void attn_aug_bwd_kernel(const bf16* __restrict__ dqa, const bf16* __restrict__ dka, const bf16* __restrict__ dvf,
                                    const bf16* __restrict__ q, bf16* __restrict__ dq, bf16* __restrict__ dv, int n, int seq,
                                    int Lp, int heads) {
  const long rows = (long)n * seq * heads;
  const int sub = threadIdx.x & 7;
  for (long row = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 3; row < rows; row += ((long)gridDim.x * blockDim.x) >> 3) {
    const int h = (int)(row % heads);
    const long bi = row / heads;
    const int i = (int)(bi % seq);
    const long b = bi / seq;
    const long krow = (b * Lp + i + 1) * heads + h;
    float a[8], c[8], qv[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(dqa + row * AUG_W + sub * 8), a);
    unpack8(*reinterpret_cast<const uint4*>(dka + krow * AUG_W + sub * 8), c);
    unpack8(*reinterpret_cast<const uint4*>(q + row * AUG_D + sub * 8), qv);
    const float ghi = __bfloat162float(dka[krow * AUG_W + 64]);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = a[e] + c[e] - ghi * qv[e];
    *reinterpret_cast<uint4*>(dq + row * AUG_D + sub * 8) = pack8(o);
    *reinterpret_cast<uint4*>(dv + row * AUG_D + sub * 8) = *reinterpret_cast<const uint4*>(dvf + krow * AUG_D + sub * 8);
  }
}

// null key / value rows (j = 0) summed over the images: grid = heads, 128 threads (64 key columns, 64 value columns).
// mode 0: dnull_k = sum_b dka[b,0,h,:64] - dka[b,0,h,64] k0 ,  dnull_v = sum_b dvf[b,0,h,:]
// mode 1: g_null_k = - sum_b dka[b,0,h,64] * wnull_k   (second derivative; the value half is zero)
__global__ void attn_aug_null_kernel(const bf16* __restrict__ dka, const bf16* __restrict__ dvf, const float* __restrict__ null_kv,
                                     const float* __restrict__ wnull, float* __restrict__ out, int n, int Lp, int heads, int mode) {
  const int h = blockIdx.x, t = threadIdx.x, c = t & 63;
  float acc = 0.f;
  if (t < 64) {
    const float k0 = mode == 0 ? bf16_round(null_kv[h * AUG_D + c]) : (wnull ? wnull[h * AUG_D + c] : 0.f);
    for (int b = 0; b < n; ++b) {
      const long row = ((long)b * Lp) * heads + h;
      const float ghi = __bfloat162float(dka[row * AUG_W + 64]);
      acc += mode == 0 ? __bfloat162float(dka[row * AUG_W + c]) - ghi * k0 : -ghi * k0;
    }
    out[h * AUG_D + c] = acc;
  } else {
    if (mode == 0)
      for (int b = 0; b < n; ++b) acc += __bfloat162float(dvf[(((long)b * Lp) * heads + h) * AUG_D + c]);
    out[(heads + h) * AUG_D + c] = acc;
  }
}

static int aug_check(int d, int Lp, int seq, const void* a, const void* b) {
  if (d != AUG_D) return gg_fail("attn_augment: dim_head %d != 64", d);
  if (Lp < seq + 1) return gg_fail("attn_augment: Lp %d < seq + 1", Lp);
  if (Lp % 4) return gg_fail("attn_augment: Lp %d must be a multiple of 4 (four 8-lane rows per warp reduce together)", Lp);
  if (((uintptr_t)a | (uintptr_t)b) & 15) return gg_fail("attn_augment: operands must be 16-byte aligned");
  return 0;
}

extern "C" {
int gg_attn_augment_fwd(const void* q, const void* v, const float* null_kv, void* qa, void* ka, void* vf, int n, int seq, int Lp,
                        int heads, int d, gg_stream_t stream) {
  if (aug_check(d, Lp, seq, q, v)) return -1;
  const long rows = (long)n * Lp * heads;
  attn_aug_build_kernel<<<gg_blocks(rows * 8, 256, 148 * 16), 256, 0, ST>>>((const bf16*)q, (const bf16*)v, null_kv, nullptr, nullptr,
                                                                           nullptr, nullptr, (bf16*)qa, (bf16*)ka, (bf16*)vf, nullptr,
                                                                           n, seq, Lp, heads, 0);
  return gg_check_launch("attn_augment_fwd");
}
int gg_attn_augment_bwd(const void* dqa, const void* dka, const void* dvf, const void* q, const float* null_kv, void* dq, void* dv,
                        float* dnull, int n, int seq, int Lp, int heads, int d, gg_stream_t stream) {
  if (aug_check(d, Lp, seq, dqa, dka)) return -1;
  const long rows = (long)n * seq * heads;
  attn_aug_bwd_kernel<<<gg_blocks(rows * 8, 256, 148 * 16), 256, 0, ST>>>((const bf16*)dqa, (const bf16*)dka, (const bf16*)dvf,
                                                                         (const bf16*)q, (bf16*)dq, (bf16*)dv, n, seq, Lp, heads);
  attn_aug_null_kernel<<<heads, 128, 0, ST>>>((const bf16*)dka, (const bf16*)dvf, null_kv, nullptr, dnull, n, Lp, heads, 0);
  return gg_check_launch("attn_augment_bwd");
}
int gg_attn_augment_bwd2(const void* wq, const void* wv, const float* wnull, const void* q, const float* null_kv, const void* dka,
                         void* g_dqa, void* g_dka, void* g_dvf, void* g_q, float* g_null, int n, int seq, int Lp, int heads, int d,
                         gg_stream_t stream) {
  if (aug_check(d, Lp, seq, wq, wv)) return -1;
  const long rows = (long)n * Lp * heads;
  attn_aug_build_kernel<<<gg_blocks(rows * 8, 256, 148 * 16), 256, 0, ST>>>((const bf16*)q, nullptr, null_kv, (const bf16*)wq,
                                                                           (const bf16*)wv, wnull, (const bf16*)dka, (bf16*)g_dqa,
                                                                           (bf16*)g_dka, (bf16*)g_dvf, (bf16*)g_q, n, seq, Lp, heads, 1);
  attn_aug_null_kernel<<<heads, 128, 0, ST>>>((const bf16*)dka, nullptr, null_kv, wnull, g_null, n, Lp, heads, 1);
  return gg_check_launch("attn_augment_bwd2");
}
}
