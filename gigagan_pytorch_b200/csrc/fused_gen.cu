// Small fused kernels of the generator / discriminator glue (HBM- or latency-bound; one launch replaces a chain of
// pointwise / reduction launches of the composed operators).  extern "C" surface declared in include/gigagan_sm100.h.
#include "../../include/gigagan_sm100.h"
#include <cstdlib>
#include "gg_common.cuh"

#define ST ((cudaStream_t)stream)

// ------------------------------------------------------------------ shared-bank AdaptiveConv2DMod (4x4 ... 16x16 layers)
// prep.  grid (cdiv(O, SB_OB) + XB), 512 threads.  The first blocks compute the demodulation statistics of SB_OB output
// channels for ALL images: (mod + 1) of an image chunk is staged in shared memory ONCE per block (that 32 KB fill is the
// fixed cost of a block: with one block per channel, or per (image, channel), the kernel took ~105 us whatever its
// size), every bank row [n][I*KK] is read once and reused for all images.  The remaining blocks write xs = x * (mod + 1)
// over a slice of the whole batch.  attn = softmax over the n <= 8 kernel logits.
#define SB_BCH 16
#define SB_OB 2       // measured (7 launches of the README-256 generator): 1 -> 0.276 ms, 2 -> 0.271, 4 -> 0.394
template <typename T, int NK>
__global__ void __launch_bounds__(512)
sbank_prep_kernel(const float* __restrict__ bank, const float* __restrict__ mod, const float* __restrict__ kmod,
                  const T* __restrict__ x, T* __restrict__ xs, float* __restrict__ attn, float* __restrict__ dinv, int B,
                  int n, int O, int I, int KK, int HW, int demod, float eps, long ldm, long ldk, int XB, int use_ssm, int ob) {
  __shared__ float sa[SB_BCH * 8];
  __shared__ float red[16][SB_BCH];
  const int OBL = (O + ob - 1) / ob;
  if ((int)blockIdx.x >= OBL) {
    const long per = (long)HW * I, tot = per * B;
    for (long e = (long)(blockIdx.x - OBL) * blockDim.x + threadIdx.x; e < tot; e += (long)XB * blockDim.x) {
      int i = (int)(e % I);
      int b = (int)(e / per);
      stf(xs + e, ldf(x + e) * (mod[(long)b * ldm + i] + 1.f));
    }
    return;
  }
  extern __shared__ float ssm[];                          // [SB_BCH][I]: (mod + 1) of the current image chunk (use_ssm)
  const int E = I * KK;
  for (int b0 = 0; b0 < B; b0 += SB_BCH) {
    const int nb = min(SB_BCH, B - b0);
    __syncthreads();
    if (use_ssm)
      for (int t = threadIdx.x; t < nb * I; t += blockDim.x) ssm[t] = mod[(long)(b0 + t / I) * ldm + (t % I)] + 1.f;
    if ((int)threadIdx.x < nb) {                          // softmax over the kernel logits of image b0 + t
      const int b = b0 + threadIdx.x;
      float m = -INFINITY, ssum = 0.f;
      for (int j = 0; j < n; ++j) m = fmaxf(m, n == 1 ? 0.f : kmod[(long)b * ldk + j]);
      for (int j = 0; j < n; ++j) ssum += n == 1 ? 1.f : expf(kmod[(long)b * ldk + j] - m);
      for (int j = 0; j < 8; ++j)
        sa[threadIdx.x * 8 + j] = j < n ? (n == 1 ? 1.f : expf(kmod[(long)b * ldk + j] - m) / ssum) : 0.f;
      if (blockIdx.x == 0) for (int j = 0; j < n; ++j) attn[b * n + j] = sa[threadIdx.x * 8 + j];
    }
    __syncthreads();
    for (int oo = 0; oo < ob; ++oo) {
      const int o = blockIdx.x * ob + oo;
      if (o >= O) break;
      if (!demod) { if ((int)threadIdx.x < nb) dinv[(long)(b0 + threadIdx.x) * O + o] = 1.f; continue; }
      float ss[SB_BCH];
#pragma unroll
      for (int t = 0; t < SB_BCH; ++t) ss[t] = 0.f;
      // three elements per thread and iteration: their 3*NK bank loads are issued before any arithmetic (with one element
      // per iteration every warp had a single load in flight and the kernel ran at ~2 GB/s per block)
      for (int e0 = threadIdx.x; e0 < E; e0 += 3 * blockDim.x) {
        float w[3][NK];
        int ii[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int e = e0 + u * blockDim.x;
          const bool ok = e < E;
          ii[u] = ok ? e / KK : 0;
#pragma unroll
          for (int j = 0; j < NK; ++j) w[u][j] = (ok && j < n) ? __ldg(bank + ((long)j * O + o) * E + e) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 3; ++u)
#pragma unroll
          for (int t = 0; t < SB_BCH; ++t)
            if (t < nb) {
              float v = 0.f;
#pragma unroll
              for (int j = 0; j < NK; ++j) v += sa[t * 8 + j] * w[u][j];
              const float um = v * (use_ssm ? ssm[t * I + ii[u]] : mod[(long)(b0 + t) * ldm + ii[u]] + 1.f);
              ss[t] += um * um;
            }
      }
#pragma unroll
      for (int t = 0; t < SB_BCH; ++t) {
        const float r = warp_sum(ss[t]);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][t] = r;
      }
      __syncthreads();
      if ((int)threadIdx.x < nb) {
        float tsum = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tsum += red[w][threadIdx.x];
        dinv[(long)(b0 + threadIdx.x) * O + o] = rsqrtf(fmaxf(tsum, eps));
      }
      __syncthreads();
    }
  }
}

// y[b,p,o] = dinv[b,o] * sum_j attn[b,j] * ycat[b,p,j*O+o]
template <typename T>
__global__ void sbank_combine_fwd_kernel(const T* __restrict__ ycat, const float* __restrict__ attn,
                                         const float* __restrict__ dinv, T* __restrict__ y, int HW, int n, int O, long tot) {
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (long)gridDim.x * blockDim.x) {
    int o = (int)(e % O);
    long bp = e / O;
    int b = (int)(bp / HW);
    float z = 0.f;
    for (int j = 0; j < n; ++j) z += attn[b * n + j] * ldf(ycat + bp * ((long)n * O) + (long)j * O + o);
    stf(y + e, z * dinv[(long)b * O + o]);
  }
}

// grid (cdiv(O,64), B), 256 threads = 64 output channels x 4 pixel lanes (lane l takes pixels l, l+4, ...)
template <typename T>
__global__ void sbank_combine_bwd_kernel(const T* __restrict__ gy, const T* __restrict__ ycat,
                                         const float* __restrict__ attn, const float* __restrict__ dinv,
                                         T* __restrict__ gyn, float* __restrict__ gdinv, float* __restrict__ gattn, int B,
                                         int HW, int n, int O) {
  __shared__ float red[8][8];
  __shared__ float sgd[4][64];
  const int b = blockIdx.y, ol = threadIdx.x & 63, pl = threadIdx.x >> 6, o = blockIdx.x * 64 + ol;
  const bool live = o < O;
  float a[8], ga[8];
  for (int j = 0; j < 8; ++j) { a[j] = j < n ? attn[b * n + j] : 0.f; ga[j] = 0.f; }
  const float d = live ? dinv[(long)b * O + o] : 0.f;
  float gd = 0.f;
  if (live)
    for (int p = pl; p < HW; p += 4) {
      const long bp = (long)b * HW + p;
      const float g = ldf(gy + bp * O + o);
      float z = 0.f;
      for (int j = 0; j < n; ++j) {
        float yj = ldf(ycat + bp * ((long)n * O) + (long)j * O + o);
        z += a[j] * yj;
        ga[j] += g * d * yj;
        stf(gyn + ((long)j * B * HW + bp) * O + o, g * d * a[j]);
      }
      gd += g * z;
    }
  sgd[pl][ol] = gd;
  for (int j = 0; j < n; ++j) {
    float t = warp_sum(ga[j]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][j] = t;
  }
  __syncthreads();
  if (pl == 0 && live) gdinv[(long)b * O + o] = sgd[0][ol] + sgd[1][ol] + sgd[2][ol] + sgd[3][ol];
  if ((int)threadIdx.x < n) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w][threadIdx.x];
    atomicAdd(gattn + b * n + threadIdx.x, t);
  }
}

// gx = gxs * (mod + 1);  dmod[b,i] += sum_p gxs[b,p,i] * x[b,p,i].   grid (cdiv(I,64), B), 64 channels x 4 pixel lanes
template <typename T>
__global__ void sbank_bwd_x_kernel(const T* __restrict__ gxs, const T* __restrict__ x, const float* __restrict__ mod,
                                   T* __restrict__ gx, float* __restrict__ dmod, int HW, int I, long ldm) {
  __shared__ float sacc[4][64];
  const int b = blockIdx.y, il = threadIdx.x & 63, pl = threadIdx.x >> 6, i = blockIdx.x * 64 + il;
  const bool live = i < I;
  const float s = live ? mod[(long)b * ldm + i] + 1.f : 0.f;
  float acc = 0.f;
  if (live)
    for (int p = pl; p < HW; p += 4) {
      const long e = ((long)b * HW + p) * I + i;
      const float g = ldf(gxs + e);
      acc += g * ldf(x + e);
      stf(gx + e, g * s);
    }
  sacc[pl][il] = acc;
  __syncthreads();
  if (pl == 0 && live) dmod[(long)b * I + i] += sacc[0][il] + sacc[1][il] + sacc[2][il] + sacc[3][il];
}

// ------------------------------------------------------------------ aux-decoder patch selection (gigagan_pytorch.py:1300-1312)
// t (B, pd*hh, pd*ww, C) viewed as pd x pd patches; out (B*nsel, hh, ww, C) row (b*nsel + s) = patch sel[b*nsel+s] of image b.
// transposed != 0: the gradient - gt (B, pd*hh, pd*ww, C) gets the rows of g scattered back, zero elsewhere (gt is fully
// written: every pixel belongs to exactly one patch, which is either selected once or not at all).
template <typename T>
__global__ void patch_select_kernel(const T* __restrict__ src, T* __restrict__ dst, const int* __restrict__ sel, int B,
                                    int nsel, int pd, int hh, int ww, int C, int transposed, long tot) {
  const int H = pd * hh, W = pd * ww;
  if (!transposed) {                                   // tot = B*nsel*hh*ww*C
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (long)gridDim.x * blockDim.x) {
      int c = (int)(e % C);
      long r = e / C;
      int xw = (int)(r % ww); r /= ww;
      int yh = (int)(r % hh); r /= hh;
      int s = (int)(r % nsel), b = (int)(r / nsel);
      int pi = sel[b * nsel + s], py = pi / pd, px = pi % pd;
      dst[e] = src[(((long)b * H + py * hh + yh) * W + px * ww + xw) * C + c];
    }
  } else {                                             // tot = B*H*W*C
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (long)gridDim.x * blockDim.x) {
      int c = (int)(e % C);
      long r = e / C;
      int X = (int)(r % W); r /= W;
      int Y = (int)(r % H);
      int b = (int)(r / H);
      int pi = (Y / hh) * pd + X / ww;
      float v = 0.f;
      for (int s = 0; s < nsel; ++s)
        if (sel[b * nsel + s] == pi)
          v = ldf(src + ((((long)b * nsel + s) * hh + Y % hh) * ww + X % ww) * C + c);
      stf(dst + e, v);
    }
  }
}

// ------------------------------------------------------------------ one-output-channel heads (to_logits 1x1 convs / Linear -> 1)
// y[r] = sum_c x[r,c] * w[c] + bias.  The batched-GEMM route pads the single output to a 16-wide tile; its weight
// gradient is then a (C x R) x (R x 16) product whose whole reduction over R = 10^4..10^5 rows lands on ONE CTA.
// Here: one warp per row forwards; the backward writes dx[r,c] = gy[r] * w[c] and accumulates dw[c] = sum_r gy[r] x[r,c]
// and dbias = sum_r gy[r] in the same pass over x (column sums in registers, one atomic per block and column).
template <typename T>
__global__ void row_linear_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                      float* __restrict__ y, long R, int C) {
  constexpr int V = 16 / sizeof(T);
  const int lane = threadIdx.x & 31;
  const long warp0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  const float b = bias ? bias[0] : 0.f;
  for (long r = warp0; r < R; r += nwarps) {
    float acc = 0.f;
    for (int c = lane * V; c < C; c += 32 * V) {
      if (sizeof(T) == 2) {
        uint4 u = *reinterpret_cast<const uint4*>(x + r * C + c);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int k = 0; k < 4; ++k) { float2 f = __bfloat1622float2(h[k]); acc += f.x * w[c + 2 * k] + f.y * w[c + 2 * k + 1]; }
      } else {
        float4 f = *reinterpret_cast<const float4*>(x + r * C + c);
        acc += f.x * w[c] + f.y * w[c + 1] + f.z * w[c + 2] + f.w * w[c + 3];
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) y[r] = acc + b;
  }
}
template <typename T>
__global__ void row_linear_bwd_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ gy,
                                      T* __restrict__ dx, float* __restrict__ dw, float* __restrict__ dbias, long R, int C,
                                      int nvec) {
  constexpr int V = 16 / sizeof(T);
  __shared__ float sm[256 * V];
  const int t = threadIdx.x, cv = t & (nvec - 1), rl = t / nvec, lanes = 256 / nvec;
  float acc[V], wv[V], gsum = 0.f;
#pragma unroll
  for (int k = 0; k < V; ++k) { acc[k] = 0.f; wv[k] = w[cv * V + k]; }
  for (long r = (long)blockIdx.x * lanes + rl; r < R; r += (long)gridDim.x * lanes) {
    const float g = gy[r];
    float xv[V], o[V];
    if (sizeof(T) == 2) {
      uint4 u = *reinterpret_cast<const uint4*>(x + r * C + cv * V);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int k = 0; k < 4; ++k) { float2 f = __bfloat1622float2(h[k]); xv[2 * k] = f.x; xv[2 * k + 1] = f.y; }
    } else {
      float4 f = *reinterpret_cast<const float4*>(x + r * C + cv * V);
      xv[0] = f.x; xv[1] = f.y; xv[2] = f.z; xv[3] = f.w;
    }
#pragma unroll
    for (int k = 0; k < V; ++k) { acc[k] = fmaf(g, xv[k], acc[k]); o[k] = g * wv[k]; }
    if (dx) {
      if (sizeof(T) == 2) {
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(o[2 * k], o[2 * k + 1]);
        *reinterpret_cast<uint4*>(dx + r * C + cv * V) = u;
      } else {
        *reinterpret_cast<float4*>(dx + r * C + cv * V) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
    if (cv == 0) gsum += g;
  }
  if (dw) {
#pragma unroll
    for (int k = 0; k < V; ++k) sm[rl * (nvec * V) + cv * V + k] = acc[k];
    __syncthreads();
    for (int c = t; c < C; c += 256) {
      float sres = 0.f;
      for (int i = 0; i < lanes; ++i) sres += sm[i * C + c];
      atomicAdd(dw + c, sres);
    }
  }
  if (dbias && cv == 0) atomicAdd(dbias, gsum);
}

extern "C" {
int gg_sbank_prep(const float* bank, const float* mod, const float* kmod, const void* x, void* xs, float* attn, float* dinv,
                  int B, int n, int O, int I, int KK, int HW, int demod, float eps, int64_t mod_ld, int64_t kmod_ld,
                  int dtype, gg_stream_t stream) {
  if (n > 8) return gg_fail("num_conv_kernels > 8 unsupported");
  if (n > 1 && !kmod) return gg_fail("gg_sbank_prep: kmod missing");
  int XB = gg_cdiv((long)B * HW * I, 512 * 4);
  if (XB < 1) XB = 1;
  if (XB > 512) XB = 512;
  size_t smem = sizeof(float) * (size_t)SB_BCH * I;
  int use_ssm = smem <= 40 * 1024;
  if (!use_ssm) smem = 0;
  static int ob = 0;                               // output channels per block (GG_SB_OB: tuning sweeps)
  if (!ob) { const char* e = getenv("GG_SB_OB"); ob = e ? atoi(e) : SB_OB; if (ob < 1) ob = 1; }
  const int OBL = gg_cdiv(O, ob);
#define SB_GO(NK) sbank_prep_kernel<T, NK><<<OBL + XB, 512, smem, ST>>>(bank, mod, kmod, (const T*)x, (T*)xs, attn, dinv, B, n, O, I, \
                                                                       KK, HW, demod, eps, (long)mod_ld, (long)kmod_ld, XB, use_ssm, ob)
  GG_DISPATCH(dtype, (n <= 1 ? SB_GO(1) : n <= 2 ? SB_GO(2) : n <= 4 ? SB_GO(4) : SB_GO(8)));
#undef SB_GO
  return gg_check_launch("sbank_prep");
}
int gg_sbank_combine_fwd(const void* ycat, const float* attn, const float* dinv, void* y, int B, int HW, int n, int O,
                         int dtype, gg_stream_t stream) {
  long tot = (long)B * HW * O;
  GG_DISPATCH(dtype, (sbank_combine_fwd_kernel<T><<<gg_blocks(tot, 256), 256, 0, ST>>>((const T*)ycat, attn, dinv, (T*)y, HW, n,
                                                                                       O, tot)));
  return gg_check_launch("sbank_combine_fwd");
}
int gg_sbank_combine_bwd(const void* gy, const void* ycat, const float* attn, const float* dinv, void* gyn, float* gdinv,
                         float* gattn_ws, int B, int HW, int n, int O, int dtype, gg_stream_t stream) {
  if (n > 8) return gg_fail("num_conv_kernels > 8 unsupported");
  cudaMemsetAsync(gattn_ws, 0, sizeof(float) * (size_t)B * n, ST);
  dim3 grid(gg_cdiv(O, 64), B);
  GG_DISPATCH(dtype, (sbank_combine_bwd_kernel<T><<<grid, 256, 0, ST>>>((const T*)gy, (const T*)ycat, attn, dinv, (T*)gyn, gdinv,
                                                                        gattn_ws, B, HW, n, O)));
  return gg_check_launch("sbank_combine_bwd");
}
int gg_sbank_bwd_x(const void* gxs, const void* x, const float* mod, void* gx, float* dmod, int B, int HW, int I,
                   int64_t mod_ld, int dtype, gg_stream_t stream) {
  dim3 grid(gg_cdiv(I, 64), B);
  GG_DISPATCH(dtype, (sbank_bwd_x_kernel<T><<<grid, 256, 0, ST>>>((const T*)gxs, (const T*)x, mod, (T*)gx, dmod, HW, I,
                                                                  (long)mod_ld)));
  return gg_check_launch("sbank_bwd_x");
}
int gg_patch_select(const void* src, void* dst, const int* sel, int B, int nsel, int pd, int hh, int ww, int C,
                    int transposed, int dtype, gg_stream_t stream) {
  long tot = transposed ? (long)B * pd * hh * pd * ww * C : (long)B * nsel * hh * ww * C;
  GG_DISPATCH(dtype, (patch_select_kernel<T><<<gg_blocks(tot, 256), 256, 0, ST>>>((const T*)src, (T*)dst, sel, B, nsel, pd, hh, ww,
                                                                                  C, transposed, tot)));
  return gg_check_launch("patch_select");
}
int gg_row_linear_fwd(const void* x, const float* w, const float* bias, float* y, int64_t R, int C, int dtype,
                      gg_stream_t stream) {
  int V = dtype == GG_F32 ? 4 : 8;
  if (C % V || (((uintptr_t)x) & 15)) return gg_fail("gg_row_linear_fwd: C %% %d != 0 or unaligned", V);
  GG_DISPATCH(dtype, (row_linear_fwd_kernel<T><<<gg_blocks(R * 32, 256, 148 * 16), 256, 0, ST>>>((const T*)x, w, bias, y, R, C)));
  return gg_check_launch("row_linear_fwd");
}
int gg_row_linear_bwd(const void* x, const float* w, const float* gy, void* dx, float* dw, float* dbias, int64_t R, int C,
                      int dtype, gg_stream_t stream) {
  int V = dtype == GG_F32 ? 4 : 8;
  int nvec = C % V == 0 ? C / V : 0;
  if (nvec <= 0 || nvec > 256 || (nvec & (nvec - 1)) || (((uintptr_t)x | (uintptr_t)dx) & 15))
    return gg_fail("gg_row_linear_bwd: C / %d must be a power of two <= 256, 16-byte aligned tensors", V);
  int lanes = 256 / nvec;
  int blocks = gg_blocks((R + lanes - 1) / lanes * 256, 256, 148 * 8);
  GG_DISPATCH(dtype, (row_linear_bwd_kernel<T><<<blocks, 256, 0, ST>>>((const T*)x, w, gy, (T*)dx, dw, dbias, R, C, nvec)));
  return gg_check_launch("row_linear_bwd");
}
}
