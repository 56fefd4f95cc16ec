// FFMA (CUDA-core) implicit-GEMM family: NHWC conv fprop / dgrad / wgrad and strided batched GEMM.
// This is the fp32 (1e-5 parity) arithmetic path and the fall-back shape coverage for bf16 layers the
// tcgen05 kernels in conv_tc.cu do not take (channel counts not a multiple of 16, Cout < 16, strides).
// Reference call sites replaced: F.conv2d in gigagan_pytorch.py:402-409 (grouped, per-sample weights),
// nn.Conv2d at :1608-1620,:292,:1454-1470,:1656 and the einsum/bmm at :574-590.
#include "gg_common.cuh"

#define BM 64
#define BN 64
#define BK 16

struct ConvP {
  int N, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad;
  long w_gstride;       // elements between per-sample weight sets (0: shared weights)
  int m_per_group;      // rows (pixels) per weight group
  int groups;
  int act;              // 0 none, 1 leaky-relu(0.2)
  float gain;
  long y_off, y_sn, y_sh, y_sw;   // fprop output addressing (elements); dense when y_sw == Cout etc.
};

// 4x4 register micro-tile FMA over one BK slab held in shared memory.
#define GG_MICRO_KERNEL()                                                     \
  _Pragma("unroll") for (int kk = 0; kk < BK; ++kk) {                         \
    float a[4], b[4];                                                         \
    _Pragma("unroll") for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];  \
    _Pragma("unroll") for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];  \
    _Pragma("unroll") for (int i = 0; i < 4; ++i)                             \
      _Pragma("unroll") for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]); \
  }

// MODE 0: fprop  y[m,co]  = sum_k x_gather[m,k] * w[co][k]          (m = output pixel, k = (ky,kx,ci))
// MODE 1: dgrad  dx[m,ci] = sum_k dy_gather[m,k] * w[co][ky][kx][ci] (m = input pixel,  k = (ky,kx,co))
template <typename T, int MODE>
__global__ void __launch_bounds__(256) conv_gemm_simt(const T* __restrict__ src, const T* __restrict__ w,
                                                      const float* __restrict__ bias, const T* __restrict__ res,
                                                      T* __restrict__ dst, ConvP p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int tiles_per_group = (p.m_per_group + BM - 1) / BM;
  const int g = blockIdx.x / tiles_per_group;
  const int m_local0 = (blockIdx.x % tiles_per_group) * BM;
  const int n0 = blockIdx.y * BN;
  const int Ncols = MODE == 0 ? p.Cout : p.Cin;
  const int Cred = MODE == 0 ? p.Cin : p.Cout;       // channel extent inside one tap of K
  const int K = p.KH * p.KW * Cred;
  const int PH = MODE == 0 ? p.OH : p.H, PW = MODE == 0 ? p.OW : p.W;   // pixel grid the rows live on
  const T* wg = w + (long)g * p.w_gstride;

  // A-load assignment: kk = tid%16, rows (tid/16)+16*j
  const int a_kk = tid & 15;
  int a_n[4], a_y[4], a_x[4];
  bool a_ok[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int ml = m_local0 + (tid >> 4) + 16 * j;
    a_ok[j] = ml < p.m_per_group;
    long m = (long)g * p.m_per_group + ml;
    int n = (int)(m / (PH * PW));
    int r = (int)(m % (PH * PW));
    a_n[j] = n; a_y[j] = r / PW; a_x[j] = r % PW;
  }
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += BK) {
    // ---- A tile
    {
      int k = k0 + a_kk;
      bool kok = k < K;
      int tap = kok ? k / Cred : 0, c = kok ? k % Cred : 0;
      int ky = tap / p.KW, kx = tap % p.KW;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = 0.f;
        if (kok && a_ok[j]) {
          if (MODE == 0) {
            int iy = a_y[j] * p.stride - p.pad + ky, ix = a_x[j] * p.stride - p.pad + kx;
            if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
              v = ldf(src + (((long)a_n[j] * p.H + iy) * p.W + ix) * p.Cin + c);
          } else {
            int ty_ = a_y[j] + p.pad - ky, tx_ = a_x[j] + p.pad - kx;
            if (ty_ >= 0 && tx_ >= 0 && ty_ % p.stride == 0 && tx_ % p.stride == 0) {
              int oy = ty_ / p.stride, ox = tx_ / p.stride;
              if (oy < p.OH && ox < p.OW) v = ldf(src + (((long)a_n[j] * p.OH + oy) * p.OW + ox) * p.Cout + c);
            }
          }
        }
        As[a_kk][(tid >> 4) + 16 * j] = v;
      }
    }
    // ---- B tile
    if (MODE == 0) {
      int k = k0 + (tid & 15);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int nn = (tid >> 4) + 16 * j, co = n0 + nn;
        float v = 0.f;
        if (k < K && co < p.Cout) v = ldf(wg + (long)co * K + k);
        Bs[tid & 15][nn] = v;
      }
    } else {
      int nn = tid & 63, ci = n0 + nn;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int kk = (tid >> 6) + 4 * j, k = k0 + kk;
        float v = 0.f;
        if (k < K && ci < p.Cin) {
          int tap = k / p.Cout, co = k % p.Cout;
          v = ldf(wg + ((long)co * p.KH * p.KW + tap) * p.Cin + ci);
        }
        Bs[kk][nn] = v;
      }
    }
    __syncthreads();
    GG_MICRO_KERNEL();
    __syncthreads();
  }
  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int ml = m_local0 + ty * 4 + i;
    if (ml >= p.m_per_group) continue;
    long m = (long)g * p.m_per_group + ml;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= Ncols) continue;
      float v = acc[i][j];
      long o = m * Ncols + n;
      if (MODE == 0) {
        int img = (int)(m / (p.OH * p.OW)), rr = (int)(m % (p.OH * p.OW));
        o = p.y_off + (long)img * p.y_sn + (long)(rr / p.OW) * p.y_sh + (long)(rr % p.OW) * p.y_sw + n;
        if (bias) v += bias[n];
        if (p.act == 1) v = v > 0.f ? v : 0.2f * v;
        if (res) v += ldf(res + o);
        v *= p.gain;
      }
      stf(dst + o, v);
    }
  }
}

// wgrad: dw[g][co][ky][kx][ci] (fp32, atomically accumulated) = sum_m dy[m,co] * x_gather[m,(ky,kx,ci)]
template <typename T>
__global__ void __launch_bounds__(256) conv_wgrad_simt(const T* __restrict__ x, const T* __restrict__ dy,
                                                       float* __restrict__ dw, ConvP p, int splits) {
  __shared__ float As[BK][BM + 4];   // dy tile   [pixel][co]
  __shared__ float Bs[BK][BN + 4];   // x-gather  [pixel][k]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int K = p.KH * p.KW * p.Cin;
  const int k0 = blockIdx.x * BN, co0 = blockIdx.y * BM;
  const int g = blockIdx.z / splits, sp = blockIdx.z % splits;
  const int chunk = ((p.m_per_group + splits - 1) / splits + BK - 1) / BK * BK;
  const int mb = sp * chunk, me = min(p.m_per_group, mb + chunk);
  const int col = tid & 63;
  const int kcol = k0 + col;
  const bool kok = kcol < K;
  const int tap = kok ? kcol / p.Cin : 0, ci = kok ? kcol % p.Cin : 0;
  const int ky = tap / p.KW, kx = tap % p.KW;
  const bool cok = co0 + col < p.Cout;
  float acc[4][4] = {};
  for (int m0 = mb; m0 < me; m0 += BK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int pp = (tid >> 6) + 4 * j, ml = m0 + pp;
      float a = 0.f, b = 0.f;
      if (ml < me) {
        long m = (long)g * p.m_per_group + ml;
        if (cok) a = ldf(dy + m * p.Cout + co0 + col);
        if (kok) {
          int n = (int)(m / (p.OH * p.OW)), r = (int)(m % (p.OH * p.OW));
          int iy = (r / p.OW) * p.stride - p.pad + ky, ix = (r % p.OW) * p.stride - p.pad + kx;
          if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) b = ldf(x + (((long)n * p.H + iy) * p.W + ix) * p.Cin + ci);
        }
      }
      As[pp][col] = a;
      Bs[pp][col] = b;
    }
    __syncthreads();
    GG_MICRO_KERNEL();
    __syncthreads();
  }
  float* dwg = dw + (long)g * p.w_gstride;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int co = co0 + ty * 4 + i;
    if (co >= p.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = k0 + tx * 4 + j;
      if (k < K) atomicAdd(dwg + (long)co * K + k, acc[i][j]);
    }
  }
}

struct BmmP {
  int M, N, K, b2;                 // batch index z -> (z / b2, z % b2)
  long sA1, sA2, rsA, csA;         // element strides
  long sB1, sB2, rsB, csB;         // B indexed [k][n]
  long sC1, sC2, rsC;
  float alpha;
};

template <typename T>
__global__ void __launch_bounds__(256) bmm_simt(const T* __restrict__ A, const T* __restrict__ B,
                                                const float* __restrict__ bias, T* __restrict__ C, BmmP p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int z1 = blockIdx.z / p.b2, z2 = blockIdx.z % p.b2;
  const T* a = A + z1 * p.sA1 + z2 * p.sA2;
  const T* b = B + z1 * p.sB1 + z2 * p.sB2;
  T* c = C + z1 * p.sC1 + z2 * p.sC2;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  // pick the load mapping whose fastest thread index walks the contiguous axis
  const bool a_kfast = p.csA <= p.rsA, b_nfast = p.csB <= p.rsB;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < p.K; k0 += BK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int kk, mm;
      if (a_kfast) { kk = tid & 15; mm = (tid >> 4) + 16 * j; } else { mm = tid & 63; kk = (tid >> 6) + 4 * j; }
      int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < p.M && k < p.K) ? ldf(a + m * p.rsA + k * p.csA) : 0.f;
      int kb, nn;
      if (b_nfast) { nn = tid & 63; kb = (tid >> 6) + 4 * j; } else { kb = tid & 15; nn = (tid >> 4) + 16 * j; }
      int n = n0 + nn; k = k0 + kb;
      Bs[kb][nn] = (n < p.N && k < p.K) ? ldf(b + k * p.rsB + n * p.csB) : 0.f;
    }
    __syncthreads();
    GG_MICRO_KERNEL();
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j] * p.alpha;
      if (bias) v += bias[n];
      stf(c + m * p.rsC + n, v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host launchers (called from the C-ABI in gg_api.cu)
// ------------------------------------------------------------------------------------------------
static int fill_convp(ConvP& p, int N, int H, int W, int Cin, int OH, int OW, int Cout, int KH, int KW, int stride,
                      int pad, int per_sample_w, int mode_rows_on_input) {
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.OH = OH; p.OW = OW; p.Cout = Cout; p.KH = KH; p.KW = KW;
  p.stride = stride; p.pad = pad; p.act = 0; p.gain = 1.f;
  p.y_off = 0; p.y_sn = (long)OH * OW * Cout; p.y_sh = (long)OW * Cout; p.y_sw = Cout;
  long rows_per_img = mode_rows_on_input ? (long)H * W : (long)OH * OW;
  if (per_sample_w) { p.groups = N; p.m_per_group = (int)rows_per_img; p.w_gstride = (long)Cout * KH * KW * Cin; }
  else {
    if (rows_per_img * N > 2147483647L) return gg_fail("conv too large");
    p.groups = 1; p.m_per_group = (int)(rows_per_img * N); p.w_gstride = 0;
  }
  return 0;
}

int ggi_simt_conv_fprop(const void* x, const void* w, const float* bias, const void* res, void* y, int N, int H, int W,
                       int Cin, int OH, int OW, int Cout, int KH, int KW, int stride, int pad, int per_sample_w, int act,
                       float gain, const long* ystr, int dtype, cudaStream_t st) {
  ConvP p;
  if (fill_convp(p, N, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad, per_sample_w, 0)) return -1;
  p.act = act; p.gain = gain;
  if (ystr) { p.y_off = ystr[0]; p.y_sn = ystr[1]; p.y_sh = ystr[2]; p.y_sw = ystr[3]; }
  dim3 grid(p.groups * gg_cdiv(p.m_per_group, BM), gg_cdiv(Cout, BN));
  GG_DISPATCH(dtype, (conv_gemm_simt<T, 0><<<grid, 256, 0, st>>>((const T*)x, (const T*)w, bias, (const T*)res, (T*)y, p)));
  return gg_check_launch("conv_fprop_simt");
}

int ggi_simt_conv_dgrad(const void* dy, const void* w, void* dx, int N, int H, int W, int Cin, int OH, int OW, int Cout,
                       int KH, int KW, int stride, int pad, int per_sample_w, int dtype, cudaStream_t st) {
  ConvP p;
  if (fill_convp(p, N, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad, per_sample_w, 1)) return -1;
  dim3 grid(p.groups * gg_cdiv(p.m_per_group, BM), gg_cdiv(Cin, BN));
  GG_DISPATCH(dtype, (conv_gemm_simt<T, 1><<<grid, 256, 0, st>>>((const T*)dy, (const T*)w, nullptr, nullptr, (T*)dx, p)));
  return gg_check_launch("conv_dgrad_simt");
}

int ggi_simt_conv_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int OH, int OW, int Cout,
                       int KH, int KW, int stride, int pad, int per_sample_w, int dtype, cudaStream_t st) {
  ConvP p;
  if (fill_convp(p, N, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad, per_sample_w, 0)) return -1;
  long K = (long)KH * KW * Cin;
  cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)p.groups * Cout * K, st);
  int tiles = gg_cdiv(K, BN) * gg_cdiv(Cout, BM) * p.groups;
  int splits = 1;
  while (tiles * splits < 148 * 4 && p.m_per_group / (splits * 2) >= 256) splits *= 2;
  dim3 grid(gg_cdiv(K, BN), gg_cdiv(Cout, BM), p.groups * splits);
  GG_DISPATCH(dtype, (conv_wgrad_simt<T><<<grid, 256, 0, st>>>((const T*)x, (const T*)dy, dw, p, splits)));
  return gg_check_launch("conv_wgrad_simt");
}

// skinny fp32 products (style network, squeeze-excite MLPs, modulation projections: M = batch <= 64): one warp per
// output element, lanes stride over K - these sit on serial dependency chains, so latency matters, not throughput.
__global__ void __launch_bounds__(256) skinny_gemm_f32(const float* __restrict__ A, const float* __restrict__ B,
                                                       const float* __restrict__ bias, float* __restrict__ C, BmmP p) {
  long w = blockIdx.x * (long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (long)p.M * p.N) return;
  int m = (int)(w / p.N), n = (int)(w % p.N), lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int k = lane; k < p.K; k += 32) acc = fmaf(A[m * p.rsA + k * p.csA], B[k * p.rsB + n * p.csB], acc);
  acc = warp_sum(acc);
  if (lane == 0) C[m * p.rsC + n] = acc * p.alpha + (bias ? bias[n] : 0.f);
}

int ggi_simt_bmm(const void* A, const void* B, const float* bias, void* C, int b1, int b2, int M, int N, int K,
                const long* sa, const long* sb, const long* sc, float alpha, int dtype, cudaStream_t st) {
  BmmP p;
  p.M = M; p.N = N; p.K = K; p.b2 = b2;
  p.sA1 = sa[0]; p.sA2 = sa[1]; p.rsA = sa[2]; p.csA = sa[3];
  p.sB1 = sb[0]; p.sB2 = sb[1]; p.rsB = sb[2]; p.csB = sb[3];
  p.sC1 = sc[0]; p.sC2 = sc[1]; p.rsC = sc[2];
  p.alpha = alpha;
  if (dtype == GG_F32 && b1 * b2 == 1 && M <= 64 && (long)M * N <= (1L << 22)) {
    skinny_gemm_f32<<<gg_cdiv((long)M * N, 8), 256, 0, st>>>((const float*)A, (const float*)B, bias, (float*)C, p);
    return gg_check_launch("skinny_gemm_f32");
  }
  if ((long)b1 * b2 > 65535) return gg_fail("bmm batch too large");
  dim3 grid(gg_cdiv(M, BM), gg_cdiv(N, BN), b1 * b2);
  GG_DISPATCH(dtype, (bmm_simt<T><<<grid, 256, 0, st>>>((const T*)A, (const T*)B, bias, (T*)C, p)));
  return gg_check_launch("bmm_simt");
}
