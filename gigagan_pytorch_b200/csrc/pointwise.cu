// Bandwidth-bound kernels: pointwise maps (3 derivative levels), broadcast multiplies/adds and their
// partner reductions, row softmax, separable sparse resampling (bilinear x2 + binomial blur, bilinear resize),
// NCHW<->NHWC, noise+activation, the AdaptiveConv2DMod weight builder, AdamW.
// Reference call sites: gigagan_pytorch.py:224-261 (RMSNorm/Blur/Upsample), :297-307 (SqueezeExcite),
// :378-400 (weight modulation/demodulation), :925-940 (Noise), :588 (softmax), optimizer.py:34 (AdamW).
#include "gg_common.cuh"

// 16-byte vector access: 8 bf16 or 4 fp32 per thread
template <typename T> struct VecN { static constexpr int N = 16 / sizeof(T); };
template <typename T> __device__ __forceinline__ void ldv(const T* p, float* out);
template <> __device__ __forceinline__ void ldv<float>(const float* p, float* out) {
  float4 v = *reinterpret_cast<const float4*>(p);
  out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
template <> __device__ __forceinline__ void ldv<bf16>(const bf16* p, float* out) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); out[2 * i] = f.x; out[2 * i + 1] = f.y; }
}
template <typename T> __device__ __forceinline__ void stv(T* p, const float* in);
template <> __device__ __forceinline__ void stv<float>(float* p, const float* in) {
  *reinterpret_cast<float4*>(p) = make_float4(in[0], in[1], in[2], in[3]);
}
template <> __device__ __forceinline__ void stv<bf16>(bf16* p, const float* in) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(in[2 * i], in[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = v;
}
static inline bool al16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

// ------------------------------------------------------------------ unary maps
enum { U_LRELU = 0, U_RELU = 1, U_GELU = 2, U_SILU = 3, U_SIGMOID = 4, U_INVNORM = 5, U_RSQRT_EPS8 = 6 };

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

template <int LEVEL>
__device__ __forceinline__ float unary_eval(int kind, float x) {
  switch (kind) {
    case U_LRELU: return LEVEL == 0 ? (x > 0.f ? x : 0.2f * x) : LEVEL == 1 ? (x > 0.f ? 1.f : 0.2f) : 0.f;
    case U_RELU: return LEVEL == 0 ? fmaxf(x, 0.f) : LEVEL == 1 ? (x > 0.f ? 1.f : 0.f) : 0.f;
    case U_GELU: {
      float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
      if (LEVEL == 0) return x * cdf;
      float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
      return LEVEL == 1 ? cdf + x * pdf : pdf * (2.f - x * x);
    }
    case U_SILU: {
      float s = sigmoidf_(x);
      return LEVEL == 0 ? x * s : LEVEL == 1 ? s * (1.f + x * (1.f - s)) : s * (1.f - s) * (2.f + x * (1.f - 2.f * s));
    }
    case U_SIGMOID: {
      float s = sigmoidf_(x);
      return LEVEL == 0 ? s : LEVEL == 1 ? s * (1.f - s) : s * (1.f - s) * (1.f - 2.f * s);
    }
    case U_INVNORM: {   // 1 / max(sqrt(x), 1e-12)  (F.normalize's denominator), x = sum of squares
      if (x <= 1e-24f) return LEVEL == 0 ? 1e12f : 0.f;
      float r = rsqrtf(x);
      return LEVEL == 0 ? r : LEVEL == 1 ? -0.5f * r * r * r : 0.75f * r * r * r * r * r;
    }
    case U_RSQRT_EPS8: {   // rsqrt(max(x, 1e-8))  (StyleGAN2 demodulation, gigagan_pytorch.py:399)
      if (x <= 1e-8f) return LEVEL == 0 ? 1e4f : 0.f;
      float r = rsqrtf(x);
      return LEVEL == 0 ? r : LEVEL == 1 ? -0.5f * r * r * r : 0.75f * r * r * r * r * r;
    }
  }
  return 0.f;
}

// level 0: out = f(x); level 1: out = a * f'(x); level 2: out = a * b * f''(x)
template <typename T, int LEVEL>
__global__ void unary_kernel(int kind, const T* __restrict__ x, const T* __restrict__ a, const T* __restrict__ b,
                             T* __restrict__ out, long n, int vec) {
  constexpr int V = VecN<T>::N;
  if (vec) {
    long nv = n / V;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nv; i += (long)gridDim.x * blockDim.x) {
      float xv[V], av[V], bv[V], o[V];
      ldv(x + i * V, xv);
      if (LEVEL >= 1) ldv(a + i * V, av);
      if (LEVEL >= 2) ldv(b + i * V, bv);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float v = unary_eval<LEVEL>(kind, xv[j]);
        if (LEVEL >= 1) v *= av[j];
        if (LEVEL >= 2) v *= bv[j];
        o[j] = v;
      }
      stv(out + i * V, o);
    }
    return;
  }
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float v = unary_eval<LEVEL>(kind, ldf(x + i));
    if (LEVEL >= 1) v *= ldf(a + i);
    if (LEVEL >= 2) v *= ldf(b + i);
    stf(out + i, v);
  }
}

int ggi_pw_unary(int kind, int level, const void* x, const void* a, const void* b, void* out, long n, int dtype,
                cudaStream_t st) {
  int esz = dtype == GG_F32 ? 4 : 2;
  int vec = (n % (16 / esz) == 0) && al16(x) && al16(out) && (level < 1 || al16(a)) && (level < 2 || al16(b));
  int blocks = gg_blocks(vec ? n / (16 / esz) : n, 256);
  GG_DISPATCH(dtype, {
    if (level == 0) unary_kernel<T, 0><<<blocks, 256, 0, st>>>(kind, (const T*)x, nullptr, nullptr, (T*)out, n, vec);
    else if (level == 1) unary_kernel<T, 1><<<blocks, 256, 0, st>>>(kind, (const T*)x, (const T*)a, nullptr, (T*)out, n, vec);
    else unary_kernel<T, 2><<<blocks, 256, 0, st>>>(kind, (const T*)x, (const T*)a, (const T*)b, (T*)out, n, vec);
  });
  return gg_check_launch("unary");
}

// ------------------------------------------------------------------ binary maps
template <typename T>
__global__ void mul_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long n, int vec) {
  constexpr int V = VecN<T>::N;
  if (vec) {
    long nv = n / V;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nv; i += (long)gridDim.x * blockDim.x) {
      float av[V], bv[V];
      ldv(a + i * V, av); ldv(b + i * V, bv);
#pragma unroll
      for (int j = 0; j < V; ++j) av[j] *= bv[j];
      stv(out + i * V, av);
    }
    return;
  }
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    stf(out + i, ldf(a + i) * ldf(b + i));
}
template <typename T>
__global__ void axpby_kernel(float alpha, const T* __restrict__ x, float beta, const T* __restrict__ y,
                             T* __restrict__ out, long n, int vec) {
  constexpr int V = VecN<T>::N;
  if (vec) {
    long nv = n / V;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nv; i += (long)gridDim.x * blockDim.x) {
      float xv[V], yv[V];
      ldv(x + i * V, xv);
      if (y) ldv(y + i * V, yv);
#pragma unroll
      for (int j = 0; j < V; ++j) xv[j] = alpha * xv[j] + (y ? beta * yv[j] : 0.f);
      stv(out + i * V, xv);
    }
    return;
  }
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float v = alpha * ldf(x + i);
    if (y) v += beta * ldf(y + i);
    stf(out + i, v);
  }
}
int ggi_pw_mul(const void* a, const void* b, void* out, long n, int dtype, cudaStream_t st) {
  int V = dtype == GG_F32 ? 4 : 8;
  int vec = (n % V == 0) && al16(a) && al16(b) && al16(out);
  GG_DISPATCH(dtype, (mul_kernel<T><<<gg_blocks(vec ? n / V : n, 256), 256, 0, st>>>((const T*)a, (const T*)b, (T*)out, n, vec)));
  return gg_check_launch("mul");
}
int ggi_pw_axpby(float alpha, const void* x, float beta, const void* y, void* out, long n, int dtype, cudaStream_t st) {
  int V = dtype == GG_F32 ? 4 : 8;
  int vec = (n % V == 0) && al16(x) && al16(out) && (!y || al16(y));
  GG_DISPATCH(dtype, (axpby_kernel<T><<<gg_blocks(vec ? n / V : n, 256), 256, 0, st>>>(alpha, (const T*)x, beta, (const T*)y, (T*)out, n, vec)));
  return gg_check_launch("axpby");
}

// ------------------------------------------------------------------ broadcasts over a [R, C] view
// mode 0 (ROWS): s[r].  mode 1 (SAMPLE_CH): s[((r / P) % Ns) * C + c]  (P rows per sample, scale-major repeat)
// op 0: out = x * s ; op 1: out = x + s.   s is always fp32.
template <typename T>
__global__ void bcast_kernel(const T* __restrict__ x, const float* __restrict__ s, T* __restrict__ out, long R, int C,
                             int P, int Ns, int mode, int op, int vec) {
  constexpr int V = VecN<T>::N;
  long n = R * C;
  if (vec) {
    long nv = n / V;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nv; i += (long)gridDim.x * blockDim.x) {
      long e = i * V, r = e / C;
      int c = (int)(e - r * C);
      float xv[V];
      ldv(x + e, xv);
      if (mode == 0) {
        float sv = s[r];
#pragma unroll
        for (int j = 0; j < V; ++j) xv[j] = op == 0 ? xv[j] * sv : xv[j] + sv;
      } else {
        const float* sp = s + ((r / P) % Ns) * (long)C + c;
#pragma unroll
        for (int j = 0; j < V; ++j) xv[j] = op == 0 ? xv[j] * sp[j] : xv[j] + sp[j];
      }
      stv(out + e, xv);
    }
    return;
  }
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    long r = i / C;
    int c = (int)(i % C);
    float sv = mode == 0 ? s[r] : s[((r / P) % Ns) * (long)C + c];
    float xv = ldf(x + i);
    stf(out + i, op == 0 ? xv * sv : xv + sv);
  }
}
int ggi_pw_bcast(const void* x, const float* s, void* out, long R, int C, int P, int Ns, int mode, int op, int dtype,
                cudaStream_t st) {
  int V = dtype == GG_F32 ? 4 : 8;
  int vec = (C % V == 0) && al16(x) && al16(out);
  GG_DISPATCH(dtype, (bcast_kernel<T><<<gg_blocks(vec ? R * C / V : R * C, 256), 256, 0, st>>>((const T*)x, s, (T*)out, R, C, P, Ns, mode, op, vec)));
  return gg_check_launch("bcast");
}

// out[r] = sum_c a[r,c] * b[r,c]  (b may be null -> row sums); one warp per row
template <typename T>
__global__ void rowdot_kernel(const T* __restrict__ a, const T* __restrict__ b, float* __restrict__ out, long R, int C) {
  long r = blockIdx.x * (long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  int lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int c = lane; c < C; c += 32) acc += ldf(a + r * C + c) * (b ? ldf(b + r * C + c) : 1.f);
  acc = warp_sum(acc);
  if (lane == 0) out[r] = acc;
}
int ggi_red_rowdot(const void* a, const void* b, float* out, long R, int C, int dtype, cudaStream_t st) {
  GG_DISPATCH(dtype, (rowdot_kernel<T><<<gg_cdiv(R, 8), 256, 0, st>>>((const T*)a, (const T*)b, out, R, C)));
  return gg_check_launch("rowdot");
}

// out[n', c] = sum_{n == n' (mod Ns)} sum_{p < P} a[(n,p), c] * b[(n,p), c]   (fp32 out, zeroed here)
template <typename T>
__global__ void dot_sc_kernel(const T* __restrict__ a, const T* __restrict__ b, float* __restrict__ out, int C, int P,
                              int Ns, long rows_per_out, int splits) {
  __shared__ float sm[8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  int ns = blockIdx.y;
  long chunk = (rows_per_out + splits - 1) / splits;
  long j0 = blockIdx.z * chunk, j1 = min(rows_per_out, j0 + chunk);
  float acc = 0.f;
  if (c < C)
    for (long j = j0 + threadIdx.y; j < j1; j += 8) {
      long rep = j / P, pp = j % P;
      long r = (rep * Ns + ns) * P + pp;
      acc += ldf(a + r * C + c) * (b ? ldf(b + r * C + c) : 1.f);
    }
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
    atomicAdd(out + (long)ns * C + c, t);
  }
}
// vector variant: each thread owns V consecutive channels (16-byte loads); block = 32 channel-vectors x 8 row lanes
template <typename T>
__global__ void dot_sc_vec_kernel(const T* __restrict__ a, const T* __restrict__ b, float* __restrict__ out, int C, int P,
                                  int Ns, long rows_per_out, int splits) {
  constexpr int V = VecN<T>::N;
  __shared__ float sm[8][32 * V + 1];
  int cv = blockIdx.x * 32 + threadIdx.x;          // channel-vector index
  int c = cv * V;
  int ns = blockIdx.y;
  long chunk = (rows_per_out + splits - 1) / splits;
  long j0 = blockIdx.z * chunk, j1 = min(rows_per_out, j0 + chunk);
  float acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = 0.f;
  if (c < C)
    for (long j = j0 + threadIdx.y; j < j1; j += 8) {
      long rep = j / P, pp = j - rep * P;
      long r = (rep * Ns + ns) * P + pp;
      float av[V], bv[V];
      ldv(a + r * C + c, av);
      if (b) {
        ldv(b + r * C + c, bv);
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] = fmaf(av[k], bv[k], acc[k]);
      } else {
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] += av[k];
      }
    }
#pragma unroll
  for (int k = 0; k < V; ++k) sm[threadIdx.y][threadIdx.x * V + k] = acc[k];
  __syncthreads();
  int t = threadIdx.y * 32 + threadIdx.x;            // 256 threads reduce 32*V columns
  for (int col = t; col < 32 * V; col += 256) {
    int cc = blockIdx.x * 32 * V + col;
    if (cc < C) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += sm[i][col];
      atomicAdd(out + (long)ns * C + cc, s);
    }
  }
}
// narrow maps (C/V = 1..32 channel vectors, a power of two): the 256 threads tile (256/nvec row lanes) x nvec vectors, so
// consecutive threads read consecutive 16-byte pieces of the row-major stream (bias gradients of the 16-64 channel
// 128^2 / 256^2 layers left 28 of 32 lanes idle in the kernel above)
template <typename T>
__global__ void dot_sc_flat_kernel(const T* __restrict__ a, const T* __restrict__ b, float* __restrict__ out, int C, int P,
                                   int Ns, long rows_per_out, int splits, int nvec) {
  constexpr int V = VecN<T>::N;
  __shared__ float sm[256 * V];
  const int t = threadIdx.x, cv = t & (nvec - 1), rl = t / nvec, lanes = 256 / nvec;
  const int ns = blockIdx.y;
  long chunk = (rows_per_out + splits - 1) / splits;
  long j0 = blockIdx.x * chunk, j1 = min(rows_per_out, j0 + chunk);
  float acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = 0.f;
  for (long j = j0 + rl; j < j1; j += lanes) {
    long rep = j / P, pp = j - rep * P;
    long r = (rep * Ns + ns) * P + pp;
    float av[V], bv[V];
    ldv(a + r * C + cv * V, av);
    if (b) {
      ldv(b + r * C + cv * V, bv);
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] = fmaf(av[k], bv[k], acc[k]);
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] += av[k];
    }
  }
#pragma unroll
  for (int k = 0; k < V; ++k) sm[rl * C + cv * V + k] = acc[k];
  __syncthreads();
  if (t < C) {
    float sres = 0.f;
    for (int i = 0; i < lanes; ++i) sres += sm[i * C + t];
    atomicAdd(out + (long)ns * C + t, sres);
  }
}
int ggi_red_dot_sc_acc(const void* a, const void* b, float* out, long R, int C, int P, int Ns, int accumulate, int dtype,
                       cudaStream_t st);
int ggi_red_dot_sc(const void* a, const void* b, float* out, long R, int C, int P, int Ns, int dtype, cudaStream_t st) {
  return ggi_red_dot_sc_acc(a, b, out, R, C, P, Ns, 0, dtype, st);
}
// accumulate != 0: out += the sums (out holds a running gradient); 0: out is overwritten
int ggi_red_dot_sc_acc(const void* a, const void* b, float* out, long R, int C, int P, int Ns, int accumulate, int dtype,
                       cudaStream_t st) {
  long N = R / P;
  long rows_per_out = (N / Ns) * P;
  if (!accumulate) cudaMemsetAsync(out, 0, sizeof(float) * (size_t)Ns * C, st);
  int V = dtype == GG_F32 ? 4 : 8;
  dim3 block(32, 8);
  if (C % V == 0 && al16(a) && (!b || al16(b)) && C / V <= 32 && ((C / V) & (C / V - 1)) == 0 && C <= 256) {
    int nvec = C / V, splits = 1;
    while (Ns * splits < 148 * 8 && rows_per_out / (splits * 2) >= 4 * (256 / nvec)) splits *= 2;
    dim3 grid(splits, Ns);
    GG_DISPATCH(dtype, (dot_sc_flat_kernel<T><<<grid, 256, 0, st>>>((const T*)a, (const T*)b, out, C, P, Ns, rows_per_out, splits, nvec)));
    return gg_check_launch("dot_sc_flat");
  }
  if (C % V == 0 && al16(a) && (!b || al16(b))) {
    int gx = gg_cdiv(C, 32 * V);
    int base = gx * Ns, splits = 1;
    while (base * splits < 148 * 8 && rows_per_out / (splits * 2) >= 32) splits *= 2;
    dim3 grid(gx, Ns, splits);
    GG_DISPATCH(dtype, (dot_sc_vec_kernel<T><<<grid, block, 0, st>>>((const T*)a, (const T*)b, out, C, P, Ns, rows_per_out, splits)));
    return gg_check_launch("dot_sc_vec");
  }
  int base = gg_cdiv(C, 32) * Ns, splits = 1;
  while (base * splits < 148 * 4 && rows_per_out / (splits * 2) >= 64) splits *= 2;
  dim3 grid(gg_cdiv(C, 32), Ns, splits);
  GG_DISPATCH(dtype, (dot_sc_kernel<T><<<grid, block, 0, st>>>((const T*)a, (const T*)b, out, C, P, Ns, rows_per_out, splits)));
  return gg_check_launch("dot_sc");
}

// ------------------------------------------------------------------ row softmax of (S + bias)
// bias (optional) is fp32 [Ns][C]; row r uses bias row (r / P) % Ns.  Fast path: one warp per row, the row lives in
// registers (16-byte loads); generic path: one CTA per row.
template <typename T, int MAXV>
__global__ void softmax_rows_warp_kernel(const T* __restrict__ s, const float* __restrict__ bias, T* __restrict__ p,
                                         long R, int C, int P, int Ns) {
  constexpr int V = VecN<T>::N;
  long r = blockIdx.x * (long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  int lane = threadIdx.x & 31;
  const T* sr = s + r * C;
  const float* br = bias ? bias + ((r / P) % Ns) * (long)C : nullptr;
  int nvec = C / V;
  float v[MAXV][V];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int vi = lane + 32 * i;
    if (vi < nvec) {
      ldv(sr + vi * V, v[i]);
#pragma unroll
      for (int j = 0; j < V; ++j) { if (br) v[i][j] += br[vi * V + j]; m = fmaxf(m, v[i][j]); }
    }
  }
  m = warp_max(m);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int vi = lane + 32 * i;
    if (vi < nvec) {
#pragma unroll
      for (int j = 0; j < V; ++j) { v[i][j] = __expf(v[i][j] - m); sum += v[i][j]; }
    }
  }
  sum = warp_sum(sum);
  float inv = 1.f / sum;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int vi = lane + 32 * i;
    if (vi < nvec) {
#pragma unroll
      for (int j = 0; j < V; ++j) v[i][j] *= inv;
      stv(p + r * C + vi * V, v[i]);
    }
  }
}

template <typename T>
__global__ void softmax_rows_kernel(const T* __restrict__ s, const float* __restrict__ bias, T* __restrict__ p, int C,
                                    int P, int Ns) {
  __shared__ float red[32];
  long r = blockIdx.x;
  const T* sr = s + r * C;
  const float* br = bias ? bias + ((r / P) % Ns) * (long)C : nullptr;
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) m = fmaxf(m, ldf(sr + c) + (br ? br[c] : 0.f));
  m = warp_max(m);
  if (lane == 0) red[wid] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < nw; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) sum += __expf(ldf(sr + c) + (br ? br[c] : 0.f) - m);
  sum = warp_sum(sum);
  if (lane == 0) red[wid] = sum;
  __syncthreads();
  sum = 0.f;
  for (int i = 0; i < nw; ++i) sum += red[i];
  float inv = 1.f / sum;
  for (int c = threadIdx.x; c < C; c += blockDim.x) stf(p + r * C + c, __expf(ldf(sr + c) + (br ? br[c] : 0.f) - m) * inv);
}
int ggi_softmax_rows(const void* s, const float* bias, void* p, long R, int C, int P, int Ns, int dtype, cudaStream_t st) {
  int V = dtype == GG_F32 ? 4 : 8;
  if (!bias) { P = 1; Ns = 1; }
  if (C % V == 0 && C / V <= 32 * 6 && al16(s) && al16(p)) {
    GG_DISPATCH(dtype, (softmax_rows_warp_kernel<T, 6><<<gg_cdiv(R, 8), 256, 0, st>>>((const T*)s, bias, (T*)p, R, C, P, Ns)));
  } else {
    GG_DISPATCH(dtype, (softmax_rows_kernel<T><<<(unsigned)R, 128, 0, st>>>((const T*)s, bias, (T*)p, C, P, Ns)));
  }
  return gg_check_launch("softmax_rows");
}

// dS = P * (gP - sum_j P_j gP_j), one warp per row, rows in registers
// gp2 (nullable): a second gradient of the same shape added to gp on the fly (the gradient-penalty pass hands the
// probabilities' own second-order gradient in this way instead of materialising the sum)
template <typename T, int MAXV>
__global__ void softmax_bwd_rows_warp_kernel(const T* __restrict__ p, const T* __restrict__ gp, const T* __restrict__ gp2,
                                             T* __restrict__ ds, long R, int C) {
  constexpr int V = VecN<T>::N;
  long r = blockIdx.x * (long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  int lane = threadIdx.x & 31, nvec = C / V;
  float pv[MAXV][V], gv[MAXV][V];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int vi = lane + 32 * i;
    if (vi < nvec) {
      ldv(p + r * C + vi * V, pv[i]);
      ldv(gp + r * C + vi * V, gv[i]);
      if (gp2) {
        float g2[V];
        ldv(gp2 + r * C + vi * V, g2);
#pragma unroll
        for (int j = 0; j < V; ++j) gv[i][j] += g2[j];
      }
#pragma unroll
      for (int j = 0; j < V; ++j) dot = fmaf(pv[i][j], gv[i][j], dot);
    }
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int vi = lane + 32 * i;
    if (vi < nvec) {
#pragma unroll
      for (int j = 0; j < V; ++j) pv[i][j] *= (gv[i][j] - dot);
      stv(ds + r * C + vi * V, pv[i]);
    }
  }
}
template <typename T>
__global__ void softmax_bwd_rows_kernel(const T* __restrict__ p, const T* __restrict__ gp, const T* __restrict__ gp2,
                                        T* __restrict__ ds, int C) {
  __shared__ float red[32];
  long r = blockIdx.x;
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float dot = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    dot += ldf(p + r * C + c) * (ldf(gp + r * C + c) + (gp2 ? ldf(gp2 + r * C + c) : 0.f));
  dot = warp_sum(dot);
  if (lane == 0) red[wid] = dot;
  __syncthreads();
  dot = 0.f;
  for (int i = 0; i < nw; ++i) dot += red[i];
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    stf(ds + r * C + c, ldf(p + r * C + c) * (ldf(gp + r * C + c) + (gp2 ? ldf(gp2 + r * C + c) : 0.f) - dot));
}
int ggi_softmax_bwd_rows(const void* p, const void* gp, const void* gp2, void* ds, long R, int C, int dtype, cudaStream_t st) {
  int V = dtype == GG_F32 ? 4 : 8;
  if (C % V == 0 && C / V <= 32 * 6 && al16(p) && al16(gp) && al16(ds) && (!gp2 || al16(gp2))) {
    GG_DISPATCH(dtype, (softmax_bwd_rows_warp_kernel<T, 6><<<gg_cdiv(R, 8), 256, 0, st>>>((const T*)p, (const T*)gp, (const T*)gp2, (T*)ds, R, C)));
  } else {
    GG_DISPATCH(dtype, (softmax_bwd_rows_kernel<T><<<(unsigned)R, 128, 0, st>>>((const T*)p, (const T*)gp, (const T*)gp2, (T*)ds, C)));
  }
  return gg_check_launch("softmax_bwd_rows");
}

// second-order softmax backward in one pass (gradient-penalty path):  given P, gP (inputs of dS = P*(gP - r)) and the
// upstream G = d/d(dS):   d_gP = P * (G - <G,P>)      d_P = G * (gP - <P,gP>) - gP * <G,P>
template <typename T, int MAXV>
__global__ void softmax_bwd2_rows_warp_kernel(const T* __restrict__ p, const T* __restrict__ gp, const T* __restrict__ G,
                                              T* __restrict__ d_p, T* __restrict__ d_gp, long R, int C) {
  constexpr int V = VecN<T>::N;
  long r = blockIdx.x * (long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  int lane = threadIdx.x & 31, nvec = C / V;
  float pv[MAXV][V], gv[MAXV][V], Gv[MAXV][V];
  float r_pg = 0.f, r_Gp = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int vi = lane + 32 * i;
    if (vi < nvec) {
      ldv(p + r * C + vi * V, pv[i]); ldv(gp + r * C + vi * V, gv[i]); ldv(G + r * C + vi * V, Gv[i]);
#pragma unroll
      for (int j = 0; j < V; ++j) { r_pg = fmaf(pv[i][j], gv[i][j], r_pg); r_Gp = fmaf(Gv[i][j], pv[i][j], r_Gp); }
    }
  }
  r_pg = warp_sum(r_pg); r_Gp = warp_sum(r_Gp);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int vi = lane + 32 * i;
    if (vi < nvec) {
      float a[V], b[V];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        a[j] = Gv[i][j] * (gv[i][j] - r_pg) - gv[i][j] * r_Gp;
        b[j] = pv[i][j] * (Gv[i][j] - r_Gp);
      }
      stv(d_p + r * C + vi * V, a);
      stv(d_gp + r * C + vi * V, b);
    }
  }
}
int ggi_softmax_bwd2_rows(const void* p, const void* gp, const void* G, void* d_p, void* d_gp, long R, int C, int dtype,
                          cudaStream_t st) {
  int V = dtype == GG_F32 ? 4 : 8;
  if (!(C % V == 0 && C / V <= 32 * 5 && al16(p) && al16(gp) && al16(G) && al16(d_p) && al16(d_gp))) return 1;
  GG_DISPATCH(dtype, (softmax_bwd2_rows_warp_kernel<T, 5><<<gg_cdiv(R, 8), 256, 0, st>>>((const T*)p, (const T*)gp, (const T*)G, (T*)d_p, (T*)d_gp, R, C)));
  return gg_check_launch("softmax_bwd2_rows");
}

// ------------------------------------------------------------------ separable sparse resampling (NHWC)
// y[n,oy,ox,c] = sum_{a<Ty} sum_{b<Tx} wy[oy,a] wx[ox,b] x[n, iy[oy,a], ix[ox,b], c]
template <typename T>
__global__ void resample_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int OH, int OW,
                                const int* __restrict__ iy, const float* __restrict__ wy, int Ty,
                                const int* __restrict__ ix, const float* __restrict__ wx, int Tx) {
  long n = (long)N * OH * OW * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long t = i / C;
    int ox = (int)(t % OW); t /= OW;
    int oy = (int)(t % OH);
    int b = (int)(t / OH);
    float acc = 0.f;
    for (int a = 0; a < Ty; ++a) {
      float wa = wy[oy * Ty + a];
      if (wa == 0.f) continue;
      const T* row = x + ((long)b * H + iy[oy * Ty + a]) * W * C + c;
      float racc = 0.f;
      for (int q = 0; q < Tx; ++q) {
        float wq = wx[ox * Tx + q];
        if (wq != 0.f) racc += wq * ldf(row + (long)ix[ox * Tx + q] * C);
      }
      acc += wa * racc;
    }
    stf(y + i, acc);
  }
}
// vector variant: one thread = 16 bytes of channels of one output pixel
template <typename T>
__global__ void resample_vec_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int OH, int OW,
                                    const int* __restrict__ iy, const float* __restrict__ wy, int Ty,
                                    const int* __restrict__ ix, const float* __restrict__ wx, int Tx) {
  constexpr int V = VecN<T>::N;
  const int CV = C / V;
  long n = (long)N * OH * OW * CV;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int cv = (int)(i % CV);
    long t = i / CV;
    int ox = (int)(t % OW); t /= OW;
    int oy = (int)(t % OH);
    int b = (int)(t / OH);
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
    for (int a = 0; a < Ty; ++a) {
      float wa = wy[oy * Ty + a];
      if (wa == 0.f) continue;
      const T* row = x + ((long)b * H + iy[oy * Ty + a]) * W * C + cv * V;
      for (int q = 0; q < Tx; ++q) {
        float wq = wx[ox * Tx + q] * wa;
        if (wq == 0.f) continue;
        float v[V];
        ldv(row + (long)ix[ox * Tx + q] * C, v);
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] = fmaf(wq, v[k], acc[k]);
      }
    }
    stv(y + i * V, acc);
  }
}
// same, taps held in registers (MT >= Ty, Tx) and the MT loads of one input row issued back to back: the loop above
// chained table load -> address -> data load for every tap (238 GB/s on the 16-channel 256x256 maps)
template <typename T, int MT>
__global__ void resample_vec2_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int OH, int OW,
                                     const int* __restrict__ iy, const float* __restrict__ wy, int Ty,
                                     const int* __restrict__ ix, const float* __restrict__ wx, int Tx) {
  constexpr int V = VecN<T>::N;
  const int CV = C / V;
  const long n = (long)N * OH * OW * CV;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    long t = i / CV;
    const int ox = (int)(t % OW); t /= OW;
    const int oy = (int)(t % OH);
    const int b = (int)(t / OH);
    int xo[MT]; float xw[MT];
#pragma unroll
    for (int q = 0; q < MT; ++q) {
      const bool on = q < Tx;
      xw[q] = on ? wx[ox * Tx + q] : 0.f;
      xo[q] = on ? ix[ox * Tx + q] * C : 0;
    }
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
    const T* img = x + (long)b * H * W * C + cv * V;
#pragma unroll
    for (int a = 0; a < MT; ++a) {
      if (a < Ty) {
        const float wa = wy[oy * Ty + a];
        if (wa != 0.f) {
          const T* row = img + (long)iy[oy * Ty + a] * W * C;
          float v[MT][V];
#pragma unroll
          for (int q = 0; q < MT; ++q) ldv(row + xo[q], v[q]);            // zero-weight taps read pixel 0 (valid memory)
#pragma unroll
          for (int q = 0; q < MT; ++q) {
            const float wq = xw[q] * wa;
#pragma unroll
            for (int k = 0; k < V; ++k) acc[k] = fmaf(wq, v[q][k], acc[k]);
          }
        }
      }
    }
    stv(y + i * V, acc);
  }
}
int ggi_resample2d(const void* x, void* y, int N, int H, int W, int C, int OH, int OW, const int* iy, const float* wy,
                  int Ty, const int* ix, const float* wx, int Tx, int dtype, cudaStream_t st) {
  long n = (long)N * OH * OW * C;
  int V = dtype == GG_F32 ? 4 : 8;
  if (C % V == 0 && al16(x) && al16(y)) {
    int blocks = gg_blocks(n / V, 256, 148 * 32);
    if (Ty <= 4 && Tx <= 4) {
      GG_DISPATCH(dtype, (resample_vec2_kernel<T, 4><<<blocks, 256, 0, st>>>((const T*)x, (T*)y, N, H, W, C, OH, OW, iy, wy, Ty, ix, wx, Tx)));
    } else if (Ty <= 8 && Tx <= 8) {
      GG_DISPATCH(dtype, (resample_vec2_kernel<T, 8><<<blocks, 256, 0, st>>>((const T*)x, (T*)y, N, H, W, C, OH, OW, iy, wy, Ty, ix, wx, Tx)));
    } else {
      GG_DISPATCH(dtype, (resample_vec_kernel<T><<<gg_blocks(n / V, 256), 256, 0, st>>>((const T*)x, (T*)y, N, H, W, C, OH, OW, iy, wy, Ty, ix, wx, Tx)));
    }
    return gg_check_launch("resample2d_vec");
  }
  GG_DISPATCH(dtype, (resample_kernel<T><<<gg_blocks(n, 256), 256, 0, st>>>((const T*)x, (T*)y, N, H, W, C, OH, OW, iy, wy, Ty, ix, wx, Tx)));
  return gg_check_launch("resample2d");
}

// ------------------------------------------------------------------ layout: NCHW fp32 <-> NHWC T (channel pad)
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int N, int C, int HW, int Cp) {
  long n = (long)N * HW * Cp;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % Cp);
    long t = i / Cp;
    int p = (int)(t % HW);
    int b = (int)(t / HW);
    stf(dst + i, c < C ? src[((long)b * C + c) * HW + p] : 0.f);
  }
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, float* __restrict__ dst, int N, int C, int HW, int Cp) {
  long n = (long)N * C * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int p = (int)(i % HW);
    long t = i / HW;
    int c = (int)(t % C);
    int b = (int)(t / C);
    dst[i] = ldf(src + ((long)b * HW + p) * Cp + c);
  }
}
int ggi_nchw_to_nhwc(const float* src, void* dst, int N, int C, int HW, int Cp, int dtype, cudaStream_t st) {
  GG_DISPATCH(dtype, (nchw_to_nhwc_kernel<T><<<gg_blocks((long)N * HW * Cp, 256), 256, 0, st>>>(src, (T*)dst, N, C, HW, Cp)));
  return gg_check_launch("nchw_to_nhwc");
}
int ggi_nhwc_to_nchw(const void* src, float* dst, int N, int C, int HW, int Cp, int dtype, cudaStream_t st) {
  GG_DISPATCH(dtype, (nhwc_to_nchw_kernel<T><<<gg_blocks((long)N * C * HW, 256), 256, 0, st>>>((const T*)src, dst, N, C, HW, Cp)));
  return gg_check_launch("nhwc_to_nchw");
}

// ------------------------------------------------------------------ Noise + LeakyReLU (generator)
// y = lrelu(x + wn[c] * noise[n,pixel])   x,y: [R=N*HW, C]; noise fp32 [R]; wn fp32 [C]
template <typename T>
__global__ void noise_act_fwd_kernel(const T* __restrict__ x, const float* __restrict__ noise,
                                     const float* __restrict__ wn, T* __restrict__ y, long R, int C) {
  long n = R * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float v = ldf(x + i) + wn[i % C] * noise[i / C];
    stf(y + i, v > 0.f ? v : 0.2f * v);
  }
}
// dx = gy * lrelu'(y); dwn[c] += sum_r dx[r,c] * noise[r]
template <typename T>
__global__ void noise_act_bwd_kernel(const T* __restrict__ y, const T* __restrict__ gy,
                                     const float* __restrict__ noise, T* __restrict__ dx, float* __restrict__ dwn,
                                     long R, int C, int splits) {
  __shared__ float sm[8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  long chunk = (R + splits - 1) / splits;
  long r0 = blockIdx.y * chunk, r1 = min(R, r0 + chunk);
  float acc = 0.f;
  if (c < C)
    for (long r = r0 + threadIdx.y; r < r1; r += 8) {
      float g = ldf(gy + r * C + c) * (ldf(y + r * C + c) > 0.f ? 1.f : 0.2f);
      stf(dx + r * C + c, g);
      acc += g * noise[r];
    }
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
    atomicAdd(dwn + c, t);
  }
}
// 16-byte vector forms (C % V == 0): thread = one channel vector of one pixel row
template <typename T>
__global__ void noise_act_fwd_vec_kernel(const T* __restrict__ x, const float* __restrict__ noise,
                                         const float* __restrict__ wn, T* __restrict__ y, long R, int C) {
  constexpr int V = VecN<T>::N;
  const int CV = C / V;
  const long n = R * CV;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    const float nz = noise[i / CV];
    float v[V], o[V];
    ldv(x + i * V, v);
#pragma unroll
    for (int k = 0; k < V; ++k) { float t = v[k] + __ldg(wn + cv * V + k) * nz; o[k] = t > 0.f ? t : 0.2f * t; }
    stv(y + i * V, o);
  }
}
// nvec = C / V channel vectors (a power of two <= 256): 256 threads = (256 / nvec row lanes) x nvec vectors
template <typename T>
__global__ void noise_act_bwd_vec_kernel(const T* __restrict__ y, const T* __restrict__ gy, const float* __restrict__ noise,
                                         T* __restrict__ dx, float* __restrict__ dwn, long R, int C, int nvec) {
  constexpr int V = VecN<T>::N;
  __shared__ float sm[256 * V];
  const int t = threadIdx.x, cv = t & (nvec - 1), rl = t / nvec, lanes = 256 / nvec;
  float acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = 0.f;
  for (long r = (long)blockIdx.x * lanes + rl; r < R; r += (long)gridDim.x * lanes) {
    float yv[V], gv[V], ov[V];
    ldv(y + r * C + cv * V, yv);
    ldv(gy + r * C + cv * V, gv);
    const float nz = noise[r];
#pragma unroll
    for (int k = 0; k < V; ++k) { ov[k] = yv[k] > 0.f ? gv[k] : 0.2f * gv[k]; acc[k] = fmaf(ov[k], nz, acc[k]); }
    stv(dx + r * C + cv * V, ov);
  }
#pragma unroll
  for (int k = 0; k < V; ++k) sm[rl * (nvec * V) + cv * V + k] = acc[k];
  __syncthreads();
  for (int c = t; c < C; c += 256) {
    float sres = 0.f;
    for (int i = 0; i < lanes; ++i) sres += sm[i * C + c];
    atomicAdd(dwn + c, sres);
  }
}
int ggi_noise_act_fwd(const void* x, const float* noise, const float* wn, void* y, long R, int C, int dtype, cudaStream_t st) {
  int V = dtype == GG_F32 ? 4 : 8;
  if (C % V == 0 && al16(x) && al16(y)) {
    GG_DISPATCH(dtype, (noise_act_fwd_vec_kernel<T><<<gg_blocks(R * (C / V), 256, 148 * 32), 256, 0, st>>>((const T*)x, noise, wn, (T*)y, R, C)));
    return gg_check_launch("noise_act_fwd_vec");
  }
  GG_DISPATCH(dtype, (noise_act_fwd_kernel<T><<<gg_blocks(R * C, 256), 256, 0, st>>>((const T*)x, noise, wn, (T*)y, R, C)));
  return gg_check_launch("noise_act_fwd");
}
int ggi_noise_act_bwd(const void* y, const void* gy, const float* noise, void* dx, float* dwn, long R, int C, int dtype,
                     cudaStream_t st) {
  cudaMemsetAsync(dwn, 0, sizeof(float) * C, st);
  int V = dtype == GG_F32 ? 4 : 8;
  int nvec = C % V == 0 ? C / V : 0;
  if (nvec > 0 && nvec <= 256 && (nvec & (nvec - 1)) == 0 && al16(y) && al16(gy) && al16(dx)) {
    int lanes = 256 / nvec;
    int blocks = gg_blocks((R + lanes - 1) / lanes * 256, 256, 148 * 8);
    GG_DISPATCH(dtype, (noise_act_bwd_vec_kernel<T><<<blocks, 256, 0, st>>>((const T*)y, (const T*)gy, noise, (T*)dx, dwn, R, C, nvec)));
    return gg_check_launch("noise_act_bwd_vec");
  }
  int splits = 1;
  while (gg_cdiv(C, 32) * splits < 148 * 4 && R / (splits * 2) >= 64) splits *= 2;
  dim3 grid(gg_cdiv(C, 32), splits), block(32, 8);
  GG_DISPATCH(dtype, (noise_act_bwd_kernel<T><<<grid, block, 0, st>>>((const T*)y, (const T*)gy, noise, (T*)dx, dwn, R, C, splits)));
  return gg_check_launch("noise_act_bwd");
}

// ------------------------------------------------------------------ AdaptiveConv2DMod weight builder (K1)
// bank [n][o][i][kk] fp32 (reference layout, kk = k*k); mod [B][I]; kmod [B][n] (null when n == 1)
// out w [B][o][kk][i] (T, kernel layout), attn [B][n], dinv [B][o] (fp32 stats kept for backward)
__global__ void adaconv_attn_kernel(const float* __restrict__ kmod, float* __restrict__ attn, int B, int n, long ldk) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (n == 1) { attn[b] = 1.f; return; }
  float m = -INFINITY;
  for (int j = 0; j < n; ++j) m = fmaxf(m, kmod[b * ldk + j]);
  float s = 0.f;
  for (int j = 0; j < n; ++j) s += expf(kmod[b * ldk + j] - m);
  for (int j = 0; j < n; ++j) attn[b * n + j] = expf(kmod[b * ldk + j] - m) / s;
}

template <typename T>
__global__ void adaconv_weights_fwd_kernel(const float* __restrict__ bank, const float* __restrict__ mod,
                                           const float* __restrict__ attn, T* __restrict__ w,
                                           float* __restrict__ dinv, int n, int O, int I, int KK, int demod, float eps,
                                           int Opad, long ldm) {
  __shared__ float red[32];
  int b = blockIdx.y, o = blockIdx.x;
  int E = I * KK;
  float ss = 0.f;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    int i = e / KK;
    float v = 0.f;
    for (int j = 0; j < n; ++j) v += attn[b * n + j] * bank[((long)j * O + o) * E + e];
    float u = v * (mod[(long)b * ldm + i] + 1.f);
    ss += u * u;
  }
  float d = 1.f;
  if (demod) {
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[i];
    d = rsqrtf(fmaxf(t, eps));
  }
  if (threadIdx.x == 0) dinv[(long)b * O + o] = d;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    int i = e / KK, kk = e % KK;
    float v = 0.f;
    for (int j = 0; j < n; ++j) v += attn[b * n + j] * bank[((long)j * O + o) * E + e];
    float u = v * (mod[(long)b * ldm + i] + 1.f);
    stf(w + (((long)b * Opad + o) * KK + kk) * I + i, u * d);
  }
}

// backward, pass 1 (demod only): q[b,o] = sum_{i,kk} gw * u        (one CTA per (b,o))
__global__ void adaconv_q_kernel(const float* __restrict__ bank, const float* __restrict__ mod,
                                 const float* __restrict__ attn, const float* __restrict__ gw, float* __restrict__ q,
                                 int n, int O, int I, int KK, int Opad, long ldm) {
  __shared__ float red[32];
  int b = blockIdx.y, o = blockIdx.x;
  int E = I * KK;
  float acc = 0.f;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    int kk = e / I, i = e - kk * I;                        // i fastest: coalesced over gw's kernel layout
    float v = 0.f;
    for (int j = 0; j < n; ++j) v += attn[b * n + j] * bank[(((long)j * O + o) * I + i) * KK + kk];
    acc += gw[(((long)b * Opad + o) * KK + kk) * I + i] * v * (mod[(long)b * ldm + i] + 1.f);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
    q[(long)b * O + o] = t;
  }
}

// backward, pass 2: thread = (kk, i) of one output channel o, loops over the batch.  dbank is owned (plain store),
// dmod gets one atomic per (b, o, i), gattn one per (block, b, n).
__global__ void adaconv_weights_bwd_kernel(const float* __restrict__ bank, const float* __restrict__ mod,
                                           const float* __restrict__ attn, const float* __restrict__ dinv,
                                           const float* __restrict__ q, const float* __restrict__ gw,
                                           float* __restrict__ dbank, float* __restrict__ dmod,
                                           float* __restrict__ gattn, int B, int n, int O, int I, int KK, int demod,
                                           float eps, int Opad, long ldm, const float* __restrict__ dw_add) {
  extern __shared__ float sm[];              // [KK*32] partial dmod terms, then [B*n] gattn partials
  float* sm_dm = sm;
  float* sm_ga = sm + KK * 32;
  const int o = blockIdx.y, i0 = blockIdx.x * 32;
  const int il = threadIdx.x & 31, kk = threadIdx.x >> 5;          // blockDim = 32 * KK
  const int i = i0 + il;
  const bool live = i < I;
  for (int t = threadIdx.x; t < B * n; t += blockDim.x) sm_ga[t] = 0.f;
  float wb[8], acc[8];
  for (int j = 0; j < n; ++j) { wb[j] = live ? bank[(((long)j * O + o) * I + i) * KK + kk] : 0.f; acc[j] = 0.f; }
  __syncthreads();
  for (int b = 0; b < B; ++b) {
    float d = dinv[(long)b * O + o];
    float v = 0.f;
    for (int j = 0; j < n; ++j) v += attn[b * n + j] * wb[j];
    float s = live ? mod[(long)b * ldm + i] + 1.f : 0.f;
    float g = (live && gw) ? gw[(((long)b * Opad + o) * KK + kk) * I + i] : 0.f;
    float gu = d * g;
    if (demod && d * d * eps < 0.999999f) gu -= d * d * d * (v * s) * q[(long)b * O + o];
    float gv = gu * s;
    sm_dm[kk * 32 + il] = gu * v;
    for (int j = 0; j < n; ++j) {
      acc[j] += attn[b * n + j] * gv;
      float t = warp_sum(gv * wb[j]);
      if (il == 0) atomicAdd(&sm_ga[b * n + j], t);
    }
    __syncthreads();
    if (kk == 0 && live) {
      float t = 0.f;
      for (int k2 = 0; k2 < KK; ++k2) t += sm_dm[k2 * 32 + il];
      atomicAdd(dmod + (long)b * I + i, t);
    }
    __syncthreads();
  }
  if (live)                 // dw_add: kernel-layout [n][O][KK][I] weight gradients of the shared-bank convolutions, folded in
    for (int j = 0; j < n; ++j)
      dbank[(((long)j * O + o) * I + i) * KK + kk] = acc[j] + (dw_add ? dw_add[(((long)j * O + o) * KK + kk) * I + i] : 0.f);
  for (int t = threadIdx.x; t < B * n; t += blockDim.x) atomicAdd(gattn + t, sm_ga[t]);
}

// Restructured pass 2 (KK = 1 or 9, batch chunks of <= 16 images): lane = input channel of a 32-channel tile, each of
// the 8 warps walks output channels o0+w, o0+w+8, ... of the block's OC-channel range.  dmod is accumulated over the
// whole o range in registers (one atomic per (block, b, i) instead of one per (b, o, i): the 512-way contention on
// dmod made the 512x512 layers take 330 us), gattn through per-warp shared-memory slots.  accumulate != 0: dbank +=.
#define AB_BCH 16
#define AB_OC 8      // one output channel per warp: the per-warp work is a serial chain (all launches took ~110 us at 32)
template <int KK, int NK>
__global__ void __launch_bounds__(256)
adaconv_weights_bwd2_kernel(const float* __restrict__ bank, const float* __restrict__ mod, const float* __restrict__ attn,
                            const float* __restrict__ dinv, const float* __restrict__ q, const float* __restrict__ gw,
                            float* __restrict__ dbank, float* __restrict__ dmod, float* __restrict__ gattn, int b0, int nb,
                            int n, int O, int I, int demod, float eps, int Opad, long ldm,
                            const float* __restrict__ dw_add, int accumulate) {
  __shared__ float sa[AB_BCH][8];
  __shared__ float sm_ga[8][AB_BCH][8];
  __shared__ float sm_dm[8][AB_BCH][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane, o0 = blockIdx.y * AB_OC;
  const bool live = i < I;
  for (int t = threadIdx.x; t < AB_BCH * 8; t += blockDim.x) {
    int bb = t >> 3, j = t & 7;
    sa[bb][j] = (bb < nb && j < n) ? attn[(b0 + bb) * n + j] : 0.f;
    for (int w = 0; w < 8; ++w) sm_ga[w][bb][j] = 0.f;
  }
  __syncthreads();
  float dm[AB_BCH], sv[AB_BCH];
#pragma unroll
  for (int t = 0; t < AB_BCH; ++t) { dm[t] = 0.f; sv[t] = (live && t < nb) ? mod[(long)(b0 + t) * ldm + i] + 1.f : 0.f; }
  for (int oo = warp; oo < AB_OC && o0 + oo < O; oo += 8) {
    const int o = o0 + oo;
    float wb[NK][KK], acc[NK][KK];
#pragma unroll
    for (int j = 0; j < NK; ++j)
#pragma unroll
      for (int kk = 0; kk < KK; ++kk) {
        wb[j][kk] = (live && j < n) ? bank[(((long)j * O + o) * I + i) * KK + kk] : 0.f;
        acc[j][kk] = 0.f;
      }
#pragma unroll
    for (int t = 0; t < AB_BCH; ++t) {
      if (t < nb) {
      const int b = b0 + t;
      float gq[KK];                          // the KK gradient loads of this (image, channel) go out back to back
#pragma unroll
      for (int kk = 0; kk < KK; ++kk) gq[kk] = (live && gw) ? __ldg(gw + (((long)b * Opad + o) * KK + kk) * I + i) : 0.f;
      const float d = dinv[(long)b * O + o], s = sv[t];
      const float qv = demod ? q[(long)b * O + o] : 0.f;
      const bool clamp = !(demod && d * d * eps < 0.999999f);
      float dmv = 0.f, gvw[NK];
#pragma unroll
      for (int j = 0; j < NK; ++j) gvw[j] = 0.f;
#pragma unroll
      for (int kk = 0; kk < KK; ++kk) {
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < NK; ++j) v += sa[t][j] * wb[j][kk];
        const float g = gq[kk];
        float gu = d * g;
        if (!clamp) gu -= d * d * d * (v * s) * qv;
        const float gv = gu * s;
        dmv += gu * v;
#pragma unroll
        for (int j = 0; j < NK; ++j) { acc[j][kk] += sa[t][j] * gv; gvw[j] += gv * wb[j][kk]; }
      }
      dm[t] += dmv;
#pragma unroll
      for (int j = 0; j < NK; ++j) {
        if (j < n) {
          const float r = warp_sum(gvw[j]);
          if (lane == 0) sm_ga[warp][t][j] += r;
        }
      }
      }
    }
    if (live) {
#pragma unroll
      for (int j = 0; j < NK; ++j)
        if (j < n) {
#pragma unroll
          for (int kk = 0; kk < KK; ++kk) {
            const long at = (((long)j * O + o) * I + i) * KK + kk;
            float r = acc[j][kk] + (dw_add ? dw_add[(((long)j * O + o) * KK + kk) * I + i] : 0.f);
            dbank[at] = accumulate ? dbank[at] + r : r;
          }
        }
    }
  }
#pragma unroll
  for (int t = 0; t < AB_BCH; ++t) sm_dm[warp][t][lane] = dm[t];
  __syncthreads();
  for (int t = warp; t < nb; t += 8) {
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) r += sm_dm[w][t][lane];
    if (live) atomicAdd(dmod + (long)(b0 + t) * I + i, r);
  }
  for (int t = threadIdx.x; t < nb * 8; t += blockDim.x) {
    const int bb = t >> 3, j = t & 7;
    if (j < n) {
      float r = 0.f;
      for (int w = 0; w < 8; ++w) r += sm_ga[w][bb][j];
      atomicAdd(gattn + (b0 + bb) * n + j, r);
    }
  }
}
template <int KK>
static void launch_bwd2(dim3 grid, cudaStream_t st, int nk, const float* bank, const float* mod, const float* attn,
                        const float* dinv, const float* q, const float* gw, float* dbank, float* dmod, float* gattn, int b0,
                        int nb, int n, int O, int I, int demod, float eps, int Opad, long ldm, const float* dw_add, int acc) {
#define AB_GO(NK) adaconv_weights_bwd2_kernel<KK, NK><<<grid, 256, 0, st>>>(bank, mod, attn, dinv, q, gw, dbank, dmod, gattn, \
                                                                           b0, nb, n, O, I, demod, eps, Opad, ldm, dw_add, acc)
  if (nk <= 1) AB_GO(1);
  else if (nk <= 2) AB_GO(2);
  else if (nk <= 4) AB_GO(4);
  else AB_GO(8);
#undef AB_GO
}

// dkmod = attn * (gattn - sum_j attn_j gattn_j)
__global__ void adaconv_kmod_bwd_kernel(const float* __restrict__ attn, const float* __restrict__ gattn,
                                        float* __restrict__ dkmod, int B, int n) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float dot = 0.f;
  for (int j = 0; j < n; ++j) dot += attn[b * n + j] * gattn[b * n + j];
  for (int j = 0; j < n; ++j) dkmod[b * n + j] = attn[b * n + j] * (gattn[b * n + j] - dot);
}

int ggi_adaconv_weights_fwd(const float* bank, const float* mod, const float* kmod, void* w, float* attn, float* dinv,
                           int B, int n, int O, int I, int KK, int demod, float eps, int Opad, long ldm, long ldk, int dtype,
                           cudaStream_t st) {
  if (n > 8) return gg_fail("num_conv_kernels > 8 unsupported");
  adaconv_attn_kernel<<<gg_cdiv(B, 128), 128, 0, st>>>(kmod, attn, B, n, ldk);
  dim3 grid(O, B);
  GG_DISPATCH(dtype, (adaconv_weights_fwd_kernel<T><<<grid, 256, 0, st>>>(bank, mod, attn, (T*)w, dinv, n, O, I, KK, demod, eps, Opad, ldm)));
  return gg_check_launch("adaconv_weights_fwd");
}
// gw != NULL: gradient of the materialised per-sample weights (q computed here);  gw == NULL: shared-bank form, only the
// demodulation chain is differentiated: q_ext[b][o] = dL/d dinv, gattn_ws[0 .. B*n) already holds the caller's direct
// dL/d attn (NOT cleared here), dw_add = kernel-layout weight gradients of the bank convolutions to fold into dbank.
int ggi_adaconv_weights_bwd(const float* bank, const float* mod, const float* attn, const float* dinv, const float* gw,
                           float* dbank, float* dmod, float* dkmod, float* gattn_ws, int B, int n, int O, int I, int KK,
                           int demod, float eps, int Opad, long ldm, const float* q_ext, const float* dw_add,
                           cudaStream_t st) {
  // gattn_ws: B*n floats of gattn followed by B*O floats for q
  float* q = gattn_ws + (size_t)B * n;
  cudaMemsetAsync(dmod, 0, sizeof(float) * (size_t)B * I, st);
  if (gw) cudaMemsetAsync(gattn_ws, 0, sizeof(float) * (size_t)B * n, st);
  if (KK > 32) return gg_fail("adaconv backward: kernel area %d unsupported", KK);
  if (demod && gw) {
    dim3 g1(O, B);
    adaconv_q_kernel<<<g1, 256, 0, st>>>(bank, mod, attn, gw, q, n, O, I, KK, Opad, ldm);
  }
  if (KK == 9 || KK == 1) {
    dim3 grid(gg_cdiv(I, 32), gg_cdiv(O, AB_OC));
    for (int b0 = 0; b0 < B; b0 += AB_BCH) {
      int nb = B - b0 < AB_BCH ? B - b0 : AB_BCH;
      if (KK == 9) launch_bwd2<9>(grid, st, n, bank, mod, attn, dinv, gw ? q : q_ext, gw, dbank, dmod, gattn_ws, b0, nb, n, O, I,
                                  demod, eps, Opad, ldm, b0 == 0 ? dw_add : nullptr, b0 > 0);
      else launch_bwd2<1>(grid, st, n, bank, mod, attn, dinv, gw ? q : q_ext, gw, dbank, dmod, gattn_ws, b0, nb, n, O, I, demod,
                          eps, Opad, ldm, b0 == 0 ? dw_add : nullptr, b0 > 0);
    }
  } else {
    dim3 grid(gg_cdiv(I, 32), O);
    size_t smem = sizeof(float) * ((size_t)KK * 32 + (size_t)B * n);
    adaconv_weights_bwd_kernel<<<grid, 32 * KK, smem, st>>>(bank, mod, attn, dinv, gw ? q : q_ext, gw, dbank, dmod, gattn_ws,
                                                           B, n, O, I, KK, demod, eps, Opad, ldm, dw_add);
  }
  if (n > 1 && dkmod) adaconv_kmod_bwd_kernel<<<gg_cdiv(B, 128), 128, 0, st>>>(attn, gattn_ws, dkmod, B, n);
  return gg_check_launch("adaconv_weights_bwd");
}

// ------------------------------------------------------------------ AdamW over a flat fp32 buffer
// chunk table: int4 {offset_lo, len, wd_flag, offset_hi}; one CTA per chunk.  step_ptr holds the 1-based step.
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, const int4* __restrict__ chunks, const int* __restrict__ step_ptr,
                             float lr, float b1, float b2, float eps, float wd, float grad_scale) {
  int4 ch = chunks[blockIdx.x];
  long off = ((long)ch.w << 31) | (unsigned)ch.x;
  int step = *step_ptr;
  float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
  float decay = ch.z ? 1.f - lr * wd : 1.f;
  float inv_sqrt_bc2 = rsqrtf(bc2), step_size = lr / bc1;
  for (int i = threadIdx.x; i < ch.y; i += blockDim.x) {
    long j = off + i;
    float gi = g[j] * grad_scale, pi = p[j] * decay;
    float mi = m[j] + (1.f - b1) * (gi - m[j]);
    float vi = b2 * v[j] + (1.f - b2) * gi * gi;
    float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[j] = pi - step_size * (mi / denom);
    m[j] = mi; v[j] = vi;
  }
}
int ggi_adamw(float* p, const float* g, float* m, float* v, const void* chunks, int nchunks, const int* step_ptr,
             float lr, float b1, float b2, float eps, float wd, float grad_scale, cudaStream_t st) {
  adamw_kernel<<<nchunks, 256, 0, st>>>(p, g, m, v, (const int4*)chunks, step_ptr, lr, b1, b2, eps, wd, grad_scale);
  return gg_check_launch("adamw");
}
__global__ void incr_kernel(int* p) { *p += 1; }
int ggi_incr(int* p, cudaStream_t st) { incr_kernel<<<1, 1, 0, st>>>(p); return gg_check_launch("incr"); }

// ------------------------------------------------------------------ multi-tensor weight re-layout (once per step)
// For every registered conv weight (fp32 master, reference layout [O][I][KK]) write BOTH kernel layouts in bf16/fp32:
//   fwd[o][kk][ipad]            (K-major B operand of the forward implicit GEMM, input channels zero-padded)
//   bwd[ipad][KK-1-kk][o]       (flipped + in/out swapped: the data gradient runs as a forward convolution)
// entries: int64 x 8 = {src_off, O, I, KK, Ipad, fwd_off, bwd_off, 0}; chunks: int32 x 4 = {entry, o0, i0, TI}:
// one block moves the (32 output channels) x (TI input channels) x KK tile through shared memory so that the master
// read ([o][i][kk]: runs of TI*KK floats) and both writes (runs of TI along i / 32 along o) are contiguous.
#define WP_MAX_ROW 380
template <typename T>
__global__ void weight_prep_multi_kernel(const float* __restrict__ master, const long* __restrict__ entries,
                                         const int4* __restrict__ chunks, T* __restrict__ fwd, T* __restrict__ bwd) {
  __shared__ float s[32 * (WP_MAX_ROW + 1)];
  const int4 ch = chunks[blockIdx.x];
  const long* e = entries + (long)ch.x * 8;
  const long src = e[0], fo = e[5], bo = e[6];
  const int O = (int)e[1], I = (int)e[2], KK = (int)e[3], Ip = (int)e[4];
  const int o0 = ch.y, i0 = ch.z, TI = ch.w;
  const int no = min(32, O - o0), ni = min(TI, Ip - i0);        // tile extents (ni counts padded channels too)
  const int nir = max(0, min(ni, I - i0));                       // channels that exist in the master
  const int pitch = TI * KK + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int ol = warp; ol < no; ol += 8) {
    const float* row = master + src + ((long)(o0 + ol) * I + i0) * KK;
    for (int j = lane; j < ni * KK; j += 32) s[ol * pitch + j] = j < nir * KK ? row[j] : 0.f;
  }
  __syncthreads();
  if (sizeof(T) == 2 && !((ni | i0 | Ip | no | o0 | O) & 1) && !((fo | bo) & 1)) {
    // bf16: two neighbouring elements per 4-byte store (the scalar loops below were store-instruction bound at ~1 TB/s)
    const int nh = ni >> 1, oh = no >> 1;
    for (int t = threadIdx.x; t < no * KK * nh; t += blockDim.x) {         // fwd[o][kk][i, i+1]
      int ih = t % nh, r = t / nh;
      int kk = r % KK, ol = r / KK;
      const float* sp = s + ol * pitch + (2 * ih) * KK + kk;
      *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<bf16*>(fwd) + fo + ((long)(o0 + ol) * KK + kk) * Ip + i0 + 2 * ih) =
          __floats2bfloat162_rn(sp[0], sp[KK]);
    }
    for (int t = threadIdx.x; t < ni * KK * oh; t += blockDim.x) {         // bwd[i][KK-1-kk][o, o+1]
      int o2 = t % oh, r = t / oh;
      int kk = r % KK, il = r / KK;
      const float* sp = s + (2 * o2) * pitch + il * KK + kk;
      *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<bf16*>(bwd) + bo + ((long)(i0 + il) * KK + (KK - 1 - kk)) * O + o0 + 2 * o2) =
          __floats2bfloat162_rn(sp[0], sp[pitch]);
    }
    return;
  }
  for (int t = threadIdx.x; t < no * KK * ni; t += blockDim.x) {           // fwd[o][kk][i]
    int il = t % ni, r = t / ni;
    int kk = r % KK, ol = r / KK;
    stf(fwd + fo + ((long)(o0 + ol) * KK + kk) * Ip + i0 + il, s[ol * pitch + il * KK + kk]);
  }
  for (int t = threadIdx.x; t < ni * KK * no; t += blockDim.x) {           // bwd[i][KK-1-kk][o]
    int ol = t % no, r = t / no;
    int kk = r % KK, il = r / KK;
    stf(bwd + bo + ((long)(i0 + il) * KK + (KK - 1 - kk)) * O + o0 + ol, s[ol * pitch + il * KK + kk]);
  }
}
int ggi_weight_prep_multi(const float* master, const void* entries, const void* chunks, int nchunks, void* fwd, void* bwd,
                          int dtype, cudaStream_t st) {
  GG_DISPATCH(dtype, (weight_prep_multi_kernel<T><<<nchunks, 256, 0, st>>>(master, (const long*)entries, (const int4*)chunks, (T*)fwd, (T*)bwd)));
  return gg_check_launch("weight_prep_multi");
}

// Weight-gradient sink: the kernel-layout fp32 gradient dw[o][kk][ipad] produced by the wgrad kernels is ACCUMULATED into
// the master-layout gradient buffer dst[o][i][kk] (the optimiser's flat .grad view) in one pass: replaces a permute
// copy plus the autograd engine's separate accumulation kernel.  One block = one output channel x 128 input channels.
__global__ void wgrad_sink_kernel(const float* __restrict__ dw, float* __restrict__ dst, int O, int I, int KK, int Ipad) {
  extern __shared__ float sk[];                     // [KK][129]
  const int o = blockIdx.y, i0 = blockIdx.x * 128;
  const int ni = min(128, I - i0);
  const float* src = dw + (long)o * KK * Ipad + i0;
  for (int t = threadIdx.x; t < KK * 128; t += blockDim.x) {
    int il = t & 127, kk = t >> 7;
    if (il < ni) sk[kk * 129 + il] = src[(long)kk * Ipad + il];
  }
  __syncthreads();
  float* d = dst + ((long)o * I + i0) * KK;
  for (int t = threadIdx.x; t < ni * KK; t += blockDim.x) {
    int il = t / KK, kk = t - il * KK;
    d[t] += sk[kk * 129 + il];
  }
}
int ggi_wgrad_sink(const float* dw, float* dst, int O, int I, int KK, int Ipad, cudaStream_t st) {
  if (KK > 64) return gg_fail("wgrad_sink: filter too large");
  dim3 grid((I + 127) / 128, O);
  wgrad_sink_kernel<<<grid, 256, KK * 129 * sizeof(float), st>>>(dw, dst, O, I, KK, Ipad);
  return gg_check_launch("wgrad_sink");
}

// ------------------------------------------------------------------ LeakyReLU backward fused with the bias gradient
// out = gy * lrelu'(y)  and  dbias[c] += sum_rows out[r, c]  in ONE pass over the activation gradient (the separate
// column reduction re-read the whole map).  nvec = C / V channel vectors (a power of two <= 256): the 256 threads tile
// (256 / nvec row lanes) x nvec vectors, every thread keeps its V column sums in registers.
template <typename T>
__global__ void lrelu_bwd_bias_kernel(const T* __restrict__ y, const T* __restrict__ gy, T* __restrict__ out,
                                      float* __restrict__ dbias, long R, int C, int nvec) {
  constexpr int V = VecN<T>::N;
  __shared__ float sm[256 * V];
  const int t = threadIdx.x, cv = t & (nvec - 1), rl = t / nvec, lanes = 256 / nvec;
  float acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = 0.f;
  for (long r = (long)blockIdx.x * lanes + rl; r < R; r += (long)gridDim.x * lanes) {
    float yv[V], gv[V], ov[V];
    ldv(y + r * C + cv * V, yv);
    ldv(gy + r * C + cv * V, gv);
#pragma unroll
    for (int k = 0; k < V; ++k) { ov[k] = yv[k] > 0.f ? gv[k] : 0.2f * gv[k]; acc[k] += ov[k]; }
    stv(out + r * C + cv * V, ov);
  }
#pragma unroll
  for (int k = 0; k < V; ++k) sm[rl * (nvec * V) + cv * V + k] = acc[k];
  __syncthreads();
  for (int c = t; c < C; c += 256) {
    float sres = 0.f;
    for (int i = 0; i < lanes; ++i) sres += sm[i * C + c];
    atomicAdd(dbias + c, sres);
  }
}
// returns 1 when the shape is not covered (caller composes unary + dot_sc)
int ggi_lrelu_bwd_bias(const void* y, const void* gy, void* out, float* dbias, long R, int C, int accumulate, int dtype,
                       cudaStream_t st) {
  int V = dtype == GG_F32 ? 4 : 8;
  if (C % V || !al16(y) || !al16(gy) || !al16(out)) return 1;
  int nvec = C / V;
  if (nvec > 256 || (nvec & (nvec - 1))) return 1;
  if (!accumulate) cudaMemsetAsync(dbias, 0, sizeof(float) * C, st);
  int lanes = 256 / nvec;
  int blocks = gg_blocks((R + lanes - 1) / lanes * 256, 256, 148 * 8);
  GG_DISPATCH(dtype, (lrelu_bwd_bias_kernel<T><<<blocks, 256, 0, st>>>((const T*)y, (const T*)gy, (T*)out, dbias, R, C, nvec)));
  return gg_check_launch("lrelu_bwd_bias");
}

// ------------------------------------------------------------------ fused ChannelRMSNorm (gigagan_pytorch.py:224-232)
// y[r,c] = x[r,c] * inv[r] * s * gamma[c],  inv[r] = 1 / max(||x[r,:]||, 1e-12);  one warp per pixel row, 16-byte
// accesses, x read twice (second time from L1/L2).  First-order fast path of the composed rowdot/invnorm/bcast chain.
template <typename T>
__global__ void rmsnorm_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma, T* __restrict__ y,
                                   float* __restrict__ inv, long R, int C, float s) {
  constexpr int V = VecN<T>::N;
  const int lane = threadIdx.x & 31;
  const long warp0 = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long r = warp0; r < R; r += nwarps) {
    const T* xr = x + r * C;
    float ss = 0.f;
    for (int c = lane * V; c < C; c += 32 * V) {
      float v[V];
      ldv<T>(xr + c, v);
#pragma unroll
      for (int i = 0; i < V; ++i) ss += v[i] * v[i];
    }
    ss = warp_sum(ss);
    const float iv = ss <= 1e-24f ? 1e12f : rsqrtf(ss);
    if (lane == 0) inv[r] = iv;
    T* yr = y + r * C;
    for (int c = lane * V; c < C; c += 32 * V) {
      float v[V], o[V];
      ldv<T>(xr + c, v);
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = v[i] * iv * (s * __ldg(gamma + c + i));
      stv<T>(yr + c, o);
    }
  }
}
// gx = inv * (t - xh * dot(t, xh)),  t = gy * s * gamma,  xh = x * inv;   dgamma[c] += sum_r gy * xh * s
template <typename T>
__global__ void rmsnorm_bwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ inv,
                                   const T* __restrict__ gy, T* __restrict__ gx, float* __restrict__ dgamma, long R, int C,
                                   float s) {
  constexpr int V = VecN<T>::N;
  extern __shared__ float dg[];                                    // [C] per-block partial of dgamma
  for (int c = threadIdx.x; c < C; c += blockDim.x) dg[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long warp0 = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  float acc[4][V];                                                 // this lane's columns: lane*V + k*32*V, k < 4
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[k][i] = 0.f;
  float gam[4][V];                                                 // s * gamma of this lane's columns
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = lane * V + k * 32 * V;
#pragma unroll
    for (int i = 0; i < V; ++i) gam[k][i] = c < C ? s * __ldg(gamma + c + i) : 0.f;
  }
  for (long r = warp0; r < R; r += nwarps) {
    const T* xr = x + r * C;
    const T* gr = gy + r * C;
    const float iv = inv[r];
    const bool clamped = iv >= 1e12f;                              // ||x|| below eps: y = x / eps, no projection term
    float vv[4][V], gg[4][V];                                      // the row is read from global memory ONCE
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = lane * V + k * 32 * V;
      if (c < C) { ldv<T>(xr + c, vv[k]); ldv<T>(gr + c, gg[k]); }
    }
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = lane * V + k * 32 * V;
      if (c < C) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float xh = vv[k][i] * iv;
          dot += gg[k][i] * gam[k][i] * xh;
          acc[k][i] += gg[k][i] * xh * s;
        }
      }
    }
    dot = clamped ? 0.f : warp_sum(dot);
    T* or_ = gx + r * C;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = lane * V + k * 32 * V;
      if (c < C) {
        float o[V];
#pragma unroll
        for (int i = 0; i < V; ++i) o[i] = iv * (gg[k][i] * gam[k][i] - vv[k][i] * iv * dot);
        stv<T>(or_ + c, o);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = lane * V + k * 32 * V;
    if (c < C) {
#pragma unroll
      for (int i = 0; i < V; ++i) atomicAdd(&dg[c + i], acc[k][i]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(dgamma + c, dg[c]);
}
int ggi_rmsnorm_fwd(const void* x, const float* gamma, void* y, float* inv, long R, int C, float s, int dtype, cudaStream_t st) {
  if (C % 8 || !al16(x) || !al16(y)) return gg_fail("rmsnorm_fwd: C %% 8 != 0 or unaligned");
  int blocks = gg_blocks(R * 32, 256, 148 * 8);
  GG_DISPATCH(dtype, (rmsnorm_fwd_kernel<T><<<blocks, 256, 0, st>>>((const T*)x, gamma, (T*)y, inv, R, C, s)));
  return gg_check_launch("rmsnorm_fwd");
}
int ggi_rmsnorm_bwd(const void* x, const float* gamma, const float* inv, const void* gy, void* gx, float* dgamma, long R, int C,
                    float s, int dtype, cudaStream_t st) {
  if (C % 8 || C > (dtype == GG_BF16 ? 1024 : 512) || !al16(x) || !al16(gy) || !al16(gx))
    return gg_fail("rmsnorm_bwd: needs C %% 8 == 0, C <= 1024 (bf16) / 512 (fp32), 16-byte aligned tensors");
  int blocks = gg_blocks(R * 32, 256, 148 * 4);
  GG_DISPATCH(dtype, (rmsnorm_bwd_kernel<T><<<blocks, 256, C * sizeof(float), st>>>((const T*)x, gamma, inv, (const T*)gy, (T*)gx, dgamma, R, C, s)));
  return gg_check_launch("rmsnorm_bwd");
}

// ------------------------------------------------------------------ UnetUpsampler extras (unet_upsampler.py)
// 2x2 max-pool of NHWC maps (ref :158 F.max_pool2d) and its gradient (first maximum in row-major order wins, as ATen)
template <typename T>
__global__ void maxpool2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C) {
  int OH = H / 2, OW = W / 2;
  long n = (long)N * OH * OW * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long t = i / C;
    int ox = (int)(t % OW); t /= OW;
    int oy = (int)(t % OH); int b = (int)(t / OH);
    const T* p0 = x + (((long)b * H + 2 * oy) * W + 2 * ox) * C + c;
    float m = fmaxf(fmaxf(ldf(p0), ldf(p0 + C)), fmaxf(ldf(p0 + (long)W * C), ldf(p0 + (long)W * C + C)));
    stf(y + i, m);
  }
}
template <typename T>
__global__ void maxpool2_bwd_kernel(const T* __restrict__ x, const T* __restrict__ gy, T* __restrict__ gx, int N, int H,
                                    int W, int C) {
  int OH = H / 2, OW = W / 2;
  long n = (long)N * OH * OW * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long t = i / C;
    int ox = (int)(t % OW); t /= OW;
    int oy = (int)(t % OH); int b = (int)(t / OH);
    long o00 = (((long)b * H + 2 * oy) * W + 2 * ox) * C + c;
    long off[4] = {o00, o00 + C, o00 + (long)W * C, o00 + (long)W * C + C};
    float v[4];
    int best = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = ldf(x + off[k]); if (v[k] > v[best]) best = k; }
    float g = ldf(gy + i);
#pragma unroll
    for (int k = 0; k < 4; ++k) stf(gx + off[k], k == best ? g : 0.f);
  }
}
int ggi_maxpool2_fwd(const void* x, void* y, int N, int H, int W, int C, int dtype, cudaStream_t st) {
  long n = (long)N * (H / 2) * (W / 2) * C;
  GG_DISPATCH(dtype, (maxpool2_fwd_kernel<T><<<gg_blocks(n, 256), 256, 0, st>>>((const T*)x, (T*)y, N, H, W, C)));
  return gg_check_launch("maxpool2_fwd");
}
int ggi_maxpool2_bwd(const void* x, const void* gy, void* gx, int N, int H, int W, int C, int dtype, cudaStream_t st) {
  long n = (long)N * (H / 2) * (W / 2) * C;
  GG_DISPATCH(dtype, (maxpool2_bwd_kernel<T><<<gg_blocks(n, 256), 256, 0, st>>>((const T*)x, (const T*)gy, (T*)gx, N, H, W, C)));
  return gg_check_launch("maxpool2_bwd");
}

// softmax over the TOKEN axis of (B, n, C) maps, per (sample, channel)  (LinearAttention k.softmax(dim=-1), ref :340)
template <typename T>
__global__ void softmax_tokens_kernel(const T* __restrict__ x, T* __restrict__ y, int n, int C) {
  __shared__ float sm[8][33];
  int c = blockIdx.x * 32 + threadIdx.x, b = blockIdx.y;
  const T* xb = x + (long)b * n * C;
  T* yb = y + (long)b * n * C;
  float m = -INFINITY;
  if (c < C) for (int r = threadIdx.y; r < n; r += 8) m = fmaxf(m, ldf(xb + (long)r * C + c));
  sm[threadIdx.y][threadIdx.x] = m;
  __syncthreads();
  m = sm[0][threadIdx.x];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, sm[i][threadIdx.x]);
  __syncthreads();
  float s = 0.f;
  if (c < C) for (int r = threadIdx.y; r < n; r += 8) s += __expf(ldf(xb + (long)r * C + c) - m);
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += sm[i][threadIdx.x];
  float inv = 1.f / s;
  if (c < C) for (int r = threadIdx.y; r < n; r += 8) stf(yb + (long)r * C + c, __expf(ldf(xb + (long)r * C + c) - m) * inv);
}
int ggi_softmax_tokens(const void* x, void* y, int B, int n, int C, int dtype, cudaStream_t st) {
  dim3 grid(gg_cdiv(C, 32), B), block(32, 8);
  GG_DISPATCH(dtype, (softmax_tokens_kernel<T><<<grid, block, 0, st>>>((const T*)x, (T*)y, n, C)));
  return gg_check_launch("softmax_tokens");
}
