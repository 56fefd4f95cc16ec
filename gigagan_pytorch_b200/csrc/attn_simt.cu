// Fused (flash-style, online-softmax) attention with a learned null key/value and either dot-product or
// shared-QK L2-distance logits.  FFMA version (fp32 accumulate, any dim_head in {8,16,32,64}).
// Replaces gigagan_pytorch.py:562-592 (SelfAttention core; sim/attn never materialised) and
// attend.py:64-110.  L2 logits use  -|q-k|^2*s == (2 q.k - |k|^2)*s - |q|^2*s ; the last term is constant
// along the softmax axis and is dropped (SURVEY.md section 7 identity).
#include "gg_common.cuh"

#define ATT_T 128   // threads per CTA == queries (fwd, dq) or keys (dk/dv) per CTA
#define ATT_TILE 32

struct AttP {
  int B, heads, nq, nk, d, has_null, mode;
  long q_rs, k_rs, v_rs, o_rs;     // row strides (elements); batch stride = n * row stride; head offset = h*d
  float scale;
};

template <typename T, int D>
__global__ void __launch_bounds__(ATT_T) attn_fwd_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                         const T* __restrict__ v, const float* __restrict__ null_kv,
                                                         T* __restrict__ o, float* __restrict__ lse, AttP p) {
  __shared__ float ks[ATT_TILE][D + 1], vs[ATT_TILE][D + 1], kb[ATT_TILE];
  const int bh = blockIdx.y, b = bh / p.heads, h = bh % p.heads;
  const int i = blockIdx.x * ATT_T + threadIdx.x;
  const bool live = i < p.nq;
  float qr[D], acc[D];
  const float ls = p.mode == 1 ? 2.f * p.scale : p.scale;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    qr[c] = live ? ldf(q + ((long)b * p.nq + i) * p.q_rs + h * D + c) * ls : 0.f;
    acc[c] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  const int nkt = p.nk + p.has_null;
  for (int j0 = 0; j0 < nkt; j0 += ATT_TILE) {
    __syncthreads();
    for (int e = threadIdx.x; e < ATT_TILE * D; e += ATT_T) {
      int jj = e / D, c = e % D, j = j0 + jj;
      float kv = 0.f, vv = 0.f;
      if (j < nkt) {
        if (p.has_null && j == 0) { kv = null_kv[h * D + c]; vv = null_kv[(p.heads + h) * D + c]; }
        else {
          long row = (long)b * p.nk + (j - p.has_null);
          kv = ldf(k + row * p.k_rs + h * D + c);
          vv = ldf(v + row * p.v_rs + h * D + c);
        }
      }
      ks[jj][c] = kv; vs[jj][c] = vv;
    }
    __syncthreads();
    if (threadIdx.x < ATT_TILE) {
      float s = 0.f;
      if (p.mode == 1)
        for (int c = 0; c < D; ++c) s += ks[threadIdx.x][c] * ks[threadIdx.x][c];
      kb[threadIdx.x] = -p.scale * s;
    }
    __syncthreads();
    int jn = min(ATT_TILE, nkt - j0);
    for (int jj = 0; jj < jn; ++jj) {
      float s = kb[jj];
#pragma unroll
      for (int c = 0; c < D; ++c) s = fmaf(qr[c], ks[jj][c], s);
      float mn = fmaxf(m, s);
      float corr = __expf(m - mn), pj = __expf(s - mn);
      l = l * corr + pj;
#pragma unroll
      for (int c = 0; c < D; ++c) acc[c] = fmaf(acc[c], corr, pj * vs[jj][c]);
      m = mn;
    }
  }
  if (live) {
    float inv = 1.f / l;
#pragma unroll
    for (int c = 0; c < D; ++c) stf(o + ((long)b * p.nq + i) * p.o_rs + h * D + c, acc[c] * inv);
    lse[(long)bh * p.nq + i] = m + __logf(l);
  }
}

// dq (and delta = rowsum(go*o)) : one thread per query
template <typename T, int D>
__global__ void __launch_bounds__(ATT_T) attn_bwd_dq_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                            const T* __restrict__ v, const float* __restrict__ null_kv,
                                                            const T* __restrict__ o, const T* __restrict__ go,
                                                            const float* __restrict__ lse, T* __restrict__ dq,
                                                            float* __restrict__ delta, AttP p) {
  __shared__ float ks[ATT_TILE][D + 1], vs[ATT_TILE][D + 1], kb[ATT_TILE];
  const int bh = blockIdx.y, b = bh / p.heads, h = bh % p.heads;
  const int i = blockIdx.x * ATT_T + threadIdx.x;
  const bool live = i < p.nq;
  const float ls = p.mode == 1 ? 2.f * p.scale : p.scale;
  float qr[D], gr[D], dacc[D];
  float dl = 0.f;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    long off = ((long)b * p.nq + i) * p.q_rs + h * D + c, offo = ((long)b * p.nq + i) * p.o_rs + h * D + c;
    qr[c] = live ? ldf(q + off) * ls : 0.f;
    gr[c] = live ? ldf(go + offo) : 0.f;
    dl += live ? gr[c] * ldf(o + offo) : 0.f;
    dacc[c] = 0.f;
  }
  const float L = live ? lse[(long)bh * p.nq + i] : 0.f;
  if (live) delta[(long)bh * p.nq + i] = dl;
  const int nkt = p.nk + p.has_null;
  for (int j0 = 0; j0 < nkt; j0 += ATT_TILE) {
    __syncthreads();
    for (int e = threadIdx.x; e < ATT_TILE * D; e += ATT_T) {
      int jj = e / D, c = e % D, j = j0 + jj;
      float kv = 0.f, vv = 0.f;
      if (j < nkt) {
        if (p.has_null && j == 0) { kv = null_kv[h * D + c]; vv = null_kv[(p.heads + h) * D + c]; }
        else {
          long row = (long)b * p.nk + (j - p.has_null);
          kv = ldf(k + row * p.k_rs + h * D + c);
          vv = ldf(v + row * p.v_rs + h * D + c);
        }
      }
      ks[jj][c] = kv; vs[jj][c] = vv;
    }
    __syncthreads();
    if (threadIdx.x < ATT_TILE) {
      float s = 0.f;
      if (p.mode == 1)
        for (int c = 0; c < D; ++c) s += ks[threadIdx.x][c] * ks[threadIdx.x][c];
      kb[threadIdx.x] = -p.scale * s;
    }
    __syncthreads();
    int jn = min(ATT_TILE, nkt - j0);
    for (int jj = 0; jj < jn; ++jj) {
      float s = kb[jj], dp = 0.f;
#pragma unroll
      for (int c = 0; c < D; ++c) { s = fmaf(qr[c], ks[jj][c], s); dp = fmaf(gr[c], vs[jj][c], dp); }
      float ds = __expf(s - L) * (dp - dl) * ls;
#pragma unroll
      for (int c = 0; c < D; ++c) dacc[c] = fmaf(ds, ks[jj][c], dacc[c]);
    }
  }
  if (live) {
#pragma unroll
    for (int c = 0; c < D; ++c) stf(dq + ((long)b * p.nq + i) * p.q_rs + h * D + c, dacc[c]);
  }
}

// dk, dv : one thread per key (key 0 = null when has_null; its gradient is atomically summed over the batch)
template <typename T, int D>
__global__ void __launch_bounds__(ATT_T) attn_bwd_dkv_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                             const T* __restrict__ v, const float* __restrict__ null_kv,
                                                             const T* __restrict__ go, const float* __restrict__ lse,
                                                             const float* __restrict__ delta, T* __restrict__ dk,
                                                             T* __restrict__ dv, float* __restrict__ dnull, AttP p) {
  __shared__ float qs[ATT_TILE][D + 1], gs[ATT_TILE][D + 1], ls_[ATT_TILE], ds_[ATT_TILE];
  const int bh = blockIdx.y, b = bh / p.heads, h = bh % p.heads;
  const int j = blockIdx.x * ATT_T + threadIdx.x;
  const int nkt = p.nk + p.has_null;
  const bool live = j < nkt;
  const bool is_null = p.has_null && j == 0;
  const float lsc = p.mode == 1 ? 2.f * p.scale : p.scale;
  float kr[D], vr[D], dka[D], dva[D];
  float ksq = 0.f;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    float kv = 0.f, vv = 0.f;
    if (live) {
      if (is_null) { kv = null_kv[h * D + c]; vv = null_kv[(p.heads + h) * D + c]; }
      else {
        long row = (long)b * p.nk + (j - p.has_null);
        kv = ldf(k + row * p.k_rs + h * D + c);
        vv = ldf(v + row * p.v_rs + h * D + c);
      }
    }
    kr[c] = kv; vr[c] = vv; dka[c] = 0.f; dva[c] = 0.f;
    ksq += kv * kv;
  }
  const float kbias = p.mode == 1 ? -p.scale * ksq : 0.f;
  float dssum = 0.f;
  for (int i0 = 0; i0 < p.nq; i0 += ATT_TILE) {
    __syncthreads();
    for (int e = threadIdx.x; e < ATT_TILE * D; e += ATT_T) {
      int ii = e / D, c = e % D, i = i0 + ii;
      float qv = 0.f, gv = 0.f;
      if (i < p.nq) {
        qv = ldf(q + ((long)b * p.nq + i) * p.q_rs + h * D + c);
        gv = ldf(go + ((long)b * p.nq + i) * p.o_rs + h * D + c);
      }
      qs[ii][c] = qv; gs[ii][c] = gv;
    }
    if (threadIdx.x < ATT_TILE) {
      int i = i0 + threadIdx.x;
      ls_[threadIdx.x] = i < p.nq ? lse[(long)bh * p.nq + i] : INFINITY;
      ds_[threadIdx.x] = i < p.nq ? delta[(long)bh * p.nq + i] : 0.f;
    }
    __syncthreads();
    int in = min(ATT_TILE, p.nq - i0);
    for (int ii = 0; ii < in; ++ii) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < D; ++c) { s = fmaf(qs[ii][c], kr[c], s); dp = fmaf(gs[ii][c], vr[c], dp); }
      float pj = __expf(s * lsc + kbias - ls_[ii]);
      float ds = pj * (dp - ds_[ii]) * lsc;
      dssum += ds;
#pragma unroll
      for (int c = 0; c < D; ++c) { dva[c] = fmaf(pj, gs[ii][c], dva[c]); dka[c] = fmaf(ds, qs[ii][c], dka[c]); }
    }
  }
  if (!live) return;
  if (p.mode == 1) {      // d/dk of -(|k|^2) * scale :  -2*scale*k*sum_i dS  ==  -lsc * k * sum_i dS/lsc*... (ds already has lsc)
#pragma unroll
    for (int c = 0; c < D; ++c) dka[c] -= dssum * kr[c];
  }
  if (is_null) {
#pragma unroll
    for (int c = 0; c < D; ++c) {
      atomicAdd(dnull + h * D + c, dka[c]);
      atomicAdd(dnull + (p.heads + h) * D + c, dva[c]);
    }
  } else {
    long row = (long)b * p.nk + (j - p.has_null);
#pragma unroll
    for (int c = 0; c < D; ++c) {
      stf(dk + row * p.k_rs + h * D + c, dka[c]);
      stf(dv + row * p.v_rs + h * D + c, dva[c]);
    }
  }
}

static AttP make_attp(int B, int heads, int nq, int nk, int d, long q_rs, long k_rs, long v_rs, long o_rs, float scale,
                      int mode, int has_null) {
  AttP p;
  p.B = B; p.heads = heads; p.nq = nq; p.nk = nk; p.d = d; p.has_null = has_null; p.mode = mode;
  p.q_rs = q_rs; p.k_rs = k_rs; p.v_rs = v_rs; p.o_rs = o_rs; p.scale = scale;
  return p;
}

#define ATT_DISPATCH_D(d, ...)                                        \
  switch (d) {                                                        \
    case 8: { constexpr int D = 8; __VA_ARGS__; break; }              \
    case 16: { constexpr int D = 16; __VA_ARGS__; break; }            \
    case 32: { constexpr int D = 32; __VA_ARGS__; break; }            \
    case 64: { constexpr int D = 64; __VA_ARGS__; break; }            \
    default: return gg_fail("attention dim_head %d unsupported (8/16/32/64)", d); \
  }

int ggi_attn_fwd(const void* q, const void* k, const void* v, const float* null_kv, void* o, float* lse, int B, int heads,
                 int nq, int nk, int d, long q_rs, long k_rs, long v_rs, long o_rs, float scale, int mode, int dtype,
                 cudaStream_t st) {
  AttP p = make_attp(B, heads, nq, nk, d, q_rs, k_rs, v_rs, o_rs, scale, mode, null_kv != nullptr);
  dim3 grid(gg_cdiv(nq, ATT_T), B * heads);
  GG_DISPATCH(dtype, ATT_DISPATCH_D(d, (attn_fwd_kernel<T, D><<<grid, ATT_T, 0, st>>>((const T*)q, (const T*)k, (const T*)v, null_kv, (T*)o, lse, p))));
  return gg_check_launch("attn_fwd");
}

int ggi_attn_bwd(const void* q, const void* k, const void* v, const float* null_kv, const void* o, const void* go,
                 const float* lse, void* dq, void* dk, void* dv, float* dnull_kv, float* delta_ws, int B, int heads,
                 int nq, int nk, int d, long q_rs, long k_rs, long v_rs, long o_rs, float scale, int mode, int dtype,
                 cudaStream_t st) {
  int has_null = null_kv != nullptr;
  AttP p = make_attp(B, heads, nq, nk, d, q_rs, k_rs, v_rs, o_rs, scale, mode, has_null);
  if (has_null) cudaMemsetAsync(dnull_kv, 0, sizeof(float) * 2 * heads * d, st);
  dim3 g1(gg_cdiv(nq, ATT_T), B * heads), g2(gg_cdiv(nk + has_null, ATT_T), B * heads);
  GG_DISPATCH(dtype, ATT_DISPATCH_D(d, {
    attn_bwd_dq_kernel<T, D><<<g1, ATT_T, 0, st>>>((const T*)q, (const T*)k, (const T*)v, null_kv, (const T*)o, (const T*)go, lse, (T*)dq, delta_ws, p);
    attn_bwd_dkv_kernel<T, D><<<g2, ATT_T, 0, st>>>((const T*)q, (const T*)k, (const T*)v, null_kv, (const T*)go, lse, delta_ws, (T*)dk, (T*)dv, dnull_kv, p);
  }));
  return gg_check_launch("attn_bwd");
}
