// Small fused kernels that replace chains of tiny launches in the training step: the GAN hinge objective over all
// logit tensors (one launch forward, one backward), the squeeze-excite gate MLP.
#include "gg_internal.h"

#define GL_MAX 8
struct GanLossArgs {
  const void* x[GL_MAX];
  void* dx[GL_MAX];
  long n[GL_MAX], row[GL_MAX], split[GL_MAX];
  int dt[GL_MAX];
  int k, mode;
  float w_ms;
};

__device__ __forceinline__ float gl_load(const void* p, int dt, long i) {
  return dt == GG_BF16 ? __bfloat162float(((const bf16*)p)[i]) : ((const float*)p)[i];
}

// one block per logit tensor j: L_j added into out (zeroed by the launcher): out[0] = L_0, out[1] = sum_{j>=1} L_j,
// out[2] = L_0 + w_ms * out[1]   (a single block walking all tensors took 57 us; 32-bit index arithmetic)
__global__ void __launch_bounds__(1024) gan_loss_fwd_kernel(GanLossArgs a, float* __restrict__ out) {
  __shared__ float red[2][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x;
  float sr = 0.f, sf = 0.f;
  const long n = a.n[j];
  const unsigned row = (unsigned)a.row[j], split = (unsigned)a.split[j];
  for (long i0 = 0; i0 < n; i0 += 1u << 30) {                       // 32-bit modulo inside 2^30-element windows
    const unsigned cnt = (unsigned)min((long)(1u << 30), n - i0), off = (unsigned)(i0 % row);
    for (unsigned i = threadIdx.x; i < cnt; i += blockDim.x) {
      const float v = gl_load(a.x[j], a.dt[j], i0 + i);
      if (a.mode == 1) sr += v;
      else if ((off + i) % row < split) sr += fmaxf(1.f + v, 0.f);
      else sf += fmaxf(1.f - v, 0.f);
    }
  }
  sr = warp_sum(sr); sf = warp_sum(sf);
  if (lane == 0) { red[0][warp] = sr; red[1][warp] = sf; }
  __syncthreads();
  if (warp == 0) {
    sr = red[0][lane]; sf = red[1][lane];              // blockDim.x == 1024: all 32 slots valid
    sr = warp_sum(sr); sf = warp_sum(sf);
    if (lane == 0) {
      float l;
      if (a.mode == 1) l = sr / (float)n;
      else {
        const long nr = n / row * split;
        l = sr / (float)nr + sf / (float)(n - nr);
      }
      if (j == 0) { atomicAdd(out + 0, l); atomicAdd(out + 2, l); }
      else { atomicAdd(out + 1, l); atomicAdd(out + 2, a.w_ms * l); }
    }
  }
}

__global__ void gan_loss_bwd_kernel(GanLossArgs a, const float* __restrict__ gout) {
  const int j = blockIdx.y;
  const long n = a.n[j], row = a.row[j], split = a.split[j];
  const float g = gout[0] * (j == 0 ? 1.f : a.w_ms);
  const long nr = n / row * split;
  const float cr = g / (float)(a.mode == 1 ? n : nr), cf = -g / (float)(n - nr);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float d;
    if (a.mode == 1) d = cr;
    else {
      const float v = gl_load(a.x[j], a.dt[j], i);
      d = (i % row < split) ? (1.f + v > 0.f ? cr : 0.f) : (1.f - v > 0.f ? cf : 0.f);
    }
    if (a.dt[j] == GG_BF16) ((bf16*)a.dx[j])[i] = __float2bfloat16_rn(d);
    else ((float*)a.dx[j])[i] = d;
  }
}

static int gl_fill(GanLossArgs& a, const void* const* x, void* const* dx, const long* meta, int k, int mode, float w_ms) {
  if (k < 1 || k > GL_MAX) return gg_fail("gan_loss: 1..%d tensors (got %d)", GL_MAX, k);
  a.k = k; a.mode = mode; a.w_ms = w_ms;
  for (int j = 0; j < k; ++j) {
    a.x[j] = x[j]; a.dx[j] = dx ? dx[j] : nullptr;
    a.n[j] = meta[4 * j]; a.row[j] = meta[4 * j + 1]; a.split[j] = meta[4 * j + 2]; a.dt[j] = (int)meta[4 * j + 3];
    if (a.n[j] < 1 || a.row[j] < 1 || a.n[j] % a.row[j]) return gg_fail("gan_loss: tensor %d has %ld elements, rows of %ld", j, a.n[j], a.row[j]);
    if (mode == 0 && (a.split[j] < 1 || a.split[j] >= a.row[j])) return gg_fail("gan_loss: tensor %d: split %ld of row %ld", j, a.split[j], a.row[j]);
  }
  return 0;
}

int ggi_gan_loss_fwd(const void* const* x, const long* meta, int k, int mode, float w_ms, float* out, cudaStream_t st) {
  GanLossArgs a;
  if (int r = gl_fill(a, x, nullptr, meta, k, mode, w_ms)) return r;
  cudaMemsetAsync(out, 0, 3 * sizeof(float), st);
  gan_loss_fwd_kernel<<<k, 1024, 0, st>>>(a, out);
  return gg_check_launch("gan_loss_fwd");
}

int ggi_gan_loss_bwd(const void* const* x, void* const* dx, const long* meta, int k, int mode, float w_ms, const float* gout,
                     cudaStream_t st) {
  GanLossArgs a;
  if (int r = gl_fill(a, x, dx, meta, k, mode, w_ms)) return r;
  long mx = 0;
  for (int j = 0; j < k; ++j) mx = a.n[j] > mx ? a.n[j] : mx;
  dim3 grid(gg_blocks(mx, 256, 148 * 4), k);
  gan_loss_bwd_kernel<<<grid, 256, 0, st>>>(a, gout);
  return gg_check_launch("gan_loss_bwd");
}
