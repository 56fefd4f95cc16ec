// tcgen05 convolution for the THIN high-resolution layers (128^2 / 256^2 maps, <= 64 channels; 3x3 or 1x1, stride 1).
//
// These layers are HBM-bound (about 10 GFLOP over 100 MB), but the generic implicit-GEMM kernel (conv_tc.cu) re-reads
// every input pixel once per filter tap through TMA boxes with 32-64 byte rows and re-loads the filter for every tile:
// it ran at ~0.5 TB/s.  Here every input row is staged in shared memory exactly ONCE and all nine taps are fed from it:
//
// * one M tile = 128 consecutive pixels of one output row; a persistent CTA walks DOWN a 128-pixel-wide strip, so each
//   new output row needs one new input row (130 pixels with the halo) in an 8-row ring; the filter stays resident.
// * the row buffer is the NO-SWIZZLE canonical UMMA layout with 8-row groups packed back to back (SBO = 128 B):
//   [channel octet c8][pixel slot j][8 channels] -> row r of the A operand lives at start + 16*r + c8*PLANE, i.e. the
//   operand is PIXEL-LINEAR and the tap (ky, kx) is simply the descriptor of ring row y+ky-1 with its start address
//   advanced by kx*16 bytes.  No im2col, no boundary code (halo pixels / rows are zero-filled by cp.async src-size 0).
// * loaders (4 warps) fill the ring with 16-byte cp.async (global side: one contiguous run of 130*Cin*2 bytes per row),
//   completion -> fence.proxy.async -> mbarrier; warp 1 issues the MMAs (M=128, N=Cout, K=16) and releases ring rows
//   with tcgen05.commit; warps 2..5 drain the double-buffered TMEM accumulator (bias / LeakyReLU / residual / gain).
// Replaces the same reference lines as conv_tc.cu for the 128^2 / 256^2 blocks (gigagan_pytorch.py:402-409, :1608-1620).
#include "tc_common.cuh"

#define TH_SLOTS 8                       // barrier slots; p.slots (<= 8) ring rows are in use
#define TH_PIX 132                       // pixel slots per ring row (128 + 2 halo, padded to a multiple of 4)
#define TH_PLANE (TH_PIX * 16)           // bytes between channel octets of one ring row
#define TH_THREADS 320
#define TH_LOADERS 128
#define TH_LAG 3                         // cp.async groups (ring rows) in flight per loader thread

struct ThP {
  int N, H, W, Cin, Cout, K, per_sample, act;
  int strips, total_rows, rows_per_cta;
  int planes, slot_bytes, tap_bytes, w_bytes, w_slots, slots;
  float gain;
  uint32_t idesc, tmem_cols;
};

__device__ __forceinline__ uint64_t th_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;          // K-major, no swizzle: stride between the two 8-channel octets of a K step
  d |= (uint64_t)(sbo_bytes >> 4) << 32;          // stride between 8-row groups (128 B = packed -> pixel-linear rows)
  d |= (uint64_t)1 << 46;
  return d;                                        // layout_type 0 = no swizzle
}
__device__ __forceinline__ void th_cp16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void th_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void th_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }
__device__ __forceinline__ void th_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(TH_THREADS, 2)
conv_thin_tc_kernel(const ThP p, const bf16* __restrict__ x, const bf16* __restrict__ w, const float* __restrict__ bias,
                    const bf16* __restrict__ res, bf16* __restrict__ y) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 127u) & ~127u;
  uint8_t* gen_base = smem_raw + (base - raw);
  const uint32_t wbase = base + p.slots * p.slot_bytes;
  const uint32_t bars = wbase + p.w_slots * p.w_bytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (TH_SLOTS + s); };
  auto wfull_bar = [&](int s) { return bars + 8u * (2 * TH_SLOTS + s); };
  auto wempty_bar = [&](int s) { return bars + 8u * (2 * TH_SLOTS + 2 + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * TH_SLOTS + 4 + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * TH_SLOTS + 6 + a); };
  uint32_t* tmem_slot = (uint32_t*)(gen_base + p.slots * p.slot_bytes + p.w_slots * p.w_bytes + 8 * (2 * TH_SLOTS + 8));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TH_SLOTS; ++s) { mbar_init(full_bar(s), TH_LOADERS / 32); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(wfull_bar(s), TH_LOADERS / 32); mbar_init(wempty_bar(s), 1);
      mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int r_begin = blockIdx.x * p.rows_per_cta;
  const int r_end = min(p.total_rows, r_begin + p.rows_per_cta);
  const int taps = p.K * p.K;

  if (warp >= 6) {
    // ===================================================== loaders: input rows (and filters) -> shared memory
    const int lt = threadIdx.x - 192;
    uint32_t pend[TH_LAG + 1];
    int head = 0, npend = 0;
    int c = 0, wcount = 0, prev_n = -1;
    auto retire_oldest = [&]() {                    // caller has made the oldest group complete (wait_group)
      th_fence_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(pend[head]);
      head = (head + 1) % (TH_LAG + 1);
      --npend;
    };
    for (int r = r_begin; r < r_end;) {
      const int yy = r % p.H, t = r / p.H;
      const int strip = t % p.strips, n = t / p.strips;
      const int seg = min(r_end - r, p.H - yy);
      const int x0 = strip * 128;
      if (wcount == 0 || (p.per_sample && n != prev_n)) {
        const int ws = p.per_sample ? (wcount & 1) : 0;
        mbar_wait(wempty_bar(ws), (uint32_t)(((wcount >> 1) & 1) ^ 1));
        const bf16* wsrc = w + (p.per_sample ? (long)n * p.Cout * taps * p.Cin : 0L);
        const uint32_t wdst = wbase + ws * p.w_bytes;
        const int pieces = p.Cout * taps * p.planes;
        for (int id = lt; id < pieces; id += TH_LOADERS) {
          int c8 = id % p.planes, q = id / p.planes;
          int tap = q % taps, co = q / taps;
          th_cp16(wdst + tap * p.tap_bytes + c8 * (p.Cout * 16) + co * 16, wsrc + ((long)co * taps + tap) * p.Cin + c8 * 8, 16);
        }
        th_commit_group();
        th_wait_group<0>();                         // a filter switch drains the row pipeline too (rare)
        while (npend > 0) retire_oldest();
        th_fence_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(wfull_bar(ws));
        ++wcount; prev_n = n;
      }
      const int iy0 = p.K == 3 ? yy - 1 : yy;
      const int nrows = p.K == 3 ? seg + 2 : seg;
      for (int i = 0; i < nrows; ++i, ++c) {
        const int iy = iy0 + i, s = c % p.slots;
        mbar_wait(empty_bar(s), (uint32_t)(((c / p.slots) & 1) ^ 1));
        const bool rowok = iy >= 0 && iy < p.H;
        const bf16* src_row = x + ((long)n * p.H + (rowok ? iy : 0)) * p.W * p.Cin;
        const uint32_t dst_row = base + s * p.slot_bytes;
        const int pieces = 130 * p.planes;
        for (int id = lt; id < pieces; id += TH_LOADERS) {
          int j = id / p.planes, c8 = id - j * p.planes;
          int xx = x0 - 1 + j;
          bool ok = rowok && xx >= 0 && xx < p.W;
          th_cp16(dst_row + c8 * TH_PLANE + j * 16, ok ? (const void*)(src_row + (long)xx * p.Cin + c8 * 8) : (const void*)x, ok ? 16 : 0);
        }
        th_commit_group();
        pend[(head + npend) % (TH_LAG + 1)] = full_bar(s);
        ++npend;
        if (npend > TH_LAG) { th_wait_group<TH_LAG>(); retire_oldest(); }
      }
      r += seg;
    }
    th_wait_group<0>();
    while (npend > 0) retire_oldest();
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    int c = 0, waited = 0, wcount = 0, prev_n = -1, cur_ws = 0;
    int acc = 0; uint32_t acc_phase = 0;
    const int ksteps = p.Cin >> 4;
    for (int r = r_begin; r < r_end;) {
      const int yy = r % p.H, t = r / p.H;
      const int n = t / p.strips;
      const int seg = min(r_end - r, p.H - yy);
      if (wcount == 0 || (p.per_sample && n != prev_n)) {
        cur_ws = p.per_sample ? (wcount & 1) : 0;
        mbar_wait(wfull_bar(cur_ws), (uint32_t)((wcount >> 1) & 1));
        ++wcount; prev_n = n;
      }
      const int rnext = r + seg;
      const bool last_use = p.per_sample && (rnext >= r_end || (rnext / p.H) / p.strips != n);
      for (int i = 0; i < seg; ++i) {
        const int need = c + i + (p.K == 3 ? 2 : 0);
        while (waited <= need) { mbar_wait(full_bar(waited % p.slots), (uint32_t)((waited / p.slots) & 1)); ++waited; }
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        if (lane == 0) {
          th_fence_async();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.Cout);
          uint32_t accum = 0;
          for (int ky = 0; ky < p.K; ++ky) {
            const uint32_t arow = base + ((c + i + ky) % p.slots) * p.slot_bytes;
            for (int kx = 0; kx < p.K; ++kx) {
              const uint32_t a0 = arow + (uint32_t)(kx + (p.K == 1 ? 1 : 0)) * 16u;
              const uint32_t b0 = wbase + cur_ws * p.w_bytes + (ky * p.K + kx) * p.tap_bytes;
              for (int ks = 0; ks < ksteps; ++ks) {
                uint64_t da = th_desc(a0 + ks * 2 * TH_PLANE, TH_PLANE, 128);
                uint64_t db = th_desc(b0 + ks * 2 * (p.Cout * 16), p.Cout * 16, 128);
                tc_mma_f16(d_tmem, da, db, p.idesc, accum);
                accum = 1;
              }
            }
          }
          tc_commit(empty_bar((c + i) % p.slots));
          if (i == seg - 1) {
            if (p.K == 3) { tc_commit(empty_bar((c + i + 1) % p.slots)); tc_commit(empty_bar((c + i + 2) % p.slots)); }
            if (last_use) tc_commit(wempty_bar(cur_ws));
          }
          tc_commit(tfull_bar(acc));
        }
        __syncwarp();
        acc ^= 1; if (acc == 0) acc_phase ^= 1u;
      }
      c += p.K == 3 ? seg + 2 : seg;
      r += seg;
    }
  } else if (warp >= 2) {
    // ===================================================== epilogue (warps 2..5 -> TMEM lane quarter warp % 4)
    const int q = warp & 3;
    const int m = q * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;
    for (int r = r_begin; r < r_end; ++r) {
      const int yy = r % p.H, t = r / p.H;
      const int strip = t % p.strips, n = t / p.strips;
      const long pix = (((long)n * p.H + yy) * p.W + strip * 128 + m) * p.Cout;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.Cout);
      for (int c0 = 0; c0 < p.Cout; c0 += 16) {
        uint32_t rr[16];
        tc_ld16(taddr + c0, rr);
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rr[j]);
        if (bias) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += __ldg(bias + c0 + j);
        }
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
        }
        if (res) {
          uint4 r0, r1;
          ld_global_256(res + pix + c0, r0, r1);
          const bf16* rb0 = (const bf16*)&r0; const bf16* rb1 = (const bf16*)&r1;
#pragma unroll
          for (int j = 0; j < 8; ++j) { v[j] += __bfloat162float(rb0[j]); v[8 + j] += __bfloat162float(rb1[j]); }
        }
        uint4 o0, o1;
        __nv_bfloat162* ob0 = (__nv_bfloat162*)&o0; __nv_bfloat162* ob1 = (__nv_bfloat162*)&o1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ob0[j] = __floats2bfloat162_rn(v[2 * j] * p.gain, v[2 * j + 1] * p.gain);
          ob1[j] = __floats2bfloat162_rn(v[8 + 2 * j] * p.gain, v[8 + 2 * j + 1] * p.gain);
        }
        st_global_256(y + pix + c0, o0, o1);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1; if (acc == 0) acc_phase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// returns 1 if the shape is not a thin layer (the caller falls through to the generic tensor-core kernel)
int ggi_tc_conv_thin(const void* x, const void* w, const float* bias, const void* res, void* y, int N, int H, int W, int Cin,
                     int OH, int OW, int Cout, int KH, int KW, int stride, int pad, int per_sample_w, int act, float gain,
                     cudaStream_t st) {
  if (stride != 1 || KH != KW || !(KH == 1 || KH == 3) || pad != (KH - 1) / 2 || OH != H || OW != W) return 1;
  if (W < 128 || W % 128 || Cin % 16 || Cin > 64 || Cout % 16 || Cout > 64 || Cin < 16 || Cout < 16) return 1;
  if (((uintptr_t)x | (uintptr_t)w) & 15) return 1;
  if (((uintptr_t)y | (uintptr_t)res) & 31) return 1;
  if ((long)N * (W / 128) * H > (1L << 30)) return 1;
  ThP p;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.K = KH; p.per_sample = per_sample_w; p.act = act; p.gain = gain;
  p.strips = W / 128; p.total_rows = N * p.strips * H;
  p.planes = Cin / 8; p.slot_bytes = p.planes * TH_PLANE;
  p.tap_bytes = Cin * Cout * 2; p.w_bytes = KH * KW * p.tap_bytes; p.w_slots = per_sample_w ? 2 : 1;
  p.slots = Cin > 32 ? 6 : TH_SLOTS;
  size_t smem = 128 + (size_t)p.slots * p.slot_bytes + (size_t)p.w_slots * p.w_bytes + 8 * (2 * TH_SLOTS + 8) + 16;
  if (smem > 200 * 1024) return 1;
  int per_sm = smem <= 100 * 1024 ? 2 : 1;
  int grid = tc_num_sms() * per_sm;
  int min_rows = 8;                                  // amortise the 2 halo rows (and the filter load) of a range
  if ((long)grid * min_rows > p.total_rows) grid = (p.total_rows + min_rows - 1) / min_rows;
  p.rows_per_cta = (p.total_rows + grid - 1) / grid;
  grid = (p.total_rows + p.rows_per_cta - 1) / p.rows_per_cta;
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Cout >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * Cout)) cols <<= 1;
  p.tmem_cols = cols;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(conv_thin_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  conv_thin_tc_kernel<<<grid, TH_THREADS, smem, st>>>(p, (const bf16*)x, (const bf16*)w, bias, (const bf16*)res, (bf16*)y);
  return gg_check_launch("conv_thin_tc");
}
