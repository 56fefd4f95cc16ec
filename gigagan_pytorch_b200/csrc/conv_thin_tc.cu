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
// * loaders (8 warps, one ring row each) fill the ring with 16-byte cp.async (global side: one contiguous run of
//   130*Cin*2 bytes per row), completion -> fence.proxy.async -> mbarrier; warp 1 issues the MMAs (M=128, N=Cout, K=16) and releases ring rows
//   with tcgen05.commit; warps 2..5 drain the double-buffered TMEM accumulator (bias / LeakyReLU / residual / gain).
// Replaces the same reference lines as conv_tc.cu for the 128^2 / 256^2 blocks (gigagan_pytorch.py:402-409, :1608-1620).
#include "tc_common.cuh"

#define TH_SLOTS 8                       // barrier slots; p.slots (<= 8) ring rows are in use
#define TH_MAXACC 8                      // TMEM accumulator ring (one output row each): hides the MMA <-> epilogue handoff
#define TH_PIX 132                       // pixel slots per ring row (128 + 2 halo, padded to a multiple of 4)
#define TH_PLANE (TH_PIX * 16)           // bytes between channel octets of one ring row
#define TH_LOADW 8                       // loader warps (6..13); warp w owns ring slot w (w < p.slots)
#define TH_THREADS (192 + 32 * TH_LOADW)

// optional timeline trace of CTA 0 (debug tool tools/trace_thin.py): (tag, clock64) pairs appended to a global buffer
__device__ unsigned long long* g_th_trace = nullptr;
__device__ unsigned int g_th_trace_n = 0;
__device__ __forceinline__ void th_trace(int role, int row, int stage) {
  // slot = (role, row mod 64, stage): plain stores, no atomics (a trace point costs a clock read and one store)
  if (g_th_trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0)
    g_th_trace[((role & 31) * 64 + (row & 63)) * 8 + (stage & 7)] = (unsigned long long)clock64();
}
int ggi_debug_thin_trace(unsigned long long* buf) {
  unsigned int zero = 0;
  cudaMemcpyToSymbol(g_th_trace_n, &zero, sizeof(zero));
  cudaMemcpyToSymbol(g_th_trace, &buf, sizeof(buf));
  return 0;
}

struct ThP {
  int N, H, W, Cin, Cout, K, per_sample, act;
  int strips, total_rows, rows_per_cta;
  int planes, slot_bytes, tap_bytes, w_bytes, w_slots, slots, nacc;
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

template <int K>
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
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * TH_SLOTS + 4 + TH_MAXACC + a); };
  uint32_t* tmem_slot = (uint32_t*)(gen_base + p.slots * p.slot_bytes + p.w_slots * p.w_bytes + 8 * (2 * TH_SLOTS + 4 + 2 * TH_MAXACC));
  float* bias_sm = (float*)(tmem_slot + 4);                   // Cout <= 64 floats, filled before the role split

  const int warp = tc_warp_idx(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TH_SLOTS; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(wfull_bar(s), 1); mbar_init(wempty_bar(s), 1); }
    for (int a = 0; a < TH_MAXACC; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (bias && threadIdx.x >= 64 && threadIdx.x < 64 + p.Cout) bias_sm[threadIdx.x - 64] = bias[threadIdx.x - 64];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int r_begin = blockIdx.x * p.rows_per_cta;
  const int r_end = min(p.total_rows, r_begin + p.rows_per_cta);
  constexpr int taps = K * K;

  if (warp >= 6) {
    // ===================================================== loaders: input rows (and filters) -> shared memory
    // every loader warp owns one ring slot (entry c -> warp c % slots): it only ever stalls on its own row, so
    // `slots` rows are in flight per CTA
    const int lw = warp - 6;
    int c = 0, cs = 0, wcount = 0, prev_n = -1;
    for (int r = r_begin; r < r_end;) {
      const int yy = r % p.H, t = r / p.H;
      const int strip = t % p.strips, n = t / p.strips;
      const int seg = min(r_end - r, p.H - yy);
      const int x0 = strip * 128;
      if (wcount == 0 || (p.per_sample && n != prev_n)) {
        if (lw == 0) {
          const int ws = p.per_sample ? (wcount & 1) : 0;
          mbar_wait(wempty_bar(ws), (uint32_t)(((wcount >> 1) & 1) ^ 1));
          const bf16* wsrc = w + (p.per_sample ? (long)n * p.Cout * taps * p.Cin : 0L);
          const uint32_t wdst = wbase + ws * p.w_bytes;
          const int pieces = p.Cout * taps * p.planes;
          for (int id = lane; id < pieces; id += 32) {
            int c8 = id % p.planes, q = id / p.planes;
            int tap = q % taps, co = q / taps;
            th_cp16(wdst + tap * p.tap_bytes + c8 * (p.Cout * 16) + co * 16, wsrc + ((long)co * taps + tap) * p.Cin + c8 * 8, 16);
          }
          th_commit_group();
          th_wait_group<0>();
          th_fence_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(wfull_bar(ws));
        }
        ++wcount; prev_n = n;
      }
      const int iy0 = K == 3 ? yy - 1 : yy;
      const int nrows = K == 3 ? seg + 2 : seg;
      for (int i = 0; i < nrows; ++i, ++c) {
        const int s = cs; if (++cs == p.slots) cs = 0;
        if (s != lw) continue;                        // one owner warp per ring slot: its parity waits stay in sequence
        const int iy = iy0 + i;
        th_trace(10 + lw, c, 0);
        mbar_wait(empty_bar(s), (uint32_t)(((c / p.slots) & 1) ^ 1));
        th_trace(10 + lw, c, 1);
        const bool rowok = iy >= 0 && iy < p.H;
        const bf16* src_row = x + ((long)n * p.H + (rowok ? iy : 0)) * p.W * p.Cin;
        const uint32_t dst_row = base + s * p.slot_bytes;
        // consecutive lanes take consecutive 16-byte pieces of the row (pixel-major, channel octet minor): a warp
        // instruction covers 512 contiguous bytes, every 32-byte sector is requested once.  (One pixel per lane with an
        // inner octet loop touched half a sector per request: ncu counted 2.3x the input bytes on the L2 -> SM path.)
        const int pieces = 130 * p.planes;
        for (int id = lane; id < pieces; id += 32) {
          const int j = id / p.planes, c8 = id - j * p.planes;
          const int xx = x0 - 1 + j;
          const bool ok = rowok && xx >= 0 && xx < p.W;
          th_cp16(dst_row + c8 * TH_PLANE + j * 16, ok ? (const void*)(src_row + (long)xx * p.Cin + c8 * 8) : (const void*)x, ok ? 16 : 0);
        }
        th_commit_group();
        th_trace(10 + lw, c, 2);
        th_wait_group<0>();
        th_trace(10 + lw, c, 3);
        th_fence_async();
        th_trace(10 + lw, c, 4);
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar(s));
      }
      r += seg;
    }
  } else if (warp == 1) {
    const uint32_t el = tc_elect_one();            // the lane that issues tcgen05.mma / commit (warp stays convergent)
    // ===================================================== MMA issuer
    // The whole instruction stream of this role runs on ONE lane, so it is kept lean: descriptors are a constant
    // template plus a 16-byte-unit address (adding kx pixels = +kx), ring slots advance by increment-and-wrap.
    int c = 0, waited = 0, wslot = 0, wphase = 0, wcount = 0, prev_n = -1, cur_ws = 0;
    int acc = 0; uint32_t acc_phase = 0;
    int s0 = 0;                                              // ring slot of entry c + i
    const int ksteps = p.Cin >> 4;
    const uint64_t dA0 = th_desc(0, TH_PLANE, 128), dB0 = th_desc(0, p.Cout * 16, 128);
    const uint32_t a_kstep = (2 * TH_PLANE) >> 4, b_kstep = (uint32_t)(2 * p.Cout), b_tap = (uint32_t)p.tap_bytes >> 4;
    const uint32_t slot16 = (uint32_t)p.slot_bytes >> 4, base16 = base >> 4;
    for (int r = r_begin; r < r_end;) {
      const int yy = r % p.H, t = r / p.H;
      const int n = t / p.strips;
      const int seg = min(r_end - r, p.H - yy);
      if (wcount == 0 || (p.per_sample && n != prev_n)) {
        cur_ws = p.per_sample ? (wcount & 1) : 0;
        mbar_wait(wfull_bar(cur_ws), (uint32_t)((wcount >> 1) & 1));
        ++wcount; prev_n = n;
      }
      const uint64_t dB = dB0 + (uint64_t)((wbase + cur_ws * p.w_bytes) >> 4);
      const int rnext = r + seg;
      const bool last_use = p.per_sample && (rnext >= r_end || (rnext / p.H) / p.strips != n);
      for (int i = 0; i < seg; ++i) {
        const int need = c + i + (K == 3 ? 2 : 0);
        th_trace(1, r + i, 0);
        while (waited <= need) {
          mbar_wait(full_bar(wslot), (uint32_t)wphase);
          ++waited;
          if (++wslot == p.slots) { wslot = 0; wphase ^= 1; }
        }
        th_trace(1, r + i, 1);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        th_trace(1, r + i, 2);
        tc_fence_after();
        int s1 = s0 + 1 == p.slots ? 0 : s0 + 1;
        int s2 = s1 + 1 == p.slots ? 0 : s1 + 1;
        {   // convergent issue: every lane computes the (warp-uniform) descriptors, the elected lane issues
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.Cout);
          uint32_t accum = 0;
#pragma unroll
          for (int ky = 0; ky < K; ++ky) {
            const int sl = ky == 0 ? s0 : ky == 1 ? s1 : s2;
            const uint64_t da_row = dA0 + (uint64_t)(base16 + (uint32_t)sl * slot16 + (K == 1 ? 1u : 0u));
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
              uint64_t da = da_row + (uint64_t)kx;
              uint64_t db = dB + (uint64_t)((ky * K + kx) * b_tap);
              for (int ks = 0; ks < ksteps; ++ks) {
                tc_mma_f16_el(d_tmem, da, db, p.idesc, accum, el);
                accum = 1;
                da += a_kstep; db += b_kstep;
              }
            }
          }
          tc_commit_el(empty_bar(s0), el);
          if (i == seg - 1) {
            if (K == 3) { tc_commit_el(empty_bar(s1), el); tc_commit_el(empty_bar(s2), el); }
            if (last_use) tc_commit_el(wempty_bar(cur_ws), el);
          }
          tc_commit_el(tfull_bar(acc), el);
        }
        th_trace(1, r + i, 3);
        __syncwarp();
        s0 = s1;
        if (++acc == p.nacc) { acc = 0; acc_phase ^= 1u; }
      }
      if (K == 3) { s0 = s0 + 2 >= p.slots ? s0 + 2 - p.slots : s0 + 2; }
      c += K == 3 ? seg + 2 : seg;
      r += seg;
    }
  } else if (warp >= 2) {
    // ===================================================== epilogue (warps 2..5 -> TMEM lane quarter warp % 4)
    const int q = warp & 3;
    const int m = q * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;
    int yy = r_begin % p.H, t = r_begin / p.H;
    for (int r = r_begin; r < r_end; ++r) {
      const int strip = t % p.strips, n = t / p.strips;
      const long pix = (((long)n * p.H + yy) * p.W + strip * 128 + m) * p.Cout;
      if (++yy == p.H) { yy = 0; ++t; }
      th_trace(2 + q, r, 0);
      mbar_wait(tfull_bar(acc), acc_phase);
      th_trace(2 + q, r, 1);
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
          for (int j = 0; j < 16; j += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(bias_sm + c0 + j);
            v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
          }
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
      th_trace(2 + q, r, 2);
      if (++acc == p.nacc) { acc = 0; acc_phase ^= 1u; }
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
  size_t smem = 128 + (size_t)p.slots * p.slot_bytes + (size_t)p.w_slots * p.w_bytes + 8 * (2 * TH_SLOTS + 4 + 2 * TH_MAXACC) + 16 + 64 * 4;
  if (smem > 200 * 1024) return 1;
  int per_sm = smem <= 100 * 1024 ? 2 : 1;
  int grid = tc_num_sms() * per_sm;
  int min_rows = 8;                                  // amortise the 2 halo rows (and the filter load) of a range
  if ((long)grid * min_rows > p.total_rows) grid = (p.total_rows + min_rows - 1) / min_rows;
  p.rows_per_cta = (p.total_rows + grid - 1) / grid;
  grid = (p.total_rows + p.rows_per_cta - 1) / p.rows_per_cta;
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Cout >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  uint32_t cols = 512 / per_sm;                      // accumulator ring: as many one-row stages as this CTA's TMEM share holds
  p.nacc = (int)cols / Cout;
  if (p.nacc > TH_MAXACC) p.nacc = TH_MAXACC;
  cols = 32;
  while (cols < (uint32_t)(p.nacc * Cout)) cols <<= 1;
  p.tmem_cols = cols;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(conv_thin_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_thin_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  if (KH == 1) conv_thin_tc_kernel<1><<<grid, TH_THREADS, smem, st>>>(p, (const bf16*)x, (const bf16*)w, bias, (const bf16*)res, (bf16*)y);
  else conv_thin_tc_kernel<3><<<grid, TH_THREADS, smem, st>>>(p, (const bf16*)x, (const bf16*)w, bias, (const bf16*)res, (bf16*)y);
  return gg_check_launch("conv_thin_tc");
}

// =================================================================================================
// Weight gradient of the thin layers:  dW[co][tap][ci] = sum_pixels dY[p, co] * X[p + tap, ci]
//   D[128 (co, rows >= Cout unused) x Cin_blk] (TMEM, one accumulator per tap) += A[co x 16 pixels] * B[ci x 16 pixels]^T
// The SAME pixel-linear ring rows serve as MN-major operands with the pixel axis as K (K 8-groups packed: LBO = 128 B,
// channel octets SBO = PLANE apart), so dY and X are again read from global memory once; a tap is the X descriptor of
// ring row y+ky-1 advanced by kx*16 bytes.  Each CTA keeps its 9 accumulators in TMEM over its whole row range and
// flushes them with fp32 red.add (per image when the filters are per-sample).  A reads 16 channel octets (M = 128):
// octets beyond Cout/8 alias whatever follows in shared memory - those accumulator rows are never read.
#define TW_GPLANE (128 * 16)

struct TwP {
  int N, H, W, Cin, Cout, K, per_sample;
  int strips, total_rows, rows_per_cta, ranges;
  int cin_blk, ci_blocks, xplanes, gplanes, xslot, gslot, slots;
  uint32_t idesc, tmem_cols;
};

__device__ __forceinline__ uint64_t th_desc_mn(uint32_t addr, uint32_t kgroup_bytes, uint32_t octet_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(kgroup_bytes >> 4) << 16;       // MN-major, no swizzle: LBO = stride between 8-pixel (K) groups
  d |= (uint64_t)(octet_bytes >> 4) << 32;        // SBO = stride between 8-channel (MN) groups
  d |= (uint64_t)1 << 46;
  return d;
}

__global__ void __launch_bounds__(TH_THREADS, 2)
conv_thin_wgrad_tc_kernel(const TwP p, const bf16* __restrict__ x, const bf16* __restrict__ gy, float* __restrict__ dw) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 127u) & ~127u;
  uint8_t* gen_base = smem_raw + (base - raw);
  const uint32_t gbase = base + p.slots * p.xslot;
  const uint32_t ring_end = gbase + (p.slots - 1) * p.gslot + 16 * TW_GPLANE;      // incl. the over-read window of the last slot
  const uint32_t bars = ring_end;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (TH_SLOTS + s); };
  const uint32_t tfull_bar = bars + 8u * (2 * TH_SLOTS), tempty_bar = bars + 8u * (2 * TH_SLOTS + 1);
  uint32_t* tmem_slot = (uint32_t*)(gen_base + (ring_end - base) + 8 * (2 * TH_SLOTS + 2));

  const int warp = tc_warp_idx(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TH_SLOTS; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1); mbar_init(tempty_bar, 4);
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

  const int range = blockIdx.x % p.ranges, cib = blockIdx.x / p.ranges;
  const int r_begin = range * p.rows_per_cta;
  const int r_end = min(p.total_rows, r_begin + p.rows_per_cta);
  const int taps = p.K * p.K, lead = p.K - 1, padk = (p.K - 1) / 2;
  // a flush (accumulators -> dW) ends the CTA's range, and every image when the filters are per-sample
  auto flush_after = [&](int r, int seg, int n) {
    int rn = r + seg;
    return rn >= r_end || (p.per_sample && (rn / p.H) / p.strips != n);
  };

  if (warp >= 6) {
    // ===================================================== loaders (entry c -> warp c % slots)
    const int lw = warp - 6;
    int c = 0, cs = 0;
    for (int r = r_begin; r < r_end;) {
      const int yy = r % p.H, t = r / p.H;
      const int strip = t % p.strips, n = t / p.strips;
      const int seg = min(r_end - r, p.H - yy);
      const int x0 = strip * 128;
      const int nent = seg + lead;
      for (int i = 0; i < nent; ++i, ++c) {
        const int s = cs; if (++cs == p.slots) cs = 0;
        if (s != lw) continue;                        // one owner warp per ring slot: its parity waits stay in sequence
        mbar_wait(empty_bar(s), (uint32_t)(((c / p.slots) & 1) ^ 1));
        if (i < seg + 2 * padk) {                                   // X row yy - pad + i (130 pixels with the halo)
          const int iy = yy - padk + i;
          const bool rowok = iy >= 0 && iy < p.H;
          const bf16* src_row = x + ((long)n * p.H + (rowok ? iy : 0)) * p.W * p.Cin + cib * p.cin_blk;
          const uint32_t dst_row = base + s * p.xslot;
          const int pieces = 130 * p.xplanes;
          for (int id = lane; id < pieces; id += 32) {
            int j = id / p.xplanes, c8 = id - j * p.xplanes;
            int xx = x0 - 1 + j;
            bool ok = rowok && xx >= 0 && xx < p.W;
            th_cp16(dst_row + c8 * TH_PLANE + j * 16, ok ? (const void*)(src_row + (long)xx * p.Cin + c8 * 8) : (const void*)x, ok ? 16 : 0);
          }
        }
        if (i >= lead) {                                            // dY row yy + i - lead (128 pixels)
          const bf16* src_row = gy + (((long)n * p.H + (yy + i - lead)) * p.W + x0) * p.Cout;
          const uint32_t dst_row = gbase + s * p.gslot;
          const int pieces = 128 * p.gplanes;
          for (int id = lane; id < pieces; id += 32) {
            int j = id / p.gplanes, c8 = id - j * p.gplanes;
            th_cp16(dst_row + c8 * TW_GPLANE + j * 16, src_row + (long)j * p.Cout + c8 * 8, 16);
          }
        }
        th_commit_group();
        th_wait_group<0>();
        th_fence_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar(s));
      }
      r += seg;
    }
  } else if (warp == 1) {
    const uint32_t el = tc_elect_one();            // the lane that issues tcgen05.mma / commit (warp stays convergent)
    // ===================================================== MMA issuer (one lane: lean descriptor arithmetic, see fprop)
    int c = 0, waited = 0, wslot = 0, wphase = 0, s0 = 0;
    uint32_t tphase = 0;
    bool fresh = true;                                 // accumulators hold nothing yet (first MMA of a tap overwrites)
    const uint64_t dA0 = th_desc_mn(0, 128, TW_GPLANE), dB0 = th_desc_mn(0, 128, TH_PLANE);
    const uint32_t xslot16 = (uint32_t)p.xslot >> 4, gslot16 = (uint32_t)p.gslot >> 4, base16 = base >> 4, gbase16 = gbase >> 4;
    for (int r = r_begin; r < r_end;) {
      const int yy = r % p.H, t = r / p.H;
      const int n = t / p.strips;
      const int seg = min(r_end - r, p.H - yy);
      const bool flush = flush_after(r, seg, n);
      for (int j = 0; j < seg; ++j) {
        const int need = c + j + lead;
        while (waited <= need) {
          mbar_wait(full_bar(wslot), (uint32_t)wphase);
          ++waited;
          if (++wslot == p.slots) { wslot = 0; wphase ^= 1; }
        }
        if (fresh) { mbar_wait(tempty_bar, tphase ^ 1u); }
        tc_fence_after();
        const int s1 = s0 + 1 == p.slots ? 0 : s0 + 1;
        const int s2 = s1 + 1 == p.slots ? 0 : s1 + 1;
        const int sg = p.K == 3 ? s2 : s0;              // dY of this output row travels with entry c + j + lead
        {     // (the loaders fence generic->async proxy before they arrive: no fence on this serial path)
          const uint64_t da_row = dA0 + (uint64_t)(gbase16 + (uint32_t)sg * gslot16);
          const uint32_t first = fresh ? 0u : 1u;
          for (int ky = 0; ky < p.K; ++ky) {
            const int sl = ky == 0 ? s0 : ky == 1 ? s1 : s2;
            const uint64_t db_row = dB0 + (uint64_t)(base16 + (uint32_t)sl * xslot16 + (p.K == 1 ? 1u : 0u));
            for (int kx = 0; kx < p.K; ++kx) {
              const uint32_t d_tmem = tmem_base + (uint32_t)((ky * p.K + kx) * p.cin_blk);
              uint64_t da = da_row, db = db_row + (uint64_t)kx;
              tc_mma_f16_el(d_tmem, da, db, p.idesc, first, el);
#pragma unroll
              for (int ks = 1; ks < 8; ++ks) {
                da += 16; db += 16;                     // 16 pixels x 16 bytes per K step
                tc_mma_f16_el(d_tmem, da, db, p.idesc, 1u, el);
              }
            }
          }
          tc_commit_el(empty_bar(s0), el);
          if (j == seg - 1) {
            if (p.K == 3) { tc_commit_el(empty_bar(s1), el); tc_commit_el(empty_bar(s2), el); }
            if (flush) tc_commit_el(tfull_bar, el);
          }
        }
        __syncwarp();
        fresh = false;
        s0 = s1;
      }
      if (flush) { fresh = true; tphase ^= 1u; }
      if (p.K == 3) { s0 = s0 + 2 >= p.slots ? s0 + 2 - p.slots : s0 + 2; }
      c += seg + lead;
      r += seg;
    }
  } else if (warp >= 2) {
    // ===================================================== epilogue: accumulators -> dW (fp32 red.add), once per flush
    const int q = warp & 3;
    const int co = q * 32 + lane;
    uint32_t tphase = 0;
    for (int r = r_begin; r < r_end;) {
      const int yy = r % p.H, t = r / p.H;
      const int n = t / p.strips;
      const int seg = min(r_end - r, p.H - yy);
      if (flush_after(r, seg, n)) {
        mbar_wait(tfull_bar, tphase);
        tc_fence_after();
        if (q * 32 < p.Cout) {
          float* dst = dw + (((long)(p.per_sample ? n : 0) * p.Cout + co) * taps) * p.Cin + cib * p.cin_blk;
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
          for (int tap = 0; tap < taps; ++tap) {
            for (int c0 = 0; c0 < p.cin_blk; c0 += 16) {
              uint32_t rr[16];
              tc_ld16(taddr + tap * p.cin_blk + c0, rr);
              if (co < p.Cout) {
#pragma unroll
                for (int k = 0; k < 16; k += 4)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + (long)tap * p.Cin + c0 + k),
                               "f"(__uint_as_float(rr[k])), "f"(__uint_as_float(rr[k + 1])), "f"(__uint_as_float(rr[k + 2])),
                               "f"(__uint_as_float(rr[k + 3])) : "memory");
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar);
        tphase ^= 1u;
      }
      r += seg;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// dw ([N,]Cout,KH,KW,Cin fp32) is zero-filled here (memset on the stream) and accumulated with red.add
int ggi_tc_conv_thin_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int OH, int OW, int Cout,
                           int KH, int KW, int stride, int pad, int per_sample_w, cudaStream_t st) {
  if (stride != 1 || KH != KW || !(KH == 1 || KH == 3) || pad != (KH - 1) / 2 || OH != H || OW != W) return 1;
  if (W < 128 || W % 128 || Cin % 16 || Cin > 64 || Cout % 16 || Cout > 64 || Cin < 16 || Cout < 16) return 1;
  if (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dw) & 15) return 1;
  if ((long)N * (W / 128) * H > (1L << 30)) return 1;
  TwP p;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.K = KH; p.per_sample = per_sample_w;
  p.strips = W / 128; p.total_rows = N * p.strips * H;
  p.cin_blk = Cin <= 32 ? Cin : (Cin % 32 == 0 ? 32 : 16);
  p.ci_blocks = Cin / p.cin_blk;
  p.xplanes = p.cin_blk / 8; p.gplanes = Cout / 8;
  p.xslot = p.xplanes * TH_PLANE; p.gslot = p.gplanes * TW_GPLANE;
  p.slots = TH_SLOTS;
  auto smem_for = [&](int slots) {
    return (size_t)128 + (size_t)slots * p.xslot + (size_t)(slots - 1) * p.gslot + 16 * TW_GPLANE + 8 * (2 * TH_SLOTS + 2) + 16;
  };
  if (smem_for(p.slots) > 200 * 1024) p.slots = 6;
  size_t smem = smem_for(p.slots);
  if (smem > 200 * 1024) return 1;
  uint32_t cols = 32;
  while (cols < (uint32_t)(KH * KW * p.cin_blk)) cols <<= 1;
  if (cols > 512) return 1;
  p.tmem_cols = cols;
  int per_sm = (smem <= 100 * 1024 && cols <= 256) ? 2 : 1;
  int ctas = tc_num_sms() * per_sm;
  int ranges = ctas / p.ci_blocks;
  if (ranges < 1) ranges = 1;
  int min_rows = 16;                                 // amortise the halo rows and the 9-accumulator flush
  if ((long)ranges * min_rows > p.total_rows) ranges = (p.total_rows + min_rows - 1) / min_rows;
  p.rows_per_cta = (p.total_rows + ranges - 1) / ranges;
  p.ranges = (p.total_rows + p.rows_per_cta - 1) / p.rows_per_cta;
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.cin_blk >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)(per_sample_w ? N : 1) * Cout * KH * KW * Cin, st);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(conv_thin_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  conv_thin_wgrad_tc_kernel<<<p.ranges * p.ci_blocks, TH_THREADS, smem, st>>>(p, (const bf16*)x, (const bf16*)dy, dw);
  return gg_check_launch("conv_thin_wgrad_tc");
}
