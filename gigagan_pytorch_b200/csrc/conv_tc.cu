// tcgen05 / TMA implicit-GEMM convolution for NHWC bf16 (sm_100a only).
//
//   D[128 pixels x Ntile] (fp32, TMEM)  +=  A_tap[128 pixels x chunk] (smem, K-major)  *  W_tap[Ntile x chunk]^T
//
// * one M tile = a (Wt x Ht x Nt) box of output pixels; for every filter tap the SAME TMA tensor map is read at
//   the box origin shifted by (kx - pad, ky - pad): the hardware's out-of-bounds zero fill is the padding, so there
//   is no im2col buffer and no boundary code.  stride-2 convolutions use one map per tap over a strided view.
// * weights are the kernel-layout tensor [Cout][taps][Cin] (optionally per sample) read by a second map.
// * warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane, tcgen05.mma cta_group::1 kind::f16, M=128),
//   warps 2..5 = epilogue (tcgen05.ld -> bias / LeakyReLU / residual / gain -> bf16 -> global).  smem ring of
//   `stages` (A,B) slabs with full/empty mbarriers; two TMEM accumulator stages so the epilogue of tile i overlaps
//   the MMAs of tile i+1; persistent CTAs stride over tiles (N-tile fastest so neighbours share the A tile in L2).
// Replaces: cuDNN conv under nn.Conv2d / F.conv2d, gigagan_pytorch.py:402-409, :1454-1470, :1608-1620, :1656.
#include "tc_common.cuh"

struct TcP {
  int N, OH, OW, Cout;
  int Wt, Ht, Nt, tiles_w, tiles_h, tiles_n, n_tiles_n, total_tiles;
  int Ntile, chunk, cchunks, KH, KW, pad, stride, per_sample;
  int stages, stage_bytes, a_bytes, b_bytes, b_slab, ts;   // ts = filter taps packed into one pipeline stage
  int mdual, m_tiles;                                       // mdual = 2: two 128-pixel M tiles share every weight slab
  int act;
  float gain;
  long y_off, y_sn, y_sh, y_sw;          // output addressing (elements): y_off + n*y_sn + oy*y_sh + ox*y_sw + co
  uint32_t idesc, sbo, layout_type, tmem_cols;
};

// ------------------------------------------------------------------------------------------------ kernel
#define CONV_THREADS 320          // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_fprop_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                     const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmA3,
                     const __grid_constant__ CUtensorMap tmB, const TcP p, const float* __restrict__ bias,
                     const bf16* __restrict__ res, bf16* __restrict__ y) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                  // SWIZZLE_128B slabs need 1024 B alignment
  uint8_t* gen_base = smem_raw + (base - raw);
  const uint32_t bars = base + p.stages * p.stage_bytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (TC_MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * TC_MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * TC_MAX_STAGES + 2 + a); };
  uint32_t* tmem_slot = (uint32_t*)(gen_base + p.stages * p.stage_bytes + 8 * (2 * TC_MAX_STAGES + 4));

  const int warp = tc_warp_idx(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
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
  const int kblocks = (p.KH * p.KW / p.ts) * p.cchunks;

  if (warp == 0) {
    // ===================================================== TMA producer
    {
      const uint32_t el = tc_elect_one();          // convergent producer: only the TMA / expect_tx instructions are predicated
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int nt = tile % p.n_tiles_n, mp = tile / p.n_tiles_n;
        int co0 = nt * p.Ntile;
        int ox0[2], oy0[2], n0[2];
        for (int d = 0; d < p.mdual; ++d) {
          int mt = mp * p.mdual + d;
          int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
          ox0[d] = tw * p.Wt; oy0[d] = th * p.Ht; n0[d] = tn * p.Nt;          // past the last tile: n0 >= N -> zero fill
        }
        const int nA = p.ts * p.mdual;                      // A slabs per stage (mdual == 2 implies ts == 1)
        for (int kb = 0; kb < kblocks; ++kb) {
          int tap0 = (kb / p.cchunks) * p.ts, cc = kb % p.cchunks;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx_el(full_bar(stage), (uint32_t)(nA * p.a_bytes + p.ts * p.b_bytes), el);
          uint32_t sa0 = base + stage * p.stage_bytes, sb0 = sa0 + nA * p.a_bytes;
          for (int t = 0; t < p.ts; ++t) {
            int tap = tap0 + t;
            uint32_t sb = sb0 + t * p.b_slab;
            for (int d = 0; d < p.mdual; ++d) {
              uint32_t sa = sa0 + (t * p.mdual + d) * p.a_bytes;
              if (p.stride == 1) {
                int ky = tap / p.KW, kx = tap - ky * p.KW;
                tma_load_4d_el(sa, &tmA0, full_bar(stage), cc * p.chunk, ox0[d] + kx - p.pad, oy0[d] + ky - p.pad, n0[d], el);
              } else {
                const CUtensorMap* m = tap == 0 ? &tmA0 : tap == 1 ? &tmA1 : tap == 2 ? &tmA2 : &tmA3;
                tma_load_4d_el(sa, m, full_bar(stage), cc * p.chunk, ox0[d], oy0[d], n0[d], el);
              }
            }
            tma_load_4d_el(sb, &tmB, full_bar(stage), cc * p.chunk, tap, co0, p.per_sample ? n0[0] : 0, el);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t el = tc_elect_one();            // the lane that issues tcgen05.mma / commit (warp stays convergent)
    // ===================================================== MMA issuer
    int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
    const int nacc = p.mdual == 2 ? 1 : 2;                 // dual-M uses all 512 TMEM columns for one tile pair
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.Ntile * p.mdual);
      const int nA = p.ts * p.mdual;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        {
          uint32_t sa0 = base + stage * p.stage_bytes, sb0 = sa0 + nA * p.a_bytes;
          int ksteps = p.chunk >> 4;
          for (int t = 0; t < p.ts; ++t) {
            uint64_t db = make_smem_desc(sb0 + t * p.b_slab, p.sbo, p.layout_type);
            for (int d = 0; d < p.mdual; ++d) {
              uint64_t da = make_smem_desc(sa0 + (t * p.mdual + d) * p.a_bytes, p.sbo, p.layout_type);
              for (int k = 0; k < ksteps; ++k)
                tc_mma_f16_el(d_tmem + d * p.Ntile, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), p.idesc, (kb | t | k) != 0 ? 1u : 0u, el);
            }
          }
          tc_commit_el(empty_bar(stage), el);
          if (kb == kblocks - 1) tc_commit_el(tfull_bar(acc), el);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      if (++acc == nacc) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ===================================================== epilogue (warps 2..9 -> TMEM lane quarter warp % 4; warps 2..5
    // take the first half of the tile's columns, warps 6..9 the second: two warps per SM sub-partition so that one's
    // tcgen05.ld latency overlaps the other's conversions / stores; 32 columns per TMEM load)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    const int ww = m % p.Wt, hh = (m / p.Wt) % p.Ht, nn = m / (p.Wt * p.Ht);
    const int cspan = p.Ntile >= 32 ? p.Ntile / 2 : p.Ntile;       // columns per warp group (Ntile 16: group 0 only)
    const int cbeg = half * cspan, cend = (p.Ntile >= 32 || half == 0) ? cbeg + cspan : cbeg;
    int acc = 0; uint32_t acc_phase = 0;
    const int nacc = p.mdual == 2 ? 1 : 2;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int nt = tile % p.n_tiles_n, mp = tile / p.n_tiles_n;
      int co0 = nt * p.Ntile;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
     for (int d = 0; d < p.mdual; ++d) {
      int mt = mp * p.mdual + d;
      int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
      int n = tn * p.Nt + nn;
      long pix = p.y_off + (long)n * p.y_sn + (long)(th * p.Ht + hh) * p.y_sh + (long)(tw * p.Wt + ww) * p.y_sw;
      bool live = n < p.N && mt < p.m_tiles;
      uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.Ntile * p.mdual + d * p.Ntile);
      for (int c0 = cbeg; c0 < cend; c0 += 32) {
        const bool two = c0 + 16 < cend;                 // warp-uniform: 32 columns, or a 16-column tail
        uint32_t r[32];
        if (two) tc_ld32(taddr + c0, r); else tc_ld16(taddr + c0, r);
        if (live) {
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            if (h2 == 1 && !two) break;
            const int cc = c0 + 16 * h2;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[16 * h2 + j]);
            if (bias) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] += __ldg(bias + co0 + cc + j);
            }
            if (p.act == 1) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
            }
            bf16* dst = y + pix + co0 + cc;
            if (res) {
              uint4 r0, r1;
              ld_global_256(res + pix + co0 + cc, r0, r1);
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
            st_global_256(dst, o0, o1);
          }
        }
      }
     }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == nacc) { acc = 0; acc_phase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qr) == cudaSuccess) fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

int tc_make_map4(CUtensorMap* m, const void* ptr, const uint64_t dims[4], const uint64_t strides_bytes[3],
                     const uint32_t box[4], int swizzle_bytes) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return gg_fail("cuTensorMapEncodeTiled unavailable");
  cuuint64_t gd[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gs[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                        : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return gg_fail("cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

static int g_num_sms = 0;
int tc_num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

// returns 1 if the shape is not eligible for the tensor-core path
int ggi_tc_conv_fprop(const void* x, const void* w, const float* bias, const void* res, void* y, int N, int H, int W,
                      int Cin, int OH, int OW, int Cout, int KH, int KW, int stride, int pad, int per_sample_w, int act,
                      float gain, const long* ystr, cudaStream_t st) {
  // ---- eligibility
  int chunk = Cin >= 64 ? 64 : Cin;
  if (!(chunk == 16 || chunk == 32 || chunk == 64) || Cin % chunk) return 1;
  if (Cout % 16 || Cout < 16) return 1;
  int Ntile = Cout > 256 ? 256 : Cout;
  if (Cout % Ntile || (Ntile & (Ntile - 1))) return 1;
  if (stride == 1) { if (OH != H + 2 * pad - KH + 1 || OW != W + 2 * pad - KW + 1) return 1; }
  else if (stride == 2) { if (pad != 0 || KH != KW || KH > 2 || H != 2 * OH || W != 2 * OW) return 1; }
  else return 1;
  if ((OW & (OW - 1)) || (OH & (OH - 1)) || OW < 2 || OH < 2) return 1;
  int Wt = OW < 16 ? OW : 16;
  int Ht = OH < 128 / Wt ? OH : 128 / Wt;
  int Nt = 128 / (Wt * Ht);
  if (Wt * Ht * Nt != 128) return 1;
  if (per_sample_w && Nt != 1) return 1;
  if (((uintptr_t)x | (uintptr_t)w) & 15) return 1;
  if (((uintptr_t)y | (uintptr_t)res) & 31) return 1;          // 256-bit epilogue accesses

  TcP p;
  p.N = N; p.OH = OH; p.OW = OW; p.Cout = Cout;
  p.Wt = Wt; p.Ht = Ht; p.Nt = Nt;
  p.tiles_w = OW / Wt; p.tiles_h = OH / Ht; p.tiles_n = (N + Nt - 1) / Nt;
  p.Ntile = Ntile; p.n_tiles_n = Cout / Ntile;
  p.m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  // under-filled grids (low-resolution layers: few pixel tiles): halve the N tile while that still fits one wave, so
  // that twice as many SMs share the layer (n256 4x4 512->512 ran 64 tiles on 148 SMs)
  while (Ntile > 32 && (long)p.m_tiles * (Cout / Ntile) * 2 <= tc_num_sms()) Ntile /= 2;
  p.Ntile = Ntile; p.n_tiles_n = Cout / Ntile;
  // L2 -> SM operand traffic bounds the 128x256 tile (85 flop/B): when the layer is wide and deep, let two pixel
  // tiles share every weight slab (256x256 effective tile, 128 flop/B), using all 512 TMEM columns
  // (only for long K loops - the epilogue is not overlapped in this mode - and when >= 2 full waves of tile pairs remain)
  // ... or when the pair schedule simply needs fewer waves: a pair tile takes ~1.6x a single tile (measured 1.25x the
  // throughput), so pairs win whenever waves(pairs) * 1.6 < waves(singles)
  {
    const int nsm = tc_num_sms();
    const long singles = (long)p.m_tiles * p.n_tiles_n, pairs = (long)((p.m_tiles + 1) / 2) * p.n_tiles_n;
    const long w1 = (singles + nsm - 1) / nsm, w2 = (pairs + nsm - 1) / nsm;
    const bool ok = Ntile == 256 && KH * KW * (Cin / chunk) >= 32 && !per_sample_w && p.m_tiles >= 2;
    p.mdual = (ok && (pairs >= 2 * nsm || w2 * 16 < w1 * 10)) ? 2 : 1;
  }
  p.total_tiles = ((p.m_tiles + p.mdual - 1) / p.mdual) * p.n_tiles_n;
  p.chunk = chunk; p.cchunks = Cin / chunk; p.KH = KH; p.KW = KW; p.pad = pad; p.stride = stride;
  p.per_sample = per_sample_w;
  p.a_bytes = 128 * chunk * 2; p.b_bytes = Ntile * chunk * 2;
  p.b_slab = (p.b_bytes + 1023) / 1024 * 1024;
  // small-channel layers: pack several filter taps into one stage so a stage carries enough bytes / MMAs to hide
  // the TMA + mbarrier round trip (the 16/32-channel 128^2 / 256^2 layers were latency-bound at 1 tap per stage)
  int taps = KH * KW, ts = 1;
  if (stride == 1 && p.cchunks == 1) {
    const int cand[4] = {taps, 7, 3, 1};
    for (int ci_ = 0; ci_ < 4; ++ci_) {
      int c = cand[ci_];
      if (c >= 1 && taps % c == 0 && c * (p.a_bytes + p.b_slab) <= 64 * 1024) { ts = c; break; }
    }
  }
  if (p.mdual == 2) ts = 1;
  p.ts = ts;
  p.stage_bytes = ts * (p.mdual * p.a_bytes + p.b_slab);
  int stages = (200 * 1024) / p.stage_bytes;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages < 2) return 1;
  p.stages = stages;
  p.act = act; p.gain = gain;
  if (ystr) {
    if ((ystr[0] | ystr[1] | ystr[2] | ystr[3]) & 15) return 1;
    p.y_off = ystr[0]; p.y_sn = ystr[1]; p.y_sh = ystr[2]; p.y_sw = ystr[3];
  } else { p.y_off = 0; p.y_sn = (long)OH * OW * Cout; p.y_sh = (long)OW * Cout; p.y_sw = Cout; }
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Ntile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.sbo = 8 * chunk * 2;                                   // 8 rows of (chunk*2) bytes
  p.layout_type = chunk == 64 ? 2u : chunk == 32 ? 4u : 6u;
  uint32_t cols = 2 * Ntile;
  p.tmem_cols = cols < 32 ? 32 : cols;

  // ---- tensor maps
  CUtensorMap tmA[4], tmB;
  int nmaps = stride == 1 ? 1 : KH * KW;
  for (int t = 0; t < 4; ++t) {
    int tt = t < nmaps ? t : 0;
    const bf16* basep = (const bf16*)x;
    uint64_t dims[4], strides[3];
    if (stride == 1) {
      dims[0] = Cin; dims[1] = W; dims[2] = H; dims[3] = N;
      strides[0] = (uint64_t)Cin * 2; strides[1] = (uint64_t)W * Cin * 2; strides[2] = (uint64_t)H * W * Cin * 2;
    } else {
      int ky = tt / KW, kx = tt % KW;
      basep += ((long)ky * W + kx) * Cin;
      dims[0] = Cin; dims[1] = OW; dims[2] = OH; dims[3] = N;
      strides[0] = (uint64_t)2 * Cin * 2; strides[1] = (uint64_t)2 * W * Cin * 2; strides[2] = (uint64_t)H * W * Cin * 2;
    }
    uint32_t box[4] = {(uint32_t)chunk, (uint32_t)Wt, (uint32_t)Ht, (uint32_t)Nt};
    if (tc_make_map4(&tmA[t], basep, dims, strides, box, chunk * 2)) return -1;
  }
  {
    int taps = KH * KW;
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)taps, (uint64_t)Cout, (uint64_t)(per_sample_w ? N : 1)};
    uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)taps * Cin * 2, (uint64_t)Cout * taps * Cin * 2};
    uint32_t box[4] = {(uint32_t)chunk, 1, (uint32_t)Ntile, 1};
    if (tc_make_map4(&tmB, w, dims, strides, box, chunk * 2)) return -1;
  }
  size_t smem = 1024 + (size_t)p.stages * p.stage_bytes + 8 * (2 * TC_MAX_STAGES + 4) + 16;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(conv_fprop_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  int grid = p.total_tiles < tc_num_sms() ? p.total_tiles : tc_num_sms();
  conv_fprop_tc_kernel<<<grid, CONV_THREADS, smem, st>>>(tmA[0], tmA[1], tmA[2], tmA[3], tmB, p, bias, (const bf16*)res, (bf16*)y);
  return gg_check_launch("conv_fprop_tc");
}

// =================================================================================================
// Weight gradient on tcgen05:  dW[co, tap, ci] += sum_pixels dY[p, co] * X[p + tap, ci]
//   D[128 co x Ntile ci] (TMEM) += A[co x 16 pixels] * B[ci x 16 pixels]^T, both operands MN-major: the TMA boxes
//   (64 channels x Pw x Ph x Pn pixels, 128 B rows, SWIZZLE_128B) are consumed directly with the pixel axis as K.
//   Work item = (co block, ci block, tap, [image], pixel-range split); partial sums are reduced with fp32 red.add.
// =================================================================================================
#define WG_PIX_MAX 256   // pixels per pipeline stage: 64 (4 UMMA K-steps) ... 256 (16 K-steps)

struct WgP {
  int N, OH, OW, Cout, Cin, taps, KW, pad, stride;
  int Wt, Ht, Nt, tiles_w, tiles_h, tiles_n, pix_tiles;   // pixel tiling of the OUTPUT grid
  int co_blocks, ci_blocks, Ntile, nsub_a, nsub_b, splits, per_sample, tiles_per_image, total_items;
  int stages, stage_bytes, a_bytes, b_bytes, pix;
  uint32_t idesc, tmem_cols;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmX1,
                     const __grid_constant__ CUtensorMap tmX2, const __grid_constant__ CUtensorMap tmX3,
                     const __grid_constant__ CUtensorMap tmDY, const WgP p, float* __restrict__ dw) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - raw);
  const uint32_t bars = base + p.stages * p.stage_bytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (TC_MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * TC_MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * TC_MAX_STAGES + 2 + a); };
  uint32_t* tmem_slot = (uint32_t*)(gen_base + p.stages * p.stage_bytes + 8 * (2 * TC_MAX_STAGES + 4));
  const int warp = tc_warp_idx(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
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

  // item -> (split, image, tap, ci block, co block);  pixel tiles [t0, t1) of the image (or of the whole batch)
  auto decode = [&](int item, int& cob, int& cib, int& tap, int& img, int& t0, int& t1) {
    int sp = item % p.splits; item /= p.splits;
    cib = item % p.ci_blocks; item /= p.ci_blocks;
    cob = item % p.co_blocks; item /= p.co_blocks;
    tap = item % p.taps; item /= p.taps;
    img = item;                                         // 0 unless per_sample
    int ntiles = p.per_sample ? p.tiles_per_image : p.pix_tiles;
    int per = (ntiles + p.splits - 1) / p.splits;
    t0 = sp * per; t1 = min(ntiles, t0 + per);
    if (p.per_sample) { t0 += img * p.tiles_per_image; t1 += img * p.tiles_per_image; }
  };

  if (warp == 0) {
    {
      const uint32_t el = tc_elect_one();          // convergent producer: only the TMA / expect_tx instructions are predicated
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        int cob, cib, tap, img, t0, t1;
        decode(item, cob, cib, tap, img, t0, t1);
        int ky = tap / p.KW, kx = tap - ky * p.KW;
        for (int t = t0; t < t1; ++t) {
          int tw = t % p.tiles_w, th = (t / p.tiles_w) % p.tiles_h, tn = t / (p.tiles_w * p.tiles_h);
          int ox0 = tw * p.Wt, oy0 = th * p.Ht, n0 = tn * p.Nt;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx_el(full_bar(stage), (uint32_t)(p.a_bytes + p.b_bytes), el);
          uint32_t sa = base + stage * p.stage_bytes, sb = sa + p.a_bytes;
          for (int s = 0; s < p.nsub_a; ++s)
            tma_load_4d_el(sa + s * (p.pix * 128), &tmDY, full_bar(stage), cob * 128 + s * 64, ox0, oy0, n0, el);
          for (int s = 0; s < p.nsub_b; ++s) {
            int c0 = cib * p.Ntile + s * 64;
            if (p.stride == 1) tma_load_4d_el(sb + s * (p.pix * 128), &tmX0, full_bar(stage), c0, ox0 + kx - p.pad, oy0 + ky - p.pad, n0, el);
            else {
              const CUtensorMap* m = tap == 0 ? &tmX0 : tap == 1 ? &tmX1 : tap == 2 ? &tmX2 : &tmX3;
              tma_load_4d_el(sb + s * (p.pix * 128), m, full_bar(stage), c0, ox0, oy0, n0, el);
            }
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t el = tc_elect_one();            // the lane that issues tcgen05.mma / commit (warp stays convergent)
    int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
    const uint32_t lbo_a = p.nsub_a > 1 ? p.pix * 128 : 0, lbo_b = p.pix * 128;
    const int ksteps = p.pix / 16;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      int cob, cib, tap, img, t0, t1;
      decode(item, cob, cib, tap, img, t0, t1);
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.Ntile);
      for (int t = t0; t < t1; ++t) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        {
          uint32_t sa = base + stage * p.stage_bytes, sb = sa + p.a_bytes;
          for (int k = 0; k < ksteps; ++k) {
            uint64_t da = make_smem_desc_mn(sa + k * 2048, lbo_a, 1024), db = make_smem_desc_mn(sb + k * 2048, lbo_b, 1024);
            tc_mma_f16_el(d_tmem, da, db, p.idesc, (t > t0 || k > 0) ? 1u : 0u, el);
          }
          tc_commit_el(empty_bar(stage), el);
          if (t == t1 - 1) tc_commit_el(tfull_bar(acc), el);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      if (t1 > t0) { if (++acc == 2) { acc = 0; acc_phase ^= 1u; } }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      int cob, cib, tap, img, t0, t1;
      decode(item, cob, cib, tap, img, t0, t1);
      if (t1 <= t0) continue;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      int co = cob * 128 + row;
      bool live = co < p.Cout;
      float* dst = dw + (((long)img * p.Cout + co) * p.taps + tap) * p.Cin + cib * p.Ntile;
      uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.Ntile);
      for (int c0 = 0; c0 < p.Ntile; c0 += 16) {
        uint32_t r[16];
        tc_ld16(taddr + c0, r);
        if (live) {
          if (p.splits == 1) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                     __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + j), "f"(__uint_as_float(r[j])),
                           "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3])) : "memory");
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// dw must be zero-filled by the caller of this launcher (done here with a memset on the stream).
int ggi_tc_conv_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int OH, int OW, int Cout,
                      int KH, int KW, int stride, int pad, int per_sample_w, cudaStream_t st) {
  if (Cin % 16 || Cout % 16 || Cin < 16 || Cout < 16) return 1;
  if (Cin > 256 && Cin % 256) return 1;
  if (Cin > 64 && Cin % 64) return 1;
  if (Cout > 64 && Cout % 64) return 1;
  if (stride == 1) { if (OH != H + 2 * pad - KH + 1 || OW != W + 2 * pad - KW + 1) return 1; }
  else if (stride == 2) { if (pad != 0 || KH != KW || KH > 2 || H != 2 * OH || W != 2 * OW) return 1; }
  else return 1;
  if ((OW & (OW - 1)) || (OH & (OH - 1)) || OW < 2 || OH < 2) return 1;
  if (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dw) & 15) return 1;
  WgP p;
  p.N = N; p.OH = OH; p.OW = OW; p.Cout = Cout; p.Cin = Cin; p.taps = KH * KW; p.KW = KW; p.pad = pad; p.stride = stride;
  p.co_blocks = (Cout + 127) / 128;
  p.nsub_a = Cout > 64 ? 2 : 1;
  p.per_sample = per_sample_w;
  // Work decomposition: item = (co block, ci block of Ntile channels, tap, [image], pixel-range split).  The widest ci
  // block (256) is the most efficient MMA shape, but a layer with few pixels and few (co, ci, tap) items cannot be
  // split into enough CTAs along the pixel axis (every split pays a 128 x Ntile fp32 red.add epilogue worth ~60-100
  // pipeline stages): such layers (stride-2 / 1x1 convolutions of the 8^2-32^2 maps) narrow the ci block instead.
  int best_items = -1, base_items = 0, splits = 1, pix = 64, Wt = 0, Ht = 0, Nt = 0;
  for (int ntile = Cin > 256 ? 256 : Cin; ntile >= 64 || ntile == Cin; ntile /= 2) {
    if (Cin % ntile || (ntile > 64 && ntile % 64)) { if (ntile <= 64) break; continue; }
    int nsub_b_ = (ntile + 63) / 64;
    int pix_ = 64;                                    // bigger pixel slabs for thin layers: ~48-64 KB per stage
    while (pix_ < WG_PIX_MAX && (p.nsub_a + nsub_b_) * (pix_ * 2) * 128 <= 64 * 1024 && (long)OH * OW * (per_sample_w ? 1 : N) >= 4L * pix_ * 2) pix_ *= 2;
    int Wt_ = OW < 16 ? OW : 16;
    int Ht_ = OH < pix_ / Wt_ ? OH : pix_ / Wt_;
    int Nt_ = pix_ / (Wt_ * Ht_);
    if (Wt_ * Ht_ * Nt_ != pix_ || (per_sample_w && Nt_ != 1) || N % Nt_) { if (ntile <= 64) break; continue; }
    int base_ = p.co_blocks * (Cin / ntile) * p.taps * (per_sample_w ? N : 1);
    int ntiles_ = per_sample_w ? (OW / Wt_) * (OH / Ht_) : (OW / Wt_) * (OH / Ht_) * (N / Nt_);
    int splits_ = 1;
    int min_stages = ntile * 3 / 8 > 8 ? ntile * 3 / 8 : 8;          // epilogue cost scales with the accumulator width
    while (base_ * splits_ < 2 * tc_num_sms() && ntiles_ / (splits_ * 2) >= min_stages) splits_ *= 2;
    int items_ = base_ * splits_;
    if (items_ > best_items) {
      best_items = items_; p.Ntile = ntile; base_items = base_; splits = splits_; pix = pix_; Wt = Wt_; Ht = Ht_; Nt = Nt_;
    }
    if (items_ >= tc_num_sms() || ntile <= 64) break;                // enough CTAs: keep the widest block that fills the GPU
  }
  if (best_items < 0) return 1;
  p.Wt = Wt; p.Ht = Ht; p.Nt = Nt; p.tiles_w = OW / Wt; p.tiles_h = OH / Ht; p.tiles_n = N / Nt;
  p.pix_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  p.tiles_per_image = p.tiles_w * p.tiles_h;
  p.ci_blocks = Cin / p.Ntile;
  p.nsub_b = (p.Ntile + 63) / 64;
  p.splits = splits;
  p.total_items = base_items * splits;
  p.pix = pix;
  p.a_bytes = p.nsub_a * pix * 128; p.b_bytes = p.nsub_b * pix * 128;
  p.stage_bytes = p.a_bytes + p.b_bytes;
  int stages = (200 * 1024) / p.stage_bytes;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  p.stages = stages;
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.Ntile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  uint32_t cols = 2 * p.Ntile;
  p.tmem_cols = cols < 32 ? 32 : cols;

  CUtensorMap tmX[4], tmDY;
  int nmaps = stride == 1 ? 1 : KH * KW;
  for (int t = 0; t < 4; ++t) {
    int tt = t < nmaps ? t : 0;
    const bf16* basep = (const bf16*)x;
    uint64_t dims[4], strides[3];
    if (stride == 1) {
      dims[0] = Cin; dims[1] = W; dims[2] = H; dims[3] = N;
      strides[0] = (uint64_t)Cin * 2; strides[1] = (uint64_t)W * Cin * 2; strides[2] = (uint64_t)H * W * Cin * 2;
    } else {
      int ky = tt / KW, kx = tt % KW;
      basep += ((long)ky * W + kx) * Cin;
      dims[0] = Cin; dims[1] = OW; dims[2] = OH; dims[3] = N;
      strides[0] = (uint64_t)2 * Cin * 2; strides[1] = (uint64_t)2 * W * Cin * 2; strides[2] = (uint64_t)H * W * Cin * 2;
    }
    uint32_t box[4] = {64, (uint32_t)Wt, (uint32_t)Ht, (uint32_t)Nt};
    if (tc_make_map4(&tmX[t], basep, dims, strides, box, 128)) return -1;
  }
  {
    uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)OW, (uint64_t)OH, (uint64_t)N};
    uint64_t strides[3] = {(uint64_t)Cout * 2, (uint64_t)OW * Cout * 2, (uint64_t)OH * OW * Cout * 2};
    uint32_t box[4] = {64, (uint32_t)Wt, (uint32_t)Ht, (uint32_t)Nt};
    if (tc_make_map4(&tmDY, dy, dims, strides, box, 128)) return -1;
  }
  if (p.splits > 1) cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)(per_sample_w ? N : 1) * Cout * KH * KW * Cin, st);
  size_t smem = 1024 + (size_t)p.stages * p.stage_bytes + 8 * (2 * TC_MAX_STAGES + 4) + 16;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  int grid = p.total_items < tc_num_sms() ? p.total_items : tc_num_sms();
  conv_wgrad_tc_kernel<<<grid, TC_THREADS, smem, st>>>(tmX[0], tmX[1], tmX[2], tmX[3], tmDY, p, dw);
  return gg_check_launch("conv_wgrad_tc");
}
