// Fused self-attention on tcgen05 (dim_head 64, bf16): S = Q K^T and O = P V on the tensor cores with TMEM
// accumulators, softmax in registers straight out of TMEM, P handed to the second MMA through swizzled shared memory.
// Learned null key/value (handled in registers), dot-product or shared-QK L2-distance logits
// ( -|q-k|^2 s == (2 q.k - |k|^2) s - |q|^2 s ; the last term is constant along the softmax axis and dropped ).
// Two passes over the key tiles instead of an online softmax: pass A only takes the row maximum of the logits
// (no exponentials), pass B recomputes S, forms P = exp2(t - max) exactly once per logit and accumulates O and the
// row sum with NO rescaling of the TMEM accumulator; the extra Q K^T costs 1/4 of the MMA work, the exponentials
// (the real bound: 16 MUFU/clk/SM) are issued once.
// Replaces gigagan_pytorch.py:562-592 (sim / attn never touch HBM) and attend.py:64-110.
#include "attn_tc_common.cuh"

// |k_j|^2 per (b, h, token): 8 lanes per 64-element row (one 16-byte load each), 4 rows per warp
__global__ void attn_ksq_kernel(const bf16* __restrict__ k, float* __restrict__ ksq, int B, int n, int heads, long k_rs) {
  const long total = (long)B * heads * n;
  const int sub = threadIdx.x & 7;
  for (long row = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 3; row < total; row += ((long)gridDim.x * blockDim.x) >> 3) {
    int j = (int)(row % n);                               // row = (b*heads + h)*n + j
    long bh = row / n;
    int h = (int)(bh % heads);
    long b = bh / heads;
    uint4 v = *reinterpret_cast<const uint4*>(k + (b * n + j) * k_rs + h * ATC_D + sub * 8);
    const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&v);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(hv[i]); s = fmaf(f.x, f.x, fmaf(f.y, f.y, s)); }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (sub == 0) ksq[row] = s;
  }
}

__global__ void __launch_bounds__(ATC_FWD_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AtcP p, const float* __restrict__ null_kv,
                   const float* __restrict__ ksq, bf16* __restrict__ o, float* __restrict__ lse2) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  // layout: Q 16K | K[2] 32K | V[2] 32K | P[2] 64K | ksq[2][128] f32 1K | null k,v 512B | barriers | tmem slot
  const uint32_t sQ = base, sK = base + 16384, sV = sK + 32768, sP = sV + 32768;
  float* ksq_sm = (float*)(gbase + 147456);
  float* null_sm = (float*)(gbase + 147456 + 1024);
  float* xchg = (float*)(gbase + 147456 + 1536);              // row max / row sum exchange between the two column halves
  const uint32_t bars = base + 147456 + 2560;
  enum { Q_FULL = 0, K_FULL = 1, K_EMPTY = 3, V_FULL = 5, V_EMPTY = 7, S_FULL = 9, S_EMPTY = 11, P_FULL = 13, P_EMPTY = 15, O_FULL = 17 };
  auto bar = [&](int i) { return bars + 8u * i; };
  uint32_t* tmem_slot = (uint32_t*)(gbase + 147456 + 2560 + 8 * 18);

  const int warp = tc_warp_idx(), lane = threadIdx.x & 31;
  const int qt = blockIdx.x % p.tiles, bh = blockIdx.x / p.tiles;
  const int b = bh / p.heads, h = bh % p.heads;
  const int T = p.tiles;

  if (threadIdx.x == 0) {
    mbar_init(bar(Q_FULL), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(K_FULL + i), 1); mbar_init(bar(K_EMPTY + i), 1);
      mbar_init(bar(V_FULL + i), 1); mbar_init(bar(V_EMPTY + i), 1);
      mbar_init(bar(S_FULL + i), 1); mbar_init(bar(S_EMPTY + i), 8);
      mbar_init(bar(P_FULL + i), 8); mbar_init(bar(P_EMPTY + i), 1);
    }
    mbar_init(bar(O_FULL), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x < 192 && p.has_null) {
    int t = threadIdx.x - 64;                       // 128 threads: k_null[64], v_null[64]
    null_sm[t] = t < 64 ? null_kv[h * ATC_D + t] : null_kv[(p.heads + h) * ATC_D + (t - 64)];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tO = tmem + 256;       // S[2] at columns 0 / 128, O at 256 (64 columns)
  const uint32_t idesc_qk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  if (warp == 0) {
    // ================================================= TMA producer
    {
      const uint32_t el = tc_elect_one();          // convergent producer: only the TMA / expect_tx instructions are predicated
      mbar_expect_tx_el(bar(Q_FULL), 16384, el);
      tma_load_4d_el(sQ, &tmQ, bar(Q_FULL), 0, qt * ATC_T, h, b, el);
      int kc = 0, vc = 0;
      for (int pass = 0; pass < 2; ++pass)
        for (int j = 0; j < T; ++j) {
          int s = kc & 1;
          mbar_wait(bar(K_EMPTY + s), ((kc >> 1) & 1) ^ 1u);
          mbar_expect_tx_el(bar(K_FULL + s), 16384, el);
          tma_load_4d_el(sK + s * 16384, &tmK, bar(K_FULL + s), 0, j * ATC_T, h, b, el);
          ++kc;
          if (pass == 1) {
            int sv = vc & 1;
            mbar_wait(bar(V_EMPTY + sv), ((vc >> 1) & 1) ^ 1u);
            mbar_expect_tx_el(bar(V_FULL + sv), 16384, el);
            tma_load_4d_el(sV + sv * 16384, &tmV, bar(V_FULL + sv), 0, j * ATC_T, h, b, el);
            ++vc;
          }
        }
    }
  } else if (warp == 1) {
    const uint32_t el = tc_elect_one();            // the lane that issues tcgen05.mma / commit (warp stays convergent)
    // ================================================= MMA issuer
    int kc = 0, sc = 0, pc = 0;
    auto issue_S = [&]() {
      int ks = kc & 1, ss = sc & 1;
      mbar_wait(bar(K_FULL + ks), (kc >> 1) & 1);
      mbar_wait(bar(S_EMPTY + ss), ((sc >> 1) & 1) ^ 1u);
      tc_fence_after();
      {
        uint64_t da = make_smem_desc(sQ, 1024, 2), db = make_smem_desc(sK + ks * 16384, 1024, 2);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_f16_el(tS + ss * 128, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_qk, k ? 1u : 0u, el);
        tc_commit_el(bar(K_EMPTY + ks), el);
        tc_commit_el(bar(S_FULL + ss), el);
      }
      __syncwarp();
      ++kc; ++sc;
    };
    mbar_wait(bar(Q_FULL), 0);
    for (int j = 0; j < T; ++j) issue_S();           // pass A: row maxima
    issue_S();                                        // pass B, S_0
    for (int j = 0; j < T; ++j) {
      if (j + 1 < T) issue_S();
      int ps = pc & 1;
      mbar_wait(bar(P_FULL + ps), (pc >> 1) & 1);
      mbar_wait(bar(V_FULL + ps), (pc >> 1) & 1);
      tc_fence_after();
      {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          uint64_t da = make_smem_desc(sP + ps * 32768 + (k >> 2) * 16384, 1024, 2) + (uint64_t)(2 * (k & 3));
          uint64_t db = make_smem_desc_mn(sV + ps * 16384 + k * 2048, 0, 1024);
          tc_mma_f16_el(tO, da, db, idesc_pv, (j | k) ? 1u : 0u, el);
        }
        tc_commit_el(bar(P_EMPTY + ps), el);
        tc_commit_el(bar(V_EMPTY + ps), el);
        if (j == T - 1) tc_commit_el(bar(O_FULL), el);
      }
      __syncwarp();
      ++pc;
    }
  } else {
    // ================================================= softmax / epilogue
    // 8 warps: warp % 4 selects the TMEM lane quarter (32 query rows), (warp - 2) / 4 the 64-column half of every
    // 128-key tile - two threads per query row, so the MUFU/FMA work of a row is issued from two schedulers.
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2;
    const int r = q * 32 + lane;                      // row inside the tile
    const int st = threadIdx.x - 64;                  // 0..255 index among the softmax threads
    const int cb = hsel * 64;                         // first column of this thread's half
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const long grow = (long)b * p.n + qt * ATC_T + r; // global token row
    float t_null = -INFINITY;
    mbar_wait(bar(Q_FULL), 0);
    if (p.has_null) {
      const uint8_t* qrow = gbase + (sQ - base) + r * 128;
      float dot = 0.f, kn2 = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 v = *reinterpret_cast<const uint4*>(qrow + ((c ^ (r & 7)) << 4));
        const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 f = __bfloat1622float2(hp[e]);
          float k0 = null_sm[c * 8 + 2 * e], k1 = null_sm[c * 8 + 2 * e + 1];
          dot = fmaf(f.x, k0, fmaf(f.y, k1, dot));
          kn2 = fmaf(k0, k0, fmaf(k1, k1, kn2));
        }
      }
      t_null = dot * p.c2 + (p.mode == 1 ? p.kb2 * kn2 : 0.f);
    }
    float m = t_null;
    int sc = 0, pc = 0;
    // ---------------- pass A: row maximum (each thread over its 64 columns)
    for (int j = 0; j < T; ++j) {
      int ss = sc & 1;
      if (p.mode == 1) {
        if (st < 128) ksq_sm[ss * 128 + st] = ksq[((long)bh * p.n) + j * ATC_T + st] * p.kb2;
        named_bar_sync(1, 256);
      }
      mbar_wait(bar(S_FULL + ss), (sc >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        tc_ld16(tS + ss * 128 + lane_addr + cb + c0, v);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float t = __uint_as_float(v[e]) * p.c2;
          if (p.mode == 1) t += ksq_sm[ss * 128 + cb + c0 + e];
          m = fmaxf(m, t);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(S_EMPTY + ss));
      ++sc;
    }
    xchg[hsel * 128 + r] = m;
    named_bar_sync(2, 256);
    m = fmaxf(m, xchg[(1 - hsel) * 128 + r]);
    // ---------------- pass B: probabilities, partial row sums, this half's P slab
    const float p_null = p.has_null ? fast_exp2(t_null - m) : 0.f;
    float l = hsel == 0 ? p_null : 0.f;
    for (int j = 0; j < T; ++j) {
      int ss = sc & 1, ps = pc & 1;
      if (p.mode == 1) {
        if (st < 128) ksq_sm[ss * 128 + st] = ksq[((long)bh * p.n) + j * ATC_T + st] * p.kb2;
        named_bar_sync(1, 256);
      }
      mbar_wait(bar(S_FULL + ss), (sc >> 1) & 1);
      mbar_wait(bar(P_EMPTY + ps), ((pc >> 1) & 1) ^ 1u);
      tc_fence_after();
      uint8_t* prow = gbase + (sP - base) + ps * 32768 + hsel * 16384 + r * 128;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        tc_ld16(tS + ss * 128 + lane_addr + cb + c0, v);
        float pv[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float t = __uint_as_float(v[e]) * p.c2;
          if (p.mode == 1) t += ksq_sm[ss * 128 + cb + c0 + e];
          pv[e] = fast_exp2(t - m);
          l += pv[e];
        }
        uint4 o0, o1;
        __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&o0);
        __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          h0[e] = __floats2bfloat162_rn(pv[2 * e], pv[2 * e + 1]);
          h1[e] = __floats2bfloat162_rn(pv[8 + 2 * e], pv[8 + 2 * e + 1]);
        }
        int ch = c0 >> 3;                              // two 16-byte chunks ch, ch+1 of this row's 128-byte slab row
        *reinterpret_cast<uint4*>(prow + (((ch) ^ (r & 7)) << 4)) = o0;
        *reinterpret_cast<uint4*>(prow + (((ch + 1) ^ (r & 7)) << 4)) = o1;
      }
      fence_async_smem();                             // generic-proxy stores -> visible to the UMMA (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar(P_FULL + ps)); mbar_arrive(bar(S_EMPTY + ss)); }
      ++sc; ++pc;
    }
    named_bar_sync(2, 256);                           // pass-A exchange fully consumed before the slots are reused
    xchg[hsel * 128 + r] = l;
    named_bar_sync(2, 256);
    l += xchg[(1 - hsel) * 128 + r];
    // ---------------- epilogue: O / l (+ null value): each thread stores 32 of the 64 output columns
    mbar_wait(bar(O_FULL), 0);
    tc_fence_after();
    const float inv = 1.f / l;
    bf16* orow = o + grow * p.o_rs + h * ATC_D;
#pragma unroll
    for (int c0 = hsel * 32; c0 < hsel * 32 + 32; c0 += 16) {
      uint32_t v[16];
      tc_ld16(tO + lane_addr + c0, v);
      float f[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        f[e] = __uint_as_float(v[e]);
        if (p.has_null) f[e] = fmaf(p_null, null_sm[64 + c0 + e], f[e]);
        f[e] *= inv;
      }
      uint4 o0, o1;
      __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&o0);
      __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        h0[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
        h1[e] = __floats2bfloat162_rn(f[8 + 2 * e], f[8 + 2 * e + 1]);
      }
      st_global_256(orow + c0, o0, o1);
    }
    if (hsel == 0) lse2[(long)bh * p.n + qt * ATC_T + r] = m + log2f(l);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

// returns 1 when the shape is not eligible (caller falls back to the FFMA kernel)
int ggi_tc_attn_fwd(const void* q, const void* k, const void* v, const float* null_kv, void* o, float* lse, float* ksq_ws,
                    int B, int heads, int nq, int nk, int d, long q_rs, long k_rs, long v_rs, long o_rs, float scale,
                    int mode, cudaStream_t st) {
  if (d != ATC_D || nq != nk || nq % ATC_T || nq < ATC_T) return 1;
  if ((q_rs % 8) || (k_rs % 8) || (v_rs % 8) || (o_rs % 16)) return 1;
  if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v) & 15) return 1;
  if (((uintptr_t)o) & 31) return 1;
  if (mode == 1 && !ksq_ws) return 1;
  AtcP p;
  p.B = B; p.heads = heads; p.n = nq; p.tiles = nq / ATC_T; p.mode = mode; p.has_null = null_kv != nullptr;
  const float log2e = 1.4426950408889634f;
  p.c2 = (mode == 1 ? 2.f * scale : scale) * log2e;
  p.kb2 = -scale * log2e;
  p.o_rs = o_rs;
  CUtensorMap tmQ, tmK, tmV;
  if (make_qkv_map(&tmQ, q, B, nq, heads, q_rs) || make_qkv_map(&tmK, k, B, nk, heads, k_rs) || make_qkv_map(&tmV, v, B, nk, heads, v_rs)) return -1;
  if (mode == 1) {
    long rows = (long)B * heads * nk;
    attn_ksq_kernel<<<gg_blocks(rows * 8, 256, 148 * 16), 256, 0, st>>>((const bf16*)k, ksq_ws, B, nk, heads, k_rs);
  }
  size_t smem = 1024 + 147456 + 2560 + 8 * 18 + 16;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  int grid = B * heads * p.tiles;
  attn_fwd_tc_kernel<<<grid, ATC_FWD_THREADS, smem, st>>>(tmQ, tmK, tmV, p, null_kv, ksq_ws, (bf16*)o, lse);
  return gg_check_launch("attn_fwd_tc");
}

// =================================================================================================
// Backward.  Two kernels (same structure as the forward: TMA producer warp, MMA warp, 4 softmax warps):
//   attn_bwd_dq_tc : CTA per query tile, loops key tiles:  S = Q K^T, dP = dO V^T, dS' = P (dP - delta) ls,
//                    dQ += dS' K.  Also delta = rowsum(dO * O) and the null key/value gradients.
//   attn_bwd_dkv_tc: CTA per key tile, loops query tiles:  S, dP as above, dV += P^T dO, dK += dS'^T [Q | 1]
//                    (the extra ones block makes column 64 of the accumulator the column sum needed by the L2 form).
// P is recomputed from the saved log-sum-exp (log2 units); every matrix product runs on tcgen05.
// =================================================================================================
__global__ void __launch_bounds__(ATC_BWD_THREADS, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, const AtbP p,
                      const float* __restrict__ null_kv, const float* __restrict__ ksq, const bf16* __restrict__ o,
                      const float* __restrict__ lse2, bf16* __restrict__ dq, float* __restrict__ delta,
                      float* __restrict__ nullrow) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t sQ = base, sDO = base + 16384, sK = base + 32768, sV = base + 65536, sDS = base + 98304;
  float* ksq_sm = (float*)(gbase + 163840);
  float* null_sm = (float*)(gbase + 164864);
  const uint32_t bars = base + 166400;
  enum { Q_FULL = 0, K_FULL = 1, K_EMPTY = 3, V_FULL = 5, V_EMPTY = 7, S_FULL = 9, S_EMPTY = 11, DP_FULL = 13, DP_EMPTY = 14,
         DS_FULL = 15, DS_EMPTY = 17, DQ_FULL = 19, NBAR = 20 };
  auto bar = [&](int i) { return bars + 8u * i; };
  uint32_t* tmem_slot = (uint32_t*)(gbase + 166400 + 8 * NBAR);
  const int warp = tc_warp_idx(), lane = threadIdx.x & 31;
  const int qt = blockIdx.x % p.tiles, bh = blockIdx.x / p.tiles;
  const int b = bh / p.heads, h = bh % p.heads;
  const int T = p.tiles;
  if (threadIdx.x == 0) {
    mbar_init(bar(Q_FULL), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(K_FULL + i), 1); mbar_init(bar(K_EMPTY + i), 1);
      mbar_init(bar(V_FULL + i), 1); mbar_init(bar(V_EMPTY + i), 1);
      mbar_init(bar(S_FULL + i), 1); mbar_init(bar(S_EMPTY + i), 8);
      mbar_init(bar(DS_FULL + i), 8); mbar_init(bar(DS_EMPTY + i), 1);
    }
    mbar_init(bar(DP_FULL), 1); mbar_init(bar(DP_EMPTY), 8); mbar_init(bar(DQ_FULL), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x >= 64) {
    int t = threadIdx.x - 64;
    if (p.has_null && t < 128) null_sm[t] = t < 64 ? null_kv[h * ATC_D + t] : null_kv[(p.heads + h) * ATC_D + (t - 64)];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tDP = tmem + 256, tDQ = tmem + 384;
  const uint32_t idesc_kk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc_dq = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  if (warp == 0) {
    {
      const uint32_t el = tc_elect_one();          // convergent producer: only the TMA / expect_tx instructions are predicated
      mbar_expect_tx_el(bar(Q_FULL), 32768, el);
      tma_load_4d_el(sQ, &tmQ, bar(Q_FULL), 0, qt * ATC_T, h, b, el);
      tma_load_4d_el(sDO, &tmDO, bar(Q_FULL), 0, qt * ATC_T, h, b, el);
      for (int j = 0; j < T; ++j) {
        int s = j & 1;
        uint32_t par = ((j >> 1) & 1) ^ 1u;
        mbar_wait(bar(K_EMPTY + s), par);
        mbar_expect_tx_el(bar(K_FULL + s), 16384, el);
        tma_load_4d_el(sK + s * 16384, &tmK, bar(K_FULL + s), 0, j * ATC_T, h, b, el);
        mbar_wait(bar(V_EMPTY + s), par);
        mbar_expect_tx_el(bar(V_FULL + s), 16384, el);
        tma_load_4d_el(sV + s * 16384, &tmV, bar(V_FULL + s), 0, j * ATC_T, h, b, el);
      }
    }
  } else if (warp == 1) {
    const uint32_t el = tc_elect_one();            // the lane that issues tcgen05.mma / commit (warp stays convergent)
    auto issue_S = [&](int j) {
      int s = j & 1;
      mbar_wait(bar(K_FULL + s), (j >> 1) & 1);
      mbar_wait(bar(S_EMPTY + s), ((j >> 1) & 1) ^ 1u);
      tc_fence_after();
      {
        uint64_t da = make_smem_desc(sQ, 1024, 2), db = make_smem_desc(sK + s * 16384, 1024, 2);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_f16_el(tS + s * 128, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_kk, k ? 1u : 0u, el);
        tc_commit_el(bar(S_FULL + s), el);
      }
      __syncwarp();
    };
    auto issue_dP = [&](int j) {
      int s = j & 1;
      mbar_wait(bar(V_FULL + s), (j >> 1) & 1);
      mbar_wait(bar(DP_EMPTY), (j & 1) ^ 1u);
      tc_fence_after();
      {
        uint64_t da = make_smem_desc(sDO, 1024, 2), db = make_smem_desc(sV + s * 16384, 1024, 2);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_f16_el(tDP, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_kk, k ? 1u : 0u, el);
        tc_commit_el(bar(DP_FULL), el);
        tc_commit_el(bar(V_EMPTY + s), el);
      }
      __syncwarp();
    };
    mbar_wait(bar(Q_FULL), 0);
    issue_S(0);
    issue_dP(0);
    for (int j = 0; j < T; ++j) {
      if (j + 1 < T) issue_S(j + 1);
      int s = j & 1;
      mbar_wait(bar(DS_FULL + s), (j >> 1) & 1);
      tc_fence_after();
      {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          uint64_t da = make_smem_desc(sDS + s * 32768 + (k >> 2) * 16384, 1024, 2) + (uint64_t)(2 * (k & 3));
          uint64_t db = make_smem_desc_mn(sK + s * 16384 + k * 2048, 0, 1024);
          tc_mma_f16_el(tDQ, da, db, idesc_dq, (j | k) ? 1u : 0u, el);
        }
        tc_commit_el(bar(DS_EMPTY + s), el);
        tc_commit_el(bar(K_EMPTY + s), el);
        if (j == T - 1) tc_commit_el(bar(DQ_FULL), el);
      }
      __syncwarp();
      if (j + 1 < T) issue_dP(j + 1);
    }
  } else {
    const int q = warp & 3;                            // TMEM lane quarter; two warps per quarter split the tile columns
    const int hsel = (warp - 2) >> 2, cb = hsel * 64;
    const int r = q * 32 + lane;
    const int st = threadIdx.x - 64;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const long grow = (long)b * p.n + qt * ATC_T + r;
    const long srow = (long)bh * p.n + qt * ATC_T + r;
    mbar_wait(bar(Q_FULL), 0);
    // streaming pass over this row of Q, dO (smem tiles) and O (global): delta = dO.O, q.k_null, dO.v_null
    float dl = 0.f, dot = 0.f, kn2 = 0.f, dpn = 0.f;
    {
      const uint8_t* qr = gbase + (sQ - base) + r * 128;
      const uint8_t* dr = gbase + (sDO - base) + r * 128;
      const bf16* orow = o + grow * p.o_rs + h * ATC_D;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 qv = *reinterpret_cast<const uint4*>(qr + ((c ^ (r & 7)) << 4));
        uint4 dv = *reinterpret_cast<const uint4*>(dr + ((c ^ (r & 7)) << 4));
        uint4 ov = __ldg(reinterpret_cast<const uint4*>(orow) + c);
        const __nv_bfloat162* qh = reinterpret_cast<const __nv_bfloat162*>(&qv);
        const __nv_bfloat162* dh = reinterpret_cast<const __nv_bfloat162*>(&dv);
        const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&ov);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 qf = __bfloat1622float2(qh[e]), df = __bfloat1622float2(dh[e]), of = __bfloat1622float2(oh[e]);
          dl = fmaf(df.x, of.x, fmaf(df.y, of.y, dl));
          if (p.has_null) {
            float k0 = null_sm[c * 8 + 2 * e], k1 = null_sm[c * 8 + 2 * e + 1];
            float v0 = null_sm[64 + c * 8 + 2 * e], v1 = null_sm[64 + c * 8 + 2 * e + 1];
            dot = fmaf(qf.x, k0, fmaf(qf.y, k1, dot));
            kn2 = fmaf(k0, k0, fmaf(k1, k1, kn2));
            dpn = fmaf(df.x, v0, fmaf(df.y, v1, dpn));
          }
        }
      }
    }
    if (hsel == 0) delta[srow] = dl;
    const float L2 = lse2[srow];
    float ds_null = 0.f, p_null = 0.f;
    if (p.has_null) {
      float tn = dot * p.c2 + (p.mode == 1 ? p.kb2 * kn2 : 0.f);
      p_null = fast_exp2(tn - L2);
      ds_null = p_null * (dpn - dl) * p.ls;
      if (hsel == 0) {
        nullrow[srow] = ds_null;                       // consumed by attn_null_grad_kernel
        nullrow[(long)p.B * p.heads * p.n + srow] = p_null;
      }
    }
    for (int j = 0; j < T; ++j) {
      int s = j & 1;
      if (p.mode == 1) {
        if (st < 128) ksq_sm[s * 128 + st] = ksq[((long)bh * p.n) + j * ATC_T + st] * p.kb2;
        named_bar_sync(1, 256);
      }
      mbar_wait(bar(S_FULL + s), (j >> 1) & 1);
      mbar_wait(bar(DP_FULL), j & 1);
      mbar_wait(bar(DS_EMPTY + s), ((j >> 1) & 1) ^ 1u);
      tc_fence_after();
      uint8_t* dstile = gbase + (sDS - base) + s * 32768;
#pragma unroll 2
      for (int c0 = cb; c0 < cb + 64; c0 += 16) {
        uint32_t sv[16], dv[16];
        tc_ld16(tS + s * 128 + lane_addr + c0, sv);
        tc_ld16(tDP + lane_addr + c0, dv);
        float f[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float t = __uint_as_float(sv[e]) * p.c2;
          if (p.mode == 1) t += ksq_sm[s * 128 + c0 + e];
          f[e] = fast_exp2(t - L2) * (__uint_as_float(dv[e]) - dl) * p.ls;
        }
        write_tile16(dstile, r, c0, f);
      }
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar(DS_FULL + s)); mbar_arrive(bar(S_EMPTY + s)); mbar_arrive(bar(DP_EMPTY)); }
    }
    mbar_wait(bar(DQ_FULL), 0);
    tc_fence_after();
    bf16* dqrow = dq + grow * (long)(p.heads * ATC_D) + h * ATC_D;
#pragma unroll
    for (int c0 = hsel * 32; c0 < hsel * 32 + 32; c0 += 16) {
      uint32_t v[16];
      tc_ld16(tDQ + lane_addr + c0, v);
      float f[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(v[e]) + (p.has_null ? ds_null * null_sm[c0 + e] : 0.f);
      store_row16(dqrow + c0, f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

__global__ void __launch_bounds__(ATC_BWD_THREADS, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, const AtbP p,
                       const float* __restrict__ ksq, const float* __restrict__ lse2, const float* __restrict__ delta,
                       bf16* __restrict__ dk, bf16* __restrict__ dv) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  // K 16K | V 16K | [Q 16K | ones 16K] x2 | dO x2 | P 32K | dS 32K
  const uint32_t sK = base, sV = base + 16384, sQO = base + 32768, sDO = base + 98304, sP = base + 131072, sDS = base + 163840;
  float* ksq_sm = (float*)(gbase + 196608);
  const uint32_t bars = base + 197120;
  enum { KV_FULL = 0, QO_FULL = 1, QO_EMPTY = 3, SDP_FULL = 5, SDP_EMPTY = 6, PDS_FULL = 7, PDS_EMPTY = 8, OUT_FULL = 9, NBAR = 10 };
  auto bar = [&](int i) { return bars + 8u * i; };
  uint32_t* tmem_slot = (uint32_t*)(gbase + 197120 + 8 * NBAR);
  const int warp = tc_warp_idx(), lane = threadIdx.x & 31;
  const int kt = blockIdx.x % p.tiles, bh = blockIdx.x / p.tiles;
  const int b = bh / p.heads, h = bh % p.heads;
  const int T = p.tiles;
  if (threadIdx.x == 0) {
    mbar_init(bar(KV_FULL), 1);
    for (int i = 0; i < 2; ++i) { mbar_init(bar(QO_FULL + i), 1); mbar_init(bar(QO_EMPTY + i), 1); }
    mbar_init(bar(SDP_FULL), 1); mbar_init(bar(SDP_EMPTY), 8);
    mbar_init(bar(PDS_FULL), 8); mbar_init(bar(PDS_EMPTY), 1); mbar_init(bar(OUT_FULL), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {   // the two "ones" slabs (bf16 1.0 everywhere; swizzle-invariant)
    uint4 one4 = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    for (int i = threadIdx.x; i < 2 * 1024; i += ATC_BWD_THREADS) {
      int buf = i >> 10, off = (i & 1023) << 4;
      *reinterpret_cast<uint4*>(gbase + (sQO - base) + buf * 32768 + 16384 + off) = one4;
    }
    fence_async_smem();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 192 && p.mode == 1) ksq_sm[threadIdx.x - 64] = ksq[(long)bh * p.n + kt * ATC_T + (threadIdx.x - 64)] * p.kb2;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tDP = tmem + 128, tDV = tmem + 256, tDK = tmem + 320;
  const uint32_t idesc_kk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc_dv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc_dk = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(80 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  if (warp == 0) {
    {
      const uint32_t el = tc_elect_one();          // convergent producer: only the TMA / expect_tx instructions are predicated
      mbar_expect_tx_el(bar(KV_FULL), 32768, el);
      tma_load_4d_el(sK, &tmK, bar(KV_FULL), 0, kt * ATC_T, h, b, el);
      tma_load_4d_el(sV, &tmV, bar(KV_FULL), 0, kt * ATC_T, h, b, el);
      for (int i = 0; i < T; ++i) {
        int s = i & 1;
        mbar_wait(bar(QO_EMPTY + s), ((i >> 1) & 1) ^ 1u);
        mbar_expect_tx_el(bar(QO_FULL + s), 32768, el);
        tma_load_4d_el(sQO + s * 32768, &tmQ, bar(QO_FULL + s), 0, i * ATC_T, h, b, el);
        tma_load_4d_el(sDO + s * 16384, &tmDO, bar(QO_FULL + s), 0, i * ATC_T, h, b, el);
      }
    }
  } else if (warp == 1) {
    const uint32_t el = tc_elect_one();            // the lane that issues tcgen05.mma / commit (warp stays convergent)
    auto issue_SdP = [&](int i) {
      int s = i & 1;
      mbar_wait(bar(QO_FULL + s), (i >> 1) & 1);
      mbar_wait(bar(SDP_EMPTY), (i & 1) ^ 1u);
      tc_fence_after();
      {
        uint64_t dq_ = make_smem_desc(sQO + s * 32768, 1024, 2), dk_ = make_smem_desc(sK, 1024, 2);
        uint64_t do_ = make_smem_desc(sDO + s * 16384, 1024, 2), dv_ = make_smem_desc(sV, 1024, 2);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_f16_el(tS, dq_ + (uint64_t)(2 * k), dk_ + (uint64_t)(2 * k), idesc_kk, k ? 1u : 0u, el);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_f16_el(tDP, do_ + (uint64_t)(2 * k), dv_ + (uint64_t)(2 * k), idesc_kk, k ? 1u : 0u, el);
        tc_commit_el(bar(SDP_FULL), el);
      }
      __syncwarp();
    };
    mbar_wait(bar(KV_FULL), 0);
    issue_SdP(0);
    for (int i = 0; i < T; ++i) {
      int s = i & 1;
      mbar_wait(bar(PDS_FULL), i & 1);
      tc_fence_after();
      {
#pragma unroll
        for (int k = 0; k < 8; ++k) {      // K axis = the 128 queries of this tile, 16 per step
          uint64_t ap = make_smem_desc_mn(sP + k * 2048, 16384, 1024);
          uint64_t bo = make_smem_desc_mn(sDO + s * 16384 + k * 2048, 0, 1024);
          tc_mma_f16_el(tDV, ap, bo, idesc_dv, (i | k) ? 1u : 0u, el);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          uint64_t as_ = make_smem_desc_mn(sDS + k * 2048, 16384, 1024);
          uint64_t bq = make_smem_desc_mn(sQO + s * 32768 + k * 2048, 16384, 1024);
          tc_mma_f16_el(tDK, as_, bq, idesc_dk, (i | k) ? 1u : 0u, el);
        }
        tc_commit_el(bar(PDS_EMPTY), el);
        tc_commit_el(bar(QO_EMPTY + s), el);
        if (i == T - 1) tc_commit_el(bar(OUT_FULL), el);
      }
      __syncwarp();
      if (i + 1 < T) issue_SdP(i + 1);
    }
  } else {
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2, cb = hsel * 64;   // two warps per TMEM lane quarter: columns [cb, cb + 64)
    const int r = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint8_t* ptile = gbase + (sP - base);
    uint8_t* dstile = gbase + (sDS - base);
    for (int i = 0; i < T; ++i) {
      const long srow = (long)bh * p.n + i * ATC_T + r;
      const float L2 = lse2[srow], dl = delta[srow];
      mbar_wait(bar(SDP_FULL), i & 1);
      mbar_wait(bar(PDS_EMPTY), (i & 1) ^ 1u);
      tc_fence_after();
#pragma unroll 2
      for (int c0 = cb; c0 < cb + 64; c0 += 16) {
        uint32_t sv[16], dv_[16];
        tc_ld16(tS + lane_addr + c0, sv);
        tc_ld16(tDP + lane_addr + c0, dv_);
        float pf[16], df[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float t = __uint_as_float(sv[e]) * p.c2;
          if (p.mode == 1) t += ksq_sm[c0 + e];
          pf[e] = fast_exp2(t - L2);
          df[e] = pf[e] * (__uint_as_float(dv_[e]) - dl) * p.ls;
        }
        write_tile16(ptile, r, c0, pf);
        write_tile16(dstile, r, c0, df);
      }
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar(PDS_FULL)); mbar_arrive(bar(SDP_EMPTY)); }
    }
    // epilogue: thread <-> key row
    mbar_wait(bar(OUT_FULL), 0);
    tc_fence_after();
    const long grow = (long)b * p.n + kt * ATC_T + r;
    bf16* dvrow = dv + grow * (long)(p.heads * ATC_D) + h * ATC_D;
    bf16* dkrow = dk + grow * (long)(p.heads * ATC_D) + h * ATC_D;
    if (hsel == 0) {                                   // first warp of the quarter stores dV, the second dK
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        float f[16];
        tc_ld16(tDV + lane_addr + c0, v);
#pragma unroll
        for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(v[e]);
        store_row16(dvrow + c0, f);
      }
    } else {
      float csum = 0.f;
      float krow[64];
      if (p.mode == 1) {
        uint32_t v[16];
        tc_ld16(tDK + lane_addr + 64, v);
        csum = __uint_as_float(v[0]);
        read_tile_row(gbase + (sK - base), r, krow);
      }
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        float f[16];
        tc_ld16(tDK + lane_addr + c0, v);
#pragma unroll
        for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(v[e]) - (p.mode == 1 ? csum * krow[c0 + e] : 0.f);
        store_row16(dkrow + c0, f);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

// dk_null[h][c] = sum_rows ds_null * (q - [l2] k_null), dv_null[h][c] = sum_rows p_null * dO   (rows = all (b, token))
__global__ void attn_null_grad_kernel(const bf16* __restrict__ q, const bf16* __restrict__ go, const float* __restrict__ nullrow,
                                      const float* __restrict__ null_kv, float* __restrict__ dnull, int B, int n, int heads,
                                      long q_rs, int mode, int rows_per_block) {
  __shared__ float sm[32][129];
  const int h = blockIdx.y, oct = threadIdx.x & 7, rl = threadIdx.x >> 3;       // 256 threads: 32 row lanes x 8 column octets
  const long total = (long)B * n, r0 = (long)blockIdx.x * rows_per_block;
  const long r1 = min(total, r0 + rows_per_block);
  float kn[8], gk[8], gv[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { kn[e] = mode == 1 ? null_kv[h * ATC_D + oct * 8 + e] : 0.f; gk[e] = 0.f; gv[e] = 0.f; }
  for (long row = r0 + rl; row < r1; row += 32) {
    long b = row / n, i = row - b * n;
    long srow = (b * heads + h) * (long)n + i;
    float dsn = nullrow[srow], pn = nullrow[(long)B * heads * n + srow];
    uint4 qv = *reinterpret_cast<const uint4*>(q + row * q_rs + h * ATC_D + oct * 8);
    uint4 gov = *reinterpret_cast<const uint4*>(go + row * (long)(heads * ATC_D) + h * ATC_D + oct * 8);
    const __nv_bfloat162* qh = reinterpret_cast<const __nv_bfloat162*>(&qv);
    const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&gov);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 qf = __bfloat1622float2(qh[e]), gf = __bfloat1622float2(gh[e]);
      gk[2 * e] = fmaf(dsn, qf.x - kn[2 * e], gk[2 * e]); gk[2 * e + 1] = fmaf(dsn, qf.y - kn[2 * e + 1], gk[2 * e + 1]);
      gv[2 * e] = fmaf(pn, gf.x, gv[2 * e]); gv[2 * e + 1] = fmaf(pn, gf.y, gv[2 * e + 1]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) { sm[rl][oct * 8 + e] = gk[e]; sm[rl][64 + oct * 8 + e] = gv[e]; }
  __syncthreads();
  if (threadIdx.x < 128) {
    float t = 0.f;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) t += sm[r][threadIdx.x];
    int cc = threadIdx.x & 63;
    atomicAdd(dnull + (threadIdx.x < 64 ? h * ATC_D + cc : (heads + h) * ATC_D + cc), t);
  }
}

// go, dq, dk, dv: dense (B, n, heads*64); delta_ws holds 3*B*heads*n floats.  Returns 1 when not eligible.
int ggi_tc_attn_bwd(const void* q, const void* k, const void* v, const float* null_kv, const void* o, const void* go,
                    const float* lse2, void* dq, void* dk, void* dv, float* dnull_kv, float* delta_ws, float* ksq_ws,
                    int B, int heads, int nq, int nk, int d, long q_rs, long k_rs, long v_rs, long o_rs, float scale,
                    int mode, cudaStream_t st) {
  if (d != ATC_D || nq != nk || nq % ATC_T || nq < ATC_T) return 1;
  if ((q_rs % 8) || (k_rs % 8) || (v_rs % 8) || (o_rs % 8)) return 1;
  if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)o | (uintptr_t)go) & 15) return 1;
  if (((uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv) & 31) return 1;
  if (mode == 1 && !ksq_ws) return 1;
  AtbP p;
  p.B = B; p.heads = heads; p.n = nq; p.tiles = nq / ATC_T; p.mode = mode; p.has_null = null_kv != nullptr;
  const float log2e = 1.4426950408889634f;
  p.ls = mode == 1 ? 2.f * scale : scale;
  p.c2 = p.ls * log2e;
  p.kb2 = -scale * log2e;
  p.o_rs = o_rs;
  CUtensorMap tmQ, tmK, tmV, tmDO;
  long hd = (long)heads * ATC_D;
  if (make_qkv_map(&tmQ, q, B, nq, heads, q_rs) || make_qkv_map(&tmK, k, B, nk, heads, k_rs) ||
      make_qkv_map(&tmV, v, B, nk, heads, v_rs) || make_qkv_map(&tmDO, go, B, nq, heads, hd)) return -1;
  if (mode == 1) {
    long rows = (long)B * heads * nk;
    attn_ksq_kernel<<<gg_blocks(rows * 8, 256, 148 * 16), 256, 0, st>>>((const bf16*)k, ksq_ws, B, nk, heads, k_rs);
  }
  if (p.has_null) cudaMemsetAsync(dnull_kv, 0, sizeof(float) * 2 * heads * ATC_D, st);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(attn_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  int grid = B * heads * p.tiles;
  size_t smem1 = 1024 + 166400 + 8 * 20 + 16, smem2 = 1024 + 197120 + 8 * 10 + 16;
  float* nullrow = delta_ws + (size_t)B * heads * nq;
  attn_bwd_dq_tc_kernel<<<grid, ATC_BWD_THREADS, smem1, st>>>(tmQ, tmK, tmV, tmDO, p, null_kv, ksq_ws, (const bf16*)o, lse2, (bf16*)dq, delta_ws, nullrow);
  if (p.has_null) {
    int rpb = 128;
    dim3 g2(gg_cdiv((long)B * nq, rpb), heads);
    attn_null_grad_kernel<<<g2, 256, 0, st>>>((const bf16*)q, (const bf16*)go, nullrow, null_kv, dnull_kv, B, nq, heads, q_rs, mode, rpb);
  }
  attn_bwd_dkv_tc_kernel<<<grid, ATC_BWD_THREADS, smem2, st>>>(tmQ, tmK, tmV, tmDO, p, ksq_ws, lse2, delta_ws, (bf16*)dk, (bf16*)dv);
  return gg_check_launch("attn_bwd_tc");
}

// launch wrappers used by attn_tc2.cu (kernels cannot be launched across translation units without -rdc)
void atc_launch_ksq(const void* k, float* ksq_ws, int B, int n, int heads, long k_rs, cudaStream_t st) {
  long rows = (long)B * heads * n;
  attn_ksq_kernel<<<gg_blocks(rows * 8, 256, 148 * 16), 256, 0, st>>>((const bf16*)k, ksq_ws, B, n, heads, k_rs);
}
void atc_launch_null_grad(const void* q, const void* go, const float* nullrow, const float* null_kv, float* dnull,
                          int B, int n, int heads, long q_rs, int mode, cudaStream_t st) {
  int rpb = 128;
  dim3 g2(gg_cdiv((long)B * n, rpb), heads);
  attn_null_grad_kernel<<<g2, 256, 0, st>>>((const bf16*)q, (const bf16*)go, nullrow, null_kv, dnull, B, n, heads, q_rs, mode, rpb);
}
