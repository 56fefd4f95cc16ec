// Fused self-attention on tcgen05, second generation of the softmax side (same MMA / TMA structure, same shared-memory
// tiles and the same maths as attn_tc.cu; replaces gigagan_pytorch.py:562-592 and its autograd backward).
//
// ncu on the first generation (profiles/r02_ncu_attention_tc.md): tensor pipe 16-23 % active, issue slots 32 % busy with
// two softmax warps per scheduler - the softmax warps were LATENCY bound, not throughput bound.  The source page
// (tools/ncu_source_regions.py) located the waits; what changed, in the order measured:
//   * a softmax thread pulls its whole column slab of S (and dP) out of TMEM with ONE round trip and immediately hands the
//     accumulator back to the MMA warp ("early release"); the MMA warp issues S(j+1) / dP(j+1) BEFORE it waits for dS(j);
//   * per-tile scalars (|k|^2, lse, delta) are prefetched one tile ahead into registers, with no consumer behind the load;
//   * stage depth follows the release point: 3 key stages in dQ (a key tile is released by dQ(j), not by S(j)), 3 query/dO
//     stages with one shared block of ones in dK/dV, 4 key stages in the one-CTA forward;
//   * |k|^2 of a key tile is staged per WARP (__syncwarp) instead of for all softmax warps behind a 512-thread barrier;
//   * the shared-QK L2 form needs no max pass (ONEP): the row maximum of -s |q_i - k_j|^2 is the diagonal;
//   * the forward runs two CTAs per SM (OCC = 2): ~6 us of prologue/epilogue per CTA around 8-16 short tile steps.
// Template parameters: NSW softmax warps (8: each thread owns 64 of the 128 tile columns, 16: 32), L2M shared-QK L2-distance
// logits, ONEP / OCC forward only.  Defaults: forward <8, *, *, 2>, backward <16, *>; gg_set_flags bits 3-7 select the
// other instantiations (and the first generation) for A/B measurements and the parity tests.
#include "attn_tc_common.cuh"

#define ATC2_FWD_KS 4      // key-tile stages of the forward kernel
#define ATC2_DQ_KS 3       // key-tile stages of the dQ kernel (a key tile is read by S(j) early and by dQ(j) late)
#define ATC2_DKV_QS 3      // query / dO tile stages of the dK,dV kernel
#define ATC2_DKV_BARS (32768 + ATC2_DKV_QS * 16384 + 16384 + ATC2_DKV_QS * 16384 + 65536 + 512)   // tiles + |k|^2 row

__device__ __forceinline__ void tc_ld32_nw(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
}
// tcgen05.wait::ld with the loaded registers threaded through the statement: nothing that consumes them can be scheduled
// above the wait
__device__ __forceinline__ void tc_wait_ld32(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :: "memory");
}
template <int CW>
__device__ __forceinline__ void tc_ld_cols(uint32_t taddr, uint32_t* r) {
#pragma unroll
  for (int c = 0; c < CW; c += 32) tc_ld32_nw(taddr + c, r + c);
}
template <int CW>
__device__ __forceinline__ void tc_wait_cols(uint32_t* r) {
#pragma unroll
  for (int c = 0; c < CW; c += 32) tc_wait_ld32(r + c);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// 16 consecutive columns [c0, c0+16) of row r of a [128 x 128] bf16 tile (two 64-column SWIZZLE_128B slabs), already packed
__device__ __forceinline__ void write_tile16_packed(uint8_t* tile, int r, int c0, const uint32_t* pk) {
  int slab = c0 >> 6, ch = (c0 & 63) >> 3;
  uint8_t* dst = tile + slab * 16384 + r * 128;
  *reinterpret_cast<uint4*>(dst + ((ch ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  *reinterpret_cast<uint4*>(dst + (((ch + 1) ^ (r & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
}
template <int CW>
__device__ __forceinline__ void load_ksq_slab(const float* src, float* kk) {
  const float4* k4 = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int i = 0; i < CW / 4; ++i) { float4 t = k4[i]; kk[4 * i] = t.x; kk[4 * i + 1] = t.y; kk[4 * i + 2] = t.z; kk[4 * i + 3] = t.w; }
}

// ================================================================================================= forward
// ONEP (shared-QK L2-distance form only, k == q): the logits are -s |q_i - k_j|^2 <= 0 with equality at j = i (and the null
// key's logit is <= 0 as well), so the row maximum is known before any product is formed - in the reduced form used here
// (the -s |q_i|^2 term dropped) it is s |q_i|^2 - and pass A disappears: half of the Q K^T products and key-tile loads.
// OCC = 2: two CTAs per SM (one buffer each for S, V and P, two key stages, 256 TMEM columns, <= 113 KB): a CTA spends ~6 us
// of un-overlapped prologue/epilogue around 8-16 tile steps of ~0.5-0.8 us; a second resident CTA fills it.
template <int NSW, bool L2M, bool ONEP, int OCC>
__global__ void __launch_bounds__(64 + NSW * 32, OCC)
attn2_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AtcP p, const float* __restrict__ null_kv,
                 const float* __restrict__ ksq, bf16* __restrict__ o, float* __restrict__ lse2) {
  constexpr int NH = NSW / 4, CW = 128 / NH, NST = NSW * 32, OC = 64 / NH;
  constexpr int KS = OCC == 2 ? 2 : ATC2_FWD_KS;    // key-tile stages: pass A only takes row maxima, i.e. runs at the speed
                                                    // the key tiles arrive - two stages left it TMA-latency bound
  constexpr int NB = OCC == 2 ? 1 : 2;              // buffers of S (TMEM), V and P (shared memory)
  constexpr uint32_t AUX = 16384 + KS * 16384 + NB * 16384 + NB * 32768;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  // layout: Q 16K | K[KS] | V[NB] | P[NB] | |k|^2 slabs [NSW][CW] 2K | null k,v 512B | exchange [NH][128] <=2K | barriers | tmem slot
  const uint32_t sQ = base, sK = base + 16384, sV = sK + KS * 16384, sP = sV + NB * 16384;
  float* ksq_sm = (float*)(gbase + AUX);
  float* null_sm = (float*)(gbase + AUX + 2048);
  float* xchg = (float*)(gbase + AUX + 2560);
  const uint32_t bars = base + AUX + 4608;
  enum { Q_FULL = 0, K_FULL = 1, K_EMPTY = K_FULL + KS, V_FULL = K_EMPTY + KS, V_EMPTY = V_FULL + NB, S_FULL = V_EMPTY + NB,
         S_EMPTY = S_FULL + NB, P_FULL = S_EMPTY + NB, P_EMPTY = P_FULL + NB, O_FULL = P_EMPTY + NB, NBAR = O_FULL + 1 };
  auto bar = [&](int i) { return bars + 8u * i; };
  uint32_t* tmem_slot = (uint32_t*)(gbase + AUX + 4608 + 8 * NBAR);
  constexpr uint32_t TMEM_COLS = NB == 2 ? 512 : 256;

  const int warp = tc_warp_idx(), lane = threadIdx.x & 31;
  const int qt = blockIdx.x % p.tiles, bh = blockIdx.x / p.tiles;
  const int b = bh / p.heads, h = bh % p.heads;
  const int T = p.tiles;

  if (threadIdx.x == 0) {
    mbar_init(bar(Q_FULL), 1);
    for (int i = 0; i < KS; ++i) { mbar_init(bar(K_FULL + i), 1); mbar_init(bar(K_EMPTY + i), 1); }
    for (int i = 0; i < NB; ++i) {
      mbar_init(bar(V_FULL + i), 1); mbar_init(bar(V_EMPTY + i), 1);
      mbar_init(bar(S_FULL + i), 1); mbar_init(bar(S_EMPTY + i), NSW);
      mbar_init(bar(P_FULL + i), NSW); mbar_init(bar(P_EMPTY + i), 1);
    }
    mbar_init(bar(O_FULL), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
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
  const uint32_t tS = tmem, tO = tmem + NB * 128;  // S[NB] at columns 0 (/ 128), O behind them (64 columns)
  const uint32_t idesc_qk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  if (warp == 0) {
    // ================================================= TMA producer (convergent; only the instructions are predicated)
    const uint32_t el = tc_elect_one();
    mbar_expect_tx_el(bar(Q_FULL), 16384, el);
    tma_load_4d_el(sQ, &tmQ, bar(Q_FULL), 0, qt * ATC_T, h, b, el);
    int kc = 0, vc = 0;
    for (int pass = ONEP ? 1 : 0; pass < 2; ++pass)
      for (int j = 0; j < T; ++j) {
        int s = kc % KS;
        mbar_wait(bar(K_EMPTY + s), ((kc / KS) & 1) ^ 1u);
        mbar_expect_tx_el(bar(K_FULL + s), 16384, el);
        tma_load_4d_el(sK + s * 16384, &tmK, bar(K_FULL + s), 0, j * ATC_T, h, b, el);
        ++kc;
        if (pass == 1) {
          int sv = vc % NB;
          mbar_wait(bar(V_EMPTY + sv), ((vc / NB) & 1) ^ 1u);
          mbar_expect_tx_el(bar(V_FULL + sv), 16384, el);
          tma_load_4d_el(sV + sv * 16384, &tmV, bar(V_FULL + sv), 0, j * ATC_T, h, b, el);
          ++vc;
        }
      }
  } else if (warp == 1) {
    // ================================================= MMA issuer
    const uint32_t el = tc_elect_one();
    int kc = 0, sc = 0, pc = 0;
    auto issue_S = [&]() {
      int ks = kc % KS, ss = sc % NB;
      mbar_wait(bar(K_FULL + ks), (kc / KS) & 1);
      mbar_wait(bar(S_EMPTY + ss), ((sc / NB) & 1) ^ 1u);
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
    if (!ONEP)
      for (int j = 0; j < T; ++j) issue_S();         // pass A: row maxima
    issue_S();                                        // pass B, S_0
    for (int j = 0; j < T; ++j) {
      if (j + 1 < T) issue_S();
      int ps = pc % NB;
      mbar_wait(bar(P_FULL + ps), (pc / NB) & 1);
      mbar_wait(bar(V_FULL + ps), (pc / NB) & 1);
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
    // warp & 3 selects the TMEM lane quarter (32 query rows), (warp - 2) / 4 the CW-column slab of every 128-key tile
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2;
    const int r = q * 32 + lane;                      // row inside the tile
    const int cb = hsel * CW;                         // first column of this thread's slab
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const long grow = (long)b * p.n + qt * ATC_T + r; // global token row
    // |k|^2 of a key tile: every WARP stages its own CW-column slab (lanes < CW/4 fetch one float4 a tile ahead, raw -
    // nothing may consume the register before the tile boundary - and store it scaled; __syncwarp instead of a barrier
    // over all softmax warps, which cost 18 % of the samples as a convoy point in front of every tile)
    float* kslab = ksq_sm + (warp - 2) * CW;
    float4 kq4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto ksq_fetch = [&](int j) {
      if (L2M && lane < CW / 4) kq4 = __ldg(reinterpret_cast<const float4*>(ksq + ((long)bh * p.n) + j * ATC_T + cb) + lane);
    };
    auto ksq_stage = [&]() {
      __syncwarp();                                   // the previous tile's reads of the slab are done
      if (lane < CW / 4) reinterpret_cast<float4*>(kslab)[lane] = make_float4(kq4.x * p.kb2, kq4.y * p.kb2, kq4.z * p.kb2, kq4.w * p.kb2);
      __syncwarp();
    };
    ksq_fetch(0);
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
      t_null = dot * p.c2 + (L2M ? p.kb2 * kn2 : 0.f);
    }
    float m = t_null;
    int sc = 0, pc = 0;
    if (ONEP) m = -p.kb2 * __ldg(ksq + (long)bh * p.n + qt * ATC_T + r);      // s log2(e) |q_r|^2: the diagonal logit
    // ---------------- pass A: row maximum (each thread over its CW columns)
    for (int j = 0; j < (ONEP ? 0 : T); ++j) {
      int ss = sc % NB;
      if (L2M) {
        ksq_stage();
        ksq_fetch(j + 1 < T ? j + 1 : 0);             // the tile after the last one of pass A is tile 0 of pass B
      }
      mbar_wait(bar(S_FULL + ss), (sc / NB) & 1);
      tc_fence_after();
      uint32_t v[CW];
      tc_ld_cols<CW>(tS + ss * 128 + lane_addr + cb, v);
      tc_wait_cols<CW>(v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(S_EMPTY + ss));  // accumulator handed back before the arithmetic
      if (L2M) {
        float kk[CW];
        load_ksq_slab<CW>(kslab, kk);
#pragma unroll
        for (int e = 0; e < CW; ++e) m = fmaxf(m, fmaf(__uint_as_float(v[e]), p.c2, kk[e]));
      } else {
        float mr = -INFINITY;
#pragma unroll
        for (int e = 0; e < CW; ++e) mr = fmaxf(mr, __uint_as_float(v[e]));
        m = fmaxf(m, mr * p.c2);                      // c2 > 0
      }
      ++sc;
    }
    if (!ONEP) {
      xchg[hsel * 128 + r] = m;
      named_bar_sync(2, NST);
#pragma unroll
      for (int hh = 0; hh < NH; ++hh) m = fmaxf(m, xchg[hh * 128 + r]);
    }
    // ---------------- pass B: probabilities, partial row sums, this thread's slab of P
    const float p_null = p.has_null ? fast_exp2(t_null - m) : 0.f;
    const float negm = -m;
    float l = hsel == 0 ? p_null : 0.f;
    for (int j = 0; j < T; ++j) {
      int ss = sc % NB, ps = pc % NB;
      if (L2M) {
        ksq_stage();
        if (j + 1 < T) ksq_fetch(j + 1);
      }
      mbar_wait(bar(S_FULL + ss), (sc / NB) & 1);
      tc_fence_after();
      uint32_t v[CW];
      tc_ld_cols<CW>(tS + ss * 128 + lane_addr + cb, v);
      tc_wait_cols<CW>(v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(S_EMPTY + ss));
      uint32_t pk[CW / 2];
      if (L2M) {
        float kk[CW];
        load_ksq_slab<CW>(kslab, kk);
#pragma unroll
        for (int e = 0; e < CW; e += 2) {
          float p0 = fast_exp2(fmaf(__uint_as_float(v[e]), p.c2, kk[e]) + negm);
          float p1 = fast_exp2(fmaf(__uint_as_float(v[e + 1]), p.c2, kk[e + 1]) + negm);
          l += p0 + p1;
          pk[e >> 1] = pack_bf16x2(p0, p1);
        }
      } else {
#pragma unroll
        for (int e = 0; e < CW; e += 2) {
          float p0 = fast_exp2(fmaf(__uint_as_float(v[e]), p.c2, negm));
          float p1 = fast_exp2(fmaf(__uint_as_float(v[e + 1]), p.c2, negm));
          l += p0 + p1;
          pk[e >> 1] = pack_bf16x2(p0, p1);
        }
      }
      mbar_wait(bar(P_EMPTY + ps), ((pc / NB) & 1) ^ 1u);
      uint8_t* ptile = gbase + (sP - base) + ps * 32768;
#pragma unroll
      for (int g = 0; g < CW / 16; ++g) write_tile16_packed(ptile, r, cb + 16 * g, pk + 8 * g);
      fence_async_smem();                             // generic-proxy stores -> visible to the UMMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(P_FULL + ps));
      ++sc; ++pc;
    }
    named_bar_sync(2, NST);                           // pass-A exchange fully consumed before the slots are reused
    xchg[hsel * 128 + r] = l;
    named_bar_sync(2, NST);
    l = 0.f;
#pragma unroll
    for (int hh = 0; hh < NH; ++hh) l += xchg[hh * 128 + r];
    // ---------------- epilogue: O / l (+ null value): each thread stores OC of the 64 output columns
    mbar_wait(bar(O_FULL), 0);
    tc_fence_after();
    const float inv = 1.f / l;
    bf16* orow = o + grow * p.o_rs + h * ATC_D;
#pragma unroll
    for (int c0 = hsel * OC; c0 < hsel * OC + OC; c0 += 16) {
      uint32_t v[16];
      tc_ld16(tO + lane_addr + c0, v);
      float f[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        f[e] = __uint_as_float(v[e]);
        if (p.has_null) f[e] = fmaf(p_null, null_sm[64 + c0 + e], f[e]);
        f[e] *= inv;
      }
      store_row16(orow + c0, f);
    }
    if (hsel == 0) lse2[(long)bh * p.n + qt * ATC_T + r] = m + log2f(l);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

// ================================================================================================= backward: dQ
template <int NSW, bool L2M>
__global__ void __launch_bounds__(64 + NSW * 32, 1)
attn2_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, const AtbP p,
                    const float* __restrict__ null_kv, const float* __restrict__ ksq, const bf16* __restrict__ o,
                    const float* __restrict__ lse2, bf16* __restrict__ dq, float* __restrict__ delta,
                    float* __restrict__ nullrow) {
  constexpr int NH = NSW / 4, CW = 128 / NH, NST = NSW * 32, OC = 64 / NH, CPT = 8 / NH;
  constexpr int KS = ATC2_DQ_KS;                    // a key tile is released only by dQ(j): with two stages the refill of
                                                    // tile j+2 (and with it S(j+2)) waited for a TMA round trip after dQ(j)
  constexpr uint32_t AUX = 32768 + KS * 16384 + 32768 + 65536;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  // Q 16K | dO 16K | K[KS] | V[2] 32K | dS[2] 64K | |k|^2 slabs [NSW][CW] 2K | null 512B | exchange [4][NH][128] <=8K | barriers
  const uint32_t sQ = base, sDO = base + 16384, sK = base + 32768, sV = sK + KS * 16384, sDS = sV + 32768;
  float* ksq_sm = (float*)(gbase + AUX);
  float* null_sm = (float*)(gbase + AUX + 2048);
  float* xchg = (float*)(gbase + AUX + 2560);
  const uint32_t bars = base + AUX + 10752;
  enum { Q_FULL = 0, K_FULL = 1, K_EMPTY = K_FULL + KS, V_FULL = K_EMPTY + KS, V_EMPTY = V_FULL + 2, S_FULL = V_EMPTY + 2,
         S_EMPTY = S_FULL + 2, DP_FULL = S_EMPTY + 2, DP_EMPTY = DP_FULL + 1, DS_FULL = DP_EMPTY + 1, DS_EMPTY = DS_FULL + 2,
         DQ_FULL = DS_EMPTY + 2, NBAR = DQ_FULL + 1 };
  auto bar = [&](int i) { return bars + 8u * i; };
  uint32_t* tmem_slot = (uint32_t*)(gbase + AUX + 10752 + 8 * NBAR);
  const int warp = tc_warp_idx(), lane = threadIdx.x & 31;
  const int qt = blockIdx.x % p.tiles, bh = blockIdx.x / p.tiles;
  const int b = bh / p.heads, h = bh % p.heads;
  const int T = p.tiles;
  if (threadIdx.x == 0) {
    mbar_init(bar(Q_FULL), 1);
    for (int i = 0; i < KS; ++i) { mbar_init(bar(K_FULL + i), 1); mbar_init(bar(K_EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(V_FULL + i), 1); mbar_init(bar(V_EMPTY + i), 1);
      mbar_init(bar(S_FULL + i), 1); mbar_init(bar(S_EMPTY + i), NSW);
      mbar_init(bar(DS_FULL + i), NSW); mbar_init(bar(DS_EMPTY + i), 1);
    }
    mbar_init(bar(DP_FULL), 1); mbar_init(bar(DP_EMPTY), NSW); mbar_init(bar(DQ_FULL), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x < 192 && p.has_null) {
    int t = threadIdx.x - 64;
    null_sm[t] = t < 64 ? null_kv[h * ATC_D + t] : null_kv[(p.heads + h) * ATC_D + (t - 64)];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tDP = tmem + 256, tDQ = tmem + 384;
  const uint32_t idesc_kk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc_dq = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  if (warp == 0) {
    const uint32_t el = tc_elect_one();
    mbar_expect_tx_el(bar(Q_FULL), 32768, el);
    tma_load_4d_el(sQ, &tmQ, bar(Q_FULL), 0, qt * ATC_T, h, b, el);
    tma_load_4d_el(sDO, &tmDO, bar(Q_FULL), 0, qt * ATC_T, h, b, el);
    for (int j = 0; j < T; ++j) {
      int s = j & 1, ks = j % KS;
      uint32_t par = ((j >> 1) & 1) ^ 1u;
      mbar_wait(bar(K_EMPTY + ks), ((j / KS) & 1) ^ 1u);
      mbar_expect_tx_el(bar(K_FULL + ks), 16384, el);
      tma_load_4d_el(sK + ks * 16384, &tmK, bar(K_FULL + ks), 0, j * ATC_T, h, b, el);
      mbar_wait(bar(V_EMPTY + s), par);
      mbar_expect_tx_el(bar(V_FULL + s), 16384, el);
      tma_load_4d_el(sV + s * 16384, &tmV, bar(V_FULL + s), 0, j * ATC_T, h, b, el);
    }
  } else if (warp == 1) {
    const uint32_t el = tc_elect_one();
    auto issue_S = [&](int j) {
      int s = j & 1, ks = j % KS;
      mbar_wait(bar(K_FULL + ks), (j / KS) & 1);
      mbar_wait(bar(S_EMPTY + s), ((j >> 1) & 1) ^ 1u);
      tc_fence_after();
      {
        uint64_t da = make_smem_desc(sQ, 1024, 2), db = make_smem_desc(sK + ks * 16384, 1024, 2);
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
      // the next tile's products first: the softmax warps release S / dP as soon as they hold them in registers, so these
      // run underneath tile j's exponentials; only then wait for tile j's dS
      if (j + 1 < T) { issue_S(j + 1); issue_dP(j + 1); }
      int s = j & 1, ks = j % KS;
      mbar_wait(bar(DS_FULL + s), (j >> 1) & 1);
      tc_fence_after();
      {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          uint64_t da = make_smem_desc(sDS + s * 32768 + (k >> 2) * 16384, 1024, 2) + (uint64_t)(2 * (k & 3));
          uint64_t db = make_smem_desc_mn(sK + ks * 16384 + k * 2048, 0, 1024);
          tc_mma_f16_el(tDQ, da, db, idesc_dq, (j | k) ? 1u : 0u, el);
        }
        tc_commit_el(bar(DS_EMPTY + s), el);
        tc_commit_el(bar(K_EMPTY + ks), el);
        if (j == T - 1) tc_commit_el(bar(DQ_FULL), el);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;                            // TMEM lane quarter
    const int hsel = (warp - 2) >> 2, cb = hsel * CW;  // column slab of this thread
    const int r = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const long grow = (long)b * p.n + qt * ATC_T + r;
    const long srow = (long)bh * p.n + qt * ATC_T + r;
    float* kslab = ksq_sm + (warp - 2) * CW;           // per-warp slab of the key tile's |k|^2 (see the forward kernel)
    float4 kq4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto ksq_fetch = [&](int j) {
      if (L2M && lane < CW / 4) kq4 = __ldg(reinterpret_cast<const float4*>(ksq + ((long)bh * p.n) + j * ATC_T + cb) + lane);
    };
    ksq_fetch(0);
    const float L2 = lse2[srow];
    uint4 ovr[CPT];                                    // this thread's chunks of the O row: in flight while Q / dO arrive
    {
      const bf16* orow = o + grow * p.o_rs + h * ATC_D;
#pragma unroll
      for (int cc = 0; cc < CPT; ++cc) ovr[cc] = __ldg(reinterpret_cast<const uint4*>(orow) + hsel * CPT + cc);
    }
    mbar_wait(bar(Q_FULL), 0);
    // streaming pass over this row of Q, dO (smem tiles) and O (global): delta = dO.O, q.k_null, dO.v_null - the NH threads
    // of a row take CPT 16-byte chunks each and exchange the partial sums
    float dl = 0.f, dot = 0.f, kn2 = 0.f, dpn = 0.f;
    {
      const uint8_t* qr = gbase + (sQ - base) + r * 128;
      const uint8_t* dr = gbase + (sDO - base) + r * 128;
#pragma unroll
      for (int cc = 0; cc < CPT; ++cc) {
        const int c = hsel * CPT + cc;
        uint4 qv = *reinterpret_cast<const uint4*>(qr + ((c ^ (r & 7)) << 4));
        uint4 dv = *reinterpret_cast<const uint4*>(dr + ((c ^ (r & 7)) << 4));
        uint4 ov = ovr[cc];
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
      xchg[(0 * NH + hsel) * 128 + r] = dl;
      xchg[(1 * NH + hsel) * 128 + r] = dot;
      xchg[(2 * NH + hsel) * 128 + r] = kn2;
      xchg[(3 * NH + hsel) * 128 + r] = dpn;
      named_bar_sync(2, NST);
      dl = dot = kn2 = dpn = 0.f;
#pragma unroll
      for (int hh = 0; hh < NH; ++hh) {
        dl += xchg[(0 * NH + hh) * 128 + r];
        dot += xchg[(1 * NH + hh) * 128 + r];
        kn2 += xchg[(2 * NH + hh) * 128 + r];
        dpn += xchg[(3 * NH + hh) * 128 + r];
      }
    }
    if (hsel == 0) delta[srow] = dl;
    float ds_null = 0.f, p_null = 0.f;
    if (p.has_null) {
      float tn = dot * p.c2 + (L2M ? p.kb2 * kn2 : 0.f);
      p_null = fast_exp2(tn - L2);
      ds_null = p_null * (dpn - dl) * p.ls;
      if (hsel == 0) {
        nullrow[srow] = ds_null;                       // consumed by attn_null_grad_kernel
        nullrow[(long)p.B * p.heads * p.n + srow] = p_null;
      }
    }
    const float negL2 = -L2, ndls = -dl * p.ls;
    for (int j = 0; j < T; ++j) {
      int s = j & 1;
      if (L2M) {
        __syncwarp();
        if (lane < CW / 4) reinterpret_cast<float4*>(kslab)[lane] = make_float4(kq4.x * p.kb2, kq4.y * p.kb2, kq4.z * p.kb2, kq4.w * p.kb2);
        __syncwarp();
        if (j + 1 < T) ksq_fetch(j + 1);
      }
      mbar_wait(bar(S_FULL + s), (j >> 1) & 1);
      mbar_wait(bar(DP_FULL), j & 1);
      tc_fence_after();
      uint32_t sv[CW], dv[CW];
      tc_ld_cols<CW>(tS + s * 128 + lane_addr + cb, sv);
      tc_ld_cols<CW>(tDP + lane_addr + cb, dv);
      tc_wait_cols<CW>(sv);
      tc_wait_cols<CW>(dv);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar(S_EMPTY + s)); mbar_arrive(bar(DP_EMPTY)); }   // both accumulators handed back early
      uint32_t pk[CW / 2];
      if (L2M) {
        float kk[CW];
        load_ksq_slab<CW>(kslab, kk);
#pragma unroll
        for (int e = 0; e < CW; e += 2) {
          float p0 = fast_exp2(fmaf(__uint_as_float(sv[e]), p.c2, kk[e]) + negL2);
          float p1 = fast_exp2(fmaf(__uint_as_float(sv[e + 1]), p.c2, kk[e + 1]) + negL2);
          pk[e >> 1] = pack_bf16x2(p0 * fmaf(__uint_as_float(dv[e]), p.ls, ndls), p1 * fmaf(__uint_as_float(dv[e + 1]), p.ls, ndls));
        }
      } else {
#pragma unroll
        for (int e = 0; e < CW; e += 2) {
          float p0 = fast_exp2(fmaf(__uint_as_float(sv[e]), p.c2, negL2));
          float p1 = fast_exp2(fmaf(__uint_as_float(sv[e + 1]), p.c2, negL2));
          pk[e >> 1] = pack_bf16x2(p0 * fmaf(__uint_as_float(dv[e]), p.ls, ndls), p1 * fmaf(__uint_as_float(dv[e + 1]), p.ls, ndls));
        }
      }
      mbar_wait(bar(DS_EMPTY + s), ((j >> 1) & 1) ^ 1u);
      uint8_t* dstile = gbase + (sDS - base) + s * 32768;
#pragma unroll
      for (int g = 0; g < CW / 16; ++g) write_tile16_packed(dstile, r, cb + 16 * g, pk + 8 * g);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(DS_FULL + s));
    }
    mbar_wait(bar(DQ_FULL), 0);
    tc_fence_after();
    bf16* dqrow = dq + grow * (long)(p.heads * ATC_D) + h * ATC_D;
#pragma unroll
    for (int c0 = hsel * OC; c0 < hsel * OC + OC; c0 += 16) {
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

// ================================================================================================= backward: dK, dV
template <int NSW, bool L2M>
__global__ void __launch_bounds__(64 + NSW * 32, 1)
attn2_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, const AtbP p,
                     const float* __restrict__ ksq, const float* __restrict__ lse2, const float* __restrict__ delta,
                     bf16* __restrict__ dk, bf16* __restrict__ dv) {
  constexpr int NH = NSW / 4, CW = 128 / NH, HALF = NH / 2, OC = 64 / HALF;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  constexpr int QS = ATC2_DKV_QS;                   // query / dO stages: a stage is released only by dV/dK(i); with two stages
                                                    // S/dP(i+2) waited for a TMA round trip behind them (28 % of the softmax
                                                    // warps' samples in ncu).  The block of ones (columns 64..79 of the dK
                                                    // product's B operand, reached through the descriptor's LBO) is shared.
  // K 16K | V 16K | Q[QS] | ones 16K | dO[QS] | P 32K | dS 32K | ksq 512B | barriers
  const uint32_t sK = base, sV = base + 16384, sQ = base + 32768, sOnes = sQ + QS * 16384, sDO = sOnes + 16384,
                 sP = sDO + QS * 16384, sDS = sP + 32768;
  float* ksq_sm = (float*)(gbase + ATC2_DKV_BARS - 512);
  const uint32_t bars = base + ATC2_DKV_BARS;
  enum { KV_FULL = 0, QO_FULL = 1, QO_EMPTY = QO_FULL + QS, SDP_FULL = QO_EMPTY + QS, SDP_EMPTY = SDP_FULL + 1,
         PDS_FULL = SDP_EMPTY + 1, PDS_EMPTY = PDS_FULL + 1, OUT_FULL = PDS_EMPTY + 1, NBAR = OUT_FULL + 1 };
  auto bar = [&](int i) { return bars + 8u * i; };
  uint32_t* tmem_slot = (uint32_t*)(gbase + ATC2_DKV_BARS + 8 * NBAR);
  const int warp = tc_warp_idx(), lane = threadIdx.x & 31;
  const int kt = blockIdx.x % p.tiles, bh = blockIdx.x / p.tiles;
  const int b = bh / p.heads, h = bh % p.heads;
  const int T = p.tiles;
  if (threadIdx.x == 0) {
    mbar_init(bar(KV_FULL), 1);
    for (int i = 0; i < QS; ++i) { mbar_init(bar(QO_FULL + i), 1); mbar_init(bar(QO_EMPTY + i), 1); }
    mbar_init(bar(SDP_FULL), 1); mbar_init(bar(SDP_EMPTY), NSW);
    mbar_init(bar(PDS_FULL), NSW); mbar_init(bar(PDS_EMPTY), 1); mbar_init(bar(OUT_FULL), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {   // the "ones" slab (bf16 1.0 everywhere; swizzle-invariant)
    uint4 one4 = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    for (int i = threadIdx.x; i < 1024; i += 64 + NSW * 32)
      *reinterpret_cast<uint4*>(gbase + (sOnes - base) + (i << 4)) = one4;
    fence_async_smem();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 192 && L2M) ksq_sm[threadIdx.x - 64] = ksq[(long)bh * p.n + kt * ATC_T + (threadIdx.x - 64)] * p.kb2;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tDP = tmem + 128, tDV = tmem + 256, tDK = tmem + 320;
  const uint32_t idesc_kk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc_dv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc_dk = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(80 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  if (warp == 0) {
    const uint32_t el = tc_elect_one();
    mbar_expect_tx_el(bar(KV_FULL), 32768, el);
    tma_load_4d_el(sK, &tmK, bar(KV_FULL), 0, kt * ATC_T, h, b, el);
    tma_load_4d_el(sV, &tmV, bar(KV_FULL), 0, kt * ATC_T, h, b, el);
    for (int i = 0; i < T; ++i) {
      int s = i % QS;
      mbar_wait(bar(QO_EMPTY + s), ((i / QS) & 1) ^ 1u);
      mbar_expect_tx_el(bar(QO_FULL + s), 32768, el);
      tma_load_4d_el(sQ + s * 16384, &tmQ, bar(QO_FULL + s), 0, i * ATC_T, h, b, el);
      tma_load_4d_el(sDO + s * 16384, &tmDO, bar(QO_FULL + s), 0, i * ATC_T, h, b, el);
    }
  } else if (warp == 1) {
    const uint32_t el = tc_elect_one();
    auto issue_SdP = [&](int i) {
      int s = i % QS;
      mbar_wait(bar(QO_FULL + s), (i / QS) & 1);
      mbar_wait(bar(SDP_EMPTY), (i & 1) ^ 1u);
      tc_fence_after();
      {
        uint64_t dq_ = make_smem_desc(sQ + s * 16384, 1024, 2), dk_ = make_smem_desc(sK, 1024, 2);
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
      if (i + 1 < T) issue_SdP(i + 1);      // needs only the early release of S / dP by the softmax warps of tile i
      int s = i % QS;
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
          uint64_t bq = make_smem_desc_mn(sQ + s * 16384 + k * 2048, (uint32_t)(QS - s) * 16384u, 1024);   // LBO -> the ones
          tc_mma_f16_el(tDK, as_, bq, idesc_dk, (i | k) ? 1u : 0u, el);
        }
        tc_commit_el(bar(PDS_EMPTY), el);
        tc_commit_el(bar(QO_EMPTY + s), el);
        if (i == T - 1) tc_commit_el(bar(OUT_FULL), el);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2, cb = hsel * CW;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint8_t* ptile = gbase + (sP - base);
    uint8_t* dstile = gbase + (sDS - base);
    const long srow0 = (long)bh * p.n + r;
    float L2n = lse2[srow0], dln = delta[srow0];       // tile 0; later tiles are fetched one tile ahead
    for (int i = 0; i < T; ++i) {
      const float negL2 = -L2n, ndls = -dln * p.ls;
      mbar_wait(bar(SDP_FULL), i & 1);
      tc_fence_after();
      uint32_t sv[CW], dv_[CW];
      tc_ld_cols<CW>(tS + lane_addr + cb, sv);
      tc_ld_cols<CW>(tDP + lane_addr + cb, dv_);
      tc_wait_cols<CW>(sv);
      tc_wait_cols<CW>(dv_);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(SDP_EMPTY));
      if (i + 1 < T) { L2n = lse2[srow0 + (long)(i + 1) * ATC_T]; dln = delta[srow0 + (long)(i + 1) * ATC_T]; }
      uint32_t pk[CW / 2], dk_[CW / 2];
      float kk[L2M ? CW : 1];
      if (L2M) load_ksq_slab<CW>(ksq_sm + cb, kk);     // this key tile's |k|^2 * kb2 (constant over the query tiles)
#pragma unroll
      for (int e = 0; e < CW; e += 2) {
        float t0 = L2M ? fmaf(__uint_as_float(sv[e]), p.c2, kk[L2M ? e : 0]) + negL2 : fmaf(__uint_as_float(sv[e]), p.c2, negL2);
        float t1 = L2M ? fmaf(__uint_as_float(sv[e + 1]), p.c2, kk[L2M ? e + 1 : 0]) + negL2 : fmaf(__uint_as_float(sv[e + 1]), p.c2, negL2);
        float p0 = fast_exp2(t0), p1 = fast_exp2(t1);
        pk[e >> 1] = pack_bf16x2(p0, p1);
        dk_[e >> 1] = pack_bf16x2(p0 * fmaf(__uint_as_float(dv_[e]), p.ls, ndls), p1 * fmaf(__uint_as_float(dv_[e + 1]), p.ls, ndls));
      }
      mbar_wait(bar(PDS_EMPTY), (i & 1) ^ 1u);
#pragma unroll
      for (int g = 0; g < CW / 16; ++g) {
        write_tile16_packed(ptile, r, cb + 16 * g, pk + 8 * g);
        write_tile16_packed(dstile, r, cb + 16 * g, dk_ + 8 * g);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(PDS_FULL));
    }
    // epilogue: thread <-> key row; the first HALF column groups store dV, the others dK
    mbar_wait(bar(OUT_FULL), 0);
    tc_fence_after();
    const long grow = (long)b * p.n + kt * ATC_T + r;
    const int part = hsel % HALF;
    if (hsel < HALF) {
      bf16* dvrow = dv + grow * (long)(p.heads * ATC_D) + h * ATC_D;
#pragma unroll
      for (int c0 = part * OC; c0 < part * OC + OC; c0 += 16) {
        uint32_t v[16];
        float f[16];
        tc_ld16(tDV + lane_addr + c0, v);
#pragma unroll
        for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(v[e]);
        store_row16(dvrow + c0, f);
      }
    } else {
      bf16* dkrow = dk + grow * (long)(p.heads * ATC_D) + h * ATC_D;
      float csum = 0.f;
      if (L2M) {
        uint32_t v[16];
        tc_ld16(tDK + lane_addr + 64, v);
        csum = __uint_as_float(v[0]);
      }
      const uint8_t* krow = gbase + (sK - base) + r * 128;
#pragma unroll
      for (int c0 = part * OC; c0 < part * OC + OC; c0 += 16) {
        uint32_t v[16];
        float f[16];
        tc_ld16(tDK + lane_addr + c0, v);
#pragma unroll
        for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(v[e]);
        if (L2M) {
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            uint4 kv = *reinterpret_cast<const uint4*>(krow + ((((c0 >> 3) + cc) ^ (r & 7)) << 4));
            const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&kv);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float2 kf = __bfloat1622float2(hp[e]);
              f[cc * 8 + 2 * e] -= csum * kf.x;
              f[cc * 8 + 2 * e + 1] -= csum * kf.y;
            }
          }
        }
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

// ================================================================================================= host side
static bool atc2_eligible(int nq, int nk, int d, long q_rs, long k_rs, long v_rs) {
  return d == ATC_D && nq == nk && nq % ATC_T == 0 && nq >= ATC_T && !(q_rs % 8) && !(k_rs % 8) && !(v_rs % 8);
}

template <int NSW, bool L2M, bool ONEP, int OCC>
static int atc2_launch_fwd(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const AtcP& p,
                           const float* null_kv, const float* ksq_ws, void* o, float* lse, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(attn2_fwd_kernel<NSW, L2M, ONEP, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  constexpr int KS = OCC == 2 ? 2 : ATC2_FWD_KS, NB = OCC == 2 ? 1 : 2;
  size_t smem = 1024 + (16384 + KS * 16384 + NB * 16384 + NB * 32768) + 4608 + 8 * (2 * KS + 6 * NB + 2) + 16;
  attn2_fwd_kernel<NSW, L2M, ONEP, OCC><<<p.B * p.heads * p.tiles, 64 + NSW * 32, smem, st>>>(tmQ, tmK, tmV, p, null_kv, ksq_ws, (bf16*)o, lse);
  return gg_check_launch("attn2_fwd");
}

// returns 1 when the shape is not eligible (caller falls back to the FFMA kernel).  nsw: 8 or 16 softmax warps.
int ggi_tc2_attn_fwd(const void* q, const void* k, const void* v, const float* null_kv, void* o, float* lse, float* ksq_ws,
                     int B, int heads, int nq, int nk, int d, long q_rs, long k_rs, long v_rs, long o_rs, float scale,
                     int mode, int nsw, cudaStream_t st) {
  if (!atc2_eligible(nq, nk, d, q_rs, k_rs, v_rs) || (o_rs % 16)) return 1;
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
  if (mode == 1) atc_launch_ksq(k, ksq_ws, B, nk, heads, k_rs, st);
  // the single-pass form needs the keys to BE the queries (the discriminator's shared-QK attention); gg_set_flags bit 6
  // (passed in as nsw + 64) keeps the two-pass kernel for A/B measurements
  const bool onep = mode == 1 && q == k && q_rs == k_rs && !(nsw & 64);
  // default: two CTAs per SM with 8 softmax warps each (1.3x the one-CTA forms at every measured shape); gg_set_flags
  // bits 4 / 5 / 7 select the one-CTA-per-SM kernels with 8 / 16 / 8 softmax warps (A/B measurements)
  const bool occ2 = !(nsw & 128) && !(nsw & 256);
  nsw &= 63;
#define ATC2_FWD(NSW_, L2_, ONEP_, OCC_) atc2_launch_fwd<NSW_, L2_, ONEP_, OCC_>(tmQ, tmK, tmV, p, null_kv, ksq_ws, o, lse, st)
  if (occ2) return mode != 1 ? ATC2_FWD(8, false, false, 2) : onep ? ATC2_FWD(8, true, true, 2) : ATC2_FWD(8, true, false, 2);
  if (nsw == 8) return mode != 1 ? ATC2_FWD(8, false, false, 1) : onep ? ATC2_FWD(8, true, true, 1) : ATC2_FWD(8, true, false, 1);
  return mode != 1 ? ATC2_FWD(16, false, false, 1) : onep ? ATC2_FWD(16, true, true, 1) : ATC2_FWD(16, true, false, 1);
#undef ATC2_FWD
}

template <int NSW, bool L2M>
static int atc2_launch_bwd(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const CUtensorMap& tmDO,
                           const AtbP& p, const void* q, const float* null_kv, const void* o, const void* go,
                           const float* lse2, void* dq, void* dk, void* dv, float* dnull_kv, float* delta_ws, float* ksq_ws,
                           long q_rs, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(attn2_bwd_dq_kernel<NSW, L2M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(attn2_bwd_dkv_kernel<NSW, L2M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  const int grid = p.B * p.heads * p.tiles, threads = 64 + NSW * 32;
  size_t smem1 = 1024 + (32768 + ATC2_DQ_KS * 16384 + 32768 + 65536) + 10752 + 8 * (2 * ATC2_DQ_KS + 16) + 16;
  size_t smem2 = 1024 + ATC2_DKV_BARS + 8 * (2 * ATC2_DKV_QS + 6) + 16;
  float* nullrow = delta_ws + (size_t)p.B * p.heads * p.n;
  attn2_bwd_dq_kernel<NSW, L2M><<<grid, threads, smem1, st>>>(tmQ, tmK, tmV, tmDO, p, null_kv, ksq_ws, (const bf16*)o, lse2, (bf16*)dq, delta_ws, nullrow);
  if (p.has_null) atc_launch_null_grad(q, go, nullrow, null_kv, dnull_kv, p.B, p.n, p.heads, q_rs, p.mode, st);
  attn2_bwd_dkv_kernel<NSW, L2M><<<grid, threads, smem2, st>>>(tmQ, tmK, tmV, tmDO, p, ksq_ws, lse2, delta_ws, (bf16*)dk, (bf16*)dv);
  return gg_check_launch("attn2_bwd");
}

// go, dq, dk, dv: dense (B, n, heads*64); delta_ws holds 3*B*heads*n floats.  Returns 1 when not eligible.
int ggi_tc2_attn_bwd(const void* q, const void* k, const void* v, const float* null_kv, const void* o, const void* go,
                     const float* lse2, void* dq, void* dk, void* dv, float* dnull_kv, float* delta_ws, float* ksq_ws,
                     int B, int heads, int nq, int nk, int d, long q_rs, long k_rs, long v_rs, long o_rs, float scale,
                     int mode, int nsw, cudaStream_t st) {
  if (!atc2_eligible(nq, nk, d, q_rs, k_rs, v_rs) || (o_rs % 8)) return 1;
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
  if (mode == 1) atc_launch_ksq(k, ksq_ws, B, nk, heads, k_rs, st);
  if (p.has_null) cudaMemsetAsync(dnull_kv, 0, sizeof(float) * 2 * heads * ATC_D, st);
#define ATC2_BWD(NSW_, L2_) atc2_launch_bwd<NSW_, L2_>(tmQ, tmK, tmV, tmDO, p, q, null_kv, o, go, lse2, dq, dk, dv, dnull_kv, delta_ws, ksq_ws, q_rs, st)
  if (nsw == 8) return mode == 1 ? ATC2_BWD(8, true) : ATC2_BWD(8, false);
  return mode == 1 ? ATC2_BWD(16, true) : ATC2_BWD(16, false);
#undef ATC2_BWD
}
