// tcgen05 strided batched GEMM (bf16 in, fp32 TMEM accumulate, bf16 out):  C[z] = alpha * A[z] @ B[z] (+ bias)
// Operands are arbitrary strided views; each may be K-major (reduction axis contiguous) or MN-major (row/column
// axis contiguous) - exactly the four cases that the attention products and their first/second derivatives
// produce from (batch, tokens, heads, dim_head) activations without any transposition copies:
//   S = Q K^T (K-major x K-major), O = P V (K-major x MN-major), dV = P^T dO (MN-major x MN-major), ...
// Reference call sites: torch.einsum at gigagan_pytorch.py:574,:579,:590 and their autograd.
#include "tc_common.cuh"

struct BmmTcP {
  int M, N, K, b2, batches;
  int a_mn, b_mn;
  int Ntile, nsub_b, m_tiles, n_tiles, total_tiles, kblocks;
  int stages, stage_bytes, a_bytes, b_bytes;
  long sC1, sC2, rsC;
  float alpha;
  uint32_t idesc, tmem_cols;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
bmm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const BmmTcP p,
              const float* __restrict__ bias, bf16* __restrict__ C) {
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
  const int tiles_per_batch = p.m_tiles * p.n_tiles;

  if (warp == 0) {
    {
      const uint32_t el = tc_elect_one();          // convergent producer: only the TMA / expect_tx instructions are predicated
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int z = tile / tiles_per_batch, r = tile - z * tiles_per_batch;
        int mt = r / p.n_tiles, nt = r - mt * p.n_tiles;
        int z1 = z / p.b2, z2 = z - z1 * p.b2;
        int m0 = mt * 128, n0 = nt * p.Ntile;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx_el(full_bar(stage), (uint32_t)(p.a_bytes + p.b_bytes), el);
          uint32_t sa = base + stage * p.stage_bytes, sb = sa + p.a_bytes;
          if (!p.a_mn) tma_load_4d_el(sa, &tmA, full_bar(stage), kb * 64, m0, z2, z1, el);
          else {
            tma_load_4d_el(sa, &tmA, full_bar(stage), m0, kb * 64, z2, z1, el);
            tma_load_4d_el(sa + 8192, &tmA, full_bar(stage), m0 + 64, kb * 64, z2, z1, el);
          }
          if (!p.b_mn) tma_load_4d_el(sb, &tmB, full_bar(stage), kb * 64, n0, z2, z1, el);
          else
            for (int s = 0; s < p.nsub_b; ++s) tma_load_4d_el(sb + s * 8192, &tmB, full_bar(stage), n0 + s * 64, kb * 64, z2, z1, el);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t el = tc_elect_one();            // the lane that issues tcgen05.mma / commit (warp stays convergent)
    int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.Ntile);
      for (int kb = 0; kb < p.kblocks; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        {
          uint32_t sa = base + stage * p.stage_bytes, sb = sa + p.a_bytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint64_t da = p.a_mn ? make_smem_desc_mn(sa + k * 2048, 8192, 1024) : make_smem_desc(sa, 1024, 2) + (uint64_t)(k * 2);
            uint64_t db = p.b_mn ? make_smem_desc_mn(sb + k * 2048, 8192, 1024) : make_smem_desc(sb, 1024, 2) + (uint64_t)(k * 2);
            tc_mma_f16_el(d_tmem, da, db, p.idesc, (kb | k) != 0 ? 1u : 0u, el);
          }
          tc_commit_el(empty_bar(stage), el);
          if (kb == p.kblocks - 1) tc_commit_el(tfull_bar(acc), el);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;
    const bool vec_ok = (p.rsC % 16 == 0) && (p.sC1 % 16 == 0) && (p.sC2 % 16 == 0) && (((uintptr_t)C & 31) == 0);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int z = tile / tiles_per_batch, r = tile - z * tiles_per_batch;
      int mt = r / p.n_tiles, nt = r - mt * p.n_tiles;
      int z1 = z / p.b2, z2 = z - z1 * p.b2;
      int m = mt * 128 + row, n0 = nt * p.Ntile;
      bool live = m < p.M;
      bf16* crow = C + z1 * p.sC1 + z2 * p.sC2 + (long)m * p.rsC;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.Ntile);
      for (int c0 = 0; c0 < p.Ntile; c0 += 16) {
        uint32_t r16[16];
        tc_ld16(taddr + c0, r16);
        int n = n0 + c0;
        if (live && n < p.N) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r16[j]) * p.alpha;
          if (bias) {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (n + j < p.N) v[j] += __ldg(bias + n + j);
          }
          if (vec_ok && n + 16 <= p.N) {
            uint4 o0, o1;
            __nv_bfloat162* ob0 = (__nv_bfloat162*)&o0; __nv_bfloat162* ob1 = (__nv_bfloat162*)&o1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              ob0[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
              ob1[j] = __floats2bfloat162_rn(v[8 + 2 * j], v[8 + 2 * j + 1]);
            }
            st_global_256(crow + n, o0, o1);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (n + j < p.N) crow[n + j] = __float2bfloat16_rn(v[j]);
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

// returns 1 if not eligible.  sa/sb: {batch1, batch2, row, col} element strides; B indexed [k][n].
int ggi_tc_bmm(const void* A, const void* B, const float* bias, void* C, int b1, int b2, int M, int N, int K,
               const long* sa, const long* sb, const long* sc, float alpha, cudaStream_t st) {
  if (M < 16 || N < 16 || K < 16) return 1;      // short M: TMA zero-fills the rest of the 128-row tile
  int a_mn, b_mn;
  long a_outer, b_outer;
  if (sa[3] == 1 && sa[2] % 8 == 0) { a_mn = 0; a_outer = sa[2]; }            // K contiguous
  else if (sa[2] == 1 && sa[3] % 8 == 0) { a_mn = 1; a_outer = sa[3]; }       // M contiguous
  else return 1;
  if (sb[2] == 1 && sb[3] % 8 == 0) { b_mn = 0; b_outer = sb[3]; }            // K contiguous (B is [k][n])
  else if (sb[3] == 1 && sb[2] % 8 == 0) { b_mn = 1; b_outer = sb[2]; }       // N contiguous
  else return 1;
  if ((sa[0] % 8) || (sa[1] % 8) || (sb[0] % 8) || (sb[1] % 8)) return 1;
  if (((uintptr_t)A | (uintptr_t)B) & 15) return 1;
  if (a_outer <= 0 || b_outer <= 0) return 1;
  BmmTcP p;
  p.M = M; p.N = N; p.K = K; p.b2 = b2; p.batches = b1 * b2; p.a_mn = a_mn; p.b_mn = b_mn;
  int Ntile = N <= 64 ? 64 : N <= 128 ? 128 : (N % 256 == 0 || (N % 256) > 128) ? 256 : 128;
  p.Ntile = Ntile; p.nsub_b = Ntile / 64;
  p.m_tiles = (M + 127) / 128; p.n_tiles = (N + Ntile - 1) / Ntile;
  p.total_tiles = p.m_tiles * p.n_tiles * p.batches;
  p.kblocks = (K + 63) / 64;
  p.a_bytes = 16384; p.b_bytes = Ntile * 128; p.stage_bytes = p.a_bytes + p.b_bytes;
  int stages = (200 * 1024) / p.stage_bytes;
  p.stages = stages > TC_MAX_STAGES ? TC_MAX_STAGES : stages;
  p.sC1 = sc[0]; p.sC2 = sc[1]; p.rsC = sc[2]; p.alpha = alpha;
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
            ((uint32_t)(Ntile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.tmem_cols = 2 * Ntile;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)(a_mn ? M : K), (uint64_t)(a_mn ? K : M), (uint64_t)b2, (uint64_t)b1};
    uint64_t strides[3] = {(uint64_t)a_outer * 2, (uint64_t)(b2 > 1 ? sa[1] : a_outer * dims[1]) * 2,
                           (uint64_t)(b1 > 1 ? sa[0] : (b2 > 1 ? sa[1] * b2 : a_outer * dims[1])) * 2};
    uint32_t box[4] = {64, (uint32_t)(a_mn ? 64 : 128), 1, 1};
    if (tc_make_map4(&tmA, A, dims, strides, box, 128)) return -1;
  }
  {
    uint64_t dims[4] = {(uint64_t)(b_mn ? N : K), (uint64_t)(b_mn ? K : N), (uint64_t)b2, (uint64_t)b1};
    uint64_t strides[3] = {(uint64_t)b_outer * 2, (uint64_t)(b2 > 1 ? sb[1] : b_outer * dims[1]) * 2,
                           (uint64_t)(b1 > 1 ? sb[0] : (b2 > 1 ? sb[1] * b2 : b_outer * dims[1])) * 2};
    uint32_t box[4] = {64, (uint32_t)(b_mn ? 64 : Ntile), 1, 1};
    if (tc_make_map4(&tmB, B, dims, strides, box, 128)) return -1;
  }
  size_t smem = 1024 + (size_t)p.stages * p.stage_bytes + 8 * (2 * TC_MAX_STAGES + 4) + 16;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(bmm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  int grid = p.total_tiles < tc_num_sms() ? p.total_tiles : tc_num_sms();
  bmm_tc_kernel<<<grid, TC_THREADS, smem, st>>>(tmA, tmB, p, bias, (bf16*)C);
  return gg_check_launch("bmm_tc");
}
