// Microbenchmark (debug only): how long does a chain of small tcgen05.mma take when consecutive MMAs accumulate into the
// SAME TMEM tile versus alternating between `nacc` tiles?  One CTA, one warp; operands are zero-filled SWIZZLE_128B
// K-major slabs (128 x 64 bf16 A, N x 64 bf16 B), every MMA is M=128, N, K=16.  Result: cycles for `iters` MMAs.
#include "tc_common.cuh"

__global__ void __launch_bounds__(32, 1)
debug_mma_chain_kernel(int N, int nacc, int iters, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* g = smem_raw + (base - raw);
  for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += 32) reinterpret_cast<uint4*>(g)[i] = make_uint4(0, 0, 0, 0);
  const uint32_t bar = base + 16384 + 32768;
  uint32_t* tmem_slot = (uint32_t*)(g + 16384 + 32768 + 16);
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncwarp();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t da = make_smem_desc(base, 1024, 2), db = make_smem_desc(base + 16384, 1024, 2);
  const uint32_t el = tc_elect_one();
  long long t0 = clock64();
  int a = 0;
  for (int i = 0; i < iters; ++i) {
    tc_mma_f16_el(tmem + (uint32_t)(a * N), da + (uint64_t)(2 * (i & 3)), db + (uint64_t)(2 * (i & 3)), idesc, i >= nacc ? 1u : 0u, el);
    if (++a == nacc) a = 0;
  }
  long long t1 = clock64();
  tc_commit_el(bar, el);
  mbar_wait(bar, 0);
  long long t2 = clock64();
  if (threadIdx.x == 0) { out[0] = (unsigned long long)(t1 - t0); out[1] = (unsigned long long)(t2 - t0); }
  tc_fence_before();
  __syncwarp();
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

// out[0] = cycles to ISSUE the chain, out[1] = cycles until the last MMA completed
int ggi_debug_mma_chain(int N, int nacc, int iters, unsigned long long* out, cudaStream_t st) {
  if (N < 16 || N > 256 || N % 16 || nacc < 1 || nacc * N > 512 || iters < 1) return gg_fail("debug_mma_chain: bad arguments");
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(debug_mma_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); attr_set = true; }
  debug_mma_chain_kernel<<<1, 32, 1024 + 16384 + 32768 + 64, st>>>(N, nacc, iters, out);
  return gg_check_launch("debug_mma_chain");
}
