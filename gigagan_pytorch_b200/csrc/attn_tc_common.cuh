// Shared pieces of the fused-attention tensor-core kernels (attn_tc.cu: 8 softmax warps, chunked TMEM reads;
// attn_tc2.cu: 16 softmax warps, whole-slab TMEM reads with early accumulator release).
#pragma once
#include "tc_common.cuh"

#define ATC_THREADS 192
#define ATC_FWD_THREADS 320   // forward: TMA warp + MMA warp + 8 softmax warps (two per TMEM lane quarter)
#define ATC_BWD_THREADS 320   // backward kernels: same split (each softmax thread owns 64 of the 128 tile columns)
#define ATC_D 64
#define ATC_T 128          // queries per CTA == keys per tile

struct AtcP {
  int B, heads, n, tiles;          // tiles = n / 128
  int mode, has_null;
  float c2;                        // logit scale * log2(e)          (dot: scale, l2: 2*scale)
  float kb2;                       // l2: -scale*log2(e), applied to |k|^2
  long o_rs;                       // row stride of O (elements)
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct AtbP {
  int B, heads, n, tiles, mode, has_null;
  float c2, kb2, ls;
  long o_rs;
};

// read the 64 bf16 of row r of a [128 x 64] SWIZZLE_128B K-major tile
__device__ __forceinline__ void read_tile_row(const uint8_t* tile, int r, float* out) {
  const uint8_t* row = tile + r * 128;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint4 v = *reinterpret_cast<const uint4*>(row + ((c ^ (r & 7)) << 4));
    const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int e = 0; e < 4; ++e) { float2 f = __bfloat1622float2(hp[e]); out[c * 8 + 2 * e] = f.x; out[c * 8 + 2 * e + 1] = f.y; }
  }
}
// write 16 consecutive columns [c0, c0+16) of row r of a [128 rows x 128 cols] bf16 tile stored as two 64-column
// SWIZZLE_128B slabs (16 KB apart)
__device__ __forceinline__ void write_tile16(uint8_t* tile, int r, int c0, const float* v) {
  uint4 o0, o1;
  __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&o0);
  __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    h0[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    h1[e] = __floats2bfloat162_rn(v[8 + 2 * e], v[8 + 2 * e + 1]);
  }
  int slab = c0 >> 6, ch = (c0 & 63) >> 3;
  uint8_t* dst = tile + slab * 16384 + r * 128;
  *reinterpret_cast<uint4*>(dst + ((ch ^ (r & 7)) << 4)) = o0;
  *reinterpret_cast<uint4*>(dst + (((ch + 1) ^ (r & 7)) << 4)) = o1;
}
__device__ __forceinline__ void store_row16(bf16* dst, const float* f) {
  uint4 o0, o1;
  __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&o0);
  __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    h0[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    h1[e] = __floats2bfloat162_rn(f[8 + 2 * e], f[8 + 2 * e + 1]);
  }
  st_global_256(dst, o0, o1);
}

static inline int make_qkv_map(CUtensorMap* m, const void* ptr, int B, int n, int heads, long rs) {
  uint64_t dims[4] = {ATC_D, (uint64_t)n, (uint64_t)heads, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)rs * 2, (uint64_t)ATC_D * 2, (uint64_t)n * rs * 2};
  uint32_t box[4] = {ATC_D, ATC_T, 1, 1};
  return tc_make_map4(m, ptr, dims, strides, box, 128);
}

// launch wrappers around the helper kernels defined in attn_tc.cu
void atc_launch_ksq(const void* k, float* ksq_ws, int B, int n, int heads, long k_rs, cudaStream_t st);
void atc_launch_null_grad(const void* q, const void* go, const float* nullrow, const float* null_kv, float* dnull,
                          int B, int n, int heads, long q_rs, int mode, cudaStream_t st);
