// Shared pieces of the two attention kernels (reference models/layers.py:368-415; 2x6x12 = 144-token windows,
// head_dim 32): the launch arguments, the swizzled [144 x 32] 16-bit tile addressing, and the warp-level
// mma.sync / ldmatrix / cp.async helpers used by the rows 128..143 tail of the tcgen05 forward kernel
// (attention_tc.cuh) and by the backward kernel (attention_bwd.cuh).  The round-1 mma.sync forward kernel that
// lived here was superseded by attention_tc.cuh and has been removed.
#pragma once
#include "common.cuh"
#include "geometry.cuh"

namespace pg {

constexpr int ATT_TOK = 144;
constexpr int ATT_D = 32;
constexpr int ATT_TILE_BYTES = ATT_TOK * ATT_D * 2;   // 9216
constexpr int ATT_BUF_BYTES = 3 * ATT_TILE_BYTES;     // q, k, v

struct AttnArgs {
  const void* qkv;     // [nLon*types*144][3C] 16-bit, window order, q pre-scaled
  const float* bias;   // [types][heads][144][144] fp32 (earth_specific_bias parameter)
  void* out;           // [Z*H*W tokens][C] 16-bit, natural order (or window order, see `natural`)
  int C, heads, types, nLon, nH;
  int H, W;            // token grid (natural-order output of the tcgen05 kernel)
  int natural;         // 1: out rows are natural tokens (pad rows dropped); 0: window order, all rows
  int roll;            // add the shifted-window mask
  int plane_rows;      // rows per (q|k|v, head) plane of the head-major qkv buffer
  int debug;           // development only: bit0 skip tail math, bit1 skip softmax math, bit2 skip stores
  long long* trace;    // development only: per-role clock64 timeline of CTA (0,0), [role 8][window 32][event 4]
};

// swizzled byte offset of 16-byte chunk c (0..3) of row r in a [144][32] 16-bit tile
__device__ __forceinline__ uint32_t att_off(int r, int c) { return uint32_t(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
template <bool kFp16>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (kFp16) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

}  // namespace pg
