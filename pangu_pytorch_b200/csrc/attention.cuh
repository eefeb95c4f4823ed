// Earth-specific 3-D window attention (reference models/layers.py:368-415) for the
// 2x6x12 = 144-token windows, head_dim 32.
//
//   S = (q*scale) k^T + bias[type, head] (+ shift mask)  ->  softmax  ->  P v
//
// One CTA owns one (window type, head) pair and walks the longitude windows that share
// that bias tile; the 144x144 fp32 bias tile (with the 0/-100 shifted-window mask folded
// in) lives in REGISTERS in accumulator-fragment layout for the whole
// CTA lifetime, so the 253 M bias parameters are read exactly once per block.  q/k/v tiles
// stream through a double-buffered cp.async ring; S and P never leave registers.
//
// v1 uses warp-level mma.sync (m16n8k16) - attention here is exp/HBM bound, not tensor
// bound (hard part 2); the tcgen05 variant is tracked in DESIGN.md.
#pragma once
#include "common.cuh"
#include "geometry.cuh"

namespace pg {

constexpr int ATT_TOK = 144;
constexpr int ATT_D = 32;
constexpr int ATT_WARPS = 9;                 // 9 x 16 query rows
constexpr int ATT_THREADS = ATT_WARPS * 32;  // 288
constexpr int ATT_TILE_BYTES = ATT_TOK * ATT_D * 2;   // 9216
constexpr int ATT_BUF_BYTES = 3 * ATT_TILE_BYTES;     // q, k, v
constexpr int ATT_STAGES = 3;                          // cp.async ring depth (one barrier per window)
constexpr int ATT_SMEM_BYTES = ATT_STAGES * ATT_BUF_BYTES;

struct AttnArgs {
  const void* qkv;     // [nLon*types*144][3C] 16-bit, window order, q pre-scaled
  const float* bias;   // [types][heads][144][144] fp32 (earth_specific_bias parameter)
  void* out;           // tcgen05 kernel: [Z*H*W tokens][C] 16-bit, NATURAL order; v1 kernel: window order
  int C, heads, types, nLon, nH;
  int H, W;            // token grid (natural-order output of the tcgen05 kernel)
  int natural;         // 1: out rows are natural tokens (pad rows dropped); 0: window order, all rows
  int roll;            // add the shifted-window mask
  int lon_per_cta;     // (mma.sync v1 kernel) longitude windows walked by one CTA
  int plane_rows;      // (tcgen05 kernel) rows per (q|k|v, head) plane of the head-major qkv buffer
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

template <bool kFp16>
__global__ void __launch_bounds__(ATT_THREADS, 1) window_attention_kernel(const AttnArgs a) {
  extern __shared__ __align__(128) uint8_t att_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q4 = lane & 3;
  const int th = blockIdx.x;                // type * heads + head
  const int t = th / a.heads, head = th % a.heads;
  const int lw0 = blockIdx.y * a.lon_per_cta;
  const int lw1 = min(a.nLon, lw0 + a.lon_per_cta);
  if (lw0 >= lw1) return;

  const size_t pitch = size_t(3) * a.C * 2;   // bytes per qkv row
  const uint8_t* qkv = reinterpret_cast<const uint8_t*>(a.qkv);
  const uint32_t sbase = smem_u32(att_smem);

  auto prefetch = [&](int lw, int buf) {
    const size_t row0 = (size_t(lw) * a.types + t) * ATT_TOK;
    // 3 matrices x 144 rows x 4 chunks = 1728 chunks, 6 per thread
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int id = i * ATT_THREADS + threadIdx.x;
      const int m = id / 576, rem = id % 576;
      const int r = rem >> 2, c = rem & 3;
      const uint8_t* src = qkv + (row0 + r) * pitch + size_t(m) * a.C * 2 + head * 64 + c * 16;
      cp_async16(sbase + buf * ATT_BUF_BYTES + m * ATT_TILE_BYTES + att_off(r, c), src);
    }
    cp_async_commit();
  };

  prefetch(lw0, 0);
  if (lw0 + 1 < lw1) prefetch(lw0 + 1, 1);

  // ---- bias (+mask) tile into registers, accumulator-fragment layout:
  //      bz[j][0..1] = row r0, cols 8j + 2*q4 + {0,1};  bz[j][2..3] = row r0 + 8
  const int r0 = warp * 16 + g;
  float bz[18][4];
  {
    const float* bt = a.bias + (size_t(t) * a.heads + head) * (ATT_TOK * ATT_TOK);
    const int zw = t / a.nH, hw = t % a.nH;
    const bool zsplit = a.roll && (zw == a.types / a.nH - 1);
    const bool hsplit = a.roll && (hw == a.nH - 1);
#pragma unroll
    for (int j = 0; j < 18; ++j) {
      const int c = 8 * j + 2 * q4;
      const float2 lo = *reinterpret_cast<const float2*>(bt + size_t(r0) * ATT_TOK + c);
      const float2 hi = *reinterpret_cast<const float2*>(bt + size_t(r0 + 8) * ATT_TOK + c);
      bz[j][0] = lo.x; bz[j][1] = lo.y; bz[j][2] = hi.x; bz[j][3] = hi.y;
      if (zsplit || hsplit) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int ri = r0 + (e >> 1) * 8, cj = c + (e & 1);
          const bool mz = zsplit && ((ri / 72) != (cj / 72));
          const bool mh = hsplit && ((((ri / 12) % 6) < 3) != (((cj / 12) % 6) < 3));
          if (mz || mh) bz[j][e] += -100.0f;
        }
      }
    }
  }

  constexpr float kLog2e = 1.4426950408889634f;
  int buf = 0;
  for (int lw = lw0; lw < lw1; ++lw, buf = (buf + 1 == ATT_STAGES ? 0 : buf + 1)) {
    // window lw has landed once at most one younger group is still in flight
    if (lw + 1 < lw1) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();   // (a) everyone's copies of window lw are visible, (b) everyone finished window lw-1
    if (lw + 2 < lw1) prefetch(lw + 2, (buf + 2) % ATT_STAGES);   // refills the buffer read in iteration lw-1
    const uint32_t sq = sbase + buf * ATT_BUF_BYTES, sk = sq + ATT_TILE_BYTES, sv = sk + ATT_TILE_BYTES;

    // ---- S = Q K^T  (16 rows x 144 keys per warp)
    uint32_t qa[2][4];
    {
      // A fragment via ldmatrix.x4: matrices (rows 0-7,k lo) (rows 8-15,k lo) (rows 0-7,k hi) (rows 8-15,k hi)
      const int r = warp * 16 + (lane & 15);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int c = ks * 2 + (lane >> 4);
        ldsm_x4(sq + att_off(r, c), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
      }
    }
    float s[18][4];
#pragma unroll
    for (int j = 0; j < 18; ++j) {
      // B fragments for keys 8j..8j+7, all four 8-wide d chunks in one ldmatrix.x4
      uint32_t b0, b1, b2, b3;
      const int r = 8 * j + (lane & 7), c = lane >> 3;
      ldsm_x4(sk + att_off(r, c), b0, b1, b2, b3);
#pragma unroll
      for (int e = 0; e < 4; ++e) s[j][e] = bz[j][e];
      mma16816<kFp16>(s[j], qa[0], b0, b1);
      mma16816<kFp16>(s[j], qa[1], b2, b3);
    }

    // ---- softmax over the 144 keys of rows r0 (e=0,1) and r0+8 (e=2,3)
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 18; ++j) {
      m0 = fmaxf(m0, fmaxf(s[j][0], s[j][1]));
      m1 = fmaxf(m1, fmaxf(s[j][2], s[j][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    const float ms0 = m0 * kLog2e, ms1 = m1 * kLog2e;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int j = 0; j < 18; ++j) {
      s[j][0] = fast_exp2(fmaf(s[j][0], kLog2e, -ms0));
      s[j][1] = fast_exp2(fmaf(s[j][1], kLog2e, -ms0));
      s[j][2] = fast_exp2(fmaf(s[j][2], kLog2e, -ms1));
      s[j][3] = fast_exp2(fmaf(s[j][3], kLog2e, -ms1));
      l0 += s[j][0] + s[j][1];
      l1 += s[j][2] + s[j][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);

    // ---- O = P V   (P re-used from the S accumulators as A fragments)
    float o[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 9; ++kk) {
      uint32_t pa[4];
      pa[0] = pack16<kFp16>(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack16<kFp16>(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack16<kFp16>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack16<kFp16>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        // matrices: (keys lo, d tile 2np) (keys hi, d tile 2np) (keys lo, 2np+1) (keys hi, 2np+1), transposed
        uint32_t b0, b1, b2, b3;
        const int r = 16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = 2 * np + (lane >> 4);
        ldsm_x4_t(sv + att_off(r, c), b0, b1, b2, b3);
        mma16816<kFp16>(o[2 * np], pa, b0, b1);
        mma16816<kFp16>(o[2 * np + 1], pa, b2, b3);
      }
    }
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;

    // ---- stage O in this warp's (now dead) Q rows, then 64 B-per-row coalesced stores
    __syncwarp();
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      // row r0: cols 8n + 2*q4 -> chunk n, byte 4*q4
      *reinterpret_cast<uint32_t*>(att_smem + buf * ATT_BUF_BYTES + att_off(r0, n) + 4 * q4) =
          pack16<kFp16>(o[n][0] * i0, o[n][1] * i0);
      *reinterpret_cast<uint32_t*>(att_smem + buf * ATT_BUF_BYTES + att_off(r0 + 8, n) + 4 * q4) =
          pack16<kFp16>(o[n][2] * i1, o[n][3] * i1);
    }
    __syncwarp();
    {
      const size_t row0 = (size_t(lw) * a.types + t) * ATT_TOK + warp * 16;
      uint8_t* outp = reinterpret_cast<uint8_t*>(a.out);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int id = i * 32 + lane;
        const int r = id >> 2, c = id & 3;
        const uint4 v = *reinterpret_cast<const uint4*>(att_smem + buf * ATT_BUF_BYTES + att_off(warp * 16 + r, c));
        stg16(outp + (row0 + r) * (size_t(a.C) * 2) + head * 64 + c * 16, v);
      }
    }
  }
}

}  // namespace pg
