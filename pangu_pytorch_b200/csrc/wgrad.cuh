// Weight-gradient GEMM on tcgen05 / TMEM:
//
//   dW[n, k] += alpha * sum_m dY[m, n] * X[m, k]        dY: [M, N] 16-bit, X: [M, K] 16-bit, dW fp32
//
// The reduction runs over the TOKEN dimension m (521 280 / 131 040 rows), the output is tiny
// (e.g. 768 x 192).  Both operands are therefore "MN-major" for the tensor core: a TMA box of
// 64 feature columns x 64 token rows (SWIZZLE_128B) is consumed directly as a
// [64 (M or N) x 64 (K = tokens)] operand tile through an MN-major shared-memory descriptor --
// no transposed copy of the activations is ever made.
//
// Work unit = (128-row tile of dW, KB-column tile of dW, slice of the token range).  A CTA
// accumulates its slice in TMEM (fp32) and adds the tile into dW with vector red.global.add.
// Warps: 0 TMA producer, 1 MMA issuer, 2..5 epilogue (one TMEM lane quadrant each).
#pragma once
#include "common.cuh"

namespace pg {

constexpr int WG_THREADS = 192;
constexpr int WG_STAGES = 4;
constexpr int WG_TOK = 64;                       // token rows per pipeline stage
constexpr int WG_BOX_BYTES = 64 * WG_TOK * 2;    // one [64 cols x 64 rows] box = 8192 B
constexpr int WG_MAX_KB = 256;
constexpr int WG_STAGE_BYTES = (2 + WG_MAX_KB / 64) * WG_BOX_BYTES;      // 2 dY boxes + up to 4 X boxes
constexpr int WG_SMEM_BYTES = 1024 + WG_STAGES * WG_STAGE_BYTES + 256;

struct WgradArgs {
  float* dw;          // [N, ldw] fp32, accumulated into
  int ldw;            // row pitch of dw (elements)
  int k_off;          // column offset inside dw (cat(skip, x) halves of the recovery weight)
  int N, K;           // valid rows / columns of this dW tile set (n < N, k < K are written)
  int KB;             // columns of dW per work unit: 64, 128, 192 or 256
  int n_tiles, k_tiles, splits;
  int chunks;         // ceil(M / 64)
  float alpha;
};

// MN-major operand tile: [64 x j columns (MN)] x [8 x k token rows], 128 B rows, SWIZZLE_128B.
//   LBO = distance between consecutive 64-column boxes, SBO = distance between 8-row groups (1024 B).
__device__ __forceinline__ uint64_t make_sdesc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFF) >> 4);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <bool kFp16>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const WgradArgs a) {
  extern __shared__ uint8_t wg_raw[];
  uint8_t* smem = wg_raw + ((1024u - (smem_u32(wg_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + WG_STAGES;
  uint64_t* tfull_bar = bars + 2 * WG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // unit decomposition: blockIdx.x = (split * k_tiles + kt) * n_tiles + nt
  const int nt = blockIdx.x % a.n_tiles;
  const int kt = (blockIdx.x / a.n_tiles) % a.k_tiles;
  const int sp = blockIdx.x / (a.n_tiles * a.k_tiles);
  const int c_begin = int((long long)a.chunks * sp / a.splits);
  const int c_end = int((long long)a.chunks * (sp + 1) / a.splits);
  const int nchunks = c_end - c_begin;
  const int xboxes = a.KB / 64;
  const uint32_t stage_tx = uint32_t(2 + xboxes) * WG_BOX_BYTES;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nchunks; ++i) {
        const int st = i % WG_STAGES;
        mbar_wait(&empty_bar[st], ((i / WG_STAGES) & 1) ^ 1);
        uint8_t* sa = smem + st * WG_STAGE_BYTES;
        const int row = (c_begin + i) * WG_TOK;
        mbar_arrive_expect_tx(&full_bar[st], stage_tx);
        tma_load_2d(&tmDY, &full_bar[st], sa, nt * 128, row);
        tma_load_2d(&tmDY, &full_bar[st], sa + WG_BOX_BYTES, nt * 128 + 64, row);
        for (int j = 0; j < xboxes; ++j)
          tma_load_2d(&tmX, &full_bar[st], sa + (2 + j) * WG_BOX_BYTES, kt * a.KB + j * 64, row);
      }
    }
  } else if (warp == 1) {
    if (nchunks > 0) {       // whole warp converged, one elected lane issues (a single-lane divergent issuer is ~2x slower, see gemm.cuh)
      // both operands MN-major: a_major (bit 15) and b_major (bit 16) set
      const uint32_t idesc = make_idesc_f16(128, a.KB, kFp16) | (1u << 15) | (1u << 16);
      for (int i = 0; i < nchunks; ++i) {
        const int st = i % WG_STAGES;
        mbar_wait(&full_bar[st], (i / WG_STAGES) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + st * WG_STAGE_BYTES);
        const uint64_t da = make_sdesc_mn_sw128(sa, WG_BOX_BYTES);
        const uint64_t db = make_sdesc_mn_sw128(sa + 2 * WG_BOX_BYTES, WG_BOX_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < WG_TOK / 16; ++k)      // 16 token rows = 2048 B per MMA K step
            umma_f16_ss(tmem, da + uint64_t(k * 128), db + uint64_t(k * 128), idesc, (i | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[st]);
          if (i == nchunks - 1) umma_commit(tfull_bar);
        }
        __syncwarp();
      }
    }
  } else if (nchunks > 0) {
    const int quad = warp & 3;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const int n = nt * 128 + quad * 32 + lane;
    float* rowp = a.dw + size_t(n) * a.ldw + a.k_off + kt * a.KB;
    for (int c0 = 0; c0 < a.KB; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem + (uint32_t(quad * 32) << 16) + c0, r);
      tmem_ld_wait();
      if (n < a.N) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int k = kt * a.KB + c0 + j;
          if (k + 3 < a.K) {
            red_add_v4(rowp + c0 + j, a.alpha * __uint_as_float(r[j]), a.alpha * __uint_as_float(r[j + 1]),
                       a.alpha * __uint_as_float(r[j + 2]), a.alpha * __uint_as_float(r[j + 3]));
          } else {
            for (int e = 0; e < 4; ++e)
              if (k + e < a.K) atomicAdd(rowp + c0 + j + e, a.alpha * __uint_as_float(r[j + e]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem);
  }
}

}  // namespace pg
