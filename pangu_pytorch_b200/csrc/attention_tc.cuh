// Earth-specific window attention on tcgen05 / TMEM (reference models/layers.py:368-415).
//
// One CTA owns one (window type, head) pair and walks the longitude windows that share its
// 144x144 bias tile.  Per window:
//
//   TMA (SWIZZLE_64B boxes of the window-ordered qkv buffer)  ->  smem ring {Q, K, V} [144][32]
//   S[128x144] = Q[0:128] K^T          tcgen05.mma, A/B from smem (K-major), fp32 accum in TMEM
//   softmax:  x = S + bias(+mask)      bias tile is RESIDENT IN TMEM (144 columns) for the CTA's life
//             P = exp2((x - max) log2e) as packed 16-bit, stored back into TMEM over S
//   O[128x32]  = P V                   tcgen05.mma, A = P from TMEM, B = V from smem (MN-major)
//   rows 128..143 (144 = 128 + 16 does not fit an MMA M) are done by two mma.sync "tail" warps that
//   read the same smem tiles (the SWIZZLE_64B pattern equals the ldmatrix-friendly XOR swizzle).
//
// TMEM columns (512): [0,144) bias+mask | [144,288) S0/P0 | [288,432) S1/P1 | [432,464) O0 | [464,496) O1.
// Warps (512 threads): 0 TMA producer, 1 MMA issuer, 2-7 tail warps (window i -> warp 2 + i%6; a lone
// mma.sync warp needs ~4-5k cycles per window, six of them keep up with the ~1.3k-cycle window period),
// 8-15 softmax: two warpgroups alternate windows; thread = one full query row (TMEM lane 32*(warp%4)+lane).
#pragma once
#include "attention.cuh"

namespace pg {

constexpr int ATC_THREADS = 512;
constexpr int ATC_TAIL_WARPS = 6;                                // warps 2..7, window i -> warp 2 + i % 6
constexpr int ATC_TB_PITCH = 148;                                // floats per row of the tail bias tile in smem
constexpr int ATC_STAGES = 6;                                    // == ATC_TAIL_WARPS: a tail warp always reuses the
                                                                 // same stage and therefore sees every phase of its barrier
constexpr int ATC_STAGE_BYTES = ATT_BUF_BYTES;                    // 27648: Q, K, V tiles of 9216 B
constexpr int ATC_SMEM_BYTES = 1024 + ATC_STAGES * ATC_STAGE_BYTES + 16384;
constexpr uint32_t ATC_COL_BIAS = 0, ATC_COL_S = 144, ATC_COL_O = 432;

__device__ __forceinline__ uint64_t make_sdesc_sw64(uint32_t smem_addr) {
  // K-major or MN-major operand whose rows are 64 B (32 x 16-bit), 8-row groups of 512 B, SWIZZLE_64B
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFF) >> 4);
  d |= uint64_t(1) << 16;               // LBO (unused: a single 64 B block along the leading dimension)
  d |= uint64_t(512 >> 4) << 32;        // SBO: next group of 8 rows
  d |= uint64_t(1) << 46;               // descriptor version (sm_100)
  d |= uint64_t(4) << 61;               // SWIZZLE_64B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_f16_ex(int M, int N, bool fp16, bool b_mn_major) {
  return make_idesc_f16(M, N, fp16) | (b_mn_major ? (1u << 16) : 0u);
}
// D[tmem] (+)= A[tmem, 16-bit packed] * B[smem]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

static_assert(ATC_STAGES == ATC_TAIL_WARPS, "tail warp w must always wait on the same stage (no skipped mbarrier phases)");

template <bool kFp16>
__global__ void __launch_bounds__(ATC_THREADS, 1)
window_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnArgs a) {
  extern __shared__ uint8_t atc_raw[];
  uint8_t* smem = atc_raw + ((1024u - (smem_u32(atc_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem;
  uint8_t* misc = smem + ATC_STAGES * ATC_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(misc);      // [ATC_STAGES]
  uint64_t* empty_bar = full_bar + ATC_STAGES;                 // [ATC_STAGES]  count 2: PV commit + tail warp
  uint64_t* sfull_bar = empty_bar + ATC_STAGES;                         // [2]  S ready in TMEM
  uint64_t* pfull_bar = sfull_bar + 2;                         // [2]  P written (128 threads of the owning warpgroup)
  uint64_t* ofull_bar = pfull_bar + 2;                         // [2]  O ready
  uint64_t* oempty_bar = ofull_bar + 2;                        // [2]  O read out (128 threads of the owning warpgroup)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(oempty_bar + 2);
  float* s_max = reinterpret_cast<float*>(misc + 256);         // [2 buf][2 half][128]
  float* s_sum = s_max + 512;                                  // [2 buf][2 half][128]
  float* s_tbias = s_sum + 512;                                // [16][ATC_TB_PITCH] bias(+mask) rows 128..143

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int th = blockIdx.x;
  const int t = th / a.heads, head = th % a.heads;
  const int lw0 = blockIdx.y * a.lon_per_cta;
  const int lw1 = min(a.nLon, lw0 + a.lon_per_cta);
  const int nwin = lw1 - lw0;
  if (nwin <= 0) return;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int s = 0; s < ATC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 2); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sfull_bar[b], 1); mbar_init(&pfull_bar[b], 128);
      mbar_init(&ofull_bar[b], 1); mbar_init(&oempty_bar[b], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int zw = t / a.nH, hw = t % a.nH;
  const bool zsplit = a.roll && (zw == a.types / a.nH - 1);
  const bool hsplit = a.roll && (hw == a.nH - 1);
  const float* bt = a.bias + (size_t(t) * a.heads + head) * (ATT_TOK * ATT_TOK);
  const bool tracing = a.trace != nullptr && blockIdx.x == 5 && blockIdx.y == 0;
  auto TR = [&](int role, int i, int ev) { if (tracing && i < 32) a.trace[(role * 32 + i) * 4 + ev] = clock64(); };
  constexpr float kLog2e = 1.4426950408889634f;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      for (int i = 0; i < nwin; ++i) {
        const int st = i % ATC_STAGES;
        mbar_wait(&empty_bar[st], ((i / ATC_STAGES) & 1) ^ 1);
        TR(0, i, 0);
        uint8_t* dst = ring + st * ATC_STAGE_BYTES;
        const int row0 = ((lw0 + i) * a.types + t) * ATT_TOK;
        mbar_arrive_expect_tx(&full_bar[st], ATC_STAGE_BYTES);
        tma_load_2d(&tmQKV, &full_bar[st], dst, head * 32, row0);
        tma_load_2d(&tmQKV, &full_bar[st], dst + ATT_TILE_BYTES, a.C + head * 32, row0);
        tma_load_2d(&tmQKV, &full_bar[st], dst + 2 * ATT_TILE_BYTES, 2 * a.C + head * 32, row0);
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16_ex(128, 144, kFp16, false);
      constexpr uint32_t idesc_o = make_idesc_f16_ex(128, 32, kFp16, true);
      auto issue_s = [&](int i) {
        const int st = i % ATC_STAGES, b = i & 1;
        mbar_wait(&full_bar[st], (i / ATC_STAGES) & 1);
        TR(1, i, 0);
        tc_fence_after();
        const uint32_t sq = smem_u32(ring + st * ATC_STAGE_BYTES);
        const uint64_t dq = make_sdesc_sw64(sq), dk = make_sdesc_sw64(sq + ATT_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < 2; ++k)      // head_dim 32 = 2 x K16; +32 B inside the 64 B swizzle row
          umma_f16_ss(tmem + ATC_COL_S + 144 * b, dq + uint64_t(k * 2), dk + uint64_t(k * 2), idesc_s, k);
        umma_commit(&sfull_bar[b]);
      };
      issue_s(0);
      for (int i = 0; i < nwin; ++i) {
        const int st = i % ATC_STAGES, b = i & 1;
        if (i + 1 < nwin) issue_s(i + 1);
        TR(1, i, 1);
        mbar_wait(&pfull_bar[b], (i >> 1) & 1);
        TR(1, i, 2);
        mbar_wait(&oempty_bar[b], ((i >> 1) & 1) ^ 1);
        TR(1, i, 3);
        tc_fence_after();
        const uint32_t sv = smem_u32(ring + st * ATC_STAGE_BYTES + 2 * ATT_TILE_BYTES);
        const uint64_t dv = make_sdesc_sw64(sv);
#pragma unroll
        for (int kk = 0; kk < 9; ++kk)   // 144 keys = 9 x K16: P advances 8 TMEM columns, V 16 rows = 1024 B
          umma_f16_ts(tmem + ATC_COL_O + 32 * b, tmem + ATC_COL_S + 144 * b + 8 * kk, dv + uint64_t(kk * 64), idesc_o,
                      kk);
        TR(5, i, 0);
        umma_commit(&ofull_bar[b]);
        umma_commit(&empty_bar[st]);
        TR(5, i, 1);
      }
    }
  } else if (warp < 2 + ATC_TAIL_WARPS) {
    // ============================== tail warps: rows 128..143 with mma.sync ==============================
    const int g = lane >> 2, q4 = lane & 3;
    const int r0 = 128 + g;
    // bias(+mask) rows 128..143 -> smem once (each tail warp fills a slice; visibility via the named barrier)
    {
      // 16 x 144 floats = 576 float4; 192 threads x 3 independent 16-byte loads each (rows 128..143 are contiguous)
      const float4* src = reinterpret_cast<const float4*>(bt + size_t(128) * ATT_TOK);
      const int tl = (warp - 2) * 32 + lane;
      float4 v[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) v[k] = __ldg(src + tl + k * (ATC_TAIL_WARPS * 32));
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int f4 = tl + k * (ATC_TAIL_WARPS * 32);
        const int rr = f4 / 36, cj0 = (f4 % 36) * 4, ri = 128 + rr;
        float e[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int cj = cj0 + q;
          const bool mz = zsplit && ((ri / 72) != (cj / 72));
          const bool mh = hsplit && ((((ri / 12) % 6) < 3) != (((cj / 12) % 6) < 3));
          if (mz || mh) e[q] += -100.0f;
        }
        *reinterpret_cast<float4*>(s_tbias + rr * ATC_TB_PITCH + cj0) = make_float4(e[0], e[1], e[2], e[3]);
      }
    }
    named_bar_sync(2, ATC_TAIL_WARPS * 32);
    for (int i = warp - 2; i < nwin; i += ATC_TAIL_WARPS) {
      const int st = i % ATC_STAGES;
      mbar_wait(&full_bar[st], (i / ATC_STAGES) & 1);
      uint8_t* tile = ring + st * ATC_STAGE_BYTES;
      if (lane == 0) TR(2, i, 0);
      if (a.debug & 1) { __syncwarp(); if (lane == 0) mbar_arrive(&empty_bar[st]); continue; }
      const uint32_t sq = smem_u32(tile), sk = sq + ATT_TILE_BYTES, sv = sk + ATT_TILE_BYTES;
      uint32_t qa[2][4];
      {
        const int r = 128 + (lane & 15);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ldsm_x4(sq + att_off(r, ks * 2 + (lane >> 4)), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
      }
      float s[18][4];
#pragma unroll
      for (int j = 0; j < 18; ++j) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(sk + att_off(8 * j + (lane & 7), lane >> 3), b0, b1, b2, b3);
        {
          const float2 lo = *reinterpret_cast<const float2*>(s_tbias + g * ATC_TB_PITCH + 8 * j + 2 * q4);
          const float2 hi = *reinterpret_cast<const float2*>(s_tbias + (g + 8) * ATC_TB_PITCH + 8 * j + 2 * q4);
          s[j][0] = lo.x; s[j][1] = lo.y; s[j][2] = hi.x; s[j][3] = hi.y;
        }
        mma16816<kFp16>(s[j], qa[0], b0, b1);
        mma16816<kFp16>(s[j], qa[1], b2, b3);
      }
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 18; ++j) {
        m0 = fmaxf(m0, fmaxf(s[j][0], s[j][1]));
        m1 = fmaxf(m1, fmaxf(s[j][2], s[j][3]));
      }
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
      const float ms0 = m0 * kLog2e, ms1 = m1 * kLog2e;
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int j = 0; j < 18; ++j) {
        s[j][0] = fast_exp2(fmaf(s[j][0], kLog2e, -ms0)); s[j][1] = fast_exp2(fmaf(s[j][1], kLog2e, -ms0));
        s[j][2] = fast_exp2(fmaf(s[j][2], kLog2e, -ms1)); s[j][3] = fast_exp2(fmaf(s[j][3], kLog2e, -ms1));
        l0 += s[j][0] + s[j][1];
        l1 += s[j][2] + s[j][3];
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
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
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(sv + att_off(16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * np + (lane >> 4)), b0, b1, b2, b3);
          mma16816<kFp16>(o[2 * np], pa, b0, b1);
          mma16816<kFp16>(o[2 * np + 1], pa, b2, b3);
        }
      }
      const float i0 = 1.0f / l0, i1 = 1.0f / l1;
      // stage O in this warp's private Q rows 128..143 (not read by the M=128 MMA), then 64 B stores
      __syncwarp();
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        *reinterpret_cast<uint32_t*>(tile + att_off(r0, n) + 4 * q4) = pack16<kFp16>(o[n][0] * i0, o[n][1] * i0);
        *reinterpret_cast<uint32_t*>(tile + att_off(r0 + 8, n) + 4 * q4) = pack16<kFp16>(o[n][2] * i1, o[n][3] * i1);
      }
      __syncwarp();
      {
        const size_t row0 = (size_t(lw0 + i) * a.types + t) * ATT_TOK + 128;
        uint8_t* outp = reinterpret_cast<uint8_t*>(a.out);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int id = k * 32 + lane, r = id >> 2, c = id & 3;
          const uint4 v = *reinterpret_cast<const uint4*>(tile + att_off(128 + r, c));
          stg16(outp + (row0 + r) * (size_t(a.C) * 2) + head * 64 + c * 16, v);
        }
      }
      __syncwarp();
      if (lane == 0) { TR(2, i, 1); mbar_arrive(&empty_bar[st]); }
    }
    // windows handled by the other tail warp still need this warp's half of the "tail" arrival? No:
    // exactly one tail warp arrives per window (count 2 = PV commit + that warp).
  } else {
    // ============================== softmax + O epilogue (2 warpgroups) ==============================
    // Warpgroup wg owns the windows i = wg (mod 2) and the S/P/O buffers b = wg; a thread owns one
    // full query row (TMEM lane).  Two passes over TMEM (max, then exp) keep the register footprint
    // small; the other warpgroup's window hides this one's TMEM / MUFU latency.
    const int quad = warp & 3, wg = (warp - 8) >> 2;
    const int r = quad * 32 + lane;
    const uint32_t lane_addr = tmem + (uint32_t(quad * 32) << 16);
    // ---- bias (+mask) -> TMEM, once: this warpgroup fills columns [72*wg, 72*wg + 72) of every row
    {
      const int c0 = 72 * wg;
      const float4* brow = reinterpret_cast<const float4*>(bt + size_t(r) * ATT_TOK + c0);
      float4 bv[18];
#pragma unroll
      for (int q = 0; q < 18; ++q) bv[q] = __ldg(brow + q);
#pragma unroll
      for (int q = 0; q < 18; q += 2) {
        uint32_t v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float4 f = bv[q + (e >> 2)];
          float bb = (e & 3) == 0 ? f.x : (e & 3) == 1 ? f.y : (e & 3) == 2 ? f.z : f.w;
          const int cj = c0 + 4 * q + e;
          const bool mz = zsplit && ((r / 72) != (cj / 72));
          const bool mh = hsplit && ((((r / 12) % 6) < 3) != (((cj / 12) % 6) < 3));
          if (mz || mh) bb += -100.0f;
          v[e] = __float_as_uint(bb);
        }
        tmem_st8(lane_addr + ATC_COL_BIAS + c0 + 4 * q, v);
      }
      tmem_st_wait();
      tc_fence_before();
      named_bar_sync(1, 256);      // the other warpgroup's half of every bias row is in TMEM too
      tc_fence_after();
    }
    const int b = wg;
    const uint32_t s_addr = lane_addr + ATC_COL_S + 144 * b;
    const uint32_t bias_addr = lane_addr + ATC_COL_BIAS;
    float l_prev = 1.f;
    auto epilogue = [&](int j, float l) {
      mbar_wait(&ofull_bar[b], (j >> 1) & 1);
      tc_fence_after();
      uint32_t o[32];
      tmem_ld32(lane_addr + ATC_COL_O + 32 * b, o);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&oempty_bar[b]);
      if (a.debug & 4) return;
      const float inv = 1.0f / l;
      const size_t row = (size_t(lw0 + j) * a.types + t) * ATT_TOK + r;
      uint8_t* dst = reinterpret_cast<uint8_t*>(a.out) + row * (size_t(a.C) * 2) + head * 64;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 v;
        v.x = pack16<kFp16>(__uint_as_float(o[8 * q + 0]) * inv, __uint_as_float(o[8 * q + 1]) * inv);
        v.y = pack16<kFp16>(__uint_as_float(o[8 * q + 2]) * inv, __uint_as_float(o[8 * q + 3]) * inv);
        v.z = pack16<kFp16>(__uint_as_float(o[8 * q + 4]) * inv, __uint_as_float(o[8 * q + 5]) * inv);
        v.w = pack16<kFp16>(__uint_as_float(o[8 * q + 6]) * inv, __uint_as_float(o[8 * q + 7]) * inv);
        stg16(dst + 16 * q, v);
      }
    };
    int last = -1;
    for (int i = wg; i < nwin; i += 2) {
      if (r == 0) TR(3, i, 0);
      mbar_wait(&sfull_bar[b], (i >> 1) & 1);
      if (r == 0) TR(3, i, 1);
      tc_fence_after();
      if (a.debug & 2) {
        if (last >= 0) epilogue(last, l_prev);
        tc_fence_before(); mbar_arrive(&pfull_bar[b]); last = i; continue;
      }
      // ---- pass 1: row maximum of S + bias over the 144 keys.  TMEM loads are software pipelined:
      //      the loads of piece p+1 are in flight while piece p is reduced (tcgen05.wait::ld is global).
      float pm = -INFINITY;
      {
        uint32_t sa[2][16], ba[2][16];
        tmem_ld16(s_addr, sa[0]);
        tmem_ld16(bias_addr, ba[0]);
        tmem_ld_wait();
#pragma unroll
        for (int part = 0; part < 9; ++part) {
          const int cur = part & 1;
          if (part + 1 < 9) {
            tmem_ld16(s_addr + 16 * (part + 1), sa[cur ^ 1]);
            tmem_ld16(bias_addr + 16 * (part + 1), ba[cur ^ 1]);
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float a0, a1;
            unpack2(add2(pack2(__uint_as_float(sa[cur][2 * e]), __uint_as_float(sa[cur][2 * e + 1])),
                         pack2(__uint_as_float(ba[cur][2 * e]), __uint_as_float(ba[cur][2 * e + 1]))), a0, a1);
            pm = max3(pm, a0, a1);
          }
          tmem_ld_wait();
        }
      }
      if (r == 0) TR(3, i, 2);
      // ---- output of this warpgroup's previous window (its PV finished long ago)
      if (last >= 0) epilogue(last, l_prev);
      if (r == 0) TR(3, i, 3);
      // ---- pass 2: P = exp2((S + bias - max) log2e), packed 16-bit, written over S
      const float m = pm * kLog2e;
      const f32x2 l2e2 = pack2(kLog2e, kLog2e), negm2 = pack2(-m, -m);
      f32x2 lsum = pack2(0.f, 0.f);
      {
        uint32_t sa[2][16], ba[2][16];
        tmem_ld16(s_addr, sa[0]);
        tmem_ld16(bias_addr, ba[0]);
        tmem_ld_wait();
#pragma unroll
        for (int part = 0; part < 9; ++part) {
          const int cur = part & 1;
          uint32_t pk[8];
          if (part + 1 < 9) {
            tmem_ld16(s_addr + 16 * (part + 1), sa[cur ^ 1]);
            tmem_ld16(bias_addr + 16 * (part + 1), ba[cur ^ 1]);
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float a0, a1;
            const f32x2 x = add2(pack2(__uint_as_float(sa[cur][2 * e]), __uint_as_float(sa[cur][2 * e + 1])),
                                 pack2(__uint_as_float(ba[cur][2 * e]), __uint_as_float(ba[cur][2 * e + 1])));
            unpack2(fma2(x, l2e2, negm2), a0, a1);
            const float p0 = ex2_approx(a0), p1 = ex2_approx(a1);
            lsum = add2(lsum, pack2(p0, p1));
            pk[e] = pack16<kFp16>(p0, p1);
          }
          tmem_ld_wait();   // piece p+1 has landed: S columns [16(p+1), +16) are in registers before ...
          // ... P columns [8p, +8) overwrite S columns that this thread has already consumed (8p+8 <= 16(p+1))
          tmem_st8(s_addr + 8 * part, pk);
        }
      }
      {
        float a0, a1;
        unpack2(lsum, a0, a1);
        l_prev = a0 + a1;
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&pfull_bar[b]);
      if (r == 0) TR(4, i, 0);
      last = i;
    }
    if (last >= 0) epilogue(last, l_prev);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

}  // namespace pg
