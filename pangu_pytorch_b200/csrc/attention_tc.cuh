// Earth-specific window attention on tcgen05 / TMEM (reference models/layers.py:368-415).
//
// Persistent kernel, one CTA per SM.  The work list is the linear index
//     u = (type * heads + head) * nLon + lon_window
// cut into gridDim.x equal contiguous ranges, so every SM gets the same number of (window, head)
// units.  A range is walked in segments that share one (type, head) pair and therefore one 144x144
// fp32 bias tile.  Everything a segment needs streams through ONE shared-memory ring of 27 648-byte
// slots, filled by TMA in order:
//
//   [bias tile: 3 slots = 9 boxes of 16 columns] [window 0: Q,K,V] [window 1] ... [next bias tile] ...
//
// so the bias of the next segment is in flight while the current one computes and no load is ever
// waited for with an empty memory pipe.  When a segment starts, the softmax warps move the bias tile
// from its ring slots into TMEM (x log2e, with the 0/-100 shifted-window mask of SURVEY.md A2 folded
// in), where it stays resident for the segment; the tail warps keep rows 128..143 in shared memory.
// Per window:
//
//   S[128x144] = Q[0:128] K^T          tcgen05.mma, A/B from smem (K-major), fp32 accum in TMEM
//   max warp:  y = S*log2e + bias' written back over S, m = max_j y      (16-column pieces through registers)
//   exp warp:  P = exp2(y - m) packed 16-bit, written over y columns the thread has already consumed
//   O[128x32]  = P V                   tcgen05.mma, A = P from TMEM, B = V from smem (MN-major)
//   rows 128..143 (144 = 128 + 16 does not fit an MMA M) are done by mma.sync "tail" warps that
//   read the same smem tiles (the SWIZZLE_64B pattern equals the ldmatrix-friendly XOR swizzle).
//
// qkv is stored head-major by the QKV GEMM ([3*heads planes][Tp_pad rows][32]), so every Q/K/V tile is
// one contiguous 9 KB burst in HBM.
// TMEM (512 columns = the whole SM, so the allocation starts at address 0):
//   [0,144) bias' | [144,288) S0/P0 | [288,432) S1/P1 | [432,464) O0 | [464,496) O1 | [496,498) row sums l (one per S buffer).
// S of window g+2 is queued right behind PV of window g (the tensor pipe executes in issue order, so PV
// has read P before the next S overwrites it).  The MMA warp runs converged with warp-uniform operands
// (only the tcgen05 instructions are elected) and polls "next S" / "next PV" without blocking on either.
// Warps (448 threads) and the SM sub-partition (warp % 4) they run on -- placement matters, measured with the clock64 trace:
//   0, 1, 4, 5  tail warps (sub-partitions 0, 1, 0, 1; ring stage s belongs to tail warp s % 4)
//   2 TMA producer, 3 MMA issuer (sub-partitions 2, 3: they share a scheduler with no tail warp)
//   6-9 "exp" warps, 10-13 "max" warps (one of each per sub-partition = per TMEM lane quadrant)
// With six tail warps two sub-partitions hosted two of them, with the MMA issuer next to a tail warp its sub-partition hosted
// four busy warps: in both layouts the exp warp of the crowded sub-partition ran 1.5-2x slower than the others and,
// since P of a window is published by all four, set the pace of the whole CTA.
// (see softmax_row_max / softmax_row_exp; warps w and w + 4 own TMEM lane quadrant w % 4): thread = one query row (TMEM lane).
// Measured on B200 (tools/tmem_bench.cu): a dependent tcgen05.ld + wait costs ~38 clk and 8 warps read ~1 KB/clk, so
// re-reading S and bias' in the exp pass is cheap; what is NOT cheap is (a) a register spill: with 232 KB of shared memory
// there is no L1 left and every spill reload is an L2 round trip, (b) code size: see the note above tail_rows.
#pragma once
#include "attention.cuh"

namespace pg {

constexpr int ATC_THREADS = 448;
constexpr int ATC_SOFT_WARP0 = 6;                                // first exp warp (6..9), then the max warps (10..13)
constexpr int ATC_STAGES = 8;
constexpr int ATC_TMA_WARP = 2, ATC_MMA_WARP = 3;                // on the sub-partitions that host no tail warp
constexpr int ATC_TAIL_WARPS = 4;                                // warps 0, 1, 4, 5: ring stage s belongs to tail warp s % 4 (two stages each), so
                                                                 // every phase of a stage's mbarriers is seen by one warp, in order
constexpr int ATC_STAGE_BYTES = ATT_BUF_BYTES;                   // 27648: Q, K, V tiles (or 3 bias boxes) of 9216 B
constexpr int ATC_TB_PITCH = 148;                                // floats per row of the tail bias tile in smem
constexpr int ATC_TB_BYTES = 16 * ATC_TB_PITCH * 4;              // rows 128..143
constexpr int ATC_XM_BYTES = 2 * 128 * 4;                        // row-maximum mailbox: [S buffer][row] fp32 (max warp -> exp warp)
constexpr int ATC_SMEM_BYTES = ATC_STAGES * ATC_STAGE_BYTES + ATC_TB_BYTES + ATC_XM_BYTES + 256;   // base is 1024 B aligned (checked)
constexpr uint32_t ATC_COL_BIAS = 0, ATC_COL_S = 144, ATC_COL_O = 432, ATC_COL_L = 496;
static_assert(ATC_SMEM_BYTES <= 232448, "attention shared memory budget");
static_assert(3 * ATC_STAGE_BYTES == ATT_TOK * ATT_TOK * 4, "a bias tile is exactly three ring slots");

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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// pointer forms: r must index registers after full unrolling (constant offsets into a local array)
__device__ __forceinline__ void tmem_st32p(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st4p(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_ld32p(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8p(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t (&r)[1]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t (&r)[2]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Code size matters here: five roles run different code on every SM sub-partition at the same time, and a fully unrolled
// version of the two functions below (58 KB of SASS) made instruction fetch the top stall reason of the kernel (ncu:
// no_inst 21 % of the samples; the instruction cache holds 32 KB).  Both are therefore written as ROLLED loops over
// key ranges with register-resident pieces of fixed size.

// Tail rows 128..143 of a window on mma.sync, flash-style over three ranges of 48 keys (6 key tiles, 3 K=16 steps):
// running row maxima m, per-thread partial sums l and the output accumulators o are rescaled in place.
// Fragment layout as in mma.m16n8k16: this thread holds rows gq (e = 0,1) and gq + 8 (e = 2,3), columns 2*q4 + {0,1}.
template <bool kFp16>
__device__ __forceinline__ void tail_rows(uint32_t sk, uint32_t sv, const uint32_t (&qa)[2][4], const float* tb0,
                                          const float* tb1, float (&o)[4][4], float& l0, float& l1, int lane) {
  constexpr float kLog2e = 1.4426950408889634f;
  constexpr int NT = 6;
  float m0 = -INFINITY, m1 = -INFINITY;
  l0 = l1 = 0.f;
#pragma unroll
  for (int n = 0; n < 4; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
#pragma unroll 1
  for (int j0 = 0; j0 < 18; j0 += NT) {
    float s[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(sk + att_off(8 * (j0 + j) + (lane & 7), lane >> 3), b0, b1, b2, b3);
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      mma16816<kFp16>(s[j], qa[0], b0, b1);
      mma16816<kFp16>(s[j], qa[1], b2, b3);
    }
    float x0 = -INFINITY, x1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const float2 lo = *reinterpret_cast<const float2*>(tb0 + 8 * (j0 + j));
      const float2 hi = *reinterpret_cast<const float2*>(tb1 + 8 * (j0 + j));
      s[j][0] = fmaf(s[j][0], kLog2e, lo.x); s[j][1] = fmaf(s[j][1], kLog2e, lo.y);
      s[j][2] = fmaf(s[j][2], kLog2e, hi.x); s[j][3] = fmaf(s[j][3], kLog2e, hi.y);
      x0 = max3(x0, s[j][0], s[j][1]);
      x1 = max3(x1, s[j][2], s[j][3]);
    }
    x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 1)); x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 2));
    x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 1)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 2));
    const float n0 = fmaxf(m0, x0), n1 = fmaxf(m1, x1);
    const float c0 = fast_exp2(m0 - n0), c1 = fast_exp2(m1 - n1);      // first range: m = -inf -> factor 0 on zeros
    m0 = n0; m1 = n1;
    float p0 = 0.f, p1 = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      s[j][0] = fast_exp2(s[j][0] - n0); s[j][1] = fast_exp2(s[j][1] - n0);
      s[j][2] = fast_exp2(s[j][2] - n1); s[j][3] = fast_exp2(s[j][3] - n1);
      p0 += s[j][0] + s[j][1];
      p1 += s[j][2] + s[j][3];
    }
    l0 = fmaf(l0, c0, p0);
    l1 = fmaf(l1, c1, p1);
#pragma unroll
    for (int n = 0; n < 4; ++n) { o[n][0] *= c0; o[n][1] *= c0; o[n][2] *= c1; o[n][3] *= c1; }
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
      uint32_t pa[4];
      pa[0] = pack16<kFp16>(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack16<kFp16>(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack16<kFp16>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack16<kFp16>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(sv + att_off(8 * j0 + 16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * np + (lane >> 4)), b0, b1, b2, b3);
        mma16816<kFp16>(o[2 * np], pa, b0, b1);
        mma16816<kFp16>(o[2 * np + 1], pa, b2, b3);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
}

// The softmax of a window is split by ROLE between the two warps that own a TMEM lane quadrant (they share an SM
// sub-partition, i.e. one MUFU and one issue port): the "max" warp computes the row maximum of window g+1 (FMA / ALU work)
// while the "exp" warp turns window g into probabilities (MUFU work).  With both warps running the same phase of the same
// window (key-split variants of this kernel) the MUFU sat idle during the max / epilogue phases and was oversubscribed
// during the exp phase: 3 600 clk per window against a MUFU bound of 1 300.
// Both functions stream the 144 score columns of one query row (TMEM lane) and the matching columns of the resident
// bias' tile in 16-column pieces through two register buffers (piece p + 1 in flight while piece p is processed), as
// ROLLED loops: five roles run different code on every sub-partition and the instruction cache holds 32 KB.

// max warp: y = S*log2e + bias' written back over S (fp32, in place) and its row maximum over the 144 keys
__device__ __forceinline__ float softmax_row_max(uint32_t sp, uint32_t bp) {
  constexpr float kLog2e = 1.4426950408889634f;
  const f32x2 l2e2 = pack2(kLog2e, kLog2e);
  uint32_t sa[16], ba[16], sb[16], bb[16];
  float m0 = -INFINITY, m1 = -INFINITY;        // two chains: the reduction is latency bound
  auto reduce = [&](uint32_t (&sx)[16], const uint32_t (&bx)[16]) {      // sx: S in, y out
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      float a0, a1, a2, a3;
      unpack2(fma2(pack2(__uint_as_float(sx[2 * e]), __uint_as_float(sx[2 * e + 1])), l2e2,
                   pack2(__uint_as_float(bx[2 * e]), __uint_as_float(bx[2 * e + 1]))), a0, a1);
      unpack2(fma2(pack2(__uint_as_float(sx[2 * e + 2]), __uint_as_float(sx[2 * e + 3])), l2e2,
                   pack2(__uint_as_float(bx[2 * e + 2]), __uint_as_float(bx[2 * e + 3]))), a2, a3);
      m0 = max3(m0, a0, a1);
      m1 = max3(m1, a2, a3);
      sx[2 * e] = __float_as_uint(a0); sx[2 * e + 1] = __float_as_uint(a1);
      sx[2 * e + 2] = __float_as_uint(a2); sx[2 * e + 3] = __float_as_uint(a3);
    }
  };
  tmem_ld16(sp, sa);
  tmem_ld16(bp, ba);
  tmem_ld_wait();
#pragma unroll 1
  for (int c = 0; c < 8; c += 2) {             // pieces 0..7 in pairs, piece 8 after the loop
    tmem_ld16(sp + 16 * (c + 1), sb);
    tmem_ld16(bp + 16 * (c + 1), bb);
    reduce(sa, ba);
    tmem_ld_wait();
    tmem_st16(sp + 16 * c, sa);
    tmem_st_wait();                            // sa is reloaded below: the store must have read it
    tmem_ld16(sp + 16 * (c + 2), sa);
    tmem_ld16(bp + 16 * (c + 2), ba);
    reduce(sb, bb);
    tmem_ld_wait();
    tmem_st16(sp + 16 * (c + 1), sb);
    tmem_st_wait();
  }
  reduce(sa, ba);
  tmem_st16(sp + 128, sa);
  tmem_st_wait();
  return fmaxf(m0, m1);
}

// exp warp: P = exp2(y - m), packed 16-bit, written over columns [0, 72) of the row (piece p of P lands on columns
// [8p, 8p+8), which this thread consumed in piece p/2 <= p; nobody else reads the row any more); returns the row sum of
// the unrounded p.  Per pair of keys: one FADD2, two MUFU, one FADD2 (sum), one F2FP -- a warp issues in order, so every
// non-MUFU instruction here adds to the 8 clk each MUFU.EX2 holds the pipe (the bias add lives in the max warp for that reason).
template <bool kFp16>
__device__ __forceinline__ float softmax_row_exp(uint32_t sp, float m) {
  // Software pipelined over the 16-column pieces: the MUFUs of piece c+1 (stage A) share a basic block with the row-sum
  // adds, the packing and the store of piece c (stage B), so that ptxas can interleave them.  Piece by piece
  // (A then B of the same piece, separated by the TMEM wait / store) the pass took 222 clk per piece against 128 clk of
  // MUFU time: fill and drain of the MUFU pipeline were paid nine times per row.
  const f32x2 negm2 = pack2(-m, -m);
  uint32_t ya[16], yb[16];
  float pa[16], pb[16];
  f32x2 l0 = pack2(0.f, 0.f), l1 = pack2(0.f, 0.f);
  auto stage_a = [&](const uint32_t (&yx)[16], float (&px)[16]) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float a0, a1;
      unpack2(add2(pack2(__uint_as_float(yx[2 * e]), __uint_as_float(yx[2 * e + 1])), negm2), a0, a1);
      px[2 * e] = ex2_approx(a0);
      px[2 * e + 1] = ex2_approx(a1);
    }
  };
  auto stage_b = [&](const float (&px)[16], uint32_t taddr) {
    uint32_t pk[8];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      l0 = add2(l0, pack2(px[2 * e], px[2 * e + 1]));
      l1 = add2(l1, pack2(px[2 * e + 2], px[2 * e + 3]));
      pk[e] = pack16<kFp16>(px[2 * e], px[2 * e + 1]);
      pk[e + 1] = pack16<kFp16>(px[2 * e + 2], px[2 * e + 3]);
    }
    tmem_st8(taddr, pk);
  };
  tmem_ld16(sp, ya);
  tmem_ld_wait();
  tmem_ld16(sp + 16, yb);
  stage_a(ya, pa);
#pragma unroll 1
  for (int c = 0; c < 8; c += 2) {      // invariant: pa = p of piece c, piece c+1 in flight into yb
    tmem_ld_wait();
    tmem_ld16(sp + 16 * (c + 2), ya);
    stage_a(yb, pb);
    stage_b(pa, sp + 8 * c);            // columns [8c, 8c+8) were consumed in piece c/2
    tmem_ld_wait();
    if (c + 3 < 9) tmem_ld16(sp + 16 * (c + 3), yb);
    stage_a(ya, pa);
    stage_b(pb, sp + 8 * (c + 1));
  }
  stage_b(pa, sp + 64);
  float a0, a1;
  unpack2(add2(l0, l1), a0, a1);
  return a0 + a1;
}

template <bool kFp16>
__global__ void __launch_bounds__(ATC_THREADS, 1)
window_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmBias,
                           const AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t atc_raw[];
  if ((smem_u32(atc_raw) & 1023u) != 0u) __trap();   // the swizzled tiles need 1024 B alignment and there is no slack to fix it up
  uint8_t* smem = atc_raw;
  uint8_t* ring = smem;
  float* s_tbias = reinterpret_cast<float*>(smem + ATC_STAGES * ATC_STAGE_BYTES);     // [16][ATC_TB_PITCH]
  float* s_xm = reinterpret_cast<float*>(smem + ATC_STAGES * ATC_STAGE_BYTES + ATC_TB_BYTES);   // [2][128]
  uint8_t* misc = smem + ATC_STAGES * ATC_STAGE_BYTES + ATC_TB_BYTES + ATC_XM_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(misc);      // [ATC_STAGES]  TMA bytes landed
  uint64_t* empty_bar = full_bar + ATC_STAGES;                 // [ATC_STAGES]  count 2 (window: PV commit + tail warp;
                                                               //                        bias: MMA warp for the softmax warps + tail group)
  uint64_t* sfull_bar = empty_bar + ATC_STAGES;                // [2]  S ready in TMEM
  uint64_t* pfull_bar = sfull_bar + 2;                         // [2]  P written (the 128 threads of the exp warps)
  uint64_t* ofull_bar = pfull_bar + 2;                         // [2]  O ready
  uint64_t* oempty_bar = ofull_bar + 2;                        // [2]  O read out (the 128 threads of the max warps)
  uint64_t* sbias_bar = oempty_bar + 2;                        // [1]  a segment's bias tile has been read out of the ring by all
                                                               //      8 softmax warps (the MMA warp then frees the 3 slots for them)
  uint64_t* mfull_bar = sbias_bar + 1;                         // [2]  row maxima of a window are in the mailbox (128 max-warp threads)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mfull_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == ATC_TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmBias);
    for (int s = 0; s < ATC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 2); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sfull_bar[b], 1); mbar_init(&pfull_bar[b], 128);
      mbar_init(&ofull_bar[b], 1); mbar_init(&oempty_bar[b], 128);
      mbar_init(&mfull_bar[b], 128);
    }
    mbar_init(sbias_bar, 8);
    fence_barrier_init();
  }
  if (warp == ATC_MMA_WARP) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (*tmem_slot != 0u) __trap();     // all 512 columns: the allocation can only start at TMEM address 0
  constexpr uint32_t tmem = 0u;       // compile-time constant keeps the MMA operands in uniform registers
  pdl_trigger();
  pdl_wait();                         // qkv of the preceding GEMM is read from here on

  // balanced contiguous range of (type, head, lon window) units for this CTA
  const long long total = (long long)a.types * a.heads * a.nLon;
  const int u_begin = int(total * blockIdx.x / gridDim.x);
  const int u_end = int(total * (blockIdx.x + 1) / gridDim.x);
  const int nunits = u_end - u_begin;
  const int lw_first = u_begin % a.nLon;
#ifdef PANGU_ATTN_TRACE     // development builds only: per-role clock64 timeline of CTA 5 ([role 8][window 64][event 4])
  const bool tracing = a.trace != nullptr && blockIdx.x == 5;
  auto TR = [&](int role, int g, int ev, bool who) { if (tracing && who && g < 64) a.trace[(role * 64 + g) * 4 + ev] = clock64(); };
#else
  auto TR = [](int, int, int, bool) {};
#endif
  constexpr float kLog2e = 1.4426950408889634f;
#ifdef PANGU_DEV_SWITCHES   // timing ablations (results invalid): bit0 no tail-warp math, bit1 no exp pass, bit2 no output stores, bit3 no max pass
  const int dbg = a.debug;
#else
  constexpr int dbg = 0;
#endif

  if (warp == ATC_TMA_WARP) {
    // ============================== TMA producer ==============================
    // ring slot sequence: per segment 3 bias slots, then one slot per window
    if (lane == 0) {
      int th = u_begin / a.nLon, lw = lw_first;
      int q = 0;
      for (int u = u_begin; u < u_end;) {
        const int nwin = min(a.nLon - lw, u_end - u);
        const int t = th / a.heads, head = th % a.heads;
        for (int j = 0; j < 3; ++j, ++q) {
          const int st = q % ATC_STAGES;
          mbar_wait(&empty_bar[st], ((q / ATC_STAGES) & 1) ^ 1);
          uint8_t* dst = ring + st * ATC_STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[st], ATC_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < 3; ++k)      // box = 16 key columns x 144 query rows of the fp32 bias tile
            tma_load_2d(&tmBias, &full_bar[st], dst + k * ATT_TILE_BYTES, 16 * (3 * j + k), th * ATT_TOK);
        }
        for (int i = 0; i < nwin; ++i, ++q) {
          const int st = q % ATC_STAGES;
          mbar_wait(&empty_bar[st], ((q / ATC_STAGES) & 1) ^ 1);
          TR(0, u - u_begin + i, 0, true);
          uint8_t* dst = ring + st * ATC_STAGE_BYTES;
          const int row0 = ((lw + i) * a.types + t) * ATT_TOK;
          mbar_arrive_expect_tx(&full_bar[st], ATC_STAGE_BYTES);
          tma_load_2d(&tmQKV, &full_bar[st], dst, 0, head * a.plane_rows + row0);
          tma_load_2d(&tmQKV, &full_bar[st], dst + ATT_TILE_BYTES, 0, (a.heads + head) * a.plane_rows + row0);
          tma_load_2d(&tmQKV, &full_bar[st], dst + 2 * ATT_TILE_BYTES, 0, (2 * a.heads + head) * a.plane_rows + row0);
        }
        u += nwin; lw = 0; ++th;
      }
    }
  } else if (warp == ATC_MMA_WARP) {
    // ============================== MMA issuer ==============================
    // The whole warp walks the loop (uniform control flow and operands); one elected lane issues.
    constexpr uint32_t idesc_s = make_idesc_f16_ex(128, 144, kFp16, false);
    constexpr uint32_t idesc_o = make_idesc_f16_ex(128, 32, kFp16, true);
    const uint32_t ring_u32 = smem_u32(ring);
    // Two cursors over the ring: cs = next S to issue, cp = next PV to issue (g = window count, q = ring slot,
    // rem = windows left in the segment, bias = bias slots of the upcoming segment still to be passed).
    // cs observes the full barrier of EVERY slot in ring order, bias slots included, so it never probes a
    // stage whose previous phase it has not seen complete.  All probes are the non-blocking test_wait: the warp
    // must keep issuing PVs while it waits for data (a blocking wait here can deadlock short segments).
    struct Cur { int g, q, rem, bias; };
    auto advance = [&](Cur& c, bool observe_bias) {
      ++c.g; ++c.q;
      if (--c.rem == 0) {
        if (observe_bias) c.bias = 3; else c.q += 3;
        c.rem = min(a.nLon, nunits - c.g);
      }
    };
    Cur cs{0, 0, min(a.nLon - lw_first, nunits), 3}, cp{0, 3, cs.rem, 0};
    // third cursor: segments whose bias slots are still to be handed back to the producer on behalf of the softmax warps
    int br_seg = 0, br_q = 0, br_left = nunits, br_win = cs.rem;
    auto release_bias = [&]() {
      if (lane < 3) mbar_arrive(&empty_bar[(br_q + lane) % ATC_STAGES]);
      __syncwarp();
      br_q += 3 + br_win; br_left -= br_win; br_win = min(a.nLon, br_left); ++br_seg;
    };
    while (cp.g < nunits) {
      bool progress = false;
      if (br_left > 0 && __any_sync(0xffffffffu, mbar_test_wait(sbias_bar, br_seg & 1))) {
        release_bias();
        progress = true;
      }
      if (cs.g < nunits && cs.bias > 0) {
        if (__any_sync(0xffffffffu, mbar_test_wait(&full_bar[cs.q % ATC_STAGES], (cs.q / ATC_STAGES) & 1))) {
          ++cs.q; --cs.bias;
          progress = true;
        }
      } else if (cs.g < nunits && cs.g < cp.g + 2) {  // buffer cs.g & 1 is free: PV of window cs.g - 2 has been issued
        const int st = cs.q % ATC_STAGES, b = cs.g & 1;
        if (__any_sync(0xffffffffu, mbar_test_wait(&full_bar[st], (cs.q / ATC_STAGES) & 1))) {
          TR(1, cs.g, 0, lane == 0);
          tc_fence_after();
          const uint32_t sq = ring_u32 + st * ATC_STAGE_BYTES;
          const uint64_t dq = make_sdesc_sw64(sq), dk = make_sdesc_sw64(sq + ATT_TILE_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 2; ++k)      // head_dim 32 = 2 x K16; +32 B inside the 64 B swizzle row
              umma_f16_ss(tmem + ATC_COL_S + 144 * b, dq + uint64_t(k * 2), dk + uint64_t(k * 2), idesc_s, k);
            umma_commit(&sfull_bar[b]);
          }
          __syncwarp();
          advance(cs, true);
          progress = true;
        }
      }
      {
        const int st = cp.q % ATC_STAGES, b = cp.g & 1;
        if (__any_sync(0xffffffffu, mbar_test_wait(&pfull_bar[b], (cp.g >> 1) & 1))) {
          mbar_wait(&oempty_bar[b], ((cp.g >> 1) & 1) ^ 1);   // short: the max warps read O(g-2) right after posting the maxima of window g,
                                                              // which the exp warps needed before they could publish P(g)
          TR(1, cp.g, 2, lane == 0);
          tc_fence_after();
          const uint64_t dv = make_sdesc_sw64(ring_u32 + st * ATC_STAGE_BYTES + 2 * ATT_TILE_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 9; ++kk)   // 144 keys = 9 x K16: P advances 8 TMEM columns, V 16 rows = 1024 B
              umma_f16_ts(tmem + ATC_COL_O + 32 * b, tmem + ATC_COL_S + 144 * b + 8 * kk, dv + uint64_t(kk * 64), idesc_o, kk);
            umma_commit(&ofull_bar[b]);
            umma_commit(&empty_bar[st]);
          }
          __syncwarp();
          TR(1, cp.g, 3, lane == 0);
          advance(cp, false);
          progress = true;
        }
      }
      // Nothing to issue: park on the event that comes next in the steady state (P of window cp.g) instead of spinning.  A
      // polling loop here costs the exp warp of this sub-partition issue slots: measured 1.5-2x longer exp passes in the
      // lane quadrant that shares a scheduler with this warp, which set the pace of the CTA.
      if (!progress) (void)mbar_try_wait_hint(&pfull_bar[cp.g & 1], (cp.g >> 1) & 1, 300);
    }
    while (br_left > 0) {          // (every segment's tile was staged before its first P, so these are already complete)
      mbar_wait(sbias_bar, br_seg & 1);
      release_bias();
    }
  } else {
    // ============================== tail + softmax warps: walk the segments ==============================
    int gbase = 0;     // windows this CTA has finished: barrier phases and the buffer rotation run on
    int lw0 = lw_first;
    int th = u_begin / a.nLon;
    for (int u = u_begin, seg = 0; u < u_end; ++seg) {
      const int nwin = min(a.nLon - lw0, u_end - u);
      const int t = th / a.heads, head = th % a.heads;
      const int zw = t / a.nH, hw = t % a.nH;
      const bool zsplit = a.roll && (zw == a.types / a.nH - 1);
      const bool hsplit = a.roll && (hw == a.nH - 1);
      const int qb = gbase + 3 * seg;            // ring slots qb..qb+2 hold this segment's bias tile; windows follow
      // additive mask (reference gen_mask, models/layers.py:153-181) for query row ri and the 4 keys cj..cj+3
      auto mask_of = [&](int ri, int cj) {
        const bool mz = zsplit && ((ri / 72) != (cj / 72));
        const bool mh = hsplit && ((((ri / 12) % 6) < 3) != (((cj / 12) % 6) < 3));
        return (mz || mh) ? -100.0f : 0.0f;
      };
      // box p (16 key columns) of the bias tile sits in ring slot qb + p/3 at tile offset (p%3)
      auto box_ptr = [&](int p) { return ring + ((qb + p / 3) % ATC_STAGES) * ATC_STAGE_BYTES + (p % 3) * ATT_TILE_BYTES; };

      // Output goes out in NATURAL token order (window reverse + un-roll + crop of models/layers.py:227-243 folded
      // into the store address): row k of a window of this type is token (zp, hp, wp) with zp, hp fixed for the
      // segment and wp advancing by 12 per longitude window; pad rows (hp >= H) are dropped.
      const int Hp_ = a.H + 5;
      auto row_base = [&](int k) {            // (zp * H + hp) * W, or -1 for a pad row
        const int zl = k / 72, hl = (k / 12) % 6;
        int zp = 2 * zw + zl, hp = 6 * hw + hl;
        if (a.roll) { zp = (zp + 1) & 7; hp = (hp + 3) % Hp_; }
        return hp >= a.H ? -1 : (zp * a.H + hp) * a.W;
      };
      auto out_row = [&](int base, int k, int lw) {   // output row of window row k in longitude window lw (-1: dropped)
        if (!a.natural) return (lw * a.types + t) * ATT_TOK + k;      // window order, pad rows included
        int wp = 12 * lw + (k % 12) + (a.roll ? 6 : 0);
        wp = wp >= a.W ? wp - a.W : wp;
        return base < 0 ? -1 : base + wp;
      };
      if (warp < ATC_SOFT_WARP0) {
        // ============================== tail warps: rows 128..143 with mma.sync ==============================
        const int tix = warp < 2 ? warp : warp - 2;   // tail warp index 0..3 (warps 0, 1, 4, 5)
        const int tid = tix * 32 + lane;         // 0..127
        named_bar_sync(2, ATC_TAIL_WARPS * 32);  // every tail warp is done with the previous tail bias rows
#pragma unroll
        for (int j = 0; j < 3; ++j) mbar_wait(&full_bar[(qb + j) % ATC_STAGES], ((qb + j) / ATC_STAGES) & 1);
#pragma unroll 1
        for (int id = tid; id < 16 * 36; id += ATC_TAIL_WARPS * 32) {      // 16 rows x 36 float4 = 576 pieces
          const int rr = id / 36, c4 = id % 36, ri = 128 + rr, cj = 4 * c4;
          const float4 v = *reinterpret_cast<const float4*>(box_ptr(c4 >> 2) + att_off(ri, c4 & 3));
          const float m = mask_of(ri, cj);
          *reinterpret_cast<float4*>(s_tbias + rr * ATC_TB_PITCH + cj) =
              make_float4((v.x + m) * kLog2e, (v.y + m) * kLog2e, (v.z + m) * kLog2e, (v.w + m) * kLog2e);
        }
        named_bar_sync(2, ATC_TAIL_WARPS * 32);
        if (tid < 3) mbar_arrive(&empty_bar[(qb + tid) % ATC_STAGES]);    // the tail group's release of the bias slots

        const int gq = lane >> 2, q4 = lane & 3;
        const int r0 = 128 + gq;
        const float* tb0 = s_tbias + gq * ATC_TB_PITCH + 2 * q4;
        const float* tb1 = tb0 + 8 * ATC_TB_PITCH;
        // this warp owns the ring stages s with s % 4 == tix: it takes the windows whose slot falls on them
        for (int i = 0; i < nwin; ++i) {
          const int g = gbase + i, q = qb + 3 + i, st = q % ATC_STAGES;
          if (st % ATC_TAIL_WARPS != tix) continue;
          mbar_wait(&full_bar[st], (q / ATC_STAGES) & 1);
          uint8_t* tile = ring + st * ATC_STAGE_BYTES;
          TR(2, g, 0, lane == 0);
          const uint32_t sq = smem_u32(tile), sk = sq + ATT_TILE_BYTES, sv = sk + ATT_TILE_BYTES;
          uint32_t qa[2][4];
          {
            const int r = 128 + (lane & 15);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) ldsm_x4(sq + att_off(r, ks * 2 + (lane >> 4)), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
          }
          // three ranges of 48 keys with a running maximum: 24 scores live at a time (the tail warps share the
          // 128-register budget with everybody else) and ONE copy of the code
          float o[4][4], l0, l1;
          if (dbg & 1) {
#pragma unroll
            for (int n = 0; n < 4; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
            l0 = l1 = 1.f;
          } else
          tail_rows<kFp16>(sk, sv, qa, tb0, tb1, o, l0, l1, lane);
          const float i0 = 1.0f / l0, i1 = 1.0f / l1;
          // stage O in this window's Q rows 128..143 (not read by the M=128 MMA), then 64 B stores
          __syncwarp();
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            *reinterpret_cast<uint32_t*>(tile + att_off(r0, n) + 4 * q4) = pack16<kFp16>(o[n][0] * i0, o[n][1] * i0);
            *reinterpret_cast<uint32_t*>(tile + att_off(r0 + 8, n) + 4 * q4) = pack16<kFp16>(o[n][2] * i1, o[n][3] * i1);
          }
          __syncwarp();
          {
            uint8_t* outp = reinterpret_cast<uint8_t*>(a.out);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int id = k * 32 + lane, r = id >> 2, c = id & 3;
              const int orow = out_row(row_base(128 + r), 128 + r, lw0 + i);
              if (orow < 0) continue;
              const uint4 v = *reinterpret_cast<const uint4*>(tile + att_off(128 + r, c));
              stg16(outp + size_t(orow) * (size_t(a.C) * 2) + head * 64 + c * 16, v);
            }
          }
          __syncwarp();
          TR(2, g, 1, lane == 0);
          if (lane == 0) mbar_arrive(&empty_bar[st]);
        }
      } else {
        // ============================== softmax + O epilogue (8 warps, two roles) ==============================
        // Warps w (6..9, "exp") and w + 4 (10..13, "max") own TMEM lane quadrant `quad` = w % 4; a thread = one query row.
        // The max warp runs one window ahead: row maximum of window i -> mailbox, then the output of window i - 2
        // (O / l read back, normalised, stored); the exp warp turns window i into P and the row sums.
        const int quad = warp & 3;
        const bool is_max = warp >= ATC_SOFT_WARP0 + 4;
        const int half = is_max ? 1 : 0;           // which boxes of the bias tile this warp stages
        const int r = quad * 32 + lane;
        const uint32_t lane_addr = tmem + (uint32_t(quad * 32) << 16);
        const int pair_bar = 3 + quad;           // named barrier of the two warps that share this lane quadrant

        // ---- segment start: bias tile -> TMEM.  The 32 bias rows of a lane quadrant are read and written only by the two
        //      warps of that quadrant, so the pair barrier is all the synchronisation the overwrite needs (the max warp gets
        //      here after its last epilogue of the segment, i.e. after the exp warp has published the last P: nobody reads
        //      the old tile any more); the ring slots are handed back by the MMA warp once all 8 warps have signalled sbias_bar.
        TR(6, seg, 0, threadIdx.x == 256 /* warp 8, lane 0: the exp warp of quadrant 0 */);
        named_bar_sync(pair_bar, 64);
#pragma unroll
        for (int j = 0; j < 3; ++j) mbar_wait(&full_bar[(qb + j) % ATC_STAGES], ((qb + j) / ATC_STAGES) & 1);
#pragma unroll 1
        for (int p = half; p < 9; p += 2) {
          const uint8_t* box = box_ptr(p);
          uint32_t v[16];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 f = *reinterpret_cast<const float4*>(box + att_off(r, c));
            const float m = mask_of(r, 16 * p + 4 * c);
            v[4 * c + 0] = __float_as_uint((f.x + m) * kLog2e); v[4 * c + 1] = __float_as_uint((f.y + m) * kLog2e);
            v[4 * c + 2] = __float_as_uint((f.z + m) * kLog2e); v[4 * c + 3] = __float_as_uint((f.w + m) * kLog2e);
          }
          tmem_st16(lane_addr + ATC_COL_BIAS + 16 * p, v);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(sbias_bar);   // this warp has read its share of the tile out of the ring
        tmem_st_wait();
        tc_fence_before();
        named_bar_sync(pair_bar, 64);            // the partner's boxes of these rows are in TMEM too
        tc_fence_after();
        TR(6, seg, 1, threadIdx.x == 256 /* warp 8, lane 0: the exp warp of quadrant 0 */);

        if (!is_max) {
          // ------------------------------ exp warp ------------------------------
#pragma unroll 1
          for (int i = 0; i < nwin; ++i) {
            const int g = gbase + i, b = g & 1;
            TR(3, g, 0, r == 0);
            mbar_wait(&mfull_bar[b], (g >> 1) & 1);      // y of this window is in TMEM (which implies S was), maxima in the mailbox
            tc_fence_after();
            TR(3, g, 1, r == 0);
            const float m = s_xm[b * 128 + r];
            const float l = (dbg & 2) ? 1.f : softmax_row_exp<kFp16>(lane_addr + ATC_COL_S + 144 * b, m);
            TR(3, g, 2, r == 0);
            tmem_st1(lane_addr + ATC_COL_L + b, __float_as_uint(l));
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&pfull_bar[b]);
            TR(3, g, 3, r == 0);
            TR(4, g, quad, lane == 0);           // publish time of every exp warp
          }
        } else {
          // ------------------------------ max warp (+ output) ------------------------------
          // Output row of this thread's query row in window lw0 + j: base + off_j rows, off advancing by a fixed step per
          // window (natural order: 12 longitudes with wrap-around at W; window order: one window of rows) -- no divisions
          // in the per-window epilogue.
          const int my_base = row_base(r);
          const bool keep = !a.natural || my_base >= 0;     // false: zero pad row of the window, its output is cropped away
          const size_t rowb = size_t(a.C) * 2;
          uint8_t* const obase = reinterpret_cast<uint8_t*>(a.out) + head * 64 +
                                 (a.natural ? size_t(my_base < 0 ? 0 : my_base) : size_t(t) * ATT_TOK + r) * rowb;
          const int ostep = a.natural ? 12 : a.types * ATT_TOK;
          const int owrap = a.natural ? a.W : 0x7fffffff;
          int ooff = a.natural ? (12 * lw0 + (r % 12) + (a.roll ? 6 : 0)) % a.W : lw0 * a.types * ATT_TOK;
          auto epilogue = [&](int j) {               // windows are finished in order j = 0, 1, ...
            const int g = gbase + j, b = g & 1;
            mbar_wait(&ofull_bar[b], (g >> 1) & 1);
            tc_fence_after();
            uint32_t o0[16], o1[16], ls[1];
            tmem_ld16(lane_addr + ATC_COL_O + 32 * b, o0);
            tmem_ld16(lane_addr + ATC_COL_O + 32 * b + 16, o1);
            tmem_ld1(lane_addr + ATC_COL_L + b, ls);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&oempty_bar[b]);
            uint8_t* dst = obase + size_t(ooff) * rowb;
            ooff += ostep;
            ooff = ooff >= owrap ? ooff - owrap : ooff;
            if (!keep || (dbg & 4)) return;
            float inv;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(__uint_as_float(ls[0])));
            auto put = [&](const uint32_t (&o)[16], uint8_t* d) {
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                uint4 v;
                v.x = pack16<kFp16>(__uint_as_float(o[8 * q + 0]) * inv, __uint_as_float(o[8 * q + 1]) * inv);
                v.y = pack16<kFp16>(__uint_as_float(o[8 * q + 2]) * inv, __uint_as_float(o[8 * q + 3]) * inv);
                v.z = pack16<kFp16>(__uint_as_float(o[8 * q + 4]) * inv, __uint_as_float(o[8 * q + 5]) * inv);
                v.w = pack16<kFp16>(__uint_as_float(o[8 * q + 6]) * inv, __uint_as_float(o[8 * q + 7]) * inv);
                stg16(d + 16 * q, v);
              }
            };
            put(o0, dst);
            put(o1, dst + 32);
          };
#pragma unroll 1
          for (int i = 0; i < nwin; ++i) {
            const int g = gbase + i, b = g & 1;
            TR(5, g, 0, r == 0);
            mbar_wait(&sfull_bar[b], (g >> 1) & 1);
            tc_fence_after();
            TR(5, g, 1, r == 0);
            s_xm[b * 128 + r] = (dbg & 8) ? 0.f : softmax_row_max(lane_addr + ATC_COL_S + 144 * b, lane_addr + ATC_COL_BIAS);
            tc_fence_before();                          // the y stores (already waited for) before the hand-over
            mbar_arrive(&mfull_bar[b]);
            TR(5, g, 2, r == 0);
            TR(7, g, quad, lane == 0);           // post time of every max warp
            if (i >= 2) epilogue(i - 2);
            TR(5, g, 3, r == 0);
          }
          if (nwin >= 2) epilogue(nwin - 2);
          epilogue(nwin - 1);
        }
        TR(6, seg, 2, threadIdx.x == 256 /* warp 8, lane 0: the exp warp of quadrant 0 */);
      }
      gbase += nwin;
      u += nwin;
      lw0 = 0;
      ++th;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == ATC_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

}  // namespace pg
