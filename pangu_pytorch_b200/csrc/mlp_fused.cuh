// Fused Mlp + LayerNorm + residual of one EarthSpecificBlock (reference models/layers.py:250-251, 264-270):
//
//     x <- x + s * LN2( GELU(x16 W1^T + b1) W2^T + b2 )
//
// in ONE persistent tcgen05 kernel: the 4C-wide hidden activation never leaves the SM.
// Per 128-token tile (one CTA; CTA pairs share every weight tile by TMA multicast):
//
//   X tile [128 x C] 16-bit          TMA -> smem (resident for the tile, K-slab granular reuse)
//   for each hidden chunk c of 64 units:
//     GEMM1  Hacc[128x64]  = X W1[c]^T            tcgen05.mma SS, fp32 accumulators in TMEM (2 buffers)
//     GELU   H = gelu_erf(Hacc + b1[c]) -> 16-bit, written back INTO TMEM over its own accumulator
//     GEMM2  Y[128xC]     += H W2[:, c]^T         tcgen05.mma with A = H from TMEM, B = W2 slab from smem
//   epilogue: Y + b2 -> LayerNorm (thread = row) -> * s + residual -> fp32 stream (in place) and 16-bit
//             shadow (natural order, or scattered to window order for the next block's QKV GEMM)
//
// The MMA warp issues  G1(c+NB-1), G2(c)  alternately (NB = Hacc buffers), so the tensor pipe always has
// GEMM1 work queued while the GELU warps are busy; the tensor pipe executes in issue order, which is what
// makes the in-place reuse of the Hacc buffers safe (G2(c) has read H(c) before G1(c+NB) overwrites it).
// TMEM: Y [0,C) then NB Hacc buffers of 64 columns (C=192: 4 buffers, C=384: 2).
// Warps (448 threads): 0 TMA producer, 1 MMA issuer, 2-5 LayerNorm epilogue (one per TMEM lane quadrant),
// 6-13 GELU (two warpgroups alternate chunks).
#pragma once
#include "common.cuh"
#include "geometry.cuh"
#include "gemm.cuh"

namespace pg {


template <int C>
struct MlpTraits {
  static constexpr int KX = C / 64;                    // K slabs of X (GEMM1)
  static constexpr int NH = C / 192;                   // 192-column halves of Y (GEMM2 N per instruction)
  static constexpr int NCH = 4 * C / 64;               // hidden chunks of 64 units
  static constexpr int YB = (C == 192) ? 2 : 1;        // Y accumulators: two at C=192, so that the LayerNorm epilogue of a
                                                       // tile runs under the GEMMs of the next one (C=384: TMEM holds one)
  static constexpr int COL_H = YB * C;                 // Hacc buffers follow the Y accumulator(s) in TMEM
  static constexpr int NB = 2;                         // Hacc buffers of 64 columns: GEMM1 runs NB-1 chunks ahead of GEMM2
                                                       // (four 32-column buffers were measured much slower, 823 vs 484 us:
                                                       // twice the barrier round trips in the single MMA-issuing warp)
  static_assert(YB * C + NB * 64 <= 512 && NCH % NB == 0, "TMEM budget / buffer rotation");
  static constexpr int S1 = 2;                         // ring 1: W1 CHUNK units [64 hidden x C k] = KX x 8 KB (one barrier
                                                       // round trip per chunk in the MMA-issuing warp instead of KX)
  static constexpr int S2 = 2;                         // ring 2: W2 units [192 out x 64 k]   = 24 KB
  static constexpr int NHS = 2;                        // H operand buffers in shared memory [128 x 64] 16-bit = 16 KB, one per GELU warpgroup
#ifndef PANGU_MLP_MCAST
#define PANGU_MLP_MCAST 1
#endif
  static constexpr bool MCAST = PANGU_MLP_MCAST != 0;  // weight units shared by the CTA pair through TMA multicast (couples the two
                                                       // CTAs' rings: every slot waits for both) or fetched by each CTA on its own
  static constexpr int X_BYTES = KX * 16384;
  static constexpr int R1_UNIT = KX * 8192, R2_UNIT = 24576;
  static constexpr int STG_PITCH = 32 * 4 + 16;        // epilogue slab row pitch (bytes), conflict-free 16 B rows
  static constexpr bool RES_TMA = (C == 192);          // residual stream moved by TMA through two swizzled [32 x 32] fp32
                                                       // tiles per epilogue warp (C=384: no shared memory left for them)
  static constexpr int SLAB_BYTES = RES_TMA ? 2 * 4096 : 32 * STG_PITCH;    // per epilogue warp
  static constexpr int OFF_R1 = X_BYTES;
  static constexpr int OFF_R2 = OFF_R1 + S1 * R1_UNIT;
  static constexpr int OFF_H = OFF_R2 + S2 * R2_UNIT;
  static constexpr int OFF_SLAB = OFF_H + NHS * 16384;
  static constexpr int LNW = 4;                                   // LayerNorm epilogue warps (a second warpgroup alternating tiles
                                                                  // was measured slower: 543 vs 508 us at C=192)
  static constexpr int THREADS = 32 * (3 + LNW + 8 + 1);          // TMA (X, W1), GEMM1 issuer, GEMM2 issuer, LayerNorm warps, 8 GELU warps, TMA (W2)
  static constexpr int OFF_PAR = OFF_SLAB + LNW * SLAB_BYTES;     // b1 [4C], b2 / gamma / beta [C] fp32
  static constexpr int OFF_TAB = OFF_PAR + 7 * C * 4;             // LNW warps x 64 ints
  static constexpr int OFF_BAR = OFF_TAB + LNW * 64 * 4;
  static constexpr int NUM_BARS = 2 * KX + 2 * S1 + 2 * S2 + 2 * NB + 2 * NHS + 2 * YB + 2 * LNW;
  static constexpr int SMEM_BYTES = 1024 + OFF_BAR + ((NUM_BARS * 8 + 16 + 127) / 128) * 128;
  static_assert(C == 192, "the single-kernel Mlp is built for C = 192 (at C = 384 neither TMEM nor shared memory has room)");
  static_assert(OFF_R1 % 1024 == 0 && OFF_R2 % 1024 == 0 && OFF_SLAB % (RES_TMA ? 1024 : 512) == 0, "operand alignment");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
};

struct MlpArgs {
  const float* b1;      // [4C]
  const float* b2;      // [C]
  const float* gamma;   // LayerNorm weight [C]
  const float* beta;    // LayerNorm bias [C]
  float* x32;           // residual stream [T, C], updated in place
  void* out16;          // 16-bit shadow of the new x32
  int T;                // tokens
  int num_tiles;        // ceil(T / 128)
  int Z, H, W;          // token grid (window scatter)
  int roll_out;         // < 0: out16 in natural order, else window order of that roll state
  float res_scale;      // DropPath factor (1 in eval)
  float eps;
  int debug;            // development ablations (timing only): bit2 LayerNorm epilogue reduced to its barrier handshakes,
                        // bit3 GELU arithmetic skipped
  long long* trace;     // development only (-DPANGU_ATTN_TRACE): clock64 timeline of one CTA, [role 8][chunk 64][event 4]
};

// D[tmem] (+)= A[tmem, 16-bit packed] * B[smem]
__device__ __forceinline__ void umma_f16_ts_(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32_(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16_(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait_() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int C, bool kFp16>
__global__ void __launch_bounds__(MlpTraits<C>::THREADS, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmRes, const MlpArgs a) {
  using T = MlpTraits<C>;
  constexpr int KX = T::KX, NH = T::NH, NCH = T::NCH, S1 = T::S1, S2 = T::S2, NB = T::NB, YB = T::YB;
  extern __shared__ uint8_t mf_raw[];
  uint8_t* smem = mf_raw + ((1024u - (smem_u32(mf_raw) & 1023u)) & 1023u);
  uint8_t* xs = smem;                          // X tile: KX slabs [128 rows][128 B], SWIZZLE_128B
  uint8_t* r1 = smem + T::OFF_R1;              // W1 units
  uint8_t* r2 = smem + T::OFF_R2;              // W2 units
  uint8_t* hs = smem + T::OFF_H;               // H operand buffers
  float* s_b1 = reinterpret_cast<float*>(smem + T::OFF_PAR);
  float* s_b2 = s_b1 + 4 * C;
  float* s_gamma = s_b2 + C;
  float* s_beta = s_gamma + C;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T::OFF_BAR);
  uint64_t* xfull = bars;                      // [KX]  X slab landed
  uint64_t* xempty = xfull + KX;               // [KX]  X slab no longer read by GEMM1 of this tile
  uint64_t* r1full = xempty + KX;              // [S1]
  uint64_t* r1empty = r1full + S1;             // [S1]  count 2: tcgen05.commit of both CTAs of the pair
  uint64_t* r2full = r1empty + S1;             // [S2]
  uint64_t* r2empty = r2full + S2;             // [S2]  count 2
  uint64_t* hfull = r2empty + S2;              // [NB]  Hacc ready (GEMM1 done)
  uint64_t* hempty = hfull + NB;               // [NB]  Hacc loaded into registers by the 4 GELU warps of its warpgroup
  uint64_t* sfull = hempty + NB;               // [NHS] H (16-bit, K-major SWIZZLE_128B) written to shared memory by 4 warps
  uint64_t* sempty = sfull + T::NHS;           // [NHS] GEMM2 has read the H buffer
  uint64_t* yfull = sempty + T::NHS;           // [YB]  Y complete
  uint64_t* yempty = yfull + YB;               // [YB]  Y drained by the 4 epilogue warps
  uint64_t* rfull = yempty + YB;               // [LNW][2] residual tile landed (RES_TMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rfull + 2 * T::LNW);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta_rank = int(cluster_ctarank());
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_units = (a.num_tiles + 1) >> 1;        // a unit = two consecutive tiles, one per CTA of the pair

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    if constexpr (T::RES_TMA) tma_prefetch_desc(&tmRes);
    for (int k = 0; k < KX; ++k) { mbar_init(&xfull[k], 1); mbar_init(&xempty[k], 1); }
    for (int s = 0; s < S1; ++s) { mbar_init(&r1full[s], 1); mbar_init(&r1empty[s], T::MCAST ? 2 : 1); }
    for (int s = 0; s < S2; ++s) { mbar_init(&r2full[s], 1); mbar_init(&r2empty[s], T::MCAST ? 2 : 1); }
    for (int b = 0; b < NB; ++b) { mbar_init(&hfull[b], 1); mbar_init(&hempty[b], 4); }
    for (int b = 0; b < T::NHS; ++b) { mbar_init(&sfull[b], 4); mbar_init(&sempty[b], 1); }
    for (int y = 0; y < YB; ++y) { mbar_init(&yfull[y], 1); mbar_init(&yempty[y], 4); }
    for (int i = 0; i < 2 * T::LNW; ++i) mbar_init(&rfull[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  // epilogue parameters (shared by all tiles)
  for (int i = threadIdx.x; i < 4 * C; i += T::THREADS) s_b1[i] = a.b1[i];
  for (int i = threadIdx.x; i < C; i += T::THREADS) { s_b2[i] = a.b2[i]; s_gamma[i] = a.gamma[i]; s_beta[i] = a.beta[i]; }
  tc_fence_before();
  cluster_sync_all();          // peer barriers are initialised before any multicast / remote commit
  tc_fence_after();
  if (*tmem_slot != 0u) __trap();      // all 512 columns: the allocation starts at TMEM address 0
  constexpr uint32_t tmem = 0u;

#ifdef PANGU_ATTN_TRACE       // development builds only: per-role clock64 timeline of CTA `a.debug >> 8` ([role 8][chunk 64][event 4])
  const bool tracing = a.trace != nullptr && int(blockIdx.x) == ((a.debug >> 8) & 255);
  auto TR = [&](int role, int g, int ev) { if (tracing && g < 64) a.trace[(role * 64 + g) * 4 + ev] = clock64(); };
#else
  auto TR = [](int, int, int) {};
#endif
  // Every lane waits.  (One waiting lane + __syncwarp was measured: a wait on an already completed barrier then takes ~800 clk
  // instead of ~300, and the issuer warps, which wait two or three times per chunk, slowed the kernels down by 60 %.)
  auto warp_wait = [&](uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); };
  if (warp == 0) {
    // ================================ TMA producer ================================
    // X(tile), then W1(c) per chunk, in the order the GEMM1 issuer consumes them
    if (lane == 0) {
      int p1 = 0;                // ring position (running counter)
      int xuse = 0;              // tiles loaded so far (X barrier phases)
      auto load_w1 = [&](int c) {          // chunk c (0..NCH-1): one unit of KX slabs [64 hidden x 64 k]
        const int s = p1 % S1;
        mbar_wait(&r1empty[s], ((p1 / S1) & 1) ^ 1);
        TR(0, p1, 0);
        mbar_arrive_expect_tx(&r1full[s], T::R1_UNIT);
        for (int k = 0; k < KX; ++k) {
          if constexpr (T::MCAST)          // this CTA fetches 32 of the 64 hidden rows of every slab and multicasts them to the pair
            tma_load_2d_mcast(&tmW1, &r1full[s], r1 + s * T::R1_UNIT + k * 8192 + cta_rank * 4096, k * 64, c * 64 + cta_rank * 32,
                              uint16_t(3), kEvictLast);
          else
            tma_load_2d_hint(&tmW1, &r1full[s], r1 + s * T::R1_UNIT + k * 8192, k * 64, c * 64, kEvictLast);
        }
        TR(0, p1, 1);
        ++p1;
      };
      auto load_x = [&](int tile, int use) {   // use-th X tile of this CTA; slabs free up as the previous tile's last GEMM1 retires
        for (int k = 0; k < KX; ++k) {
          mbar_wait(&xempty[k], (use & 1) ^ 1);
          mbar_arrive_expect_tx(&xfull[k], 16384);
          tma_load_2d_hint(&tmX, &xfull[k], xs + k * 16384, k * 64, tile * 128, kEvictFirst);   // rows >= T read as zero
        }
      };
      // flat chunk sequence over this CTA's tiles, same order as the MMA warp:  G1(cgx), G2(cgx - (NB-1))
      const int my_tiles = (num_units - pair + num_pairs - 1) / num_pairs;
      const int total = my_tiles * NCH;
      for (int cgx = 0; cgx < total; ++cgx) {
        const int tu = cgx / NCH, c = cgx % NCH;
        if (c == 0) load_x(2 * (pair + tu * num_pairs) + cta_rank, tu);
        if (c == NCH / 2 && tu + 1 < my_tiles)      // the X slabs are single-buffered: the next tile's load waits for this tile's last
          for (int k = 0; k < KX; ++k)              // GEMM1 and is on the critical path, so at least let it hit L2
            tma_prefetch_2d(&tmX, k * 64, (2 * (pair + (tu + 1) * num_pairs) + cta_rank) * 128);
        load_w1(c);
      }
      (void)xuse;
    }
  } else if (warp == 3 + T::LNW + 8) {
    // ================================ TMA producer of the W2 ring ================================
    // Its own warp: with one in-order thread for both rings a W2 unit that waited for GEMM2(c - 1) held back W1(c + 2),
    // which closed a loop GEMM2 -> W2 -> W1 -> GEMM1 -> GELU -> GEMM2 over several chunks (see mlp_fused2.cuh).
    if (lane == 0) {
      const int my_tiles = (num_units - pair + num_pairs - 1) / num_pairs;
      const int total = my_tiles * NCH * NH;
      for (int p2 = 0; p2 < total; ++p2) {          // unit p2: chunk (p2 / NH) % NCH, half p2 % NH  [192 out rows x 64 k]
        const int s = p2 % S2, c = (p2 / NH) % NCH, h = p2 % NH;
        mbar_wait(&r2empty[s], ((p2 / S2) & 1) ^ 1);
        TR(1, p2, 0);
        mbar_arrive_expect_tx(&r2full[s], T::R2_UNIT);
        if constexpr (T::MCAST)
          tma_load_2d_mcast(&tmW2, &r2full[s], r2 + s * T::R2_UNIT + cta_rank * 12288, c * 64, h * 192 + cta_rank * 96,
                            uint16_t(3), kEvictLast);
        else
          tma_load_2d_hint(&tmW2, &r2full[s], r2 + s * T::R2_UNIT, c * 64, h * 192, kEvictLast);
        TR(1, p2, 1);
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ================================ MMA issuers ================================
    // Warp 1 issues every GEMM1, warp 2 every GEMM2 (whole warp converged, warp-uniform operands, one elected lane
    // issues), so neither sits in the other's mbarrier round trips.  The hand-overs between the two GEMMs all go
    // through mbarriers: Hacc back to GEMM1 as soon as the GELU warps hold it in registers (hempty), H to GEMM2 through
    // a shared-memory operand buffer (sfull / sempty).
    constexpr uint32_t idesc1 = make_idesc_f16(128, 64, kFp16);
    constexpr uint32_t idesc2 = make_idesc_f16(128, 192, kFp16);
    const uint32_t xs_u32 = smem_u32(xs), r1_u32 = smem_u32(r1), r2_u32 = smem_u32(r2);
    int p1 = 0, p2 = 0;
    int tuse = 0;              // tiles started (X / Y barrier phases)
    auto gemm1 = [&](int c_in_tile, int tile_use) {      // chunk -> Hacc[cgx & 1]; cgx = global index of that chunk
      const int cgx = tile_use * NCH + c_in_tile, hb = cgx % NB;
      const int s = p1 % S1;
      warp_wait(&hempty[hb], ((cgx / NB) & 1) ^ 1);       // the GELU warps hold the buffer's previous contents in registers
      if (lane == 0) TR(2, cgx, 0);
      tc_fence_after();
      if (c_in_tile == 0) {
        for (int k = 0; k < KX; ++k) warp_wait(&xfull[k], tile_use & 1);
      }
      warp_wait(&r1full[s], (p1 / S1) & 1);
      if (lane == 0) TR(2, cgx, 1);
      tc_fence_after();
      const uint32_t xa = xs_u32, wb = r1_u32 + s * T::R1_UNIT;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < KX; ++k) {
          const uint64_t da = make_sdesc_sw128(xa + k * 16384);
          const uint64_t db = make_sdesc_sw128(wb + k * 8192);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16_ss(tmem + T::COL_H + 64 * hb, da + uint64_t(kk * 2), db + uint64_t(kk * 2), idesc1, (k | kk) != 0 ? 1u : 0u);
        }
        if constexpr (T::MCAST) umma_commit_mcast(&r1empty[s], uint16_t(3)); else umma_commit(&r1empty[s]);
        if (c_in_tile == NCH - 1) {
          for (int k = 0; k < KX; ++k) umma_commit(&xempty[k]);     // last reader of the X slabs
        }
        umma_commit(&hfull[hb]);
        TR(2, cgx, 2);
      }
      __syncwarp();
      ++p1;
    };
    const uint32_t hs_u32 = smem_u32(hs);
    auto gemm2 = [&](int c_in_tile, int tile_use) {
      const int cgx = tile_use * NCH + c_in_tile, sb = cgx % T::NHS;
      warp_wait(&sfull[sb], (cgx / T::NHS) & 1);         // GELU(H) of this chunk is in shared memory
      if (lane == 0) TR(3, cgx, 0);
      const int yb = tile_use % YB;
      if (c_in_tile == 0) warp_wait(&yempty[yb], ((tile_use / YB) & 1) ^ 1);   // the epilogue has drained this Y buffer
      if (lane == 0) TR(3, cgx, 1);
      tc_fence_after();
      const uint64_t da = make_sdesc_sw128(hs_u32 + sb * 16384);
      for (int h = 0; h < NH; ++h, ++p2) {
        const int s = p2 % S2;
        warp_wait(&r2full[s], (p2 / S2) & 1);
        if (lane == 0) TR(3, cgx, 2);
        tc_fence_after();
        const uint64_t db = make_sdesc_sw128(r2_u32 + s * T::R2_UNIT);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)     // K = 64 hidden units
            umma_f16_ss(tmem + yb * C + h * 192, da + uint64_t(kk * 2), db + uint64_t(kk * 2), idesc2, (c_in_tile | kk) != 0 ? 1u : 0u);
          if constexpr (T::MCAST) umma_commit_mcast(&r2empty[s], uint16_t(3)); else umma_commit(&r2empty[s]);
          if (h == NH - 1) umma_commit(&sempty[sb]);
          if (c_in_tile == NCH - 1 && h == NH - 1) umma_commit(&yfull[yb]);
          TR(3, cgx, 3);
        }
        __syncwarp();
      }
    };
    const int my_tiles = (num_units - pair + num_pairs - 1) / num_pairs;
    const int total = my_tiles * NCH;
    if (warp == 1) {
      for (int cgx = 0; cgx < total; ++cgx) gemm1(cgx % NCH, cgx / NCH);
    } else {
      for (int cgx = 0; cgx < total; ++cgx) gemm2(cgx % NCH, cgx / NCH);
    }
    (void)tuse;
  } else if (warp < 3 + T::LNW) {
    // ================================ LayerNorm + residual epilogue ================================
    // one warp per TMEM lane quadrant; thread = one token row (statistics are thread-local)
    const int quad = warp & 3;
    const int ew = warp - 3;
    uint8_t* slab = smem + T::OFF_SLAB + ew * T::SLAB_BYTES;
    int* s_tok = reinterpret_cast<int*>(smem + T::OFF_TAB) + ew * 64;
    int* s_dst = s_tok + 32;
    const Geo geo = make_geo(a.Z, a.H, a.W);
    // LayerNorm warpgroup w of LNW / 4 takes this CTA's tiles w, w + LNW / 4, ...
    for (int tuse = ew >> 2; pair + tuse * num_pairs < num_units; tuse += T::LNW / 4) {
      const int unit = pair + tuse * num_pairs;
      const int tile = 2 * unit + cta_rank;
      const int yb = tuse % YB;
      const uint32_t tacc = tmem + (uint32_t(quad * 32) << 16) + yb * C;
      __syncwarp();
      {
        const int g = tile * 128 + quad * 32 + lane;
        const int tok = g < a.T ? g : -1;
        s_tok[lane] = tok;
        s_dst[lane] = (tok >= 0 && a.roll_out >= 0) ? token_to_win_row(geo, tok, a.roll_out) : tok;
      }
      __syncwarp();
      if (a.debug & 4) {
        warp_wait(&yfull[yb], (tuse / YB) & 1);
        tc_fence_after();
        tc_fence_before();
        if (lane == 0) mbar_arrive(&yempty[yb]);
        continue;
      }
      {
        // ---- residual stream by TMA: two [32 rows x 32 fp32 columns] SWIZZLE_128B tiles per warp; chunk c lives in
        // slot c & 1; the LayerNorm result is added in place and the tile leaves with one bulk store.
        constexpr int NCHK = C / 32;
        uint8_t* slots = smem + T::OFF_SLAB + ew * T::SLAB_BYTES;
        uint64_t* rf = rfull + 2 * ew;
        const int row0 = tile * 128 + quad * 32;
        auto fetch = [&](int c) {            // lane 0
          mbar_arrive_expect_tx(&rf[c & 1], 4096);
          tma_load_2d(&tmRes, &rf[c & 1], slots + (c & 1) * 4096, c * 32, row0);
        };
        if (lane == 0) {
          bulk_wait_read<0>();               // the previous tile's stores have read both slots
          fetch(0);
          fetch(1);
        }
        if (lane == 0 && warp == 3) TR(6, tuse, 0);
        warp_wait(&yfull[yb], (tuse / YB) & 1);
        if (lane == 0 && warp == 3) TR(6, tuse, 1);
        tc_fence_after();
        float mean, rstd;
        {
          float shift = 0.f;
          f32x2 s1 = pack2(0.f, 0.f), s2 = pack2(0.f, 0.f);
#pragma unroll 1
          for (int c0 = 0; c0 < C; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tacc + c0, r);
            tmem_ld_wait();
            if (c0 == 0) shift = __uint_as_float(r[0]) + s_b2[0];
            const f32x2 nshift = pack2(-shift, -shift);
            const float4* b4 = reinterpret_cast<const float4*>(s_b2 + c0);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 bb = b4[j4];
              const f32x2 v01 = add2(add2(pack2(__uint_as_float(r[4 * j4]), __uint_as_float(r[4 * j4 + 1])), pack2(bb.x, bb.y)), nshift);
              const f32x2 v23 = add2(add2(pack2(__uint_as_float(r[4 * j4 + 2]), __uint_as_float(r[4 * j4 + 3])), pack2(bb.z, bb.w)), nshift);
              s1 = add2(s1, add2(v01, v23));
              s2 = fma2(v01, v01, s2);
              s2 = fma2(v23, v23, s2);
            }
          }
          float s1a, s1b, s2a, s2b;
          unpack2(s1, s1a, s1b);
          unpack2(s2, s2a, s2b);
          const float inv_n = 1.0f / float(C);
          const float m = (s1a + s1b) * inv_n;
          const float var = fmaxf((s2a + s2b) * inv_n - m * m, 0.f);
          mean = shift + m;
          rstd = rsqrtf(var + a.eps);
        }
        if (lane == 0 && warp == 3) TR(6, tuse, 2);
        const f32x2 ln_a = pack2(rstd * a.res_scale, rstd * a.res_scale), ln_b = pack2(-mean * rstd * a.res_scale, -mean * rstd * a.res_scale);
        const f32x2 rs2 = pack2(a.res_scale, a.res_scale);
#pragma unroll 1
        for (int c = 0; c < NCHK; ++c) {
          const int c0 = c * 32;
          uint8_t* tl = slots + (c & 1) * 4096;
          uint32_t r[32];
          tmem_ld32(tacc + c0, r);
          warp_wait(&rf[c & 1], (tuse * (NCHK / 2) + (c >> 1)) & 1);     // each slot is filled NCHK / 2 times per tile
          tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(s_b2 + c0);
          const float4* g4 = reinterpret_cast<const float4*>(s_gamma + c0);
          const float4* e4 = reinterpret_cast<const float4*>(s_beta + c0);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {      // x + s * ((acc + b2 - mean) * rstd * gamma + beta), in place (row per lane)
            uint4* cell = reinterpret_cast<uint4*>(tl + lane * 128 + ((j4 ^ (lane & 7)) << 4));
            const uint4 q = *cell;
            const float4 bb = b4[j4], gg = g4[j4], ee = e4[j4];
            f32x2 v01 = add2(pack2(__uint_as_float(r[4 * j4]), __uint_as_float(r[4 * j4 + 1])), pack2(bb.x, bb.y));
            f32x2 v23 = add2(pack2(__uint_as_float(r[4 * j4 + 2]), __uint_as_float(r[4 * j4 + 3])), pack2(bb.z, bb.w));
            v01 = fma2(fma2(v01, ln_a, ln_b), pack2(gg.x, gg.y), fma2(rs2, pack2(ee.x, ee.y), pack2(__uint_as_float(q.x), __uint_as_float(q.y))));
            v23 = fma2(fma2(v23, ln_a, ln_b), pack2(gg.z, gg.w), fma2(rs2, pack2(ee.z, ee.w), pack2(__uint_as_float(q.z), __uint_as_float(q.w))));
            float v0, v1, v2, v3;
            unpack2(v01, v0, v1);
            unpack2(v23, v2, v3);
            *cell = make_uint4(__float_as_uint(v0), __float_as_uint(v1), __float_as_uint(v2), __float_as_uint(v3));
          }
          if (c == NCHK - 1) tc_fence_before();      // accumulator fully read (every lane: tcgen05.wait::ld above)
          fence_proxy_async_smem();
          __syncwarp();
          if (c == NCHK - 1 && lane == 0) mbar_arrive(&yempty[yb]);      // hand the Y buffer back to the GEMM2 issuer
          if (lane == 0 && row0 < a.T) {
            tma_store_2d(&tmRes, tl, c0, row0);     // rows beyond T are clipped
            bulk_commit();
          }
          // 16-bit shadow from the finished tile: 64 B rows, 4 lanes per row, optionally scattered to window order
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int id = it * 32 + lane;
            const int rr = id >> 2, pc = id & 3;
            const int dst = s_dst[rr];
            if (dst < 0) continue;
            const uint4 a0 = *reinterpret_cast<const uint4*>(tl + rr * 128 + (((2 * pc) ^ (rr & 7)) << 4));
            const uint4 a1 = *reinterpret_cast<const uint4*>(tl + rr * 128 + (((2 * pc + 1) ^ (rr & 7)) << 4));
            uint4 h;
            h.x = pack16<kFp16>(__uint_as_float(a0.x), __uint_as_float(a0.y));
            h.y = pack16<kFp16>(__uint_as_float(a0.z), __uint_as_float(a0.w));
            h.z = pack16<kFp16>(__uint_as_float(a1.x), __uint_as_float(a1.y));
            h.w = pack16<kFp16>(__uint_as_float(a1.z), __uint_as_float(a1.w));
            stg16(reinterpret_cast<uint16_t*>(a.out16) + size_t(dst) * C + c0 + pc * 8, h);
          }
          __syncwarp();
          if (lane == 0 && c + 2 < NCHK) {
            bulk_wait_read<0>();           // the store above has read this slot
            fetch(c + 2);
          }
        }
        if (lane == 0 && warp == 3) TR(6, tuse, 3);
      }
    }
  } else {
    // ================================ GELU warps ================================
    // Warpgroup w owns the chunks with (global chunk index & 1) == w, Hacc buffer w and H buffer w.  It pulls the fp32
    // accumulator into registers and hands the TMEM buffer straight back (GEMM1 of chunk c + 2 may start), applies
    // bias + GELU and writes the 16-bit H tile to shared memory as the K-major SWIZZLE_128B A operand of GEMM2.
    const int quad = warp & 3, wgp = (warp - (3 + T::LNW)) >> 2;
    const uint32_t haddr = tmem + (uint32_t(quad * 32) << 16) + T::COL_H + 64 * wgp;
    const int row = quad * 32 + lane;
    uint8_t* hrow = hs + wgp * 16384 + row * 128;
    int n = 0;           // chunks this warpgroup has processed
    for (int unit = pair; unit < num_units; unit += num_pairs) {
      for (int c = wgp; c < NCH; c += 2, ++n) {
        warp_wait(&hfull[wgp], n & 1);
        if (lane == 0 && quad == 3) TR(4 + wgp, 2 * n + wgp, 0);
        tc_fence_after();
        uint32_t r[2][32];
        tmem_ld32(haddr, r[0]);
        tmem_ld32(haddr + 32, r[1]);
        tmem_ld_wait();
        if (lane == 0 && quad == 3 && wgp == 0) TR(7, n, 0);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&hempty[wgp]);      // one arrival per warp: all its lanes hold their Hacc values
        if (lane == 0 && quad == 3 && wgp == 0) TR(7, n, 1);
        uint32_t pk[32];
        const float4* b4 = reinterpret_cast<const float4*>(s_b1 + c * 64);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {       // 8 hidden units at a time (two 16 B bias reads, four interleaved pair chains)
            const float4 ba = b4[hh * 8 + 2 * j8], bb = b4[hh * 8 + 2 * j8 + 1];
            float v[8] = {__uint_as_float(r[hh][8 * j8]) + ba.x, __uint_as_float(r[hh][8 * j8 + 1]) + ba.y,
                          __uint_as_float(r[hh][8 * j8 + 2]) + ba.z, __uint_as_float(r[hh][8 * j8 + 3]) + ba.w,
                          __uint_as_float(r[hh][8 * j8 + 4]) + bb.x, __uint_as_float(r[hh][8 * j8 + 5]) + bb.y,
                          __uint_as_float(r[hh][8 * j8 + 6]) + bb.z, __uint_as_float(r[hh][8 * j8 + 7]) + bb.w};
            if (!(a.debug & 8)) gelu_erf8(v);
#pragma unroll
            for (int q = 0; q < 4; ++q) pk[hh * 16 + 4 * j8 + q] = pack16<kFp16>(v[2 * q], v[2 * q + 1]);
          }
        }
        if (lane == 0 && quad == 3) TR(4 + wgp, 2 * n + wgp, 1);
        warp_wait(&sempty[wgp], (n & 1) ^ 1);        // GEMM2 of this warpgroup's previous chunk has read the buffer
        if (lane == 0 && quad == 3) TR(4 + wgp, 2 * n + wgp, 2);
        if (lane == 0 && quad == 3 && wgp == 0) TR(7, n, 2);
#pragma unroll
        for (int q = 0; q < 8; ++q)                  // 8 x 16 B = this row's 64 hidden units, XOR-swizzled 16 B chunks
          *reinterpret_cast<uint4*>(hrow + ((q ^ (row & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        if (lane == 0 && quad == 3 && wgp == 0) TR(7, n, 3);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sfull[wgp]);       // one arrival per warp: its 32 rows of H are visible to the async proxy
        if (lane == 0 && quad == 3) TR(4 + wgp, 2 * n + wgp, 3);
      }
    }
  }

  if constexpr (T::RES_TMA) {
    if (warp >= 3 && warp < 3 + T::LNW && lane == 0) bulk_wait_all();     // shared memory must outlive the bulk stores
  }
  tc_fence_before();
  cluster_sync_all();          // no multicast / remote commit may target an exited CTA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

}  // namespace pg
