// Single-kernel Mlp + LayerNorm + residual at C = 384 on CTA PAIRS (reference models/layers.py:250-251, 264-270):
//
//     x <- x + s * LN2( GELU(x16 W1^T + b1) W2^T + b2 )
//
// The C = 192 kernel (mlp_fused.cuh) keeps X, two weight rings and the H operand in one SM's shared memory and two Y
// accumulators in its TMEM; at C = 384 neither fits (X alone is 96 KB, one Y is 384 of the 512 TMEM columns).  Here a CTA
// pair works on 256 tokens with tcgen05.mma.cta_group::2: every MMA has M = 256 (rows 0..127 in CTA 0, 128..255 in CTA 1)
// and each CTA holds only HALF of every weight tile, which is what makes the shared-memory budget close:
//
//   X tile     [128 x 384] 16-bit, own rows                                  96 KB   (resident for the tile)
//   W1 ring    3 units [32 of the 64 hidden rows of a chunk x 192 k]      3 x 12 KB   (1.5 chunks: GEMM1 runs two chunks ahead anyway)
//   W2 ring    4 units [96 of the 192 output rows of an N half x 64 k]    4 x 12 KB   (2 chunks: with 3 units a load could only be
//              issued one GEMM2 half before it was needed and its ~900 clk latency stalled every chunk)
//   H operand  2 buffers [128 x 64] 16-bit (GELU output, A of GEMM2)      2 x 16 KB
//   b1, b2, gamma, beta 10.5 KB, barriers; the LayerNorm epilogue stages its chunks in the (then idle) H buffers     total 222.9 KB
//
// TMEM (per CTA, its 128 rows): Y [0,384) | Hacc0 [384,448) | Hacc1 [448,512).
// Per chunk c of 64 hidden units (24 per tile):  G1(c): Hacc[c&1] = X W1[c]^T  (M 256, N 64, K 384);  GELU warps:
// Hacc -> registers -> bias + exact GELU -> H[c&1] in shared memory (K-major SWIZZLE_128B);  G2(c): Y += H W2[:, c]^T
// (M 256, N 2 x 192, K 64).  With a single Y accumulator the LayerNorm epilogue of a tile overlaps only the first two
// G1 chunks of the next one (~10 % of a tile).
//
// Only CTA 0 (the leader) issues MMAs.  Hand-overs that involve both CTAs:
//   * TMA loads of both CTAs complete on the LEADER's full barriers (cta_group::2 form of cp.async.bulk.tensor);
//   * "slot / buffer free" and "accumulator ready" signals are tcgen05.commit.cta_group::2 multicasts to both CTAs;
//   * "my threads are done with X" goes straight to the LEADER's barrier from both CTAs (the peer's threads arrive
//     remotely): relaxed arrives where only reads have to be finished (Hacc in registers, Y drained: tcgen05.wait::ld has
//     completed them), one release.cluster arrive per GELU warp for the H tile it wrote to its shared memory.
//     (A first version relayed the peer's local barriers through its idle issuer warps: 461 us per launch.)
// Warps (512 threads): 0 TMA producer (X, W1, residual rows), 1 GEMM1 issuer (leader only), 2 GEMM2 issuer (leader only),
// 3-6 LayerNorm epilogue, 7-14 GELU (two warpgroups alternate chunks) + their share of the epilogue, 15 W2 producer.
#pragma once
#include "common.cuh"
#include "geometry.cuh"
#include "mlp_fused.cuh"

namespace pg {

struct Mlp2Traits {
  static constexpr int C = 384, KX = 6, NCH = 24, NB = 2, NHS = 2, S1 = 3, S2 = 4, LNW = 4;
  static constexpr int X_BYTES = KX * 16384;
  static constexpr int R1_UNIT = 3 * 4096;             // 32 hidden rows x 64 k per slab, 3 slabs = half of a chunk's K
  static constexpr int R2_UNIT = 96 * 128;             // 96 output rows x 64 k
  static constexpr int OFF_R1 = X_BYTES;
  static constexpr int OFF_R2 = OFF_R1 + S1 * R1_UNIT;
  static constexpr int OFF_H = OFF_R2 + S2 * R2_UNIT;
  // The LayerNorm epilogue stages its [32 rows x 32 fp32] chunks in the H operand buffers (2 x 4 KB per epilogue warp = the
  // 32 KB of the two H buffers): between "Y complete" and the end of the epilogue no GEMM2 reads H, and the GELU warps wait
  // for the epilogue (lnfree) before they write the first H chunks of the next tile.
  static constexpr int OFF_PAR = OFF_H + NHS * 16384;  // b1 [4C], b2 / gamma / beta [C] fp32
  static constexpr int OFF_STAT = OFF_PAR + 7 * C * 4; // LayerNorm partial sums [3 column thirds][128 rows] float2
  static constexpr int OFF_BAR = OFF_STAT + 3 * 128 * 8;
  static constexpr int NUM_BARS = 2 * KX + 2 * S1 + 2 * S2 + (6 + NB) + (NHS + 6) + 2 + 1 + 2 * 12;
  static constexpr int SMEM_BYTES = OFF_BAR + ((NUM_BARS * 8 + 16 + 127) / 128) * 128;
  static constexpr int THREADS = 32 * (3 + LNW + 8 + 1);     // + the W2 producer warp
  static_assert(OFF_R1 % 1024 == 0 && OFF_R2 % 1024 == 0 && OFF_H % 1024 == 0 && R1_UNIT % 1024 == 0 && R2_UNIT % 1024 == 0,
                "operand alignment");
  static_assert(NCH % 6 == 0, "three warpgroups x two buffers");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
};

// "My H rows are written" from the peer CTA's GELU warps to the leader's barrier.  The rows live in the PEER's shared memory
// and are read there, by the peer's own tensor core (cta_group::2: each CTA supplies its 128 rows of A), through the async
// proxy: fence.proxy.async + __syncwarp have made them visible before lane 0 sends the arrive, and the MMA is dispatched
// only after the arrive has reached the leader.  So the default (CTA-scope release) remote arrive that CUTLASS's
// ClusterBarrier::arrive(cta_id) uses is enough; the .release.cluster form costs an ERRBAR + CCTL.IVALL pair, ~1 400 clk
// on the GEMM2 critical path of every chunk.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t rbar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
}

template <bool kFp16>
__global__ void __launch_bounds__(Mlp2Traits::THREADS, 1)
mlp_fused2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                  const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmR, const MlpArgs a) {
  using T = Mlp2Traits;
  constexpr int C = T::C, KX = T::KX, NCH = T::NCH, S1 = T::S1, S2 = T::S2, NB = T::NB, NHS = T::NHS;
  extern __shared__ __align__(1024) uint8_t mf2_raw[];
  if ((smem_u32(mf2_raw) & 1023u) != 0u) __trap();     // SWIZZLE_128B tiles need 1024 B alignment; there is no slack to fix it up
  uint8_t* smem = mf2_raw;
  uint8_t* xs = smem;
  uint8_t* r1 = smem + T::OFF_R1;
  uint8_t* r2 = smem + T::OFF_R2;
  uint8_t* hs = smem + T::OFF_H;
  float* s_b1 = reinterpret_cast<float*>(smem + T::OFF_PAR);
  float* s_b2 = s_b1 + 4 * C;
  float* s_gamma = s_b2 + C;
  float* s_beta = s_gamma + C;
  float2* s_stat = reinterpret_cast<float2*>(smem + T::OFF_STAT);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T::OFF_BAR);
  uint64_t* xfull = bars;                      // [KX]  leader: X slabs of both CTAs landed
  uint64_t* xempty = xfull + KX;               // [KX]  each CTA: GEMM1 of the tile no longer reads the slab (commit multicast)
  uint64_t* r1full = xempty + KX;              // [S1]  leader
  uint64_t* r1empty = r1full + S1;             // [S1]  each CTA (commit multicast)
  uint64_t* r2full = r1empty + S1;             // [S2]  leader
  uint64_t* r2empty = r2full + S2;             // [S2]  each CTA
  // hfull / sempty are waited for by the GELU warpgroups, which rotate over the chunks (chunk c -> warpgroup c % 3) while
  // the buffers alternate (c % 2): one barrier per c % 6, so that a barrier always has the same waiting warpgroup and a
  // parity wait can never be more than one phase behind.
  uint64_t* hfull = r2empty + S2;              // [6]   each CTA: Hacc[c % 2] holds chunk c (commit multicast)
  uint64_t* hempty = hfull + 6;                // [NB]  leader: the 2 x 4 GELU warps of the pair hold Hacc in registers
  uint64_t* sfull = hempty + NB;               // [NHS] leader: the 2 x 4 GELU warps of the pair have written H
  uint64_t* sempty = sfull + NHS;              // [6]   each CTA: GEMM2 of chunk c has read H[c % 2] (commit multicast)
  uint64_t* yfull = sempty + 6;                // [1]   each CTA: Y complete (commit multicast)
  uint64_t* yempty = yfull + 1;                // [1]   leader: the 2 x 12 epilogue warps of the pair have drained Y
  uint64_t* lnfree = yempty + 1;               // [1]   each CTA: its 12 epilogue warps no longer use the H buffers as staging
  uint64_t* rfb = lnfree + 1;                  // [12][2] per epilogue warp: a [32 x 16] fp32 piece of its residual rows landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rfb + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta_rank = int(cluster_ctarank());
  const bool leader = cta_rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_units = (a.num_tiles + 1) >> 1;        // a unit = two consecutive 128-token tiles, one per CTA of the pair
  const int my_units = (num_units - pair + num_pairs - 1) / num_pairs;
  const int total = my_units * NCH;                    // chunks this pair walks

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmR);
    for (int k = 0; k < KX; ++k) { mbar_init(&xfull[k], 1); mbar_init(&xempty[k], 1); }
    for (int i = 0; i < 24; ++i) mbar_init(&rfb[i], 1);
    for (int s = 0; s < S1; ++s) { mbar_init(&r1full[s], 1); mbar_init(&r1empty[s], 1); }
    for (int s = 0; s < S2; ++s) { mbar_init(&r2full[s], 1); mbar_init(&r2empty[s], 1); }
    for (int b = 0; b < 6; ++b) { mbar_init(&hfull[b], 1); mbar_init(&sempty[b], 1); }
    for (int b = 0; b < NB; ++b) mbar_init(&hempty[b], 8);
    for (int b = 0; b < NHS; ++b) mbar_init(&sfull[b], 8);
    mbar_init(yfull, 1); mbar_init(yempty, 24); mbar_init(lnfree, 12);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_cta2<512>(tmem_slot);
  for (int i = threadIdx.x; i < 4 * C; i += T::THREADS) s_b1[i] = a.b1[i];
  for (int i = threadIdx.x; i < C; i += T::THREADS) { s_b2[i] = a.b2[i]; s_gamma[i] = a.gamma[i]; s_beta[i] = a.beta[i]; }
  tc_fence_before();
  cluster_sync_all();          // peer barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  if (*tmem_slot != 0u) __trap();      // all 512 columns: the allocation starts at TMEM address 0
  constexpr uint32_t tmem = 0u;
  constexpr uint32_t COL_H = 384;
#ifdef PANGU_ATTN_TRACE       // development builds only: per-role clock64 timeline of CTA `a.debug >> 8` ([role 8][chunk 64][event 4])
  const bool tracing = a.trace != nullptr && int(blockIdx.x) == ((a.debug >> 8) & 255);
  auto TR = [&](int role, int g, int ev) { if (tracing && g < 64) a.trace[(role * 64 + g) * 4 + ev] = clock64(); };
#else
  auto TR = [](int, int, int) {};
#endif
#ifdef PANGU_DEV_SWITCHES     // timing ablations (results invalid): bit2 LayerNorm epilogue reduced to its handshakes, bit3 no GELU arithmetic,
  const int dbg = a.debug;    // bit4 no GEMM1 MMAs, bit5 no GEMM2 MMAs
#else
  constexpr int dbg = 0;
#endif

  // Every lane waits.  (One waiting lane + __syncwarp was measured: a wait on an already completed barrier then takes ~800 clk
  // instead of ~300, and the issuer warps, which wait two or three times per chunk, slowed the kernels down by 60 %.)
  auto warp_wait = [&](uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); };
  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ================================
    // issue order == consumption order of the MMA warps:  X(tile), W1(0), then per chunk  W1(c+1), W2(c)
    if (lane == 0) {
      int p1 = 0;
      auto lbar = [&](uint64_t* b) { return mapa_u32(smem_u32(b), 0); };
      auto load_w1 = [&](int c) {
        for (int h = 0; h < 2; ++h, ++p1) {     // two units per chunk: k slabs 0..2 and 3..5
          const int s = p1 % S1;
          mbar_wait(&r1empty[s], ((p1 / S1) & 1) ^ 1);
          TR(0, p1 >> 1, 2 * h);
          if (leader) mbar_arrive_expect_tx(&r1full[s], 2 * T::R1_UNIT);
          for (int k = 0; k < 3; ++k)     // this CTA's 32 of the chunk's 64 hidden rows
            tma_load_2d_cta2(&tmW1, lbar(&r1full[s]), r1 + s * T::R1_UNIT + k * 4096, (3 * h + k) * 64, c * 64 + cta_rank * 32, kEvictLast);
          TR(0, p1 >> 1, 2 * h + 1);
        }
      };
      auto tile_of = [&](int tu) { return 2 * (pair + tu * num_pairs) + cta_rank; };
      auto load_x = [&](int tile, int use) {         // slabs free up as the previous tile's last GEMM1 retires
        for (int k = 0; k < KX; ++k) {
          mbar_wait(&xempty[k], (use & 1) ^ 1);
          if (leader) mbar_arrive_expect_tx(&xfull[k], 2 * 16384);
          tma_load_2d_cta2(&tmX, lbar(&xfull[k]), xs + k * 16384, k * 64, tile * 128, kEvictFirst);   // rows >= T read as zero
        }
      };
      // The LayerNorm epilogue reads and rewrites this CTA's 128 x 384 fp32 residual rows (196 KB, contiguous).  All CTAs reach
      // their epilogues at about the same time and nothing else of the kernel touches HBM then, so the rows are pulled into
      // L2 half a tile earlier; the epilogue warps fetch them from there by TMA, piece by piece (see below).
      auto prefetch_resid = [&](int tile) {
        const long long row0 = (long long)tile * 128;
        const long long rows = row0 + 128 <= a.T ? 128 : (a.T > row0 ? a.T - row0 : 0);
        const uint8_t* p = reinterpret_cast<const uint8_t*>(a.x32 + row0 * C);
        for (long long off = 0; off < rows * C * 4; off += 16384)
          bulk_prefetch_l2(p + off, uint32_t(rows * C * 4 - off < 16384 ? rows * C * 4 - off : 16384));
      };
      // X(tile), then W1(c) per chunk.  The W2 ring has its own producer warp: with one in-order thread for both rings a W2
      // unit that waited for GEMM2(g-1) held back W1(g+2), which closed a three-chunk loop GEMM2 -> W2 -> W1 -> GEMM1 -> GELU ->
      // GEMM2 (3 400 clk per chunk against 1 900 of MMA work).  Nothing here waits for the epilogue: the next tile's X and its
      // first W1 units land, and its first two GEMM1 chunks run, while the epilogue of this tile is still busy.
      for (int g = 0; g < total; ++g) {
        const int tu = g / NCH, c = g % NCH;
        if (c == 0) load_x(tile_of(tu), tu);
        if (c == NCH / 2) prefetch_resid(tile_of(tu));
        load_w1(c);
      }
    }
  } else if (warp == 3 + T::LNW + 8) {
    // ================================ W2 producer (both CTAs) ================================
    if (lane == 0) {
      auto lbar = [&](uint64_t* b) { return mapa_u32(smem_u32(b), 0); };
      for (int p2 = 0; p2 < 2 * total; ++p2) {
        const int s = p2 % S2, c = (p2 >> 1) % NCH, h = p2 & 1, tu = (p2 >> 1) / NCH;
        if (c == 0 && h == 0 && tu > 0) mbar_wait(lnfree, (tu - 1) & 1);      // the epilogue stages its residual pieces in this ring
        mbar_wait(&r2empty[s], ((p2 / S2) & 1) ^ 1);
        TR(1, p2 >> 1, 2 * h);
        if (leader) mbar_arrive_expect_tx(&r2full[s], 2 * T::R2_UNIT);
        tma_load_2d_cta2(&tmW2, lbar(&r2full[s]), r2 + s * T::R2_UNIT, c * 64, h * 192 + cta_rank * 96, kEvictLast);
        TR(1, p2 >> 1, 2 * h + 1);
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ================================ GEMM1 issuer ================================
      constexpr uint32_t idesc1 = make_idesc_f16(256, 64, kFp16);
      const uint32_t xs_u32 = smem_u32(xs), r1_u32 = smem_u32(r1);
      for (int cgx = 0; cgx < total; ++cgx) {
        const int c = cgx % NCH, tu = cgx / NCH, hb = cgx % NB;
        warp_wait(&hempty[hb], ((cgx / NB) & 1) ^ 1);        // both CTAs' GELU warps hold the buffer's previous contents
        if (lane == 0) TR(2, cgx, 0);
        if (c == 0) {
          for (int k = 0; k < KX; ++k) warp_wait(&xfull[k], tu & 1);
        }
        for (int h = 0; h < 2; ++h) {
          const int s = (2 * cgx + h) % S1;
          warp_wait(&r1full[s], ((2 * cgx + h) / S1) & 1);
          if (lane == 0 && h == 1) TR(2, cgx, 1);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              const uint64_t da = make_sdesc_sw128(xs_u32 + (3 * h + k) * 16384);
              const uint64_t db = make_sdesc_sw128(r1_u32 + s * T::R1_UNIT + k * 4096);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                if (!(dbg & 16))
                umma_f16_ss_cta2(tmem + COL_H + 64 * hb, da + uint64_t(kk * 2), db + uint64_t(kk * 2), idesc1, (h | k | kk) != 0 ? 1u : 0u);
            }
            umma_commit_cta2(&r1empty[s]);
          }
          __syncwarp();
        }
        if (elect_one()) {
          if (c == NCH - 1) {
            for (int k = 0; k < KX; ++k) umma_commit_cta2(&xempty[k]);     // last reader of the X slabs
          }
          umma_commit_cta2(&hfull[cgx % 6]);
          TR(2, cgx, 2);
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    if (leader) {
      // ================================ GEMM2 issuer ================================
      constexpr uint32_t idesc2 = make_idesc_f16(256, 192, kFp16);
      const uint32_t hs_u32 = smem_u32(hs), r2_u32 = smem_u32(r2);
      int p2 = 0;
      for (int cgx = 0; cgx < total; ++cgx) {
        const int c = cgx % NCH, tu = cgx / NCH, sb = cgx % NHS;
        if (c == 0) {                                        // both CTAs' LayerNorm warps have drained Y of the previous tile
          warp_wait(yempty, (tu & 1) ^ 1);
        }
        mbar_wait_cluster(&sfull[sb], (cgx / NHS) & 1);      // GELU(H) of this chunk is in both CTAs' shared memory
        if (lane == 0) TR(3, cgx, 0);
        tc_fence_after();
        const uint64_t da = make_sdesc_sw128(hs_u32 + sb * 16384);
        for (int h = 0; h < 2; ++h, ++p2) {
          const int s = p2 % S2;
          warp_wait(&r2full[s], (p2 / S2) & 1);
          if (lane == 0) TR(3, cgx, 1 + h);
          tc_fence_after();
          const uint64_t db = make_sdesc_sw128(r2_u32 + s * T::R2_UNIT);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)     // K = 64 hidden units
              if (!(dbg & 32))
              umma_f16_ss_cta2(tmem + h * 192, da + uint64_t(kk * 2), db + uint64_t(kk * 2), idesc2, (c | kk) != 0 ? 1u : 0u);
            umma_commit_cta2(&r2empty[s]);
            if (h == 1) umma_commit_cta2(&sempty[cgx % 6]);
            if (c == NCH - 1 && h == 1) umma_commit_cta2(yfull);
            if (h == 1) TR(3, cgx, 3);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================================ GELU + LayerNorm warps (both CTAs, own 128 rows) ================================
    // Warps 3-14: three warpgroups take the chunks in turn (chunk c -> warpgroup c % 3, Hacc / H buffer c % 2):
    // Hacc -> registers (the TMEM buffer goes straight back to GEMM1) -> bias + exact GELU -> 16-bit H tile in shared memory
    // (K-major SWIZZLE_128B A operand of GEMM2).  One chunk costs a warp ~2 200 clk (a dependent Horner chain per pair,
    // ~0.4 IPC) plus ~1 500 for the hand-over; with two warpgroups that cycle was longer than two chunks of MMA work, and the
    // three warps per scheduler also hide each other's latencies.
    // At the end of a tile the same twelve warps (three per TMEM lane quadrant) run the LayerNorm + residual epilogue.
    const int quad = warp & 3;
    const int wgp = (warp - 3) >> 2;                     // warpgroup 0..2
    const int part = wgp;                                // which third of the columns this warp owns in the epilogue
    const uint32_t tacc = tmem + (uint32_t(quad * 32) << 16);
    const int row = quad * 32 + lane;
    const Geo geo = make_geo(a.Z, a.H, a.W);
    for (int tu = 0; tu < my_units; ++tu) {
      {
        for (int c = wgp; c < NCH; c += 3) {
          const int hb = c & 1;                          // NCH is even: buffer index and use count follow from the chunk number
          const int cgx = tu * NCH + c;
          const uint32_t haddr = tacc + COL_H + 64 * hb;
          uint8_t* hrow = hs + hb * 16384 + row * 128;
          warp_wait(&hfull[c % 6], (cgx / 6) & 1);
          if (lane == 0 && quad == 3) TR(4 + hb, tu * NCH + c, 0);
          tc_fence_after();
          uint32_t r[2][32];
          tmem_ld32(haddr, r[0]);
          tmem_ld32(haddr + 32, r[1]);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {        // one arrival per warp: all its lanes hold their Hacc values in registers
            if (leader) mbar_arrive(&hempty[hb]); else mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&hempty[hb]), 0));
          }
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
              if (!(dbg & 8)) gelu_erf8(v);
#pragma unroll
              for (int q = 0; q < 4; ++q) pk[hh * 16 + 4 * j8 + q] = pack16<kFp16>(v[2 * q], v[2 * q + 1]);
            }
          }
          if (lane == 0 && quad == 3) TR(4 + hb, tu * NCH + c, 1);
          if (cgx >= 2) warp_wait(&sempty[(cgx - 2) % 6], ((cgx - 2) / 6) & 1);     // GEMM2 of chunk c - 2 has read the buffer
          if (lane == 0 && quad == 3) TR(4 + hb, tu * NCH + c, 2);
          if (c < 2 && tu > 0) warp_wait(lnfree, (tu - 1) & 1);      // ... and nobody stages the previous tile's epilogue in it
#pragma unroll
          for (int q = 0; q < 8; ++q)                  // 8 x 16 B = this row's 64 hidden units, XOR-swizzled 16 B chunks
            *reinterpret_cast<uint4*>(hrow + ((q ^ (row & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {          // one arrival per warp: its 32 rows of H are written and visible to the async proxy
            if (leader) mbar_arrive(&sfull[hb]); else mbar_arrive_remote(mapa_u32(smem_u32(&sfull[hb]), 0));
            if (quad == 3) TR(4 + hb, tu * NCH + c, 3);
          }
        }
      }
      // ------------------------------ LayerNorm + residual epilogue of tile tu ------------------------------
      // Three warps per TMEM lane quadrant (thread = row), each owns 128 of the 384 columns.  Statistics: partial sums meet in
      // shared memory.  Then the warp walks its columns in eight 16-column pieces: the [32 x 16] fp32 residual piece comes by
      // TMA (SWIZZLE_64B, two buffers per warp, fetched by the warp's own lane 0), accumulator -> registers -> normalise -> add
      // in place (row per lane, conflict-free) -> TMA store; the 16-bit shadow goes through a swizzled [32 x 32 B] tile so
      // that its global stores are 32 B row segments (2 lanes per row, window scatter).  All of this is staged in the W2 ring
      // and the H buffers, which are idle exactly from "Y complete" to "Y drained" -- NOT in the X slabs: the next tile's X
      // load and its first two GEMM1 chunks run under this epilogue (staging in the X slabs cost 5 600 clk per tile for the X
      // tile alone).
      const int tile = 2 * (pair + tu * num_pairs) + cta_rank;
      const int row0 = tile * 128 + quad * 32;
      const int g = row0 + lane;
      const int my_tok = g < a.T ? g : -1;
      const int my_dst = (my_tok >= 0 && a.roll_out >= 0) ? token_to_win_row(geo, my_tok, a.roll_out) : my_tok;
      int dsts[2];         // shadow rows of this lane's phase-B items: row rr = it * 16 + (lane >> 1)
#pragma unroll
      for (int it = 0; it < 2; ++it) dsts[it] = __shfl_sync(0xffffffffu, my_dst, it * 16 + (lane >> 1));
      const int wslot = warp - 3;
      uint8_t* stg = r2 + wslot * 6144;            // 2 x 2 KB residual pieces + 2 KB shadow tile
      uint64_t* rf = rfb + 2 * wslot;
      auto fetch = [&](int j) {                     // lane 0
        mbar_arrive_expect_tx(&rf[j & 1], 2048);
        tma_load_2d(&tmR, &rf[j & 1], stg + (j & 1) * 2048, 128 * part + 16 * j, row0);      // rows >= T read as zero
      };
      warp_wait(yfull, tu & 1);
      if (lane == 0 && warp == 6) TR(6, tu * 8, 0);
      tc_fence_after();
      if (lane == 0 && !(dbg & 4)) { fetch(0); fetch(1); }
      float mean = 0.f, rstd = 0.f;
      if (!(dbg & 4)) {
        uint32_t r0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r0) : "r"(tacc) : "memory");
        tmem_ld_wait();
        const float shift = __uint_as_float(r0) + s_b2[0];       // any value near the row mean: keeps the one-pass variance exact enough
        const f32x2 nshift = pack2(-shift, -shift);
        f32x2 s1 = pack2(0.f, 0.f), s2 = pack2(0.f, 0.f);
#pragma unroll 1
        for (int c0 = 128 * part; c0 < 128 * part + 128; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tacc + c0, r);
          tmem_ld_wait();
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
        s_stat[part * 128 + row] = make_float2(s1a + s1b, s2a + s2b);
        named_bar_sync(1 + quad, 96);                // the quadrant's three warps
        const float2 p0 = s_stat[row], p1 = s_stat[128 + row], p2 = s_stat[256 + row];
        const float inv_n = 1.0f / float(C);
        const float m = (p0.x + p1.x + p2.x) * inv_n;
        const float var = fmaxf((p0.y + p1.y + p2.y) * inv_n - m * m, 0.f);
        mean = shift + m;
        rstd = rsqrtf(var + a.eps);
      }
      if (lane == 0 && warp == 6) TR(6, tu * 8 + 1, 0);
      const f32x2 ln_a = pack2(rstd, rstd), ln_b = pack2(-mean * rstd, -mean * rstd);
      const f32x2 rs2 = pack2(a.res_scale, a.res_scale);
      uint8_t* shd = stg + 4096;
#pragma unroll 1
      for (int j = 0; j < 8; ++j) {
        const int c0 = 128 * part + 16 * j;
        uint8_t* buf = stg + (j & 1) * 2048;
        if (dbg & 4) {
          if (j == 7 && lane == 0) {
            if (leader) mbar_arrive(yempty); else mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(yempty), 0));
          }
          continue;
        }
        if (lane == 0 && warp == 6 && tu == 0) TR(7, j, 0);
        warp_wait(&rf[j & 1], (tu * 4 + (j >> 1)) & 1);        // each buffer is filled four times per tile
        if (lane == 0 && warp == 6 && tu == 0) TR(7, j, 1);
        uint32_t r[16];
        tmem_ld16(tacc + c0, r);
        tmem_ld_wait();
        if (j == 7) {          // this warp has read the accumulator for the last time: hand Y back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(yempty); else mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(yempty), 0));
          }
        }
        const float4* b4 = reinterpret_cast<const float4*>(s_b2 + c0);
        const float4* g4 = reinterpret_cast<const float4*>(s_gamma + c0);
        const float4* e4 = reinterpret_cast<const float4*>(s_beta + c0);
        uint32_t pk[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 bb = b4[q], gg = g4[q], ee = e4[q];
          uint4* cell = reinterpret_cast<uint4*>(buf + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4));
          const uint4 x = *cell;
          f32x2 v01 = add2(pack2(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1])), pack2(bb.x, bb.y));
          f32x2 v23 = add2(pack2(__uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])), pack2(bb.z, bb.w));
          v01 = fma2(fma2(v01, ln_a, ln_b), pack2(gg.x, gg.y), pack2(ee.x, ee.y));
          v23 = fma2(fma2(v23, ln_a, ln_b), pack2(gg.z, gg.w), pack2(ee.z, ee.w));
          v01 = fma2(rs2, v01, pack2(__uint_as_float(x.x), __uint_as_float(x.y)));
          v23 = fma2(rs2, v23, pack2(__uint_as_float(x.z), __uint_as_float(x.w)));
          float f0, f1, f2, f3;
          unpack2(v01, f0, f1);
          unpack2(v23, f2, f3);
          *cell = make_uint4(__float_as_uint(f0), __float_as_uint(f1), __float_as_uint(f2), __float_as_uint(f3));
          pk[2 * q] = pack16<kFp16>(f0, f1);
          pk[2 * q + 1] = pack16<kFp16>(f2, f3);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h)
          *reinterpret_cast<uint4*>(shd + lane * 32 + ((h ^ ((lane >> 2) & 1)) << 4)) =
              make_uint4(pk[4 * h], pk[4 * h + 1], pk[4 * h + 2], pk[4 * h + 3]);
        if (lane == 0 && warp == 6 && tu == 0) TR(7, j, 2);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {       // rows >= T are clipped by the tensor map
          tma_store_2d(&tmR, buf, c0, row0);
          bulk_commit();
        }
        if (lane == 0 && warp == 6 && tu == 0) TR(7, j, 3);
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          if (dsts[it] < 0) continue;
          const int rr = it * 16 + (lane >> 1), pcc = lane & 1;
          const uint4 v = *reinterpret_cast<const uint4*>(shd + rr * 32 + ((pcc ^ ((rr >> 2) & 1)) << 4));
          stg16(reinterpret_cast<uint16_t*>(a.out16) + size_t(dsts[it]) * C + c0 + pcc * 8, v);
        }
        if (lane == 0 && warp == 6 && tu == 0) TR(7, 8 + j, 0);
        if (lane == 0) {
          bulk_wait_read<0>();               // the store has read the piece: its buffer may be refilled
          if (j + 2 < 8) fetch(j + 2);
        }
        if (lane == 0 && warp == 6 && tu == 0) TR(7, 8 + j, 1);
        __syncwarp();                        // the shadow tile is rewritten by the next piece
      }
      if (lane == 0 && warp == 6) TR(6, tu * 8 + 1, 1);
      if (lane == 0) mbar_arrive(lnfree);           // this warp no longer touches the W2 ring / H buffers (the piece loop ends in a __syncwarp)
    }
  }

  if (warp >= 3 && lane == 0) bulk_wait_all();      // this thread's residual stores have been written
  tc_fence_before();
  cluster_sync_all();          // no multicast commit / remote arrive may target an exited CTA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cta2<512>(tmem);
  }
}

}  // namespace pg
