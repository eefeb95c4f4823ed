// Persistent, warp-specialised tcgen05 GEMM engine for sm_100a.
//
//   C[M, N] = A[M, K] * B[N, K]^T     A, B: 16-bit (bf16 or fp16), K contiguous; fp32 accumulate
//
// * operands arrive by TMA (cp.async.bulk.tensor, SWIZZLE_128B, 64-element K slabs) into a
//   multi-stage shared-memory ring guarded by full/empty mbarriers;
// * one elected thread issues tcgen05.mma (M=128, N<=256 per instruction, K=16) with the
//   accumulator in tensor memory; tcgen05.commit releases ring slots / publishes the tile;
// * eight epilogue warps (two warpgroups, alternating 32-column chunks) read the accumulator
//   with tcgen05.ld (one thread == one output row, so LayerNorm statistics are thread-local),
//   apply the fused epilogue (bias / q-scale / GELU / LayerNorm / residual), stage the chunk
//   in shared memory and write it out with coalesced, row-remapped 16-byte stores.  The row
//   remaps implement window-partition / roll / crop without ever
//   materialising a permuted tensor.
//
// Epilogue variants are selected by a compile-time config struct (see gemm_kernels.cu).
#pragma once
#include "common.cuh"
#include "geometry.cuh"

namespace pg {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 x 16-bit = 128 B = one swizzle row
constexpr int UMMA_K = 16;

enum RowMapKind { RM_IDENT = 0, RM_WIN2TOK = 1, RM_UPSAMPLE = 2 };
enum DstMapKind { DM_IDENT = 0, DM_TOK2WIN = 1 };
enum RecoverKind { RC_NONE = 0, RC_UPPER = 1, RC_SURFACE = 2 };

struct GemmShape {
  int M;             // valid rows of A
  int num_m_blocks;  // ceil(M / 128)
  int num_n_blocks;  // N / BN
  int num_k_blocks;  // K / 64 (A and A2 together)
  int k_split;       // k-blocks taken from A; the rest come from A2 (== num_k_blocks if unused)
};

struct EpiArgs {
  const float* bias;    // [N] or nullptr
  const float* gamma;   // LayerNorm weight [BN] (LN configs)
  const float* beta;    // LayerNorm bias   [BN]
  const float* resid;   // fp32 residual stream, natural token rows
  float* out32;         // fp32 output
  void* out16;          // 16-bit output
  int ld32, ld16;       // row pitches in elements
  int rowmap;           // RowMapKind: A row -> output token
  int dstmap;           // DstMapKind: output token -> row of the 16-bit output
  int row_base;         // RM_IDENT: token = A row + row_base
  int Z, H, W;          // token grid of the row maps
  int roll_in;          // RM_WIN2TOK: roll state of the window-ordered A rows
  int roll_out;         // DM_TOK2WIN: roll state of the window-ordered 16-bit output
  int q_cols;           // columns [0, q_cols) are multiplied by q_scale (QKV: q *= 32^-0.5)
  float q_scale;
  float res_scale;      // DropPath factor on the normalised branch (1 in eval)
  float eps;
  int lat, lon;         // RECOVER: output field extents (721, 1440)
  int plane_rows;       // HEADMAJOR: rows per 32-column plane of the 16-bit output
  int debug;            // development only: bit0 no residual loads, bit1 no stores, bit2 no epilogue math
  long long* trace;     // development only (-DPANGU_ATTN_TRACE): clock64 timeline of CTA `debug >> 8`, [role 8][index 64][event 4]
};

// ---------------------------------------------------------------------------------------
template <class Cfg>
struct GemmTraits {
  static constexpr int BN = Cfg::BN;
  static constexpr int UN = Cfg::UN;                 // N per tcgen05.mma (<= 256, % 16 == 0)
  static constexpr int NUM_B = BN / UN;
  static constexpr int CH = Cfg::CH;                 // epilogue chunk (columns)
  static constexpr int STAGES = Cfg::STAGES;
  static constexpr int CL = Cfg::CLUSTER;            // CTAs per cluster sharing each B tile by TMA multicast
  static constexpr bool NSPLIT = Cfg::NSPLIT;        // cluster pair splits the N (LayerNorm) dimension instead of M
  static constexpr bool CTA2 = Cfg::CTA2;            // CTA pair = ONE tcgen05.mma.cta_group::2 (M = 256): each CTA holds its own 128
                                                     // A rows and HALF of the B tile, so a stage is A + B/2 and the ring gets deeper
  static constexpr int ACC_STAGES = (2 * BN <= 512) ? 2 : 1;
  static constexpr int TMEM_COLS_RAW = ACC_STAGES * BN;
  static constexpr int TMEM_COLS = TMEM_COLS_RAW <= 32 ? 32 : TMEM_COLS_RAW <= 64 ? 64 : TMEM_COLS_RAW <= 128 ? 128
                                   : TMEM_COLS_RAW <= 256 ? 256 : 512;
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = (CTA2 ? BN / 2 : BN) * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_PITCH = CH * 4 + 16;      // bytes; == 16 (mod 128) -> conflict-free 16 B rows
  // per-warp staging slab: [32 rows][STG_PITCH] fp32, or (TMA16) two [32 rows][64 B] SWIZZLE_64B tiles, or
  // (RESTMA) one [32 rows][128 B] SWIZZLE_128B residual / output tile per column chunk the warp owns
  static constexpr int RES_SLOTS = BN / 64;          // chunks of 32 columns per epilogue warp and tile
  static constexpr int SLAB_BYTES = Cfg::RESTMA ? RES_SLOTS * 4096 : Cfg::TMA16 ? 4096 : ((32 * STG_PITCH + 511) / 512) * 512;
  static constexpr int PAR_BYTES = 3 * BN * 4;       // bias / gamma / beta (shared by the 8 epilogue warps)
  // Epilogue warps: 4 per column group (one per TMEM lane quadrant).  The plain 16-bit (TMA16) epilogues are bound by
  // instruction latency, not issue slots (ncu: 2 warps per scheduler at ~0.2 IPC each), so those configs may ask for
  // 3 or 4 column groups; the LayerNorm / row-mapped epilogues keep 2.
  static constexpr int EPI_WARPS = Cfg::EPI_WARPS;
  static constexpr int GROUPS = EPI_WARPS / 4;
  static constexpr int EPI_THREADS = 32 * EPI_WARPS;
  static constexpr int THREADS = 64 + EPI_THREADS;   // warp 0 TMA, warp 1 MMA, then the epilogue warps
  static constexpr int TAB_BYTES = EPI_WARPS * 64 * 4;   // per warp: row -> token, row -> 16-bit destination row
  static constexpr int EPI_BYTES = ((EPI_WARPS * SLAB_BYTES + PAR_BYTES + TAB_BYTES + 1023) / 1024) * 1024;
  static constexpr int BAR_BYTES = 256 + (Cfg::NSPLIT ? 2 * 128 * 16 : 0)    // + LN partial-stat mailboxes
                                   + (Cfg::RESTMA ? 8 * RES_SLOTS * 8 : 0);   // + per-warp residual-landed barriers
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
  static_assert(BN % UN == 0 && UN % 16 == 0 && UN <= 256, "bad N tiling");
  static_assert(CL == 1 || ((CL == 2 || CL == 4) && (BN / CL) % 8 == 0 && BN / CL <= 256), "bad cluster B split");
  static_assert(!NSPLIT || CL == 2, "NSPLIT: a CTA pair");
  static_assert(!CTA2 || (CL == 2 && !NSPLIT && UN % 32 == 0), "CTA2: a CTA pair, N/2 of every MMA % 16 == 0");
  static_assert(!NSPLIT || (CL == 2 && Cfg::LN && NUM_B == 1), "NSPLIT: LayerNorm row split over a CTA pair");
  static_assert(BN % CH == 0 && (CH == 16 || CH == 32), "bad epilogue chunk");
  static_assert(EPI_WARPS % 4 == 0 && (EPI_WARPS == 8 || Cfg::TMA16), "more than 8 epilogue warps: TMA16 epilogue only");
  static_assert(!Cfg::TMA16 || BN % (32 * GROUPS) == 0, "TMA16: every column group takes whole 32-column chunks");
  static_assert(!Cfg::RESTMA || (Cfg::LN && Cfg::RESID && Cfg::OUT32 && Cfg::OUT16 && CH == 32 && BN % 64 == 0 &&
                                 Cfg::RECOVER == 0 && !Cfg::TMA16), "RESTMA: LayerNorm + residual epilogue");
  static_assert(SLAB_BYTES % 1024 == 0 || !Cfg::RESTMA, "SWIZZLE_128B tiles need 1024 B alignment");
  static_assert(!Cfg::TMA16 || (CH == 32 && !Cfg::LN && !Cfg::OUT32 && Cfg::RECOVER == 0), "TMA16: plain 16-bit output");
  static_assert(!CTA2 || !Cfg::RESTMA, "CTA2 is wired for the TMA16 and the generic epilogues");
  static_assert(B_BYTES % 1024 == 0, "B stage must keep 1024 B alignment");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
};

template <class Cfg, bool kFp16>
__global__ void __launch_bounds__(GemmTraits<Cfg>::THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
            const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
            const GemmShape shape, const EpiArgs ep) {
  using T = GemmTraits<Cfg>;
  constexpr int BN = T::BN, UN = T::UN, CH = T::CH;
  constexpr int kEpiThreads = T::EPI_THREADS;
  extern __shared__ uint8_t smem_raw[];
  // 1024 B alignment for SWIZZLE_128B; offset arithmetic (not an int->pointer cast) so that the
  // compiler keeps the shared address space and emits LDS/STS instead of generic LD/ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem;
  uint8_t* wg_area = smem + T::STAGES * T::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wg_area + T::EPI_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + T::STAGES;          // [STAGES]
  uint64_t* tfull_bar = bars + 2 * T::STAGES;      // [ACC_STAGES]
  uint64_t* tempty_bar = tfull_bar + T::ACC_STAGES;  // [ACC_STAGES]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + T::ACC_STAGES);
  // NSPLIT: LayerNorm partial statistics of the peer CTA arrive here (one float4 per row and accumulator stage)
  [[maybe_unused]] uint64_t* xfull_bar = bars + 24;     // [2] one local arrive.expect_tx + 128 x 16 B of st.async from the peer's row owners
  [[maybe_unused]] uint64_t* xempty_bar = bars + 26;    // [2] 256 remote arrivals (peer's readers)
  [[maybe_unused]] float4* xch = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [2][128]
  // RESTMA: residual tile of (epilogue warp, slot) has landed
  [[maybe_unused]] uint64_t* rfull_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(bars) + 256 + (Cfg::NSPLIT ? 2 * 128 * 16 : 0));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Work units: (group of CL consecutive m-blocks, n-block).  The CL CTAs of a cluster take the
  // m-blocks of one group (same n-block, k-blocks in lockstep) and share each B tile by multicast.
  constexpr int CL = T::CL;
  const int cta_rank = CL == 1 ? 0 : int(cluster_ctarank());
  const int unit0 = blockIdx.x / CL, unit_stride = gridDim.x / CL;
  // NSPLIT: both CTAs of the pair take the SAME m-block and one half of the LayerNorm row each (n = rank)
  const int num_units = T::NSPLIT ? shape.num_m_blocks : ((shape.num_m_blocks + CL - 1) / CL) * shape.num_n_blocks;
  auto unit_m = [&](int unit) { return T::NSPLIT ? unit : (unit / shape.num_n_blocks) * CL + cta_rank; };
  auto unit_n = [&](int unit) { return T::NSPLIT ? cta_rank : unit % shape.num_n_blocks; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    if constexpr (Cfg::TMA16 || Cfg::RESTMA) tma_prefetch_desc(&tmOut);
    for (int s = 0; s < T::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], T::CTA2 ? 1 : CL);     // one tcgen05.commit arrival per CTA of the cluster (CTA2: one commit, multicast)
    }
    for (int a = 0; a < T::ACC_STAGES; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], T::CTA2 ? 2 * kEpiThreads : kEpiThreads);   // CTA2: both CTAs' epilogues release CTA 0's accumulator stage
      if constexpr (T::NSPLIT) { mbar_init(&xfull_bar[a], 1); mbar_init(&xempty_bar[a], kEpiThreads); }
    }
    if constexpr (Cfg::RESTMA) {
      for (int i = 0; i < 8 * T::RES_SLOTS; ++i) mbar_init(&rfull_bar[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (T::CTA2) tmem_alloc_cta2<T::TMEM_COLS>(tmem_slot); else tmem_alloc<T::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  if constexpr (CL == 1) __syncthreads(); else cluster_sync_all();   // peer barriers initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();     // persistent grid, every CTA resident: the next kernel's CTAs may take over SMs as ours retire
  pdl_wait();        // the prologue above overlapped the previous kernel's tail; its outputs are needed from here on

#ifdef PANGU_ATTN_TRACE
  const bool tracing = ep.trace != nullptr && int(blockIdx.x) == ((ep.debug >> 8) & 255);
  auto TR = [&](int role, int g, int ev) { if (tracing && g >= 0 && g < 64) ep.trace[(role * 64 + g) * 4 + ev] = clock64(); };
#else
  auto TR = [](int, int, int) {};
#endif
  constexpr int TR0 = 24;      // first k-block / tile traced is number TR0 / TR0 / 6 (steady state)
  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int kbn = 0;
      for (int unit = unit0; unit < num_units; unit += unit_stride) {
        const int m_blk = unit_m(unit), n_blk = unit_n(unit);
        for (int kb = 0; kb < shape.num_k_blocks; ++kb, ++kbn) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          TR(0, kbn - TR0, 0);
          uint8_t* sa = ring + stage * T::STAGE_BYTES;
          uint8_t* sb = sa + T::A_BYTES;
          if constexpr (T::CTA2) {
            // both CTAs' bytes are counted on CTA 0's barrier (armed by CTA 0 alone; a completion that lands before the
            // arming only makes the signed transaction count dip below zero for a moment)
            const uint32_t lbar = mapa_u32(smem_u32(&full_bar[stage]), 0);
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * T::STAGE_BYTES);
            tma_load_2d_cta2(&tmA, lbar, sa, kb * BLOCK_K, m_blk * BLOCK_M, kEvictFirst);
#pragma unroll
            for (int j = 0; j < T::NUM_B; ++j)     // this CTA's half (UN/2 rows) of the B tile of every MMA of the K step
              tma_load_2d_cta2(&tmB, lbar, sb + j * (UN / 2) * 128, kb * BLOCK_K, n_blk * BN + j * UN + cta_rank * (UN / 2), kEvictLast);
            TR(0, kbn - TR0, 1);
            if (++stage == T::STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[stage], T::STAGE_BYTES);
          if constexpr (T::NSPLIT) {
            // the pair shares the A tile: each CTA fetches 64 of its 128 rows and multicasts them;
            // the weight slice (this CTA's half of the output columns) is private
            tma_load_2d_mcast(&tmA, &full_bar[stage], sa + cta_rank * 64 * 128, kb * BLOCK_K, m_blk * BLOCK_M + cta_rank * 64,
                              uint16_t(3), kEvictFirst);
            tma_load_2d_hint(&tmB, &full_bar[stage], sb, kb * BLOCK_K, n_blk * BN, kEvictLast);
          } else {
          if (kb < shape.k_split)
            tma_load_2d(&tmA, &full_bar[stage], sa, kb * BLOCK_K, m_blk * BLOCK_M);
          else
            tma_load_2d(&tmA2, &full_bar[stage], sa, (kb - shape.k_split) * BLOCK_K, m_blk * BLOCK_M);
          if constexpr (CL == 1) {
#pragma unroll
            for (int j = 0; j < T::NUM_B; ++j)
              tma_load_2d_hint(&tmB, &full_bar[stage], sb + j * UN * 128, kb * BLOCK_K, n_blk * BN + j * UN, kEvictLast);
          } else {
            // this CTA fetches its 1/CL slice of the B tile and multicasts it to the whole cluster
            constexpr int SL = BN / CL;
            tma_load_2d_mcast(&tmB, &full_bar[stage], sb + cta_rank * SL * 128, kb * BLOCK_K, n_blk * BN + cta_rank * SL,
                              uint16_t((1 << CL) - 1), kEvictLast);
          }
          }
          TR(0, kbn - TR0, 1);
          if (++stage == T::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // The whole warp runs the loop converged (every lane waits on the barriers) and one elected lane issues: tcgen05.mma /
    // commit take their operands from uniform registers, and in a single-lane divergent branch the issue of one k-block
    // (4 MMAs + commit) took ~400 clk and a wait on an already completed barrier ~250 -- 770 clk per k-block against 384 clk
    // of MMA work (QKV at C = 384: tensor pipe 52 % busy with the ring full and the epilogue idle).
    if (!T::CTA2 || cta_rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(T::CTA2 ? 2 * BLOCK_M : BLOCK_M, UN, kFp16);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int kbn = 0, tn = 0;
      for (int unit = unit0; unit < num_units; unit += unit_stride, ++tn) {
        if (lane == 0) TR(2, tn - TR0 / 6, 0);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        if (lane == 0) TR(2, tn - TR0 / 6, 1);
        tc_fence_after();
        for (int kb = 0; kb < shape.num_k_blocks; ++kb, ++kbn) {
          if (lane == 0) TR(1, kbn - TR0, 0);
          mbar_wait(&full_bar[stage], phase);
          if (lane == 0) TR(1, kbn - TR0, 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + stage * T::STAGE_BYTES);
          const uint64_t da = make_sdesc_sw128(sa);
          if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
#pragma unroll
            for (int j = 0; j < T::NUM_B; ++j) {
              const uint64_t db = make_sdesc_sw128(sa + T::A_BYTES + j * (T::CTA2 ? UN / 2 : UN) * 128);
              // advancing K by 16 elements = 32 B inside the swizzle row: +2 in the (addr>>4) field
              if constexpr (T::CTA2)
                umma_f16_ss_cta2(tmem_base + acc * BN + j * UN, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
              else
              umma_f16_ss(tmem_base + acc * BN + j * UN, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc,
                          (kb | k) != 0 ? 1u : 0u);
            }
          }
          // frees the ring slot (in every CTA of the cluster: the peer multicasts into it) once read
          if constexpr (T::CTA2) umma_commit_cta2(&empty_bar[stage]);
          else if constexpr (CL == 1) umma_commit(&empty_bar[stage]); else umma_commit_mcast(&empty_bar[stage], uint16_t((1 << CL) - 1));
          if (kb == shape.num_k_blocks - 1) {      // accumulator complete -> epilogue (same lane: the commit tracks its MMAs)
            if constexpr (T::CTA2) umma_commit_cta2(&tfull_bar[acc]); else umma_commit(&tfull_bar[acc]);
          }
          TR(1, kbn - TR0, 2);
          }
          __syncwarp();
          if (++stage == T::STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == T::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================================ epilogue ================================
    // Eight autonomous warps.  Warp w may touch TMEM lanes [32*(w%4), +32): it owns those 32 tile rows;
    // the two warps of a lane quadrant take alternate column chunks.  No cross-warp barrier is needed in
    // the steady state: tables, staging slab and output stores are all per warp (__syncwarp only).
    const int quad = warp & 3;                 // TMEM lane quadrant
    const int half = (warp - 2) >> 2;          // column group 0 .. GROUPS-1: which column chunks
    const int wslot = warp - 2;                // 0 .. EPI_WARPS-1
    uint8_t* slab = wg_area + wslot * T::SLAB_BYTES;                       // [32 rows][STG_PITCH] or 2 x 2 KB (TMA16)
    float* s_bias = reinterpret_cast<float*>(wg_area + T::EPI_WARPS * T::SLAB_BYTES); // shared by the epilogue warps
    float* s_gamma = s_bias + BN;
    float* s_beta = s_gamma + BN;
    int* s_tok = reinterpret_cast<int*>(s_beta + BN) + wslot * 64;         // per warp: row -> token
    int* s_dst = s_tok + 32;                                               //           row -> 16-bit dest row
    const Geo geo = make_geo(ep.Z, ep.H, ep.W);

    int acc = 0;
    uint32_t acc_phase = 0;
    int loaded_n_blk = -1;
    [[maybe_unused]] int sbuf = 0;   // TMA16: slab parity, alternates across chunks AND tiles
    [[maybe_unused]] int res_tiles = 0;   // RESTMA: tiles finished by this warp (residual barrier phase)
    for (int unit = unit0; unit < num_units; unit += unit_stride) {
      const int m_blk = unit_m(unit), n_blk = unit_n(unit);
      // ---- epilogue parameters of this n-block (uniform decision across the 8 warps; LN kernels: once)
      if (loaded_n_blk != n_blk) {
        named_bar_sync(1, kEpiThreads);
        for (int c = threadIdx.x - 64; c < BN; c += kEpiThreads) {
          s_bias[c] = ep.bias ? ep.bias[n_blk * BN + c] : 0.f;
          if constexpr (Cfg::LN) {
            const int gc = (T::NSPLIT ? n_blk * BN : 0) + c;     // LN affine is indexed by the row position
            s_gamma[c] = ep.gamma[gc]; s_beta[c] = ep.beta[gc];
          }
        }
        named_bar_sync(1, kEpiThreads);
        loaded_n_blk = n_blk;
      }
      const uint32_t tacc = tmem_base + (uint32_t(quad * 32) << 16) + acc * BN;

      if constexpr (Cfg::TMA16) {
        // ------- 16-bit row-major output, identity row map: registers -> private swizzled [32 x 64 B]
        // slab -> one TMA bulk store per slab (double buffered).  The TMEM load of the next chunk is in
        // flight while the current one is being converted.
        const int tn = (unit - unit0) / unit_stride - TR0 / 6;
        if (lane == 0 && (wslot & 3) == 0) TR(3 + (wslot >> 2), tn, 0);
        mbar_wait(&tfull_bar[acc], acc_phase);
        if (lane == 0 && (wslot & 3) == 0) TR(3 + (wslot >> 2), tn, 1);
        tc_fence_after();
        constexpr int CSTEP = 32 * T::GROUPS;      // column distance between consecutive chunks of one warp
        constexpr int NCH = BN / CSTEP;            // chunks of 32 columns per warp
        uint32_t rb[2][32];
        tmem_ld32(tacc + half * 32, rb[0]);
        tmem_ld_wait();
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          const int c0 = half * 32 + CSTEP * ci;
          uint32_t (&r)[32] = rb[ci & 1];
          if (ci + 1 < NCH) tmem_ld32(tacc + c0 + CSTEP, rb[(ci + 1) & 1]);
          if (lane == 0) bulk_wait_read<1>();      // the store issued two chunks ago has read this slab
          __syncwarp();
          const int ncol0 = n_blk * BN + c0;
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + c0);
          uint8_t* tile = slab + sbuf * 2048;
#pragma unroll
          for (int q = 0; q < 4; ++q) {          // 4 x 16 B (8 columns each) per 64 B row
            float v[8];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float4 bb = b4[2 * q + h];
              v[4 * h + 0] = __uint_as_float(r[8 * q + 4 * h + 0]) + bb.x;
              v[4 * h + 1] = __uint_as_float(r[8 * q + 4 * h + 1]) + bb.y;
              v[4 * h + 2] = __uint_as_float(r[8 * q + 4 * h + 2]) + bb.z;
              v[4 * h + 3] = __uint_as_float(r[8 * q + 4 * h + 3]) + bb.w;
            }
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
              if constexpr (Cfg::SCALEQ) {
                if (ncol0 + 8 * q + e < ep.q_cols) { v[e] *= ep.q_scale; v[e + 1] *= ep.q_scale; }   // q_cols is even
              }
              if constexpr (Cfg::GELU) gelu_erf2(v[e], v[e + 1]);
            }
            uint4 h16;
            h16.x = pack16<kFp16>(v[0], v[1]); h16.y = pack16<kFp16>(v[2], v[3]);
            h16.z = pack16<kFp16>(v[4], v[5]); h16.w = pack16<kFp16>(v[6], v[7]);
            // SWIZZLE_64B: 16 B chunk index XOR bits [7,9) of the byte address (= (slab row >> 1) & 3)
            *reinterpret_cast<uint4*>(tile + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) = h16;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && m_blk < shape.num_m_blocks) {   // an odd m-block count leaves the pair's last tile out of range
            if constexpr (Cfg::HEADMAJOR)     // out16 is [N/32 planes][plane_rows][32]: a 32-column chunk is one plane
              tma_store_2d(&tmOut, tile, 0, (ncol0 >> 5) * ep.plane_rows + m_blk * BLOCK_M + quad * 32);
            else
              tma_store_2d(&tmOut, tile, ncol0, m_blk * BLOCK_M + quad * 32);
          }
          if (lane == 0) bulk_commit();
          sbuf ^= 1;
          if (ci + 1 < NCH) tmem_ld_wait();
        }
        tc_fence_before();
        if (lane == 0 && (wslot & 3) == 0) TR(3 + (wslot >> 2), tn, 2);
        if constexpr (T::CTA2) {          // the issuing CTA (rank 0) waits for both halves of the M = 256 accumulator to be drained
          if (cta_rank == 0) mbar_arrive(&tempty_bar[acc]);
          else mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
        } else
        mbar_arrive(&tempty_bar[acc]);
        if (++acc == T::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        continue;
      }

      if constexpr (Cfg::RESTMA) {
        // ------- LayerNorm + residual with the fp32 residual stream moved by TMA (natural token rows only).
        // The warp owns RES_SLOTS tiles [32 rows x 32 fp32 columns] (SWIZZLE_128B, private smem): the residual
        // of the NEXT tile is fetched into them while the mainloop of that tile runs, the result is formed in
        // place and leaves with one TMA store per tile; no global load is ever waited for in this loop.
        constexpr int NS = T::RES_SLOTS;
        const int row0 = m_blk * BLOCK_M + quad * 32;
        if (ep.debug & 4) {          // development ablation: mainloop only (no epilogue work at all)
          mbar_wait(&tfull_bar[acc], acc_phase);
          tc_fence_after();
          tc_fence_before();
          mbar_arrive(&tempty_bar[acc]);
          if (++acc == T::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
          continue;
        }
        auto fetch_resid = [&](int mb, int nb) {      // lane 0: residual tiles of this warp for tile (mb, nb)
          if (ep.debug & 1) return;                   // development ablation: no residual traffic
#pragma unroll
          for (int sl = 0; sl < NS; ++sl) {
            mbar_arrive_expect_tx(&rfull_bar[wslot * NS + sl], 4096);
            tma_load_2d(&tmOut, &rfull_bar[wslot * NS + sl], slab + sl * 4096, nb * BN + (half + 2 * sl) * 32, mb * BLOCK_M + quad * 32);
          }
        };
        if (res_tiles == 0 && lane == 0) fetch_resid(m_blk, n_blk);      // first tile of this CTA
        // 16-bit destination rows of this warp's 32 tokens
        __syncwarp();
        {
          const int tok = row0 + lane < shape.M ? row0 + lane + ep.row_base : -1;
          s_dst[lane] = (tok >= 0 && ep.dstmap == DM_TOK2WIN) ? token_to_win_row(geo, tok, ep.roll_out) : tok;
        }
        __syncwarp();
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        float mean = 0.f, rstd = 1.f;
        {
          // thread-local LayerNorm statistics over the row (shifted sums, packed fp32x2 math)
          float shift = 0.f;
          f32x2 s1 = pack2(0.f, 0.f), s2 = pack2(0.f, 0.f);
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tacc + c0, r);
            tmem_ld_wait();
            if (c0 == 0) shift = __uint_as_float(r[0]) + s_bias[0];
            const f32x2 nshift = pack2(-shift, -shift);
            const float4* b4 = reinterpret_cast<const float4*>(s_bias + c0);
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
          if constexpr (T::NSPLIT) {
            // exchange (sum, sum of squares, shift) of this half row with the peer CTA through DSMEM
            const float my1 = s1a + s1b, my2 = s2a + s2b;
            const uint32_t peer = uint32_t(cta_rank ^ 1);
            if (half == 0) {
              if (wslot == 0 && lane == 0) mbar_arrive_expect_tx(&xfull_bar[acc], 128 * 16);   // this CTA's inbox: 128 rows x 16 B
              mbar_wait_cluster(&xempty_bar[acc], acc_phase ^ 1);  // peer has consumed my previous message in this slot
              st_async_cluster_f4(mapa_u32(smem_u32(&xch[acc * 128 + quad * 32 + lane]), peer), my1, my2, shift, 0.f,
                                  mapa_u32(smem_u32(&xfull_bar[acc]), peer));
            }
            mbar_wait_cluster(&xfull_bar[acc], acc_phase);
            const float4 o = xch[acc * 128 + quad * 32 + lane];
            mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&xempty_bar[acc]), peer));
            const float n = float(BN), inv_n2 = 1.0f / float(2 * BN);
            const float mu = (my1 + n * shift + o.x + n * o.z) * inv_n2;
            const float d0 = shift - mu, d1 = o.z - mu;
            const float m2 = (my2 + 2.f * d0 * my1 + n * d0 * d0) + (o.y + 2.f * d1 * o.x + n * d1 * d1);
            mean = mu;
            rstd = rsqrtf(fmaxf(m2 * inv_n2, 0.f) + ep.eps);
          } else {
            const float inv_n = 1.0f / float(BN);
            const float m = (s1a + s1b) * inv_n;
            const float var = fmaxf((s2a + s2b) * inv_n - m * m, 0.f);
            mean = shift + m;
            rstd = rsqrtf(var + ep.eps);
          }
        }
        const f32x2 ln_a = pack2(rstd * ep.res_scale, rstd * ep.res_scale), ln_b = pack2(-mean * rstd * ep.res_scale, -mean * rstd * ep.res_scale);
        const f32x2 rs2 = pack2(ep.res_scale, ep.res_scale);
#pragma unroll 1
        for (int sl = 0; sl < NS; ++sl) {
          const int c0 = (half + 2 * sl) * 32;
          uint8_t* tile = slab + sl * 4096;
          uint32_t r[32];
          tmem_ld32(tacc + c0, r);
          if (!(ep.debug & 1)) mbar_wait(&rfull_bar[wslot * NS + sl], res_tiles & 1);
          tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + c0);
          const float4* g4 = reinterpret_cast<const float4*>(s_gamma + c0);
          const float4* e4 = reinterpret_cast<const float4*>(s_beta + c0);
          // phase A (row per lane): x + s * ((acc + b - mean) * rstd * gamma + beta), formed in place in the residual tile
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            uint4* cell = reinterpret_cast<uint4*>(tile + lane * 128 + ((j4 ^ (lane & 7)) << 4));
            const uint4 q = *cell;
            const float4 bb = b4[j4], gg = g4[j4], ee = e4[j4];
            f32x2 v01 = add2(pack2(__uint_as_float(r[4 * j4]), __uint_as_float(r[4 * j4 + 1])), pack2(bb.x, bb.y));
            f32x2 v23 = add2(pack2(__uint_as_float(r[4 * j4 + 2]), __uint_as_float(r[4 * j4 + 3])), pack2(bb.z, bb.w));
            // s * (v * rstd - mean * rstd) * gamma + s * beta + x
            v01 = fma2(fma2(v01, ln_a, ln_b), pack2(gg.x, gg.y), fma2(rs2, pack2(ee.x, ee.y), pack2(__uint_as_float(q.x), __uint_as_float(q.y))));
            v23 = fma2(fma2(v23, ln_a, ln_b), pack2(gg.z, gg.w), fma2(rs2, pack2(ee.z, ee.w), pack2(__uint_as_float(q.z), __uint_as_float(q.w))));
            float v0, v1, v2, v3;
            unpack2(v01, v0, v1);
            unpack2(v23, v2, v3);
            *cell = make_uint4(__float_as_uint(v0), __float_as_uint(v1), __float_as_uint(v2), __float_as_uint(v3));
          }
          if (sl == NS - 1) {        // accumulator drained by this thread
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          // fp32 stream: one TMA store of the [32 x 32] tile (rows beyond M are clipped)
          if (lane == 0 && m_blk < shape.num_m_blocks && !(ep.debug & 2)) {
            tma_store_2d(&tmOut, tile, n_blk * BN + c0, row0);
            bulk_commit();
          }
          // phase B: 16-bit shadow, coalesced 64 B rows (4 lanes per row), optionally scattered to window order
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int id = it * 32 + lane;
            const int rr = id >> 2, pc = id & 3;
            const int dst = s_dst[rr];
            if (dst < 0 || (ep.debug & 2)) continue;
            const uint4 a0 = *reinterpret_cast<const uint4*>(tile + rr * 128 + (((2 * pc) ^ (rr & 7)) << 4));
            const uint4 a1 = *reinterpret_cast<const uint4*>(tile + rr * 128 + (((2 * pc + 1) ^ (rr & 7)) << 4));
            uint4 h;
            h.x = pack16<kFp16>(__uint_as_float(a0.x), __uint_as_float(a0.y));
            h.y = pack16<kFp16>(__uint_as_float(a0.z), __uint_as_float(a0.w));
            h.z = pack16<kFp16>(__uint_as_float(a1.x), __uint_as_float(a1.y));
            h.w = pack16<kFp16>(__uint_as_float(a1.z), __uint_as_float(a1.w));
            stg16(reinterpret_cast<uint16_t*>(ep.out16) + size_t(dst) * ep.ld16 + n_blk * BN + c0 + pc * 8, h);
          }
        }
        if (++acc == T::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        ++res_tiles;
        // residual tiles of this CTA's next work unit: the TMA stores above must have read the slots first
        __syncwarp();
        if (lane == 0 && unit + unit_stride < num_units) {
          bulk_wait_read<0>();
          fetch_resid(unit_m(unit + unit_stride), unit_n(unit + unit_stride));
        }
        continue;
      }

      // ---- per-warp row tables (overlap the mainloop of this tile)
      __syncwarp();
      {
        const int g = m_blk * BLOCK_M + quad * 32 + lane;  // logical A row of this lane
        int tok = -1;
        if (g < shape.M) {
          if (ep.rowmap == RM_IDENT) tok = g + ep.row_base;
          else if (ep.rowmap == RM_WIN2TOK) tok = win_row_to_token(geo, g, ep.roll_in);
          else tok = upsample_row_to_token(geo, g, n_blk);
        }
        int dst = tok;
        if (ep.dstmap == DM_TOK2WIN && tok >= 0) dst = token_to_win_row(geo, tok, ep.roll_out);
        s_tok[lane] = tok;
        s_dst[lane] = dst;
      }
      __syncwarp();

      constexpr int PPR = CH / 4;                // 16 B fp32 pieces per row of a chunk
      // phase-B item (it, lane) -> (row rr of this warp's 32, piece pc): PPR lanes cover one row
      auto load_resid = [&](int cc, uint4 (&dst)[PPR]) {
#pragma unroll
        for (int it = 0; it < PPR; ++it) {
          const int id = it * 32 + lane;
          const int rr = id / PPR, pc = id % PPR;
          const int tok = s_tok[rr];
          dst[it] = make_uint4(0u, 0u, 0u, 0u);
          if (tok >= 0 && !(ep.debug & 1)) dst[it] = ldg16(ep.resid + size_t(tok) * ep.ld32 + (n_blk * BN + cc + pc * 4));
        }
      };
      [[maybe_unused]] uint4 resq[PPR];
      if constexpr (Cfg::RESID) load_resid(half * CH, resq);   // first chunk: latency hidden by the mainloop wait
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (ep.debug & 4) {
        tc_fence_before();
        mbar_arrive(&tempty_bar[acc]);
        if (++acc == T::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        continue;
      }

      float mean = 0.f, rstd = 1.f;
      if constexpr (Cfg::LN) {
        // thread-local LayerNorm statistics over the whole row (shifted sums, packed fp32x2 math)
        float shift = 0.f;
        f32x2 s1 = pack2(0.f, 0.f), s2 = pack2(0.f, 0.f);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tacc + c0, r);
          tmem_ld_wait();
          if (c0 == 0) shift = __uint_as_float(r[0]) + s_bias[0];
          const f32x2 nshift = pack2(-shift, -shift);
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + c0);
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
        if constexpr (T::NSPLIT) {
          // exchange (sum, sum of squares, shift) of this half row with the peer CTA through DSMEM
          const float my1 = s1a + s1b, my2 = s2a + s2b;
          const uint32_t peer = uint32_t(cta_rank ^ 1);
          if (half == 0) {
            if (wslot == 0 && lane == 0) mbar_arrive_expect_tx(&xfull_bar[acc], 128 * 16);   // this CTA's inbox: 128 rows x 16 B
            mbar_wait_cluster(&xempty_bar[acc], acc_phase ^ 1);  // peer has consumed my previous message in this slot
            st_async_cluster_f4(mapa_u32(smem_u32(&xch[acc * 128 + quad * 32 + lane]), peer), my1, my2, shift, 0.f,
                                mapa_u32(smem_u32(&xfull_bar[acc]), peer));
          }
          mbar_wait_cluster(&xfull_bar[acc], acc_phase);
          const float4 o = xch[acc * 128 + quad * 32 + lane];
          mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&xempty_bar[acc]), peer));
          const float n = float(BN), inv_n2 = 1.0f / float(2 * BN);
          const float mu = (my1 + n * shift + o.x + n * o.z) * inv_n2;
          const float d0 = shift - mu, d1 = o.z - mu;
          const float m2 = (my2 + 2.f * d0 * my1 + n * d0 * d0) + (o.y + 2.f * d1 * o.x + n * d1 * d1);
          mean = mu;
          rstd = rsqrtf(fmaxf(m2 * inv_n2, 0.f) + ep.eps);
        } else {
        const float inv_n = 1.0f / float(BN);
        const float m = (s1a + s1b) * inv_n;
        const float var = fmaxf((s2a + s2b) * inv_n - m * m, 0.f);
        mean = shift + m;
        rstd = rsqrtf(var + ep.eps);
        }
      }
      // y = ((acc + bias) - mean) * rstd * gamma + beta  ==  (acc + bias) * a + b  then * gamma + beta
      const f32x2 ln_a = pack2(rstd, rstd), ln_b = pack2(-mean * rstd, -mean * rstd);

#pragma unroll 1
      for (int c0 = half * CH; c0 < BN; c0 += 2 * CH) {
        // ---------------- residual prefetch for the NEXT chunk's phase-B items
        [[maybe_unused]] uint4 resn[PPR];
        if constexpr (Cfg::RESID) {
          if (c0 + 2 * CH < BN) load_resid(c0 + 2 * CH, resn);
        }
        // ---------------- phase A: TMEM -> registers -> math -> slab (row per lane)
        {
          uint32_t r[CH];
          if constexpr (CH == 32) tmem_ld32(tacc + c0, r); else tmem_ld16(tacc + c0, r);
          tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + c0);
          const float4* g4 = reinterpret_cast<const float4*>(s_gamma + c0);
          const float4* e4 = reinterpret_cast<const float4*>(s_beta + c0);
          uint4* dstp = reinterpret_cast<uint4*>(slab + lane * T::STG_PITCH);
#pragma unroll
          for (int j4 = 0; j4 < CH / 4; ++j4) {
            const float4 bb = b4[j4];
            f32x2 v01 = add2(pack2(__uint_as_float(r[4 * j4]), __uint_as_float(r[4 * j4 + 1])), pack2(bb.x, bb.y));
            f32x2 v23 = add2(pack2(__uint_as_float(r[4 * j4 + 2]), __uint_as_float(r[4 * j4 + 3])), pack2(bb.z, bb.w));
            if constexpr (Cfg::LN) {
              const float4 gg = g4[j4], ee = e4[j4];
              v01 = fma2(fma2(v01, ln_a, ln_b), pack2(gg.x, gg.y), pack2(ee.x, ee.y));
              v23 = fma2(fma2(v23, ln_a, ln_b), pack2(gg.z, gg.w), pack2(ee.z, ee.w));
            }
            float v0, v1, v2, v3;
            unpack2(v01, v0, v1);
            unpack2(v23, v2, v3);
            dstp[j4] = make_uint4(__float_as_uint(v0), __float_as_uint(v1), __float_as_uint(v2), __float_as_uint(v3));
          }
        }
        __syncwarp();
        // ---------------- phase B: slab -> global, coalesced (PPR lanes per row), row-remapped
        if constexpr (Cfg::RECOVER != RC_NONE) {
          // lane = token row, one 16 B piece (4 longitudes of one (c, dz, dh) plane) per iteration: the 32 consecutive tokens of a
          // warp lie next to each other along the longitude, so each store instruction writes one contiguous 512 B run of an
          // output plane (8 lanes per row would scatter every instruction over 32 different planes)
#pragma unroll
          for (int it = 0; it < PPR; ++it) {
            const int rr = lane, pc = it;
            const int tok = s_tok[rr];
            if (tok < 0) continue;
            const int grp = (c0 >> 2) + pc;        // (c, dz, dh) for upper, (c, dh) for surface
            const int wt = tok % geo.W, ht = (tok / geo.W) % geo.H, zt = tok / (geo.W * geo.H);
            const int dh = grp & 3;
            const int la = 4 * ht + dh;
            if (la >= ep.lat) continue;
            size_t off;
            if constexpr (Cfg::RECOVER == RC_UPPER) {
              const int dz = (grp >> 2) & 1, c = grp >> 3;
              const int lev = 2 * zt + dz;
              if (lev >= 13) continue;
              off = ((size_t(c) * 13 + lev) * ep.lat + la) * ep.lon + 4 * wt;
            } else {
              const int c = grp >> 2;
              off = (size_t(c) * ep.lat + la) * ep.lon + 4 * wt;
            }
            const uint4 v = *reinterpret_cast<const uint4*>(slab + rr * T::STG_PITCH + pc * 16);
            stg16(ep.out32 + off, v);
          }
        } else if constexpr (Cfg::OUT32) {
#pragma unroll
          for (int it = 0; it < PPR; ++it) {
            const int id = it * 32 + lane;
            const int rr = id / PPR, pc = id % PPR;
            const int tok = s_tok[rr];
            if (tok < 0 || (ep.debug & 2)) continue;
            const int col = (Cfg::GROUPCOL ? 0 : n_blk * BN) + c0 + pc * 4;
            uint4 v = *reinterpret_cast<const uint4*>(slab + rr * T::STG_PITCH + pc * 16);
            float f0 = __uint_as_float(v.x), f1 = __uint_as_float(v.y), f2 = __uint_as_float(v.z),
                  f3 = __uint_as_float(v.w);
            if constexpr (Cfg::RESID) {
              const uint4 q = resq[it];
              f0 = fmaf(ep.res_scale, f0, __uint_as_float(q.x));
              f1 = fmaf(ep.res_scale, f1, __uint_as_float(q.y));
              f2 = fmaf(ep.res_scale, f2, __uint_as_float(q.z));
              f3 = fmaf(ep.res_scale, f3, __uint_as_float(q.w));
            }
            stg16(ep.out32 + size_t(tok) * ep.ld32 + col,
                  make_uint4(__float_as_uint(f0), __float_as_uint(f1), __float_as_uint(f2), __float_as_uint(f3)));
            if constexpr (Cfg::OUT16) {
              uint2 h = make_uint2(pack16<kFp16>(f0, f1), pack16<kFp16>(f2, f3));
              *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(ep.out16) + size_t(s_dst[rr]) * ep.ld16 + col) = h;
            }
          }
        } else {
          // 16-bit only: 8 columns (32 B of the slab) -> one 16 B store
          constexpr int PPR8 = CH / 8;
#pragma unroll
          for (int it = 0; it < PPR8; ++it) {
            const int id = it * 32 + lane;
            const int rr = id / PPR8, pc = id % PPR8;
            const int tok = s_tok[rr];
            if (tok < 0) continue;
            const int col = (Cfg::GROUPCOL ? 0 : n_blk * BN) + c0 + pc * 8;
            const uint4 a = *reinterpret_cast<const uint4*>(slab + rr * T::STG_PITCH + pc * 32);
            const uint4 b = *reinterpret_cast<const uint4*>(slab + rr * T::STG_PITCH + pc * 32 + 16);
            uint4 h;
            h.x = pack16<kFp16>(__uint_as_float(a.x), __uint_as_float(a.y));
            h.y = pack16<kFp16>(__uint_as_float(a.z), __uint_as_float(a.w));
            h.z = pack16<kFp16>(__uint_as_float(b.x), __uint_as_float(b.y));
            h.w = pack16<kFp16>(__uint_as_float(b.z), __uint_as_float(b.w));
            stg16(reinterpret_cast<uint16_t*>(ep.out16) + size_t(s_dst[rr]) * ep.ld16 + col, h);
          }
        }
        __syncwarp();  // slab free again
        if constexpr (Cfg::RESID) {
#pragma unroll
          for (int it = 0; it < PPR; ++it) resq[it] = resn[it];
        }
      }
      // accumulator drained by this thread
      tc_fence_before();
      if constexpr (T::CTA2) {          // the issuing CTA (rank 0) waits for both halves of the M = 256 accumulator
        if (cta_rank == 0) mbar_arrive(&tempty_bar[acc]);
        else mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
      } else
      mbar_arrive(&tempty_bar[acc]);
      if (++acc == T::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }

  if constexpr (Cfg::TMA16 || Cfg::RESTMA) {
    if (warp >= 2 && lane == 0) bulk_wait_all();   // smem must outlive the bulk stores
  }
  tc_fence_before();
  if constexpr (CL == 1) __syncthreads(); else cluster_sync_all();   // no remote arrive / multicast may target an exited CTA
  if (warp == 1) {
    tc_fence_after();
    if constexpr (T::CTA2) tmem_dealloc_cta2<T::TMEM_COLS>(tmem_base); else tmem_dealloc<T::TMEM_COLS>(tmem_base);
  }
}

}  // namespace pg
