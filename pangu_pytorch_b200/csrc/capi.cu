// Host side of the C ABI declared in include/pangu_b200.h: argument validation, TMA
// tensor-map encoding, kernel configuration and launches.  No allocation, no syncs.
#include "../../include/pangu_b200.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "attention.cuh"
#include "attention_tc.cuh"
#include "attention_bwd.cuh"
#include "backward.cuh"
#include "elementwise.cuh"
#include "gemm.cuh"
#include "mlp_fused.cuh"
#include "mlp_fused2.cuh"
#include "wgrad.cuh"

using namespace pg;

// ---------------------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define PG_REQUIRE(cond, ...) do { if (!(cond)) return fail(-1, __VA_ARGS__); } while (0)
#define PG_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) \
  return fail(-2, "%s failed: %s", #expr, cudaGetErrorString(e__)); } while (0)
#define PG_TRY(expr) do { int r__ = (expr); if (r__ != 0) return r__; } while (0)

extern "C" const char* pangu_last_error(void) { return g_err; }
extern "C" int pangu_version(void) { return 100; }

// ---------------------------------------------------------------------------------------
// device / driver entry points
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
// Per-DEVICE state (one process may drive several GPUs): SM count, compute capability and the "attribute already set" flags of
// every kernel instantiation are indexed by the CUDA device that is current when the entry point is called.
constexpr int kMaxDevices = 64;
static int g_sms[kMaxDevices] = {0};
static int g_ccs[kMaxDevices] = {0};
static std::mutex g_dev_mutex;
static inline int cur_dev() { int d = 0; cudaGetDevice(&d); return (d >= 0 && d < kMaxDevices) ? d : 0; }
#define g_num_sms (g_sms[cur_dev()])
static const bool g_pdl = getenv("PANGU_B200_PDL") != nullptr;     // programmatic dependent launch: measured +-0 (17.62 vs 17.56 ms), off by default
static std::once_flag g_once;
static int g_init_status = 0;

static void init_once() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
    g_init_status = -4;
    return;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
}

static int ensure_init() {
  std::call_once(g_once, init_once);
  if (g_init_status != 0) return fail(g_init_status, "pangu_b200: CUDA device / driver entry point unavailable");
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return fail(-3, "pangu_b200: no current CUDA device");
  if (g_ccs[dev] == 0) {
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    cudaDeviceGetAttribute(&g_sms[dev], cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&g_ccs[dev], cudaDevAttrComputeCapabilityMajor, dev);
  }
  if (g_ccs[dev] != 10) return fail(-5, "pangu_b200: requires an sm_100 (B200) device, found sm_%d* on device %d", g_ccs[dev], dev);
  return 0;
}

extern "C" int pangu_check_device(void) { return ensure_init(); }

// 2-D K-major operand map: dim0 = K (contiguous), dim1 = rows; box = 64 x box_rows, 128 B swizzle,
// out-of-bounds rows read as zero (this is what pads ragged M tails).
// 2-D map of a row-major 16-bit output [rows, cols]: box = 32 columns x 32 rows (one warp's slab), 64 B swizzle
static int make_out_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems) {
  PG_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (pitch_elems * 2) % 16 == 0 && cols % 32 == 0,
             "output not TMA-storable (base %p, pitch %llu, cols %llu)", base, (unsigned long long)pitch_elems,
             (unsigned long long)cols);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(-6, "cuTensorMapEncodeTiled(out) failed (%d)", int(r));
  return 0;
}

static int make_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t k, uint64_t pitch_elems,
                    uint32_t box_rows) {
  PG_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "operand base %p not 16 B aligned", base);
  PG_REQUIRE((pitch_elems * 2) % 16 == 0, "operand pitch %llu not a multiple of 16 B", (unsigned long long)pitch_elems * 2);
  PG_REQUIRE(box_rows <= 256 && k % 64 == 0, "bad box (%u rows, K=%llu)", box_rows, (unsigned long long)k);
  cuuint64_t dims[2] = {k, rows};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(-6, "cuTensorMapEncodeTiled failed (%d)", int(r));
  return 0;
}

// ---------------------------------------------------------------------------------------
// GEMM configurations (epilogue variants)
// ---------------------------------------------------------------------------------------
#ifndef PANGU_EPI_WARPS_MLP1
#define PANGU_EPI_WARPS_MLP1 16     // BN 256: 4 column groups x 2 chunks
#endif
// CTAs per cluster sharing each weight (B) tile by TMA multicast.  The GEMMs are bound by L2 -> SM throughput
// (~6300 B/clk chip-wide = 42.5 B/clk per SM, B300_MICROARCH.md "LTS throughput cap"): a 128 x N tile fed from L2 reaches
// at most (16384 N) / (16 KB + N * 128 B / CL) FLOP per L2 byte, and tcgen05 needs 193 for full rate.
#ifndef PANGU_CLUSTER_M
#define PANGU_CLUSTER_M 2     // measured: 4 is 4-8 % slower (37 four-SM clusters do not all fit the GPCs; B sharing was not the limiter)
#endif
#ifndef PANGU_EPI_WARPS_QKV
#define PANGU_EPI_WARPS_QKV 8       // BN 192: measured no gain from 12 warps (HBM / MMA bound)
#endif
struct CfgBase {
  static constexpr bool LN = false, GELU = false, SCALEQ = false, RESID = false, OUT32 = false, OUT16 = false,
                        GROUPCOL = false;
  static constexpr int RECOVER = RC_NONE;
  static constexpr int CH = 32;
  static constexpr bool TMA16 = false;   // 16-bit row-major output written with TMA bulk stores
  static constexpr int CLUSTER = 1;      // 2: CTA pairs share every weight (B) tile by TMA multicast
  static constexpr bool NSPLIT = false;  // CTA pair splits the LayerNorm row (N) instead of M; stats via DSMEM
  static constexpr bool CTA2 = false;    // CTA pair runs ONE tcgen05.mma.cta_group::2 (M = 256), B tile split between the two SMs
  static constexpr bool HEADMAJOR = false;  // TMA16: output stored as [N/32 planes][plane_rows][32]
  static constexpr bool RESTMA = false;  // LN + residual epilogue whose fp32 stream moves by TMA (identity row map)
  static constexpr int EPI_WARPS = 8;    // 4 x column groups (TMA16 configs may use 12 / 16: latency-bound epilogues)
};
#ifndef PANGU_CTA2
#define PANGU_CTA2 1            // 1: CfgQKV / CfgMLP1 / CfgLin16 use tcgen05.mma.cta_group::2 (see profiles/r02_cta2.md); 0: cta_group::1 + multicast
#endif
struct CfgQKV : CfgBase {      // linear1 of attention: bias, q-scale, 16-bit out
  static constexpr int BN = 192, UN = 192, STAGES = PANGU_CTA2 ? 6 : 4;
  static constexpr bool SCALEQ = true, OUT16 = true, TMA16 = true, HEADMAJOR = true;
  static constexpr int CLUSTER = 2;
  static constexpr bool CTA2 = PANGU_CTA2 != 0;
  static constexpr int EPI_WARPS = PANGU_EPI_WARPS_QKV;
};
struct CfgMLP1 : CfgBase {     // Mlp.linear1: bias + exact GELU, 16-bit out
  static constexpr int BN = 256, UN = 256, STAGES = PANGU_CTA2 ? 4 : 3;
  static constexpr bool GELU = true, OUT16 = true, TMA16 = true;
  static constexpr int CLUSTER = 2;
  static constexpr bool CTA2 = PANGU_CTA2 != 0;
  static constexpr int EPI_WARPS = PANGU_EPI_WARPS_MLP1;
};
struct CfgLNRes192 : CfgBase { // bias + LayerNorm(192) + residual, fp32 + 16-bit out
  static constexpr int BN = 192, UN = 192, STAGES = 3;
  static constexpr bool LN = true, RESID = true, OUT32 = true, OUT16 = true, RESTMA = true;
  static constexpr int CLUSTER = PANGU_CLUSTER_M;
};
struct CfgLNRes384 : CfgBase { // bias + LayerNorm(384) + residual: a CTA pair, 192 columns each, stats over DSMEM
  static constexpr int BN = 192, UN = 192, STAGES = 3;
  static constexpr bool LN = true, RESID = true, OUT32 = true, OUT16 = true, RESTMA = true;
  static constexpr int CLUSTER = 2;
  static constexpr bool NSPLIT = true;
};
struct CfgLNRes384F : CfgBase { // bias + LayerNorm(384) + residual with the FULL 384-wide row per CTA: a CTA pair runs M = 256 x N = 2 x 192
                                // cta_group::2 MMAs, each CTA streams its 128 A rows and half of every weight tile (157 FLOP per L2
                                // byte instead of 96), LayerNorm statistics are thread-local (no DSMEM exchange); one accumulator stage
  static constexpr int BN = 384, UN = 192, STAGES = 4;
  static constexpr bool LN = true, RESID = true, OUT32 = true, OUT16 = true;
  static constexpr int CLUSTER = 2;
  static constexpr bool CTA2 = true;
};
struct CfgPlain192 : CfgBase { // (bias) -> fp32 + 16-bit (embed, downsample.linear, upsample.linear2, tests)
  static constexpr int BN = 192, UN = 192, STAGES = 4;
  static constexpr bool OUT32 = true, OUT16 = true;
};
struct CfgUpLN : CfgBase {     // upsample.linear1: pixel shuffle + crop + LayerNorm(192) -> 16-bit
  static constexpr int BN = 192, UN = 192, STAGES = 4;
  static constexpr bool LN = true, OUT16 = true, GROUPCOL = true;
};
struct CfgRecU : CfgBase {     // _output_layer.conv: bias + un-patchify + crop (upper-air)
  static constexpr int BN = 160, UN = 160, STAGES = 4;
  static constexpr int RECOVER = RC_UPPER;
};
struct CfgRecS : CfgBase {     // _output_layer.conv_surface
  static constexpr int BN = 64, UN = 64, STAGES = 4;
  static constexpr int RECOVER = RC_SURFACE;
};

struct CfgLin16 : CfgBase {    // (bias) -> 16-bit row-major (pre-activation recompute, d hidden)
  static constexpr int BN = 256, UN = 256, STAGES = PANGU_CTA2 ? 4 : 3;
  static constexpr bool OUT16 = true, TMA16 = true;
  static constexpr int CLUSTER = 2;
  static constexpr bool CTA2 = PANGU_CTA2 != 0;
  static constexpr int EPI_WARPS = PANGU_EPI_WARPS_MLP1;
};
struct CfgAcc192 : CfgBase {   // fp32 out = residual + acc (dgrad accumulating into the gradient stream), row maps
  static constexpr int BN = 192, UN = 192, STAGES = PANGU_CTA2 ? 6 : 4;
  static constexpr bool RESID = true, OUT32 = true;
  static constexpr int CLUSTER = 2;
  static constexpr bool CTA2 = PANGU_CTA2 != 0;
};
struct CfgOut16Map : CfgBase { // 16-bit out, destination row map (dgrad scattered into window order)
  static constexpr int BN = 192, UN = 192, STAGES = PANGU_CTA2 ? 6 : 4;
  static constexpr bool OUT16 = true;
  static constexpr int CLUSTER = 2;
  static constexpr bool CTA2 = PANGU_CTA2 != 0;
};

struct GemmOperands {
  const void* a; uint64_t a_pitch;     // rows = M
  const void* a2; uint64_t a2_pitch;   // optional second K range
  int k1, k2;                          // K taken from a / a2 (multiples of 64)
  const void* b; uint64_t b_pitch;     // [N, K]
  int M, N;
};

static inline int sh_rows_padded(int M) { return (M + BLOCK_M - 1) / BLOCK_M * BLOCK_M; }

template <class Cfg, bool kFp16>
static int launch_gemm_t(const GemmOperands& o, const EpiArgs& ep, cudaStream_t stream) {
  using T = GemmTraits<Cfg>;
  PG_REQUIRE(o.N % Cfg::BN == 0, "N=%d not a multiple of the %d-column tile", o.N, Cfg::BN);
  PG_REQUIRE(!Cfg::NSPLIT || (o.N == 2 * Cfg::BN && o.k2 == 0), "NSPLIT GEMM needs N == 2 * BN");
  PG_REQUIRE(o.k1 % 64 == 0 && o.k2 % 64 == 0 && o.k1 > 0, "K (%d + %d) must be multiples of 64", o.k1, o.k2);
  PG_REQUIRE(o.M > 0, "empty M");
  CUtensorMap ma, ma2, mb;
  PG_TRY(make_map(&ma, o.a, o.M, o.k1, o.a_pitch, Cfg::NSPLIT ? 64 : BLOCK_M));
  if (o.k2 > 0) PG_TRY(make_map(&ma2, o.a2, o.M, o.k2, o.a2_pitch, BLOCK_M)); else ma2 = ma;
  PG_TRY(make_map(&mb, o.b, o.N, o.k1 + o.k2, o.b_pitch,
                  Cfg::CTA2 ? Cfg::UN / 2 : (Cfg::CLUSTER > 1 && !Cfg::NSPLIT) ? Cfg::BN / Cfg::CLUSTER : Cfg::UN));
  CUtensorMap mo = ma;
  if constexpr (Cfg::RESTMA) {
    // residual stream == fp32 output, [M, ld32] fp32 in natural row order: 32-column x 32-row SWIZZLE_128B tiles
    PG_REQUIRE(ep.rowmap == RM_IDENT && ep.row_base == 0 && ep.resid == ep.out32 && ep.out32 != nullptr && ep.out16 != nullptr,
               "TMA residual epilogue needs an identity row map and an in-place fp32 stream");
    PG_REQUIRE((reinterpret_cast<uintptr_t>(ep.out32) & 15) == 0 && (size_t(ep.ld32) * 4) % 16 == 0, "fp32 stream not TMA-addressable");
    cuuint64_t dims[2] = {cuuint64_t(ep.ld32), cuuint64_t(o.M)};
    cuuint64_t strides[1] = {cuuint64_t(ep.ld32) * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(&mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ep.out32, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-6, "cuTensorMapEncodeTiled(residual) failed (%d)", int(r));
  }
  if constexpr (Cfg::TMA16) {
    PG_REQUIRE(ep.rowmap == RM_IDENT && ep.dstmap == DM_IDENT && ep.row_base == 0 && ep.out16 != nullptr,
               "TMA-store epilogue needs an identity row map");
    if constexpr (Cfg::HEADMAJOR) {
      PG_REQUIRE(ep.plane_rows >= sh_rows_padded(o.M), "head-major output planes too short");
      PG_TRY(make_out_map(&mo, ep.out16, uint64_t(o.N / 32) * ep.plane_rows, 32, 32));
    } else {
      PG_TRY(make_out_map(&mo, ep.out16, o.M, o.N, ep.ld16));
    }
  }
  GemmShape sh;
  sh.M = o.M;
  sh.num_m_blocks = (o.M + BLOCK_M - 1) / BLOCK_M;
  sh.num_n_blocks = o.N / Cfg::BN;
  sh.num_k_blocks = (o.k1 + o.k2) / 64;
  sh.k_split = o.k1 / 64;
  auto kern = gemm_kernel<Cfg, kFp16>;
  static bool attr_done[kMaxDevices] = {false};   // per instantiation and device (cudaFuncSetAttribute is per device)
  if (!attr_done[cur_dev()]) {
    PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES));
    attr_done[cur_dev()] = true;
  }
  constexpr int CL = Cfg::CLUSTER;
  const int units = Cfg::NSPLIT ? sh.num_m_blocks : ((sh.num_m_blocks + CL - 1) / CL) * sh.num_n_blocks;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(T::THREADS);
  cfg.dynamicSmemBytes = T::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // prologue overlaps the previous kernel's tail
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 2 : 1;
  // persistent grid: as many clusters as can be co-resident (GPC boundaries may leave a few SMs out for CL = 4)
  static int max_clusters_dev[kMaxDevices] = {0};     // per instantiation and device
  int& max_clusters = max_clusters_dev[cur_dev()];
  if (max_clusters == 0) {
    cfg.gridDim = dim3((g_num_sms / CL) * CL);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = g_num_sms / CL; }
    max_clusters = n < g_num_sms / CL ? n : g_num_sms / CL;
  }
  const int grid = (units < max_clusters ? units : max_clusters) * CL;
  cfg.gridDim = dim3(grid);
  PG_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, ma2, mb, mo, sh, ep));
  PG_CUDA(cudaGetLastError());
  return 0;
}
template <class Cfg>
static int launch_gemm(const GemmOperands& o, const EpiArgs& ep, int fp16, cudaStream_t s) {
  return fp16 ? launch_gemm_t<Cfg, true>(o, ep, s) : launch_gemm_t<Cfg, false>(o, ep, s);
}

template <int C, bool kFp16>
static int launch_mlp_fused_t(const void* x16_in, const void* w1_16, const void* w2_16, const MlpArgs& a, cudaStream_t stream) {
  using T = MlpTraits<C>;
  CUtensorMap mx, m1, m2;
  PG_TRY(make_map(&mx, x16_in, a.T, C, C, 128));
  PG_TRY(make_map(&m1, w1_16, 4 * C, C, C, T::MCAST ? 32 : 64));     // W1 [4C, C]: a chunk's 64 hidden rows (half of them with multicast)
  PG_TRY(make_map(&m2, w2_16, C, 4 * C, 4 * C, T::MCAST ? 96 : 192));   // W2 [C, 4C]: 192 output rows (96 with multicast)
  CUtensorMap mr = mx;
  if (T::RES_TMA) {     // fp32 residual stream [T, C]: 32-column x 32-row SWIZZLE_128B tiles (loaded, updated in place, stored)
    PG_REQUIRE((reinterpret_cast<uintptr_t>(a.x32) & 15) == 0, "mlp: residual stream not 16 B aligned");
    cuuint64_t dims[2] = {cuuint64_t(C), cuuint64_t(a.T)};
    cuuint64_t strides[1] = {cuuint64_t(C) * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(&mr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a.x32, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-6, "cuTensorMapEncodeTiled(mlp residual) failed (%d)", int(r));
  }
  auto kern = mlp_fused_kernel<C, kFp16>;
  static bool attr_done[kMaxDevices] = {false};
  if (!attr_done[cur_dev()]) {
    PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES));
    attr_done[cur_dev()] = true;
  }
  const int units = (a.num_tiles + 1) / 2;
  const int max_pairs = g_num_sms / 2;
  const int pairs = units < max_pairs ? units : max_pairs;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(T::THREADS);
  cfg.dynamicSmemBytes = T::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PG_CUDA(cudaLaunchKernelEx(&cfg, kern, mx, m1, m2, mr, a));
  PG_CUDA(cudaGetLastError());
  return 0;
}

// C = 384: CTA pairs on tcgen05.mma.cta_group::2 (csrc/mlp_fused2.cuh)
template <bool kFp16>
static int launch_mlp_fused2_t(const void* x16_in, const void* w1_16, const void* w2_16, const MlpArgs& a, cudaStream_t stream) {
  using T = Mlp2Traits;
  constexpr int C = T::C;
  CUtensorMap mx, m1, m2;
  PG_TRY(make_map(&mx, x16_in, a.T, C, C, 128));
  PG_TRY(make_map(&m1, w1_16, 4 * C, C, C, 32));          // W1 [4C, C]: this CTA's 32 of a chunk's 64 hidden rows
  PG_TRY(make_map(&m2, w2_16, C, 4 * C, 4 * C, 96));      // W2 [C, 4C]: this CTA's 96 of the 192 output rows of an N half
  PG_REQUIRE((reinterpret_cast<uintptr_t>(a.x32) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.out16) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(a.b2) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.gamma) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(a.beta) & 15) == 0, "mlp: residual stream / LayerNorm parameters not 16 B aligned");
  CUtensorMap mr;       // fp32 residual stream [T, C]: 16-column x 32-row SWIZZLE_64B pieces (loaded, updated in place, stored)
  {
    cuuint64_t dims[2] = {cuuint64_t(C), cuuint64_t(a.T)};
    cuuint64_t strides[1] = {cuuint64_t(C) * 4};
    cuuint32_t box[2] = {16, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(&mr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a.x32, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-6, "cuTensorMapEncodeTiled(mlp residual) failed (%d)", int(r));
  }
  auto kern = mlp_fused2_kernel<kFp16>;
  static bool attr_done[kMaxDevices] = {false};
  if (!attr_done[cur_dev()]) {
    PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES));
    attr_done[cur_dev()] = true;
  }
  const int units = (a.num_tiles + 1) / 2;
  int max_pairs = g_num_sms / 2;
#ifdef PANGU_DEV_SWITCHES     // per-SM or chip-wide limit?  run the same per-pair schedule on fewer SMs
  if (const char* e = getenv("PANGU_B200_MLP_PAIRS")) max_pairs = atoi(e);
#endif
  const int pairs = units < max_pairs ? units : max_pairs;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(T::THREADS);
  cfg.dynamicSmemBytes = T::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PG_CUDA(cudaLaunchKernelEx(&cfg, kern, mx, m1, m2, mr, a));
  PG_CUDA(cudaGetLastError());
  return 0;
}

static EpiArgs epi_defaults() {
  EpiArgs e;
  memset(&e, 0, sizeof(e));
  e.Z = 8; e.H = 1; e.W = 12;
  e.res_scale = 1.f; e.q_scale = 1.f; e.eps = 1e-5f;
  e.rowmap = RM_IDENT; e.dstmap = DM_IDENT;
#ifdef PANGU_DEV_SWITCHES     // development builds only (-DPANGU_DEV_SWITCHES): timing ablations, results invalid
  if (const char* d = getenv("PANGU_B200_GEMM_DEBUG")) e.debug = atoi(d);
#endif
#ifdef PANGU_ATTN_TRACE
  if (const char* t = getenv("PANGU_B200_GEMM_TRACE")) e.trace = reinterpret_cast<long long*>(strtoull(t, nullptr, 0));
#endif
  return e;
}

static int check_grid(int Z, int H, int W, int C, int heads) {
  PG_REQUIRE(Z == 8, "Z must be 8 (got %d)", Z);
  PG_REQUIRE((H + 5) % 6 == 0, "H+5 must be a multiple of 6 (got H=%d)", H);
  PG_REQUIRE(W % 12 == 0 && W > 0, "W must be a positive multiple of 12 (got %d)", W);
  PG_REQUIRE(C == 192 || C == 384, "C must be 192 or 384 (got %d)", C);
  PG_REQUIRE(heads == 0 || heads * 32 == C, "heads*32 must equal C");
  return 0;
}

// ---------------------------------------------------------------------------------------
// entry points
// ---------------------------------------------------------------------------------------
extern "C" int pangu_cast16(const float* src, void* dst, int rows, int k_src, int k_dst, int fp16, void* stream) {
  PG_TRY(ensure_init());
  PG_REQUIRE(rows > 0 && k_src > 0 && k_dst >= k_src, "bad cast16 shape");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t n = size_t(rows) * k_dst;
  const int grid = int((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  if (fp16) cast16_kernel<true><<<grid, 256, 0, s>>>(src, static_cast<uint16_t*>(dst), rows, k_src, k_dst);
  else cast16_kernel<false><<<grid, 256, 0, s>>>(src, static_cast<uint16_t*>(dst), rows, k_src, k_dst);
  PG_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pangu_to_window16(const float* x32, void* x16w, int Z, int H, int W, int C, int roll, int fp16,
                                 void* stream) {
  PG_TRY(ensure_init());
  PG_TRY(check_grid(Z, H, W, C, 0));
  const Geo g = make_geo(Z, H, W);
  const int rows = roll < 0 ? Z * H * W : g.nLon * g.types * 144;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int blocks = (rows * 32 + 255) / 256;
  if (fp16) to_window16_kernel<true><<<blocks, 256, 0, s>>>(x32, static_cast<uint16_t*>(x16w), g, C, roll, rows);
  else to_window16_kernel<false><<<blocks, 256, 0, s>>>(x32, static_cast<uint16_t*>(x16w), g, C, roll, rows);
  PG_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pangu_patch_embed(const float* upper, const float* surface, const float* surface_mean,
                                 const float* surface_std, const float* upper_mean, const float* upper_std,
                                 const float* maps, const float* const_h, const void* w_upper16,
                                 const float* b_upper, const void* w_surface16, const float* b_surface,
                                 void* ws_a_upper, void* ws_a_surface, float* x32, void* x16w, int lat, int lon,
                                 int fp16, void* stream) {
  PG_TRY(ensure_init());
  PG_REQUIRE(lon % 48 == 0 && lon > 0, "lon must be a positive multiple of 48 (got %d)", lon);
  const int Hh = (lat + 3) / 4, Ww = lon / 4;
  PG_TRY(check_grid(8, Hh, Ww, 192, 0));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  EmbedArgs ea{upper, surface, surface_mean, surface_std, upper_mean, upper_std, maps, const_h,
               ws_a_upper, ws_a_surface, lat, lon, Hh, Ww, 1.f};
  dim3 grid((Ww + EMB_TT - 1) / EMB_TT, Hh, 8);
  if (fp16) embed_im2col_kernel<true><<<grid, EMB_THREADS, 0, s>>>(ea);
  else embed_im2col_kernel<false><<<grid, EMB_THREADS, 0, s>>>(ea);
  PG_CUDA(cudaGetLastError());
  const int plane = Hh * Ww;
  EpiArgs ep = epi_defaults();
  ep.Z = 8; ep.H = Hh; ep.W = Ww;
  ep.ld32 = 192; ep.ld16 = 192;
  ep.dstmap = DM_TOK2WIN; ep.roll_out = 0;
  // surface plane: tokens [0, plane)
  {
    GemmOperands o{ws_a_surface, 128, nullptr, 0, 128, 0, w_surface16, 128, plane, 192};
    ep.bias = b_surface; ep.out32 = x32; ep.out16 = x16w; ep.row_base = 0;
    PG_TRY(launch_gemm<CfgPlain192>(o, ep, fp16, s));
  }
  // upper-air planes: tokens [plane, 8*plane)
  {
    GemmOperands o{ws_a_upper, 192, nullptr, 0, 192, 0, w_upper16, 192, 7 * plane, 192};
    ep.bias = b_upper; ep.row_base = plane;
    PG_TRY(launch_gemm<CfgPlain192>(o, ep, fp16, s));
  }
  return 0;
}

extern "C" int pangu_qkv(const void* x16w, const void* w16, const float* bias, void* qkv16, int Z, int H, int W,
                         int C, int fp16, void* stream) {
  PG_TRY(ensure_init());
  PG_TRY(check_grid(Z, H, W, C, 0));
  const Geo g = make_geo(Z, H, W);
  const int Tp = g.nLon * g.types * 144;
  GemmOperands o{x16w, uint64_t(C), nullptr, 0, C, 0, w16, uint64_t(C), Tp, 3 * C};
  EpiArgs ep = epi_defaults();
  ep.bias = bias; ep.out16 = qkv16; ep.ld16 = 3 * C;
  ep.plane_rows = sh_rows_padded(Tp);    // head-major: [3*heads planes][plane_rows][32]
  ep.q_cols = C; ep.q_scale = 0.17677669529663687f;   // 32^-0.5, models/layers.py:289
  return launch_gemm<CfgQKV>(o, ep, fp16, static_cast<cudaStream_t>(stream));
}

extern "C" int pangu_window_attention(const void* qkv16, const float* earth_bias, void* att16, int Z, int H, int W,
                                      int C, int heads, int roll, int window_order_out, int fp16, void* stream) {
  PG_TRY(ensure_init());
  PG_TRY(check_grid(Z, H, W, C, heads));
  const Geo g = make_geo(Z, H, W);
  AttnArgs a;
  a.qkv = qkv16; a.bias = earth_bias; a.out = att16;
  a.C = C; a.heads = heads; a.types = g.types; a.nLon = g.nLon; a.nH = g.nH; a.roll = roll ? 1 : 0;
  a.H = H; a.W = W; a.natural = window_order_out ? 0 : 1;
  a.debug = 0;
#ifdef PANGU_DEV_SWITCHES
  if (const char* e = getenv("PANGU_B200_ATTN_DEBUG")) a.debug = atoi(e);
#endif
  a.trace = nullptr;
#ifdef PANGU_ATTN_TRACE       // development builds only (-DPANGU_ATTN_TRACE): device pointer of a clock64 timeline buffer
  if (const char* e = getenv("PANGU_B200_ATTN_TRACE")) a.trace = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
#endif
  const int Tp = g.nLon * g.types * 144;
  a.plane_rows = sh_rows_padded(Tp);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PG_REQUIRE((reinterpret_cast<uintptr_t>(qkv16) & 15) == 0 && (reinterpret_cast<uintptr_t>(earth_bias) & 15) == 0,
             "qkv / bias base not 16 B aligned");
  // head-major qkv: [3*heads planes][plane_rows][32] viewed as a 2-D tensor of 64 B rows; box = one 144-row tile
  CUtensorMap mq, mb;
  {
    cuuint64_t dims[2] = {32, cuuint64_t(3) * heads * a.plane_rows};
    cuuint64_t strides[1] = {64};
    cuuint32_t box[2] = {32, 144};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(&mq, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(qkv16), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-6, "cuTensorMapEncodeTiled(qkv) failed (%d)", int(r));
  }
  // earth_specific_bias [types*heads*144 rows][144] fp32; box = 16 key columns x 144 query rows (9216 B)
  {
    cuuint64_t dims[2] = {144, cuuint64_t(g.types) * heads * 144};
    cuuint64_t strides[1] = {144 * 4};
    cuuint32_t box[2] = {16, 144};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(earth_bias), dims, strides, box,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-6, "cuTensorMapEncodeTiled(bias) failed (%d)", int(r));
  }
  // persistent: one CTA per SM, each takes an equal contiguous share of the (type, head, lon window) units
  const long long units = (long long)g.types * heads * g.nLon;
  const int pgrid = int(units < g_num_sms ? units : g_num_sms);
  static bool tc_attr_dev[kMaxDevices][2] = {{false, false}};
  bool* tc_attr = tc_attr_dev[cur_dev()];
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(pgrid);
  cfg.blockDim = dim3(ATC_THREADS);
  cfg.dynamicSmemBytes = ATC_SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  if (fp16) {
    if (!tc_attr[1]) {
      PG_CUDA(cudaFuncSetAttribute(window_attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM_BYTES));
      tc_attr[1] = true;
    }
    PG_CUDA(cudaLaunchKernelEx(&cfg, window_attention_tc_kernel<true>, mq, mb, a));
  } else {
    if (!tc_attr[0]) {
      PG_CUDA(cudaFuncSetAttribute(window_attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM_BYTES));
      tc_attr[0] = true;
    }
    PG_CUDA(cudaLaunchKernelEx(&cfg, window_attention_tc_kernel<false>, mq, mb, a));
  }
  PG_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pangu_proj_ln_residual(const void* att16, const void* w16, const float* bias, const float* gamma,
                                      const float* beta, float* x32, void* x16, int Z, int H, int W, int C, int roll,
                                      float res_scale, int fp16, void* stream) {
  PG_TRY(ensure_init());
  PG_TRY(check_grid(Z, H, W, C, 0));
  (void)roll;    // att16 arrives in natural token order: the window reverse / un-roll / crop happened in the attention store
  GemmOperands o{att16, uint64_t(C), nullptr, 0, C, 0, w16, uint64_t(C), Z * H * W, C};
  EpiArgs ep = epi_defaults();
  ep.Z = Z; ep.H = H; ep.W = W;
  ep.bias = bias; ep.gamma = gamma; ep.beta = beta;
  ep.resid = x32; ep.out32 = x32; ep.out16 = x16; ep.ld32 = C; ep.ld16 = C;
  ep.rowmap = RM_IDENT; ep.dstmap = DM_IDENT;
  ep.res_scale = res_scale;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#ifdef PANGU_DEV_SWITCHES
  if (const char* e = getenv("PANGU_B200_LN384"))
    if (C == 384 && (atoi(e) & 2)) return launch_gemm<CfgLNRes384F>(o, ep, fp16, s);
#endif
  return C == 192 ? launch_gemm<CfgLNRes192>(o, ep, fp16, s) : launch_gemm<CfgLNRes384>(o, ep, fp16, s);
}

extern "C" int pangu_mlp_ln_residual(const void* x16_in, const void* w1_16, const float* b1, const void* w2_16,
                                     const float* b2, const float* gamma, const float* beta, void* ws_hidden,
                                     float* x32, void* x16_out, int Z, int H, int W, int C, int roll_out,
                                     float res_scale, int fp16, void* stream) {
  PG_TRY(ensure_init());
  PG_TRY(check_grid(Z, H, W, C, 0));
  const int T = Z * H * W;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // One-kernel Mlp (hidden activation kept in TMEM, csrc/mlp_fused.cuh): wins at C = 192 (500 vs 577 us: double-buffered
  // Y accumulator), loses at C = 384 (TMEM holds a single Y).  It is taken when the caller does not ask for the hidden
  // activation (ws_hidden == NULL; the training tape does ask).  PANGU_B200_MLP_FUSED = 0: never, 1: whenever allowed.
  static const int fused_mode = getenv("PANGU_B200_MLP_FUSED") ? atoi(getenv("PANGU_B200_MLP_FUSED")) : -1;
  const bool fused = ws_hidden == nullptr && fused_mode != 0;
  PG_REQUIRE(fused || ws_hidden != nullptr, "mlp_ln_residual: ws_hidden is required on the two-kernel path (C=%d)", C);
  if (fused) {
    // one kernel: the hidden activation stays in tensor memory (ws_hidden is not touched)
    MlpArgs a;
    a.b1 = b1; a.b2 = b2; a.gamma = gamma; a.beta = beta;
    a.x32 = x32; a.out16 = x16_out;
    a.T = T; a.num_tiles = (T + 127) / 128;
    a.Z = Z; a.H = H; a.W = W;
    a.roll_out = roll_out < 0 ? -1 : (roll_out > 0 ? 1 : 0);
    a.res_scale = res_scale; a.eps = 1e-5f;
    a.debug = 0;
    a.trace = nullptr;
#ifdef PANGU_ATTN_TRACE
    if (const char* e = getenv("PANGU_B200_MLP_TRACE")) a.trace = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
#endif
#ifdef PANGU_DEV_SWITCHES
    if (const char* d = getenv("PANGU_B200_GEMM_DEBUG")) a.debug = atoi(d);
#endif
    if (C == 384) return fp16 ? launch_mlp_fused2_t<true>(x16_in, w1_16, w2_16, a, s) : launch_mlp_fused2_t<false>(x16_in, w1_16, w2_16, a, s);
    return fp16 ? launch_mlp_fused_t<192, true>(x16_in, w1_16, w2_16, a, s) : launch_mlp_fused_t<192, false>(x16_in, w1_16, w2_16, a, s);
  }
  {
    GemmOperands o{x16_in, uint64_t(C), nullptr, 0, C, 0, w1_16, uint64_t(C), T, 4 * C};
    EpiArgs ep = epi_defaults();
    ep.bias = b1; ep.out16 = ws_hidden; ep.ld16 = 4 * C;
    PG_TRY(launch_gemm<CfgMLP1>(o, ep, fp16, s));
  }
  {
    GemmOperands o{ws_hidden, uint64_t(4 * C), nullptr, 0, 4 * C, 0, w2_16, uint64_t(4 * C), T, C};
    EpiArgs ep = epi_defaults();
    ep.Z = Z; ep.H = H; ep.W = W;
    ep.bias = b2; ep.gamma = gamma; ep.beta = beta;
    ep.resid = x32; ep.out32 = x32; ep.out16 = x16_out; ep.ld32 = C; ep.ld16 = C;
    ep.rowmap = RM_IDENT;
    ep.dstmap = roll_out < 0 ? DM_IDENT : DM_TOK2WIN;
    ep.roll_out = roll_out > 0 ? 1 : 0;
    ep.res_scale = res_scale;
    // CfgLNRes384F (full 384-wide row per CTA, cta_group::2) was measured at the same speed as the N-split pair for
    // Mlp.linear2 (228 vs 231 us: the lower L2 traffic is paid for with an epilogue that no longer overlaps the mainloop) and
    // slower for the projection (165 vs 120 us), profiles/r02_cta2.md; development builds can select it with PANGU_B200_LN384.
#ifdef PANGU_DEV_SWITCHES
    if (const char* e = getenv("PANGU_B200_LN384"))
      if (C == 384 && (atoi(e) & 1)) return launch_gemm<CfgLNRes384F>(o, ep, fp16, s);
#endif
    PG_TRY(C == 192 ? launch_gemm<CfgLNRes192>(o, ep, fp16, s) : launch_gemm<CfgLNRes384>(o, ep, fp16, s));
  }
  return 0;
}

extern "C" int pangu_downsample(const float* x32_in, const float* gamma, const float* beta, const void* w16,
                                void* ws_a, float* x32_out, void* x16w_out, int Z, int H, int W, int C, int fp16,
                                void* stream) {
  PG_TRY(ensure_init());
  PG_REQUIRE(C == 192 && Z == 8 && W % 24 == 0, "downsample: C must be 192, Z 8, W a multiple of 24");
  const int H2 = (H + 1) / 2, W2 = W / 2;
  PG_TRY(check_grid(Z, H2, W2, 384, 0));
  const int T2 = Z * H2 * W2;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DownArgs da{x32_in, gamma, beta, ws_a, Z, H, W, C, 1e-5f};
  const int blocks = (T2 * 32 + 255) / 256;
  if (fp16) downsample_gather_ln_kernel<true><<<blocks, 256, 0, s>>>(da);
  else downsample_gather_ln_kernel<false><<<blocks, 256, 0, s>>>(da);
  PG_CUDA(cudaGetLastError());
  GemmOperands o{ws_a, 768, nullptr, 0, 768, 0, w16, 768, T2, 384};
  EpiArgs ep = epi_defaults();
  ep.Z = Z; ep.H = H2; ep.W = W2;
  ep.out32 = x32_out; ep.out16 = x16w_out; ep.ld32 = 384; ep.ld16 = 384;
  ep.dstmap = DM_TOK2WIN; ep.roll_out = 0;
  return launch_gemm<CfgPlain192>(o, ep, fp16, s);
}

extern "C" int pangu_upsample(const void* x16_in, const void* w1_16, const float* gamma, const float* beta,
                              const void* w2_16, void* ws_a, float* x32_out, void* x16w_out, int Z, int H, int W,
                              int C_in, int C_out, int fp16, void* stream) {
  PG_TRY(ensure_init());
  PG_REQUIRE(C_in == 384 && C_out == 192 && Z == 8 && W % 24 == 0, "upsample: expects 384 -> 192 on a W%%24 grid");
  PG_TRY(check_grid(Z, H, W, C_out, 0));
  const int H2 = (H + 1) / 2, W2 = W / 2;
  const int T2 = Z * H2 * W2, T = Z * H * W;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  {
    GemmOperands o{x16_in, 384, nullptr, 0, 384, 0, w1_16, 384, T2, 768};
    EpiArgs ep = epi_defaults();
    ep.Z = Z; ep.H = H; ep.W = W;             // HIGH-res grid for the pixel-shuffle map
    ep.gamma = gamma; ep.beta = beta;
    ep.out16 = ws_a; ep.ld16 = 192;
    ep.rowmap = RM_UPSAMPLE;
    PG_TRY(launch_gemm<CfgUpLN>(o, ep, fp16, s));
  }
  {
    GemmOperands o{ws_a, 192, nullptr, 0, 192, 0, w2_16, 192, T, 192};
    EpiArgs ep = epi_defaults();
    ep.Z = Z; ep.H = H; ep.W = W;
    ep.out32 = x32_out; ep.out16 = x16w_out; ep.ld32 = 192; ep.ld16 = 192;
    ep.dstmap = DM_TOK2WIN; ep.roll_out = 0;
    PG_TRY(launch_gemm<CfgPlain192>(o, ep, fp16, s));
  }
  return 0;
}

extern "C" int pangu_patch_recover(const void* skip16, const void* x16, const void* w_upper16, const float* b_upper,
                                   const void* w_surface16, const float* b_surface, float* out_upper,
                                   float* out_surface, int Z, int H, int W, int C, int lat, int lon, int fp16,
                                   void* stream) {
  PG_TRY(ensure_init());
  PG_REQUIRE(C == 192 && Z == 8, "recover: C must be 192 per source (384 concatenated), Z 8");
  PG_REQUIRE(lon == 4 * W && (lat + 3) / 4 == H, "recover: field extents do not match the token grid");
  const int plane = H * W;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const uint16_t* sk = static_cast<const uint16_t*>(skip16);
  const uint16_t* xx = static_cast<const uint16_t*>(x16);
  EpiArgs ep = epi_defaults();
  ep.Z = Z - 1; ep.H = H; ep.W = W; ep.lat = lat; ep.lon = lon;
  {
    GemmOperands o{sk + size_t(plane) * C, uint64_t(C), xx + size_t(plane) * C, uint64_t(C), C, C,
                   w_upper16, uint64_t(2 * C), 7 * plane, 160};
    ep.bias = b_upper; ep.out32 = out_upper;
    PG_TRY(launch_gemm<CfgRecU>(o, ep, fp16, s));
  }
  {
    GemmOperands o{sk, uint64_t(C), xx, uint64_t(C), C, C, w_surface16, uint64_t(2 * C), plane, 64};
    ep.bias = b_surface; ep.out32 = out_surface;
    PG_TRY(launch_gemm<CfgRecS>(o, ep, fp16, s));
  }
  return 0;
}

extern "C" int pangu_denorm_fields(float* upper, float* surface, const float* surface_mean, const float* surface_std,
                                   const float* upper_mean, const float* upper_std, int lat, int lon, void* stream) {
  PG_TRY(ensure_init());
  PG_REQUIRE(lat > 0 && lon > 0 && (size_t(lat) * lon) % 4 == 0, "denorm: lat*lon must be a multiple of 4");
  const int plane4 = int(size_t(lat) * lon / 4);
  dim3 grid((plane4 + 255) / 256 < 64 ? (plane4 + 255) / 256 : 64, 69);
  denorm_fields_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(upper, surface, surface_mean, surface_std,
                                                                            upper_mean, upper_std, plane4);
  PG_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pangu_l1_loss(const float* out_upper, const float* out_surface, const float* tgt_upper,
                             const float* tgt_surface, const float* surface_mean, const float* surface_std,
                             const float* upper_mean, const float* upper_std, const float* upper_weights_host,
                             const float* surface_weights_host, float* loss, double* ws_acc, float* grad_upper,
                             float* grad_surface, int lat, int lon, void* stream) {
  PG_TRY(ensure_init());
  PG_REQUIRE(lat > 0 && lon > 0 && (size_t(lat) * lon) % 4 == 0, "l1_loss: lat*lon must be a multiple of 4");
  PG_REQUIRE(upper_weights_host && surface_weights_host && loss && ws_acc, "l1_loss: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  L1Args a;
  a.out_u = out_upper; a.out_s = out_surface; a.tgt_u = tgt_upper; a.tgt_s = tgt_surface;
  a.s_mean = surface_mean; a.s_std = surface_std; a.u_mean = upper_mean; a.u_std = upper_std;
  a.acc = ws_acc; a.grad_u = grad_upper; a.grad_s = grad_surface;
  for (int i = 0; i < 5; ++i) a.wu[i] = upper_weights_host[i];
  for (int i = 0; i < 4; ++i) a.ws[i] = surface_weights_host[i];
  a.plane4 = int(size_t(lat) * lon / 4);
  const double nu = 65.0 * lat * lon, ns = 4.0 * lat * lon;
  a.inv_nu = float(1.0 / nu); a.inv_ns = float(1.0 / ns);
  PG_CUDA(cudaMemsetAsync(ws_acc, 0, 2 * sizeof(double), s));
  dim3 grid((a.plane4 + 255) / 256 < 32 ? (a.plane4 + 255) / 256 : 32, 69);
  l1_loss_kernel<<<grid, 256, 0, s>>>(a);
  l1_finalize_kernel<<<1, 1, 0, s>>>(ws_acc, loss, 1.0 / nu, 1.0 / ns);
  PG_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pangu_scores(const float* out_upper, const float* out_surface, const float* tgt_upper,
                            const float* tgt_surface, const float* surface_mean, const float* surface_std,
                            const float* upper_mean, const float* upper_std, const float* lat_weights, double* ws_acc,
                            float* rmse, float* acc, int lat, int lon, int normalised, void* stream) {
  PG_TRY(ensure_init());
  PG_REQUIRE(lat > 0 && lon > 0 && lon % 4 == 0, "scores: lon must be a positive multiple of 4");
  PG_REQUIRE(out_upper && out_surface && tgt_upper && tgt_surface && lat_weights && ws_acc && rmse && acc, "scores: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ScoreArgs a{out_upper, out_surface, tgt_upper, tgt_surface, surface_mean, surface_std, upper_mean, upper_std,
              lat_weights, ws_acc, lat, lon, normalised};
  PG_CUDA(cudaMemsetAsync(ws_acc, 0, 69 * 4 * sizeof(double), s));
  dim3 grid(lat < 32 ? lat : 32, 69);
  scores_kernel<<<grid, 256, 0, s>>>(a);
  scores_finalize_kernel<<<1, 96, 0, s>>>(ws_acc, rmse, acc, 1.0 / (double(lat) * lon));
  PG_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pangu_linear(const void* a16, const void* w16, const float* bias, float* out32, void* out16, int M,
                            int N, int K, int gelu, int fp16, void* stream) {
  PG_TRY(ensure_init());
  GemmOperands o{a16, uint64_t(K), nullptr, 0, K, 0, w16, uint64_t(K), M, N};
  EpiArgs ep = epi_defaults();
  ep.bias = bias; ep.out32 = out32; ep.out16 = out16; ep.ld32 = N; ep.ld16 = N;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (gelu) {
    PG_REQUIRE(out16 != nullptr && N % 256 == 0, "pangu_linear(gelu): needs a 16-bit output and N %% 256 == 0");
    return launch_gemm<CfgMLP1>(o, ep, fp16, s);
  }
  if (out32 == nullptr) {
    PG_REQUIRE(out16 != nullptr && N % 256 == 0, "pangu_linear(16-bit only): needs a 16-bit output and N %% 256 == 0");
    return launch_gemm<CfgLin16>(o, ep, fp16, s);
  }
  PG_REQUIRE(out32 != nullptr && out16 != nullptr && N % 192 == 0, "pangu_linear: needs both outputs and N %% 192 == 0");
  return launch_gemm<CfgPlain192>(o, ep, fp16, s);
}

// ---------------------------------------------------------------------------------------
// backward pass
// ---------------------------------------------------------------------------------------
extern "C" int pangu_cast16_t(const float* src, void* dst, int rows, int cols, int rows_pad, int cols_pad, int fp16,
                              void* stream) {
  PG_TRY(ensure_init());
  PG_REQUIRE(rows > 0 && cols > 0 && rows_pad >= rows && cols_pad >= cols, "bad cast16_t shape");
  dim3 grid((cols_pad + 31) / 32, (rows_pad + 31) / 32);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (fp16) cast16_t_kernel<true><<<grid, 256, 0, s>>>(src, static_cast<uint16_t*>(dst), rows, cols, rows_pad, cols_pad);
  else cast16_t_kernel<false><<<grid, 256, 0, s>>>(src, static_cast<uint16_t*>(dst), rows, cols, rows_pad, cols_pad);
  PG_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pangu_dgrad(const void* a16, const void* wt16, const float* bias, const float* resid32, float* out32,
                           void* out16, int M, int N, int K, int kind, int Z, int H, int W, int roll, int fp16,
                           void* stream) {
  PG_TRY(ensure_init());
  PG_REQUIRE(kind >= 0 && kind <= 3, "dgrad: kind must be 0..3");
  GemmOperands o{a16, uint64_t(K), nullptr, 0, K, 0, wt16, uint64_t(K), M, N};
  EpiArgs ep = epi_defaults();
  ep.bias = bias;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (kind == 1) {
    PG_REQUIRE(out16 != nullptr && N % 256 == 0, "dgrad(kind 1): needs a 16-bit output and N %% 256 == 0");
    ep.out16 = out16; ep.ld16 = N;
    return launch_gemm<CfgLin16>(o, ep, fp16, s);
  }
  PG_REQUIRE(N % 192 == 0, "dgrad: N must be a multiple of 192");
  if (kind == 3) {
    PG_REQUIRE(out16 != nullptr, "dgrad(kind 3): needs a 16-bit output");
    PG_TRY(check_grid(Z, H, W, N, 0));
    PG_REQUIRE(M == Z * H * W, "dgrad(kind 3): M must be the token count");
    ep.Z = Z; ep.H = H; ep.W = W;
    ep.out16 = out16; ep.ld16 = N;
    ep.dstmap = DM_TOK2WIN; ep.roll_out = roll ? 1 : 0;
    return launch_gemm<CfgOut16Map>(o, ep, fp16, s);
  }
  PG_REQUIRE(out32 != nullptr, "dgrad: needs an fp32 output");
  ep.out32 = out32; ep.ld32 = N; ep.resid = resid32;
  if (resid32 == nullptr) ep.debug |= 1;      // no residual: plain fp32 store
  if (kind == 2) {
    PG_TRY(check_grid(Z, H, W, N, 0));
    const Geo g = make_geo(Z, H, W);
    PG_REQUIRE(M == g.nLon * g.types * 144, "dgrad(kind 2): M must be the window-padded row count");
    ep.Z = Z; ep.H = H; ep.W = W;
    ep.rowmap = RM_WIN2TOK; ep.roll_in = roll ? 1 : 0;
  }
  return launch_gemm<CfgAcc192>(o, ep, fp16, s);
}

static int make_map_cols(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems) {
  PG_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (pitch_elems * 2) % 16 == 0 && cols % 8 == 0,
             "wgrad operand not TMA-addressable (base %p, pitch %llu)", base, (unsigned long long)pitch_elems);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {64, WG_TOK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(-6, "cuTensorMapEncodeTiled(wgrad) failed (%d)", int(r));
  return 0;
}

extern "C" int pangu_wgrad(const void* dy16, int ld_dy, const void* x16, int ld_x, float* dw, int ldw, int k_off,
                           int M, int N, int K, float alpha, int fp16, void* stream) {
  PG_TRY(ensure_init());
  PG_REQUIRE(M > 0 && N > 0 && K > 0 && N % 8 == 0 && K % 4 == 0, "wgrad: bad shape (M=%d N=%d K=%d)", M, N, K);
  PG_REQUIRE((reinterpret_cast<uintptr_t>(dw) & 15) == 0 && ldw % 4 == 0 && k_off % 4 == 0, "wgrad: dW not 16 B aligned");
  CUtensorMap mdy, mx;
  PG_TRY(make_map_cols(&mdy, dy16, M, uint64_t((N + 7) / 8 * 8), ld_dy));
  PG_TRY(make_map_cols(&mx, x16, M, uint64_t((K + 7) / 8 * 8), ld_x));
  WgradArgs a;
  a.dw = dw; a.ldw = ldw; a.k_off = k_off; a.N = N; a.K = K; a.alpha = alpha;
  const int Kp = (K + 63) / 64 * 64;
  a.KB = Kp % 192 == 0 ? 192 : (Kp % 256 == 0 ? 256 : (Kp % 128 == 0 ? 128 : 64));
  a.n_tiles = (N + 127) / 128;
  a.k_tiles = Kp / a.KB;
  a.chunks = (M + WG_TOK - 1) / WG_TOK;
  int splits = (2 * g_num_sms) / (a.n_tiles * a.k_tiles);
  if (splits > a.chunks / 4) splits = a.chunks / 4;
  if (splits < 1) splits = 1;
  a.splits = splits;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = a.n_tiles * a.k_tiles * a.splits;
  static bool attr_done_dev[kMaxDevices][2] = {{false, false}};
  bool* attr_done = attr_done_dev[cur_dev()];
  if (fp16) {
    if (!attr_done[1]) { PG_CUDA(cudaFuncSetAttribute(wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES)); attr_done[1] = true; }
    wgrad_kernel<true><<<grid, WG_THREADS, WG_SMEM_BYTES, s>>>(mdy, mx, a);
  } else {
    if (!attr_done[0]) { PG_CUDA(cudaFuncSetAttribute(wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES)); attr_done[0] = true; }
    wgrad_kernel<false><<<grid, WG_THREADS, WG_SMEM_BYTES, s>>>(mdy, mx, a);
  }
  PG_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pangu_colsum16(const void* src16, int ld, float* out, int M, int N, int n_valid, float alpha, int fp16,
                              void* stream) {
  PG_TRY(ensure_init());
  PG_REQUIRE(M > 0 && N > 0 && N % 8 == 0 && N <= 2048 && n_valid <= N && ld % 8 == 0, "colsum16: bad shape");
  PG_REQUIRE((reinterpret_cast<uintptr_t>(src16) & 15) == 0, "colsum16: source not 16 B aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int rpi = 256 / (N / 8);
  int grid = (M + rpi * 64 - 1) / (rpi * 64);
  if (grid > 4 * g_num_sms) grid = 4 * g_num_sms;
  if (grid < 1) grid = 1;
  if (fp16) colsum16_kernel<true><<<grid, 256, N * sizeof(float), s>>>(static_cast<const uint16_t*>(src16), out, M, N, ld, n_valid, alpha);
  else colsum16_kernel<false><<<grid, 256, N * sizeof(float), s>>>(static_cast<const uint16_t*>(src16), out, M, N, ld, n_valid, alpha);
  PG_CUDA(cudaGetLastError());
  return 0;
}

template <bool kFp16>
static int launch_ln_bwd(const LnBwdArgs& a, int mode, cudaStream_t s) {
  int grid = (a.rows + 7) / 8;                 // 8 warps (rows) per block, grid-stride
  if (grid > 8 * g_num_sms) grid = 8 * g_num_sms;
  if (mode == LNB_IDENT && a.C == 192) ln_bwd_kernel<kFp16, LNB_IDENT, 192><<<grid, 256, 0, s>>>(a);
  else if (mode == LNB_IDENT && a.C == 384) ln_bwd_kernel<kFp16, LNB_IDENT, 384><<<grid, 256, 0, s>>>(a);
  else if (mode == LNB_UP && a.C == 192) ln_bwd_kernel<kFp16, LNB_UP, 192><<<grid, 256, 0, s>>>(a);
  else if (mode == LNB_DOWN && a.C == 768) ln_bwd_kernel<kFp16, LNB_DOWN, 768><<<grid, 256, 0, s>>>(a);
  else return fail(-1, "layernorm_bwd: unsupported (mode %d, C %d)", mode, a.C);
  PG_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pangu_layernorm_bwd(const float* y, const float* g, const float* gamma, void* dx16, float* dx32,
                                   float* dgamma, float* dbeta, float* dbias, int rows, int C, int mode, int Z, int H, int W,
                                   float scale, float palpha, int fp16, void* stream) {
  PG_TRY(ensure_init());
  PG_REQUIRE(rows > 0 && y && g && gamma, "layernorm_bwd: null argument");
  PG_REQUIRE(mode == LNB_DOWN ? dx32 != nullptr : dx16 != nullptr, "layernorm_bwd: missing output");
  if (mode == LNB_UP) PG_REQUIRE(rows == Z * H * W && W % 2 == 0, "layernorm_bwd(up): rows must be the high-res token count");
  if (mode == LNB_DOWN) PG_REQUIRE(rows == Z * ((H + 1) / 2) * (W / 2), "layernorm_bwd(down): rows must be the low-res token count");
  LnBwdArgs a;
  a.y = y; a.g = g; a.gamma = gamma; a.dx16 = dx16; a.dx32 = dx32; a.dgamma = dgamma; a.dbeta = dbeta; a.dbias = dbias;
  a.rows = rows; a.C = C; a.Z = Z; a.H = H; a.W = W; a.scale = scale; a.palpha = palpha; a.eps = 1e-5f;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return fp16 ? launch_ln_bwd<true>(a, mode, s) : launch_ln_bwd<false>(a, mode, s);
}

extern "C" int pangu_gelu_bwd(void* dh16, const void* pre16, int M, int N, float* dbias, float alpha, int fp16, void* stream) {
  PG_TRY(ensure_init());
  PG_REQUIRE(M > 0 && N > 0 && N % 8 == 0 && N <= 2048, "gelu_bwd: N must be a positive multiple of 8, at most 2048");
  PG_REQUIRE((reinterpret_cast<uintptr_t>(dh16) & 15) == 0 && (reinterpret_cast<uintptr_t>(pre16) & 15) == 0, "gelu_bwd: misaligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int tpr = N / 8;
  const int rpi = 256 / tpr > 0 ? 256 / tpr : 1;
  const int block = tpr * rpi;                  // 192 for N = 768 / 1536
  int grid = (M + 2 * rpi - 1) / (2 * rpi);
  if (grid > 16 * g_num_sms) grid = 16 * g_num_sms;
  const size_t smem = dbias ? N * sizeof(float) : 0;
  if (fp16) gelu_bwd_kernel<true><<<grid, block, smem, s>>>(static_cast<uint16_t*>(dh16), static_cast<const uint16_t*>(pre16), M, N, dbias, alpha);
  else gelu_bwd_kernel<false><<<grid, block, smem, s>>>(static_cast<uint16_t*>(dh16), static_cast<const uint16_t*>(pre16), M, N, dbias, alpha);
  PG_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pangu_window_attention_bwd(const void* qkv16, const void* datt16w, const float* earth_bias, void* dqkv16,
                                          float* dbias, float* dbqkv, int Z, int H, int W, int C, int heads, int roll,
                                          float palpha, int fp16, void* stream) {
  PG_TRY(ensure_init());
  PG_TRY(check_grid(Z, H, W, C, heads));
  const Geo g = make_geo(Z, H, W);
  AttnBwdArgs a;
  a.qkv = qkv16; a.datt = datt16w; a.bias = earth_bias; a.dqkv = dqkv16; a.dbias = dbias; a.dbqkv = dbqkv;
  a.C = C; a.heads = heads; a.types = g.types; a.nLon = g.nLon; a.nH = g.nH; a.roll = roll ? 1 : 0;
  a.plane_rows = sh_rows_padded(g.nLon * g.types * 144);
  a.q_scale = 0.17677669529663687f;
  a.palpha = palpha;
  PG_REQUIRE((reinterpret_cast<uintptr_t>(qkv16) & 15) == 0 && (reinterpret_cast<uintptr_t>(datt16w) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(earth_bias) & 7) == 0 && (reinterpret_cast<uintptr_t>(dqkv16) & 3) == 0,
             "attention_bwd: misaligned argument");
  const long long units = (long long)g.types * heads * g.nLon;
  const int grid = int(units < g_num_sms ? units : g_num_sms);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static bool attr_done_dev[kMaxDevices][2] = {{false, false}};
  bool* attr_done = attr_done_dev[cur_dev()];
  if (fp16) {
    if (!attr_done[1]) { PG_CUDA(cudaFuncSetAttribute(window_attention_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATB_SMEM_BYTES)); attr_done[1] = true; }
    window_attention_bwd_kernel<true><<<grid, ATB_THREADS, ATB_SMEM_BYTES, s>>>(a);
  } else {
    if (!attr_done[0]) { PG_CUDA(cudaFuncSetAttribute(window_attention_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATB_SMEM_BYTES)); attr_done[0] = true; }
    window_attention_bwd_kernel<false><<<grid, ATB_THREADS, ATB_SMEM_BYTES, s>>>(a);
  }
  PG_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pangu_recover_grad_gather(const float* d_upper, const float* d_surface, void* dy_upper, void* dy_surface,
                                         int lat, int lon, float scale, int fp16, void* stream) {
  PG_TRY(ensure_init());
  PG_REQUIRE(lon % 48 == 0 && lon > 0, "lon must be a positive multiple of 48 (got %d)", lon);
  const int Hh = (lat + 3) / 4, Ww = lon / 4;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  EmbedArgs ea{d_upper, d_surface, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, dy_upper, dy_surface, lat, lon, Hh, Ww, scale};
  dim3 grid((Ww + EMB_TT - 1) / EMB_TT, Hh, 8);
  if (fp16) embed_im2col_kernel<true><<<grid, EMB_THREADS, 0, s>>>(ea);
  else embed_im2col_kernel<false><<<grid, EMB_THREADS, 0, s>>>(ea);
  PG_CUDA(cudaGetLastError());
  return 0;
}
