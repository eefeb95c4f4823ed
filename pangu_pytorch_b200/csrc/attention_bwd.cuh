// Backward of the Earth-specific window attention (autograd of models/layers.py:368-415).
//
// Per (window, head), with q already scaled by 32^-0.5 as stored by the QKV GEMM:
//   S = q k^T + bias[type, head] (+ mask)       P = softmax(S)
//   dP = dO v^T      D_i = sum_j P_ij dP_ij     dS = P o (dP - D)
//   dV = P^T dO      dK = dS^T q                dQ = 32^-0.5 * dS k        dBias[type, head] += dS
//
// v1 on warp-level mma.sync (m16n8k16): the work is softmax/HBM bound like the forward.  Nine warps;
// phase 1: warp w owns query rows 16w..16w+15 (S, P, dP, dS in registers; dQ; P and dS are also written
// to shared memory as 16-bit); phase 2: warp w owns key rows 16w..16w+15 and forms dV / dK from the
// transposed P / dS tiles (ldmatrix.trans).  dBias is accumulated in shared memory (fp32) over the
// longitude windows of a (type, head) segment and flushed with atomics once per segment, so the
// 253 M-entry bias gradient costs one atomic per entry and CTA-segment instead of one per window.
// Persistent CTAs walk equal contiguous ranges of the (type, head, lon window) list, as in the forward.
#pragma once
#include "attention.cuh"

namespace pg {

constexpr int ATB_THREADS = 288;
constexpr int ATB_P_PITCH = 304;                    // bytes per row of the 16-bit P / dS tiles (19 x 16 B: conflict-free ldmatrix)
constexpr int ATB_DB_PITCH = 152;                   // floats per row of the dBias accumulator
constexpr int ATB_TILES_BYTES = 4 * ATT_TILE_BYTES; // q, k, v, dO
constexpr int ATB_P_BYTES = ATT_TOK * ATB_P_PITCH;
constexpr int ATB_DB_BYTES = ATT_TOK * ATB_DB_PITCH * 4;
constexpr int ATB_CS_BYTES = 3 * 32 * 4;              // column sums of dq / dk / dv of the current head (linear1.bias gradient)
constexpr int ATB_SMEM_BYTES = ATB_TILES_BYTES + 2 * ATB_P_BYTES + ATB_DB_BYTES + ATB_CS_BYTES;
static_assert(ATB_SMEM_BYTES <= 232448, "attention backward shared memory budget");

struct AttnBwdArgs {
  const void* qkv;      // head-major [3*heads planes][plane_rows][32] 16-bit, q pre-scaled
  const void* datt;     // [Tp][C] 16-bit, window order (pad rows zero): gradient w.r.t. the merged-head attention output
  const float* bias;    // [types][heads][144][144] fp32
  void* dqkv;           // [Tp][3C] 16-bit, window order, column s*C + head*32 + d; dq is w.r.t. the UNscaled q
  float* dbias;         // [types][heads][144][144] fp32, accumulated into (nullptr: frozen table, skipped)
  float* dbqkv;         // [3C] fp32 += palpha * column sums of dqkv: bias gradient of attention.linear1 (nullable)
  int C, heads, types, nLon, nH, roll, plane_rows;
  float q_scale;
  float palpha;         // factor on the bias gradient (1 / loss scale)
};

template <bool kFp16>
__global__ void __launch_bounds__(ATB_THREADS, 1) window_attention_bwd_kernel(const AttnBwdArgs a) {
  extern __shared__ __align__(128) uint8_t atb_smem[];
  uint8_t* tq = atb_smem;
  uint8_t* tk = tq + ATT_TILE_BYTES;
  uint8_t* tv = tk + ATT_TILE_BYTES;
  uint8_t* tdo = tv + ATT_TILE_BYTES;
  uint8_t* s_p = atb_smem + ATB_TILES_BYTES;
  uint8_t* s_ds = s_p + ATB_P_BYTES;
  float* s_db = reinterpret_cast<float*>(s_ds + ATB_P_BYTES);
  float* s_cs = s_db + ATT_TOK * ATB_DB_PITCH;       // [3][32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, q4 = lane & 3;
  constexpr float kLog2e = 1.4426950408889634f;

  const long long total = (long long)a.types * a.heads * a.nLon;
  const int u_begin = int(total * blockIdx.x / gridDim.x);
  const int u_end = int(total * (blockIdx.x + 1) / gridDim.x);

  for (int i = threadIdx.x; i < ATT_TOK * ATB_DB_PITCH; i += ATB_THREADS) s_db[i] = 0.f;
  if (threadIdx.x < 96) s_cs[threadIdx.x] = 0.f;
  // column sums of a 16-row fragment block: reduce over the 8 row groups of the warp, one shared atomic per column
  auto add_cols = [&](int x, int n, float c0, float c1) {
    c0 += __shfl_xor_sync(0xffffffffu, c0, 4); c0 += __shfl_xor_sync(0xffffffffu, c0, 8); c0 += __shfl_xor_sync(0xffffffffu, c0, 16);
    c1 += __shfl_xor_sync(0xffffffffu, c1, 4); c1 += __shfl_xor_sync(0xffffffffu, c1, 8); c1 += __shfl_xor_sync(0xffffffffu, c1, 16);
    if (gq == 0) { atomicAdd(&s_cs[x * 32 + 8 * n + 2 * q4], c0); atomicAdd(&s_cs[x * 32 + 8 * n + 2 * q4 + 1], c1); }
  };
  __syncthreads();

  const uint32_t sq = smem_u32(tq), sk = smem_u32(tk), sv = smem_u32(tv), sdo = smem_u32(tdo);
  const uint32_t sp_u = smem_u32(s_p), sds_u = smem_u32(s_ds);
  const int i0 = 16 * warp;            // phase 1: query rows; phase 2: key rows

  for (int u = u_begin; u < u_end; ++u) {
    const int th = u / a.nLon, lw = u % a.nLon;
    const int t = th / a.heads, head = th % a.heads;
    const int zw = t / a.nH, hw = t % a.nH;
    const bool zsplit = a.roll && (zw == a.types / a.nH - 1);
    const bool hsplit = a.roll && (hw == a.nH - 1);
    const int row0 = (lw * a.types + t) * ATT_TOK;

    // ---- load q, k, v (contiguous 9216 B planes) and dO (64 B per row) into swizzled tiles
    {
      const uint8_t* qkvb = reinterpret_cast<const uint8_t*>(a.qkv);
      for (int c = threadIdx.x; c < 4 * 576; c += ATB_THREADS) {
        const int tile = c / 576, cc = c % 576, r = cc >> 2, ch = cc & 3;
        const uint8_t* src;
        if (tile < 3) src = qkvb + (size_t(tile * a.heads + head) * a.plane_rows + row0) * 64 + cc * 16;
        else src = reinterpret_cast<const uint8_t*>(a.datt) + (size_t(row0 + r) * a.C + head * 32) * 2 + ch * 16;
        cp_async16(sq + tile * ATT_TILE_BYTES + att_off(r, ch), src);
      }
      cp_async_commit();
    }
    // bias + mask rows of this warp go straight into the S accumulators while the tiles are in flight
    float s[18][4];
    {
      const float* b0 = a.bias + (size_t(th) * ATT_TOK + i0 + gq) * ATT_TOK + 2 * q4;
      const float* b1 = b0 + 8 * ATT_TOK;
      const int ri0 = i0 + gq, ri1 = ri0 + 8;
#pragma unroll
      for (int j = 0; j < 18; ++j) {
        const float2 lo = *reinterpret_cast<const float2*>(b0 + 8 * j);
        const float2 hi = *reinterpret_cast<const float2*>(b1 + 8 * j);
        const int cj = 8 * j + 2 * q4;     // cj and cj + 1 share zl and hl (12-token rows, even cj)
        const bool mz0 = zsplit && ((ri0 / 72) != (cj / 72)), mz1 = zsplit && ((ri1 / 72) != (cj / 72));
        const bool ch = ((cj / 12) % 6) < 3;
        const bool mh0 = hsplit && ((((ri0 / 12) % 6) < 3) != ch), mh1 = hsplit && ((((ri1 / 12) % 6) < 3) != ch);
        const float m0 = (mz0 || mh0) ? -100.f : 0.f, m1 = (mz1 || mh1) ? -100.f : 0.f;
        s[j][0] = lo.x + m0; s[j][1] = lo.y + m0; s[j][2] = hi.x + m1; s[j][3] = hi.y + m1;
      }
    }
    cp_async_wait<0>();
    __syncthreads();

    // =============================== phase 1: query rows i0..i0+15 ===============================
    uint32_t qa[2][4], da[2][4];
    {
      const int r = i0 + (lane & 15);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        ldsm_x4(sq + att_off(r, ks * 2 + (lane >> 4)), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
        ldsm_x4(sdo + att_off(r, ks * 2 + (lane >> 4)), da[ks][0], da[ks][1], da[ks][2], da[ks][3]);
      }
    }
    float dp[18][4];
#pragma unroll
    for (int j = 0; j < 18; ++j) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(sk + att_off(8 * j + (lane & 7), lane >> 3), b0, b1, b2, b3);
      mma16816<kFp16>(s[j], qa[0], b0, b1);
      mma16816<kFp16>(s[j], qa[1], b2, b3);
      ldsm_x4(sv + att_off(8 * j + (lane & 7), lane >> 3), b0, b1, b2, b3);
      dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
      mma16816<kFp16>(dp[j], da[0], b0, b1);
      mma16816<kFp16>(dp[j], da[1], b2, b3);
    }
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 18; ++j) {
      m0 = max3(m0, s[j][0], s[j][1]);
      m1 = max3(m1, s[j][2], s[j][3]);
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int j = 0; j < 18; ++j) {
      s[j][0] = fast_exp2((s[j][0] - m0) * kLog2e); s[j][1] = fast_exp2((s[j][1] - m0) * kLog2e);
      s[j][2] = fast_exp2((s[j][2] - m1) * kLog2e); s[j][3] = fast_exp2((s[j][3] - m1) * kLog2e);
      l0 += s[j][0] + s[j][1];
      l1 += s[j][2] + s[j][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float il0 = 1.0f / l0, il1 = 1.0f / l1;
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int j = 0; j < 18; ++j) {
      s[j][0] *= il0; s[j][1] *= il0; s[j][2] *= il1; s[j][3] *= il1;
      d0 += s[j][0] * dp[j][0] + s[j][1] * dp[j][1];
      d1 += s[j][2] * dp[j][2] + s[j][3] * dp[j][3];
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    {
      float* db0 = s_db + (i0 + gq) * ATB_DB_PITCH + 2 * q4;
      float* db1 = db0 + 8 * ATB_DB_PITCH;
      uint8_t* p0 = s_p + (i0 + gq) * ATB_P_PITCH + 4 * q4;
      uint8_t* p1 = p0 + 8 * ATB_P_PITCH;
      uint8_t* e0 = s_ds + (i0 + gq) * ATB_P_PITCH + 4 * q4;
      uint8_t* e1 = e0 + 8 * ATB_P_PITCH;
#pragma unroll
      for (int j = 0; j < 18; ++j) {
        dp[j][0] = s[j][0] * (dp[j][0] - d0); dp[j][1] = s[j][1] * (dp[j][1] - d0);
        dp[j][2] = s[j][2] * (dp[j][2] - d1); dp[j][3] = s[j][3] * (dp[j][3] - d1);
        if (a.dbias) {
          float2 x = *reinterpret_cast<float2*>(db0 + 8 * j);
          x.x += dp[j][0]; x.y += dp[j][1];
          *reinterpret_cast<float2*>(db0 + 8 * j) = x;
          float2 y = *reinterpret_cast<float2*>(db1 + 8 * j);
          y.x += dp[j][2]; y.y += dp[j][3];
          *reinterpret_cast<float2*>(db1 + 8 * j) = y;
        }
        *reinterpret_cast<uint32_t*>(p0 + 16 * j) = pack16<kFp16>(s[j][0], s[j][1]);
        *reinterpret_cast<uint32_t*>(p1 + 16 * j) = pack16<kFp16>(s[j][2], s[j][3]);
        *reinterpret_cast<uint32_t*>(e0 + 16 * j) = pack16<kFp16>(dp[j][0], dp[j][1]);
        *reinterpret_cast<uint32_t*>(e1 + 16 * j) = pack16<kFp16>(dp[j][2], dp[j][3]);
      }
    }
    uint16_t* outp = reinterpret_cast<uint16_t*>(a.dqkv);
    const size_t ld = size_t(3) * a.C;
    {
      // dQ = scale * dS k   (A = dS fragments, B = k rows through ldmatrix.trans)
      float o[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 9; ++kk) {
        uint32_t pa[4];
        pa[0] = pack16<kFp16>(dp[2 * kk][0], dp[2 * kk][1]);
        pa[1] = pack16<kFp16>(dp[2 * kk][2], dp[2 * kk][3]);
        pa[2] = pack16<kFp16>(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
        pa[3] = pack16<kFp16>(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(sk + att_off(16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * np + (lane >> 4)), b0, b1, b2, b3);
          mma16816<kFp16>(o[2 * np], pa, b0, b1);
          mma16816<kFp16>(o[2 * np + 1], pa, b2, b3);
        }
      }
      uint16_t* r0p = outp + size_t(row0 + i0 + gq) * ld + head * 32 + 2 * q4;
      uint16_t* r1p = r0p + 8 * ld;
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        *reinterpret_cast<uint32_t*>(r0p + 8 * n) = pack16<kFp16>(o[n][0] * a.q_scale, o[n][1] * a.q_scale);
        *reinterpret_cast<uint32_t*>(r1p + 8 * n) = pack16<kFp16>(o[n][2] * a.q_scale, o[n][3] * a.q_scale);
        if (a.dbqkv) add_cols(0, n, (o[n][0] + o[n][2]) * a.q_scale, (o[n][1] + o[n][3]) * a.q_scale);
      }
    }
    __syncthreads();

    // =============================== phase 2: key rows i0..i0+15 ===============================
    {
      float dv[4][4], dk[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
        dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
      }
      // transposed A fragments: matrix (lane >> 3): rows i = 16 ib + (lane & 7) + 8 * (lane >> 4), columns j = i0 + 8 * ((lane >> 3) & 1)
      const uint32_t a_off = uint32_t(((lane & 7) + ((lane >> 4) & 1) * 8) * ATB_P_PITCH + (i0 + ((lane >> 3) & 1) * 8) * 2);
#pragma unroll
      for (int ib = 0; ib < 9; ++ib) {
        uint32_t pa[4], ea[4];
        ldsm_x4_t(sp_u + a_off + ib * 16 * ATB_P_PITCH, pa[0], pa[1], pa[2], pa[3]);
        ldsm_x4_t(sds_u + a_off + ib * 16 * ATB_P_PITCH, ea[0], ea[1], ea[2], ea[3]);
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t b0, b1, b2, b3;
          const uint32_t boff = att_off(16 * ib + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * np + (lane >> 4));
          ldsm_x4_t(sdo + boff, b0, b1, b2, b3);
          mma16816<kFp16>(dv[2 * np], pa, b0, b1);
          mma16816<kFp16>(dv[2 * np + 1], pa, b2, b3);
          ldsm_x4_t(sq + boff, b0, b1, b2, b3);
          mma16816<kFp16>(dk[2 * np], ea, b0, b1);
          mma16816<kFp16>(dk[2 * np + 1], ea, b2, b3);
        }
      }
      uint16_t* k0p = outp + size_t(row0 + i0 + gq) * ld + a.C + head * 32 + 2 * q4;
      uint16_t* k1p = k0p + 8 * ld;
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        *reinterpret_cast<uint32_t*>(k0p + 8 * n) = pack16<kFp16>(dk[n][0], dk[n][1]);
        *reinterpret_cast<uint32_t*>(k1p + 8 * n) = pack16<kFp16>(dk[n][2], dk[n][3]);
        *reinterpret_cast<uint32_t*>(k0p + a.C + 8 * n) = pack16<kFp16>(dv[n][0], dv[n][1]);
        *reinterpret_cast<uint32_t*>(k1p + a.C + 8 * n) = pack16<kFp16>(dv[n][2], dv[n][3]);
        if (a.dbqkv) {
          add_cols(1, n, dk[n][0] + dk[n][2], dk[n][1] + dk[n][3]);
          add_cols(2, n, dv[n][0] + dv[n][2], dv[n][1] + dv[n][3]);
        }
      }
    }
    __syncthreads();     // tiles and P / dS are free again

    // ---- end of a (type, head) segment (or of this CTA's range): flush the column sums (linear1.bias gradient) ...
    if (a.dbqkv && (lw == a.nLon - 1 || u + 1 == u_end)) {     // (the trailing __syncthreads of phase 2 ordered the shared atomics)
      if (threadIdx.x < 96) {
        atomicAdd(a.dbqkv + (threadIdx.x >> 5) * a.C + head * 32 + (threadIdx.x & 31), a.palpha * s_cs[threadIdx.x]);
        s_cs[threadIdx.x] = 0.f;
      }
      __syncthreads();
    }
    // ... and the earth_specific_bias gradient
    if (a.dbias && (lw == a.nLon - 1 || u + 1 == u_end)) {
      float* g = a.dbias + size_t(th) * ATT_TOK * ATT_TOK;
      for (int i = threadIdx.x; i < ATT_TOK * ATT_TOK; i += ATB_THREADS) {
        const int r = i / ATT_TOK, c = i % ATT_TOK;
        float* p = s_db + r * ATB_DB_PITCH + c;
        atomicAdd(g + i, a.palpha * *p);
        *p = 0.f;
      }
      __syncthreads();
    }
  }
}

}  // namespace pg
