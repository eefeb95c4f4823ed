// Memory-bound gather kernels that build GEMM A-operands without materialising the
// reference's permute/pad/cat chains:
//   * embed_im2col_kernel      (models/layers.py:48-85)
//   * downsample_gather_ln_kernel (models/layers.py:436-454)
//   * cast16_kernel            fp32 -> bf16/fp16 (weight preparation, zero padded K)
#pragma once
#include "common.cuh"
#include "geometry.cuh"

namespace pg {

constexpr int EMB_TT = 120;                // tokens per CTA tile (480 longitudes)
constexpr int EMB_PITCH = 192 * 2 + 8;     // bytes per staged token row (+8: conflict-free 8 B column writes)
constexpr int EMB_THREADS = 256;

struct EmbedArgs {
  const float* upper;     // [5][13][lat][lon]
  const float* surface;   // [4][lat][lon]
  const float* s_mean; const float* s_std;   // [4]
  const float* u_mean; const float* u_std;   // [13][5]  (level axis reversed w.r.t. data, layers.py:73-76)
  const float* maps;      // [3][4*Hh][lon]          (nullptr: zeros -- gradient patchify mode)
  const float* const_h;   // [13][lat][lon]          (nullptr: zeros; s_mean / u_mean nullptr: no normalisation)
  void* a_upper;          // [7*Hh*Ww][192] 16-bit, feature ((c*2+dz)*4+dh)*4+dw
  void* a_surface;        // [Hh*Ww][128]   16-bit, feature (c*4+dh)*4+dw, zero for f >= 112
  int lat, lon, Hh, Ww;
  float gscale;           // gradient patchify mode: factor on the values (loss scale); unused otherwise
};

template <bool kFp16>
__global__ void __launch_bounds__(EMB_THREADS) embed_im2col_kernel(const EmbedArgs a) {
  __shared__ __align__(16) uint8_t tile[EMB_TT * EMB_PITCH];
  const int zp = blockIdx.z;                 // 0 = surface plane, 1..7 = upper patch level zt + 1
  const int ht = blockIdx.y;
  const int wt0 = blockIdx.x * EMB_TT;
  const int ntok = min(EMB_TT, a.Ww - wt0);
  const int nfeat = zp == 0 ? 128 : 192;
  const int ngrp = zp == 0 ? 28 : 48;        // (c, [dz,] dh) groups of 4 dw
  const size_t plane = size_t(a.lat) * a.lon;

  // zero fill for the surface K padding (features 112..127)
  if (zp == 0) {
    for (int i = threadIdx.x; i < ntok * 4; i += EMB_THREADS)
      *reinterpret_cast<uint2*>(tile + (i >> 2) * EMB_PITCH + 224 + (i & 3) * 8) = make_uint2(0u, 0u);
  }
  // Four items per thread and iteration, all four loads issued before the first use: with one 16 B load in flight per thread
  // the kernel ran at 2.9 TB/s (latency-bound: 4 KB in flight per CTA).
  const int nitem = ngrp * ntok;
  for (int base = threadIdx.x; base < nitem; base += 4 * EMB_THREADS) {
    float4 v[4];
    float mm[4], ss[4];
    int off[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = base + u * EMB_THREADS;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      mm[u] = 0.f; ss[u] = 1.f; off[u] = -1;
      if (idx >= nitem) continue;
      const int grp = idx / ntok, tk = idx - grp * ntok;     // consecutive threads -> consecutive longitudes
      off[u] = tk * EMB_PITCH + grp * 8;
      const int dh = grp & 3;
      const int la = 4 * ht + dh;
      const int lo = 4 * (wt0 + tk);
      const float* src = nullptr;
      if (zp == 0) {
        const int c = grp >> 2;                        // 0..6
        if (c < 4) {
          if (la < a.lat) {
            src = a.surface + c * plane + size_t(la) * a.lon + lo;
            if (a.s_mean) { mm[u] = a.s_mean[c]; ss[u] = a.s_std[c]; }
          }
        } else if (a.maps) {
          src = a.maps + (size_t(c - 4) * (4 * a.Hh) + la) * a.lon + lo;
        }
      } else {
        const int dz = (grp >> 2) & 1, c = grp >> 3;   // 0..5
        const int lev = 2 * (zp - 1) + dz;
        if (lev < 13 && la < a.lat) {
          if (c < 5) {
            src = a.upper + (size_t(c) * 13 + lev) * plane + size_t(la) * a.lon + lo;
            if (a.u_mean) { mm[u] = a.u_mean[(12 - lev) * 5 + c]; ss[u] = a.u_std[(12 - lev) * 5 + c]; }
          } else if (a.const_h) {
            src = a.const_h + size_t(lev) * plane + size_t(la) * a.lon + lo;
          }
        }
      }
      if (src) v[u] = *reinterpret_cast<const float4*>(src);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (off[u] < 0) continue;
      float4 w = v[u];
      const float m = mm[u], sd = ss[u];
      w.x = (w.x - m) / sd; w.y = (w.y - m) / sd; w.z = (w.z - m) / sd; w.w = (w.w - m) / sd;      // m = 0, sd = 1: exact identity
      if (a.s_mean == nullptr) { w.x *= a.gscale; w.y *= a.gscale; w.z *= a.gscale; w.w *= a.gscale; }
      *reinterpret_cast<uint2*>(tile + off[u]) = make_uint2(pack16<kFp16>(w.x, w.y), pack16<kFp16>(w.z, w.w));
    }
  }
  __syncthreads();
  // the tile's token rows are contiguous in the output: flat 8-byte copy
  const int p8 = nfeat / 4;                          // 8 B pieces per token
  uint2* dst;
  if (zp == 0) dst = reinterpret_cast<uint2*>(a.a_surface) + (size_t(ht) * a.Ww + wt0) * p8;
  else dst = reinterpret_cast<uint2*>(a.a_upper) + ((size_t(zp - 1) * a.Hh + ht) * a.Ww + wt0) * p8;
  for (int i = threadIdx.x; i < ntok * p8; i += EMB_THREADS)
    dst[i] = *reinterpret_cast<const uint2*>(tile + (i / p8) * EMB_PITCH + (i % p8) * 8);
}

// ---------------------------------------------------------------------------------------
struct DownArgs {
  const float* x;       // [Z*H*W][C] fp32
  const float* gamma;   // [4C]
  const float* beta;    // [4C]
  void* out;            // [Z*H2*W2][4C] 16-bit
  int Z, H, W, C;       // C == 192
  float eps;
};

// one warp per output token: 2x2 merge (zero row when 2*h2+dh == H) + LayerNorm(4C)
template <bool kFp16>
__global__ void __launch_bounds__(256) downsample_gather_ln_kernel(const DownArgs a) {
  const int H2 = (a.H + 1) / 2, W2 = a.W / 2;
  const int ntok = a.Z * H2 * W2;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= ntok) return;
  const int w2 = warp % W2, h2 = (warp / W2) % H2, z = warp / (W2 * H2);
  // 4C = 768 values = 192 float4; lane handles float4 index lane + 32*i (i < 6):
  //   index q -> dh = q / 96, within-row offset (q % 96) covers the two adjacent tokens (dw, c)
  float4 v[6];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int q = lane + 32 * i;
    const int dh = q / 96, r = q % 96;
    const int h = 2 * h2 + dh;
    if (h < a.H) {
      v[i] = *reinterpret_cast<const float4*>(a.x + (size_t(z * a.H + h) * a.W + 2 * w2) * a.C + r * 4);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / 768.0f);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
    ss += dx * dx + dy * dy + dz * dz + dw * dw;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss * (1.0f / 768.0f) + a.eps);
  uint2* dst = reinterpret_cast<uint2*>(a.out) + size_t(warp) * 192;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int q = lane + 32 * i;
    const float4 g = *reinterpret_cast<const float4*>(a.gamma + q * 4);
    const float4 b = *reinterpret_cast<const float4*>(a.beta + q * 4);
    const float y0 = (v[i].x - mean) * rstd * g.x + b.x, y1 = (v[i].y - mean) * rstd * g.y + b.y;
    const float y2 = (v[i].z - mean) * rstd * g.z + b.z, y3 = (v[i].w - mean) * rstd * g.w + b.w;
    dst[q] = make_uint2(pack16<kFp16>(y0, y1), pack16<kFp16>(y2, y3));
  }
}

// ---------------------------------------------------------------------------------------
// x32 [T][C] natural order -> x16w [Tp][C] window order for roll state `roll` (pad rows = 0).
// Only used when a block is entered from a plain fp32 tensor (module-level API); inside the
// model the producing GEMM epilogue writes this layout directly.
template <bool kFp16>
__global__ void __launch_bounds__(256) to_window16_kernel(const float* __restrict__ x, uint16_t* __restrict__ out,
                                                          Geo g, int C, int roll, int rows) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int tok = roll < 0 ? row : win_row_to_token(g, row, roll);
  uint2* dst = reinterpret_cast<uint2*>(out + size_t(row) * C);
  for (int q = lane; q < C / 4; q += 32) {
    uint2 h = make_uint2(0u, 0u);
    if (tok >= 0) {
      const float4 v = *reinterpret_cast<const float4*>(x + size_t(tok) * C + q * 4);
      h = make_uint2(pack16<kFp16>(v.x, v.y), pack16<kFp16>(v.z, v.w));
    }
    dst[q] = h;
  }
}

// ---------------------------------------------------------------------------------------
// normBackData (era5_data/utils_data.py:324-330) in place on the model's normalised outputs:
// plane p < 65 is upper-air (c = p / 13, level l = p % 13, statistics row 12 - l of the
// (13,1,1,5) input-order arrays, i.e. weatherStatistics_output's level reversal), p >= 65 surface.
__global__ void __launch_bounds__(256) denorm_fields_kernel(float* __restrict__ upper, float* __restrict__ surface,
                                                            const float* __restrict__ s_mean,
                                                            const float* __restrict__ s_std,
                                                            const float* __restrict__ u_mean,
                                                            const float* __restrict__ u_std, int plane4) {
  const int p = blockIdx.y;
  float m, s;
  float4* base;
  if (p < 65) {
    const int c = p / 13, l = p % 13;
    m = u_mean[(12 - l) * 5 + c]; s = u_std[(12 - l) * 5 + c];
    base = reinterpret_cast<float4*>(upper) + size_t(p) * plane4;
  } else {
    m = s_mean[p - 65]; s = s_std[p - 65];
    base = reinterpret_cast<float4*>(surface) + size_t(p - 65) * plane4;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < plane4; i += gridDim.x * blockDim.x) {
    float4 v = base[i];
    v.x = fmaf(v.x, s, m); v.y = fmaf(v.y, s, m); v.z = fmaf(v.z, s, m); v.w = fmaf(v.w, s, m);
    base[i] = v;
  }
}

// ---------------------------------------------------------------------------------------
// Training loss of the reference (models/pangu_sample.py:57-67): the target is normalised with the
// output-order statistics (normData, era5_data/utils_data.py:315-321), then
//   L = mean(|o - t| * w_upper[c]) + 0.25 * mean(|o_s - t_s| * w_surface[c]).
// Plane p < 65: upper (c = p/13, level l = p%13, statistics row 12-l); p >= 65: surface.  Partial sums go
// to two fp64 accumulators; optionally dL/d(output) is written (seed of the backward pass).
struct L1Args {
  const float* out_u; const float* out_s; const float* tgt_u; const float* tgt_s;
  const float* s_mean; const float* s_std; const float* u_mean; const float* u_std;
  double* acc;            // [2] zero-initialised: sum_upper, sum_surface
  float* grad_u; float* grad_s;   // nullable
  float wu[5], ws[4];
  int plane4;             // lat*lon / 4
  float inv_nu, inv_ns;   // 1 / numel
};
__global__ void __launch_bounds__(256) l1_loss_kernel(const L1Args a) {
  const int p = blockIdx.y;
  const bool up = p < 65;
  float m, s, w;
  const float4 *o, *t;
  float4* g;
  if (up) {
    const int c = p / 13, l = p % 13;
    m = a.u_mean[(12 - l) * 5 + c]; s = a.u_std[(12 - l) * 5 + c]; w = a.wu[c];
    o = reinterpret_cast<const float4*>(a.out_u) + size_t(p) * a.plane4;
    t = reinterpret_cast<const float4*>(a.tgt_u) + size_t(p) * a.plane4;
    g = a.grad_u ? reinterpret_cast<float4*>(a.grad_u) + size_t(p) * a.plane4 : nullptr;
  } else {
    m = a.s_mean[p - 65]; s = a.s_std[p - 65]; w = a.ws[p - 65];
    o = reinterpret_cast<const float4*>(a.out_s) + size_t(p - 65) * a.plane4;
    t = reinterpret_cast<const float4*>(a.tgt_s) + size_t(p - 65) * a.plane4;
    g = a.grad_s ? reinterpret_cast<float4*>(a.grad_s) + size_t(p - 65) * a.plane4 : nullptr;
  }
  const float gs = up ? w * a.inv_nu : 0.25f * w * a.inv_ns;
  float sum = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.plane4; i += gridDim.x * blockDim.x) {
    const float4 ov = o[i], tv = t[i];
    const float d0 = ov.x - (tv.x - m) / s, d1 = ov.y - (tv.y - m) / s, d2 = ov.z - (tv.z - m) / s,
                d3 = ov.w - (tv.w - m) / s;
    sum += (fabsf(d0) + fabsf(d1)) + (fabsf(d2) + fabsf(d3));
    if (g) {
      g[i] = make_float4(d0 > 0.f ? gs : (d0 < 0.f ? -gs : 0.f), d1 > 0.f ? gs : (d1 < 0.f ? -gs : 0.f),
                         d2 > 0.f ? gs : (d2 < 0.f ? -gs : 0.f), d3 > 0.f ? gs : (d3 < 0.f ? -gs : 0.f));
    }
  }
  sum *= w;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int k = 0; k < 8; ++k) tot += double(part[k]);
    atomicAdd(a.acc + (up ? 0 : 1), tot);
  }
}
__global__ void l1_finalize_kernel(const double* acc, float* loss, double inv_nu, double inv_ns) {
  loss[0] = float(acc[0] * inv_nu + 0.25 * acc[1] * inv_ns);
}

// ---------------------------------------------------------------------------------------
// Latitude-weighted RMSE and ACC of the reference's evaluation loop (era5_data/score.py:92-105, 123-135 as called
// from models/pangu_sample.py:236-270), one value per plane p (65 upper-air (variable, level) planes, 4 surface):
//   w_j   = num_lat * cos(3.1416/180 * lat_j) / sum_j cos(3.1416/180 * lat_j),   lat_j = 90 - j * 180 / (num_lat - 1)
//   rmse  = sqrt(mean_{lat,lon} w (pred - tgt)^2)
//   acc   = sum w a b / sqrt(sum w a^2 * sum w b^2),   a = pred - mean_p, b = tgt - mean_p   (mean_p: the scalar
//           statistics mean of the plane, models/pangu_sample.py:250-254)
// pred may be the model's NORMALISED output (normalised != 0): pred_phys = out * std_p + mean_p is formed on the
// fly, so the scores need no separate normBackData pass.  Partial sums in fp64 (4 per plane).
struct ScoreArgs {
  const float* out_u; const float* out_s; const float* tgt_u; const float* tgt_s;
  const float* s_mean; const float* s_std; const float* u_mean; const float* u_std;
  const float* wlat;      // [lat] latitude weights
  double* acc;            // [69][4] zero-initialised: sum w d^2, sum w a b, sum w a^2, sum w b^2
  int lat, lon, normalised;
};
__global__ void __launch_bounds__(256) scores_kernel(const ScoreArgs a) {
  const int p = blockIdx.y;
  float m, s;
  const float *o, *t;
  const size_t plane = size_t(a.lat) * a.lon;
  if (p < 65) {
    const int c = p / 13, l = p % 13;
    m = a.u_mean[(12 - l) * 5 + c]; s = a.u_std[(12 - l) * 5 + c];
    o = a.out_u + size_t(p) * plane; t = a.tgt_u + size_t(p) * plane;
  } else {
    m = a.s_mean[p - 65]; s = a.s_std[p - 65];
    o = a.out_s + size_t(p - 65) * plane; t = a.tgt_s + size_t(p - 65) * plane;
  }
  const float ps = a.normalised ? s : 1.f, pm = a.normalised ? 0.f : m;   // anomaly a = out * std (normalised) or pred - mean
  float sd = 0.f, sab = 0.f, saa = 0.f, sbb = 0.f;
  const int lon4 = a.lon / 4;
  for (int row = blockIdx.x; row < a.lat; row += gridDim.x) {
    const float w = a.wlat[row];
    const float4* o4 = reinterpret_cast<const float4*>(o + size_t(row) * a.lon);
    const float4* t4 = reinterpret_cast<const float4*>(t + size_t(row) * a.lon);
    float rd = 0.f, rab = 0.f, raa = 0.f, rbb = 0.f;
    for (int i = threadIdx.x; i < lon4; i += blockDim.x) {
      const float4 ov = o4[i], tv = t4[i];
      const float av[4] = {ov.x * ps - pm, ov.y * ps - pm, ov.z * ps - pm, ov.w * ps - pm};
      const float bv[4] = {tv.x - m, tv.y - m, tv.z - m, tv.w - m};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float d = av[k] - bv[k];
        rd = fmaf(d, d, rd); rab = fmaf(av[k], bv[k], rab); raa = fmaf(av[k], av[k], raa); rbb = fmaf(bv[k], bv[k], rbb);
      }
    }
    sd = fmaf(w, rd, sd); sab = fmaf(w, rab, sab); saa = fmaf(w, raa, saa); sbb = fmaf(w, rbb, sbb);
  }
  __shared__ float part[8][4];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    sd += __shfl_xor_sync(0xffffffffu, sd, off); sab += __shfl_xor_sync(0xffffffffu, sab, off);
    saa += __shfl_xor_sync(0xffffffffu, saa, off); sbb += __shfl_xor_sync(0xffffffffu, sbb, off);
  }
  if ((threadIdx.x & 31) == 0) {
    part[threadIdx.x >> 5][0] = sd; part[threadIdx.x >> 5][1] = sab; part[threadIdx.x >> 5][2] = saa; part[threadIdx.x >> 5][3] = sbb;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double tot = 0.0;
    for (int k = 0; k < 8; ++k) tot += double(part[k][threadIdx.x]);
    atomicAdd(a.acc + p * 4 + threadIdx.x, tot);
  }
}
__global__ void scores_finalize_kernel(const double* acc, float* rmse, float* accs, double inv_n) {
  const int p = threadIdx.x;
  if (p >= 69) return;
  rmse[p] = float(sqrt(acc[p * 4] * inv_n));
  accs[p] = float(acc[p * 4 + 1] / sqrt(acc[p * 4 + 2] * acc[p * 4 + 3]));
}

// ---------------------------------------------------------------------------------------
// dst[r][0..kd) = cast(src[r][0..ks)), zero for columns ks..kd (K padding of conv_surface)
template <bool kFp16>
__global__ void cast16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int rows, int ks, int kd) {
  const size_t n = size_t(rows) * kd;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const int r = int(i / kd), c = int(i % kd);
    dst[i] = c < ks ? cvt16<kFp16>(src[size_t(r) * ks + c]) : uint16_t(0);
  }
}

}  // namespace pg
