// Memory-bound kernels of the backward pass (reference: autograd through models/layers.py,
// driven by models/pangu_sample.py:57-71):
//   * ln_bwd_kernel     LayerNorm backward (dx, d gamma, d beta) with the row maps of the plain block
//                       LayerNorms, UpSample.norm (pixel-shuffle rows) and DownSample.norm (2x2 merge rows)
//   * gelu_bwd_kernel   d pre = d hidden * gelu'(pre)       (exact erf GELU, models/layers.py:261)
//   * colsum16_kernel   bias gradients: column sums of a 16-bit [M, N] gradient matrix
//   * cast16_t_kernel   fp32 [R, C] -> 16-bit [C, R_pad] transposed weight copies (B operands of the dgrad GEMMs)
#pragma once
#include "common.cuh"
#include "geometry.cuh"

namespace pg {

enum LnBwdMode { LNB_IDENT = 0, LNB_UP = 1, LNB_DOWN = 2 };

struct LnBwdArgs {
  const float* y;       // pre-LayerNorm values.  IDENT: [rows, C]; UP: upsample.linear1 output [T2, 4C]; DOWN: x_hi [T, C/4]
  const float* g;       // gradient w.r.t. the LayerNorm output, [rows, C] fp32
  const float* gamma;   // [C]
  void* dx16;           // IDENT: [rows, C] 16-bit; UP: [T2, 4C] 16-bit (cropped positions are never written)
  float* dx32;          // DOWN: [T, C/4] fp32, read-modify-write (+=)
  float* dgamma;        // [C] += scale * sum_rows g * xhat
  float* dbeta;         // [C] += scale * sum_rows g
  float* dbias;         // [C] += sum_rows dx: bias gradient of the linear that produced y (IDENT mode; nullable)
  int rows, C;
  int Z, H, W;          // UP / DOWN: HIGH-resolution token grid
  float scale;          // DropPath factor of the branch (1 in eval)
  float palpha;         // factor on the parameter gradients (1 / loss scale)
  float eps;
};

// One warp per row, grid-stride over rows; lane owns float4 pieces q = lane + 32*i of the row.
// dx = rstd * (gh - mean(gh) - xhat * mean(gh * xhat)),  gh = scale * g * gamma,  xhat = (y - mean) * rstd
template <bool kFp16, int kMode, int kC>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const LnBwdArgs a) {
  constexpr int NP = (kC / 4 + 31) / 32;           // float4 pieces per lane
  __shared__ float s_dg[kC], s_db[kC], s_dy[kC];
  for (int i = threadIdx.x; i < kC; i += blockDim.x) { s_dg[i] = 0.f; s_db[i] = 0.f; s_dy[i] = 0.f; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  float4 gam[NP], adg[NP], adb[NP], ady[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int q = lane + 32 * i;
    gam[i] = q < kC / 4 ? *reinterpret_cast<const float4*>(a.gamma + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    adg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    adb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ady[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int W2 = a.W / 2, H2 = (a.H + 1) / 2;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < a.rows; row += warps_total) {
    // ---- addressing of this row
    const float* ysrc = nullptr;
    size_t out_off = 0;          // element offset of the row in dx16 (IDENT / UP)
    int z = 0, h2 = 0, w2 = 0;
    if constexpr (kMode == LNB_IDENT) {
      ysrc = a.y + size_t(row) * kC;
      out_off = size_t(row) * kC;
    } else if constexpr (kMode == LNB_UP) {
      const int w = row % a.W, h = (row / a.W) % a.H, zz = row / (a.W * a.H);
      const size_t tok2 = size_t(zz * H2 + (h >> 1)) * W2 + (w >> 1);
      const int grp = (h & 1) * 2 + (w & 1);
      out_off = tok2 * (4 * kC) + grp * kC;
      ysrc = a.y + out_off;
    } else {
      w2 = row % W2; h2 = (row / W2) % H2; z = row / (W2 * H2);
    }
    float4 yv[NP], gv[NP];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int q = lane + 32 * i;
      yv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      gv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < kC / 4) {
        if constexpr (kMode == LNB_DOWN) {
          constexpr int QH = kC / 8;               // float4 pieces per dh half-row (two adjacent tokens)
          const int dh = q / QH, r = q % QH;
          const int h = 2 * h2 + dh;
          if (h < a.H) yv[i] = *reinterpret_cast<const float4*>(a.y + (size_t(z * a.H + h) * a.W + 2 * w2) * (kC / 4) + r * 4);
        } else {
          yv[i] = *reinterpret_cast<const float4*>(ysrc + q * 4);
        }
        gv[i] = *reinterpret_cast<const float4*>(a.g + size_t(row) * kC + q * 4);
      }
      s += (yv[i].x + yv[i].y) + (yv[i].z + yv[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / kC);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int q = lane + 32 * i;
      if (q < kC / 4) {
        yv[i].x -= mean; yv[i].y -= mean; yv[i].z -= mean; yv[i].w -= mean;
        ss += (yv[i].x * yv[i].x + yv[i].y * yv[i].y) + (yv[i].z * yv[i].z + yv[i].w * yv[i].w);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss * (1.0f / kC) + a.eps);
    float m1 = 0.f, m2 = 0.f;      // sum gh, sum gh * xhat
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      // xhat in yv, scaled upstream gradient in gv, gh = g * gamma
      yv[i].x *= rstd; yv[i].y *= rstd; yv[i].z *= rstd; yv[i].w *= rstd;
      gv[i].x *= a.scale; gv[i].y *= a.scale; gv[i].z *= a.scale; gv[i].w *= a.scale;
      adg[i].x += gv[i].x * yv[i].x; adg[i].y += gv[i].y * yv[i].y; adg[i].z += gv[i].z * yv[i].z; adg[i].w += gv[i].w * yv[i].w;
      adb[i].x += gv[i].x; adb[i].y += gv[i].y; adb[i].z += gv[i].z; adb[i].w += gv[i].w;
      gv[i].x *= gam[i].x; gv[i].y *= gam[i].y; gv[i].z *= gam[i].z; gv[i].w *= gam[i].w;
      m1 += (gv[i].x + gv[i].y) + (gv[i].z + gv[i].w);
      m2 += (gv[i].x * yv[i].x + gv[i].y * yv[i].y) + (gv[i].z * yv[i].z + gv[i].w * yv[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m1 += __shfl_xor_sync(0xffffffffu, m1, o);
      m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    }
    m1 *= (1.0f / kC); m2 *= (1.0f / kC);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int q = lane + 32 * i;
      if (q >= kC / 4) continue;
      const float d0 = rstd * (gv[i].x - m1 - yv[i].x * m2), d1 = rstd * (gv[i].y - m1 - yv[i].y * m2);
      const float d2 = rstd * (gv[i].z - m1 - yv[i].z * m2), d3 = rstd * (gv[i].w - m1 - yv[i].w * m2);
      ady[i].x += d0; ady[i].y += d1; ady[i].z += d2; ady[i].w += d3;
      if constexpr (kMode == LNB_DOWN) {
        constexpr int QH = kC / 8;
        const int dh = q / QH, r = q % QH;
        const int h = 2 * h2 + dh;
        if (h < a.H) {
          float4* p = reinterpret_cast<float4*>(a.dx32 + (size_t(z * a.H + h) * a.W + 2 * w2) * (kC / 4) + r * 4);
          float4 o = *p;
          o.x += d0; o.y += d1; o.z += d2; o.w += d3;
          *p = o;
        }
      } else {
        reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(a.dx16) + out_off)[q] =
            make_uint2(pack16<kFp16>(d0, d1), pack16<kFp16>(d2, d3));
      }
    }
  }
  // ---- parameter gradients: lanes of every warp own the same columns -> shared atomics, then one global atomic per column
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int q = lane + 32 * i;
    if (q >= kC / 4) continue;
    atomicAdd(&s_dg[q * 4 + 0], adg[i].x); atomicAdd(&s_dg[q * 4 + 1], adg[i].y);
    atomicAdd(&s_dg[q * 4 + 2], adg[i].z); atomicAdd(&s_dg[q * 4 + 3], adg[i].w);
    atomicAdd(&s_db[q * 4 + 0], adb[i].x); atomicAdd(&s_db[q * 4 + 1], adb[i].y);
    atomicAdd(&s_db[q * 4 + 2], adb[i].z); atomicAdd(&s_db[q * 4 + 3], adb[i].w);
    if (a.dbias) {
      atomicAdd(&s_dy[q * 4 + 0], ady[i].x); atomicAdd(&s_dy[q * 4 + 1], ady[i].y);
      atomicAdd(&s_dy[q * 4 + 2], ady[i].z); atomicAdd(&s_dy[q * 4 + 3], ady[i].w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kC; i += blockDim.x) {
    if (a.dgamma) atomicAdd(a.dgamma + i, a.palpha * s_dg[i]);
    if (a.dbeta) atomicAdd(a.dbeta + i, a.palpha * s_db[i]);
    if (a.dbias) atomicAdd(a.dbias + i, a.palpha * s_dy[i]);
  }
}

// ---------------------------------------------------------------------------------------
// d pre = d hidden * gelu'(pre),  gelu'(x) = Phi(x) + x phi(x)   (in place on dh); 8 elements per thread.
// Phi through the same degree-7 fit of log2(erfc(t))/t as the forward GELU (common.cuh): erfc(t) = 2^(t P7(t)),
// Phi(x) = 1 - erfc(|x|/sqrt 2)/2 for x >= 0, erfc(|x|/sqrt 2)/2 for x < 0; phi through one more ex2.approx.
// max |error| of gelu' ~ 6e-7, far below the 16-bit rounding of the result.
__device__ __forceinline__ float gelu_grad(float x) {
  const float ax = fabsf(x);
  const float t = fminf(ax * 0.70710678118654752440f, 4.0f);
  float p = -5.904116739e-06f;
  p = fmaf(p, t, 6.987359289e-05f);
  p = fmaf(p, t, -6.779016748e-05f);
  p = fmaf(p, t, -3.477876114e-03f);
  p = fmaf(p, t, 3.092580434e-02f);
  p = fmaf(p, t, -1.497507845e-01f);
  p = fmaf(p, t, -9.181910519e-01f);
  p = fmaf(p, t, -1.627914489e+00f);
  const float half_erfc = 0.5f * ex2_approx(p * t);
  const float cdf = x >= 0.f ? 1.0f - half_erfc : half_erfc;
  const float pdf = 0.3989422804014327f * ex2_approx(-0.72134752044448170368f * x * x);   // exp(-x^2/2) = 2^(-x^2 log2(e)/2)
  return fmaf(x, pdf, cdf);
}

// dh, pre: [M, N] 16-bit.  Thread = 8 consecutive columns of a row (fixed column group, rows strided), so that the
// bias gradient of linear1, db[n] += alpha * sum_m dpre[m, n], falls out of the same pass (nullable).
template <bool kFp16>
__global__ void __launch_bounds__(256) gelu_bwd_kernel(uint16_t* __restrict__ dh, const uint16_t* __restrict__ pre, int M, int N,
                                                       float* __restrict__ db, float alpha) {
  extern __shared__ float s_acc[];               // [N]
  if (db) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
  }
  const int tpr = N / 8;                         // threads per row
  const int rpi = blockDim.x / tpr;              // rows per iteration of this block
  const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (rl < rpi) {
    const int step = gridDim.x * rpi;
    for (int m = blockIdx.x * rpi + rl; m < M; m += 2 * step) {
      uint4 d[2], p[2];
      bool ok[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int mm = m + u * step;
        ok[u] = mm < M;
        if (ok[u]) {
          d[u] = ldg16(dh + size_t(mm) * N + cg * 8);         // coherent: dh is rewritten in place
          p[u] = ldg_nc16(pre + size_t(mm) * N + cg * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (!ok[u]) continue;
        uint32_t* dw = reinterpret_cast<uint32_t*>(&d[u]);
        const uint32_t* pw = reinterpret_cast<const uint32_t*>(&p[u]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float x0 = unpack16_lo<kFp16>(pw[k]), x1 = unpack16_hi<kFp16>(pw[k]);
          const float g0 = unpack16_lo<kFp16>(dw[k]) * gelu_grad(x0), g1 = unpack16_hi<kFp16>(dw[k]) * gelu_grad(x1);
          acc[2 * k] += g0; acc[2 * k + 1] += g1;
          dw[k] = pack16<kFp16>(g0, g1);
        }
        stg16(dh + size_t(m + u * step) * N + cg * 8, d[u]);
      }
    }
    if (db) {
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(&s_acc[cg * 8 + k], acc[k]);
    }
  }
  if (db) {
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) atomicAdd(db + i, alpha * s_acc[i]);
  }
}

// ---------------------------------------------------------------------------------------
// out[n] += alpha * sum_m src[m][n] for n < n_valid.  src: 16-bit [M, ld]; N = columns scanned (multiple of 8).
template <bool kFp16>
__global__ void __launch_bounds__(256) colsum16_kernel(const uint16_t* __restrict__ src, float* __restrict__ out, int M, int N,
                                                       int ld, int n_valid, float alpha) {
  extern __shared__ float s_acc[];               // [N]
  for (int i = threadIdx.x; i < N; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int tpr = N / 8;                         // threads per row
  const int rpi = blockDim.x / tpr;              // rows per iteration of this block
  const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (rl < rpi) {
    const int step = gridDim.x * rpi;
    for (int m = blockIdx.x * rpi + rl; m < M; m += 4 * step) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {          // four independent 16 B loads in flight per thread
        const int mm = m + u * step;
        v[u] = mm < M ? ldg_nc16(src + size_t(mm) * ld + cg * 8) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(&v[u]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          acc[2 * k] += unpack16_lo<kFp16>(w[k]);
          acc[2 * k + 1] += unpack16_hi<kFp16>(w[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&s_acc[cg * 8 + k], acc[k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_valid; i += blockDim.x) atomicAdd(out + i, alpha * s_acc[i]);
}

// ---------------------------------------------------------------------------------------
// dst[c][r] = cast(src[r][c]) for r < R, c < C; dst is [C_pad rows][R_pad] with zero padding.
template <bool kFp16>
__global__ void __launch_bounds__(256) cast16_t_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int R, int C,
                                                       int R_pad, int C_pad) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    const int r = r0 + k, c = c0 + tx;
    tile[k][tx] = (r < R && c < C) ? src[size_t(r) * C + c] : 0.f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k, r = r0 + tx;
    if (c < C_pad && r < R_pad) dst[size_t(c) * R_pad + r] = cvt16<kFp16>(tile[tx][k]);
  }
}

}  // namespace pg
