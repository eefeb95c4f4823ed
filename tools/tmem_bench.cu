// Micro-benchmarks that size the softmax / LayerNorm epilogues: tcgen05.ld throughput per SM
// as a function of the number of reading warps, MUFU ex2 rate, packed-FMA rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_bench tools/tmem_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../pangu_pytorch_b200/csrc/common.cuh"
using namespace pg;

__device__ __forceinline__ void tmem_ld8b(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
        "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
        "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
        "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
        "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr)
      : "memory");
}

// mode 0: x32 loads, wait after each; 1: x32 loads, 4 in flight; 2: x16, 2: x64; 3: x8
template <int MODE>
__global__ void ld_kernel(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if constexpr (MODE == 0) {
    for (int i = 0; i < iters; ++i) {
      uint32_t r[32];
      tmem_ld32(base + ((i * 32) & 255), r);
      tmem_ld_wait();
      acc += r[0] ^ r[31];
    }
  } else if constexpr (MODE == 1) {
    for (int i = 0; i < iters; i += 4) {
      uint32_t r0[32], r1[32], r2[32], r3[32];
      tmem_ld32(base + 0, r0);
      tmem_ld32(base + 32, r1);
      tmem_ld32(base + 64, r2);
      tmem_ld32(base + 96, r3);
      tmem_ld_wait();
      acc += r0[0] ^ r1[31] ^ r2[5] ^ r3[7];
    }
  } else if constexpr (MODE == 2) {
    for (int i = 0; i < iters; i += 2) {
      uint32_t r0[64], r1[64];
      tmem_ld64(base + 0, r0);
      tmem_ld64(base + 64, r1);
      tmem_ld_wait();
      acc += r0[0] ^ r1[63];
    }
  } else {
    for (int i = 0; i < iters; i += 4) {
      uint32_t r0[8], r1[8], r2[8], r3[8];
      tmem_ld8b(base + 0, r0);
      tmem_ld8b(base + 8, r1);
      tmem_ld8b(base + 16, r2);
      tmem_ld8b(base + 24, r3);
      tmem_ld_wait();
      acc += r0[0] ^ r1[7] ^ r2[5] ^ r3[7];
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x % 32 == 0) out[blockIdx.x * 32 + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(slot); }
}

// MUFU ex2 throughput, optionally mixed with FFMA2
template <int MODE>
__global__ void ex2_kernel(int iters, long long* out, float* sink) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 0.1f, x2 = x0 + 0.2f, x3 = x0 + 0.3f;
  float x4 = x0 + .4f, x5 = x0 + .5f, x6 = x0 + .6f, x7 = x0 + .7f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    x0 = ex2_approx(x0); x1 = ex2_approx(x1); x2 = ex2_approx(x2); x3 = ex2_approx(x3);
    x4 = ex2_approx(x4); x5 = ex2_approx(x5); x6 = ex2_approx(x6); x7 = ex2_approx(x7);
    if constexpr (MODE == 1) {
      x0 = fmaf(x0, -1.f, 0.5f); x1 = fmaf(x1, -1.f, .5f); x2 = fmaf(x2, -1.f, .5f); x3 = fmaf(x3, -1.f, .5f);
      x4 = fmaf(x4, -1.f, 0.5f); x5 = fmaf(x5, -1.f, .5f); x6 = fmaf(x6, -1.f, .5f); x7 = fmaf(x7, -1.f, .5f);
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x % 32 == 0) out[blockIdx.x * 32 + (threadIdx.x >> 5)] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int main() {
  long long* d_out; uint32_t* d_sink;
  cudaMalloc(&d_out, 148 * 32 * 8);
  cudaMalloc(&d_sink, 148 * 1024 * 4);
  long long h[32];
  const int iters = 4096;
  auto report = [&](const char* name, int warps, double bytes_per_iter_per_warp) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
    printf("%-28s warps=%2d cycles=%8lld  -> %.1f B/clk/SM (%.1f B/clk/warp)\n", name, warps, mx,
           bytes_per_iter_per_warp * iters * warps / mx, bytes_per_iter_per_warp * iters / mx);
  };
  for (int warps : {1, 2, 4, 8, 16}) {
    ld_kernel<0><<<1, warps * 32>>>(iters, d_out, d_sink); report("ld x32 wait-each", warps, 4096);
    ld_kernel<1><<<1, warps * 32>>>(iters, d_out, d_sink); report("ld x32 4-in-flight", warps, 4096);
    ld_kernel<2><<<1, warps * 32>>>(iters, d_out, d_sink); report("ld x64 2-in-flight", warps, 8192);
    ld_kernel<3><<<1, warps * 32>>>(iters, d_out, d_sink); report("ld x8 4-in-flight", warps, 1024);
  }
  // all SMs at once (one CTA per SM), 8 warps
  ld_kernel<1><<<148, 256>>>(iters, d_out, d_sink); report("ld x32 4-in-flight 148 CTAs", 8, 4096);
  for (int warps : {4, 8, 16, 32}) {
    ex2_kernel<0><<<1, warps * 32>>>(iters, d_out, (float*)d_sink);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
    printf("ex2 only        warps=%2d cycles=%8lld -> %.2f ex2/clk/SM\n", warps, mx, 8.0 * 32 * iters * warps / mx);
    ex2_kernel<1><<<1, warps * 32>>>(iters, d_out, (float*)d_sink);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    mx = 0; for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
    printf("ex2 + fma       warps=%2d cycles=%8lld -> %.2f ex2/clk/SM\n", warps, mx, 8.0 * 32 * iters * warps / mx);
  }
  return 0;
}
