// Issue-rate microbenchmark for the GELU inner loop: scalar FFMA (immediate addend) vs packed FFMA2 vs MUFU.EX2,
// N independent chains per thread, W warps per CTA, one CTA per SM.   nvcc -arch=sm_100a -O3 -o tools/bin/fma_bench tools/fma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE, int CH>
__global__ void k(float* out, long long* clk, int iters, float y) {
  float x[CH]; f32x2 p[CH];
  for (int i = 0; i < CH; ++i) { x[i] = threadIdx.x * 0.001f + i; p[i] = pack2(x[i], x[i] + 1.f); }
  const f32x2 y2 = pack2(y, y);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (MODE == 0) x[i] = fmaf(x[i], y, 0.25f);                       // FFMA, immediate addend
      if (MODE == 1) p[i] = fma2(p[i], y2, pack2(0.25f, 0.25f));         // FFMA2, constant addend
      if (MODE == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (MODE == 3) x[i] = fmaf(x[i], y, x[(i + 1) % CH]);              // FFMA, 3 registers
      if (MODE == 4) p[i] = fma2(p[i], y2, p[(i + 1) % CH]);             // FFMA2, 3 register pairs
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < CH; ++i) { float a, b; unpack2(p[i], a, b); s += x[i] + a + b; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int MODE>
void run(const char* name, int warps) {
  float* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  const int iters = 2000; constexpr int CH = 8;
  k<MODE, CH><<<148, warps * 32>>>(out, clk, iters, 0.999f);
  k<MODE, CH><<<148, warps * 32>>>(out, clk, iters, 0.999f);
  long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  const double per_smsp = double(h) / (double(iters) * CH * (warps / 4.0 < 1 ? 1 : warps / 4.0));
  printf("%-28s warps/CTA %2d: %.2f clk per warp-instruction per scheduler\n", name, warps, per_smsp);
  cudaFree(out); cudaFree(clk);
}
int main() {
  for (int w : {4, 8, 16}) {
    run<0>("FFMA imm", w); run<3>("FFMA 3-reg", w); run<1>("FFMA2 const", w); run<4>("FFMA2 3-reg", w); run<2>("MUFU.EX2", w);
  }
  return 0;
}
