// Cost of waiting on an mbarrier whose phase has ALREADY completed, per wait, for the ways a warp can do it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mbar_bench tools/mbar_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
template <int MODE>
__global__ void k(long long* out, int iters, int active_warps) {
  __shared__ uint64_t bar[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp >= active_warps) return;
  // a fresh barrier is in phase 0: waiting for parity 1 ("the phase before") succeeds immediately
  uint64_t* b = &bar[warp & 7];
  unsigned acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) { while (!try_wait(b, 1)) {} }                                             // every lane
    if (MODE == 1) { if (lane == 0) { while (!try_wait(b, 1)) {} } __syncwarp(); }            // one lane + warp sync
    if (MODE == 2) { while (!test_wait(b, 1)) {} }                                            // every lane, test_wait
    if (MODE == 3) { if (lane == 0) { while (!test_wait(b, 1)) {} } __syncwarp(); }
    if (MODE == 4) { bool ok = false; while (!ok) ok = __any_sync(0xffffffffu, lane == 0 ? try_wait(b, 1) : false); }
    acc += i;
  }
  long long t1 = clock64();
  if (lane == 0) out[warp] = (t1 - t0) / iters + (acc == 12345u);
}
template <int MODE>
void run(const char* name) {
  long long* out; cudaMalloc(&out, 64 * 8);
  for (int w : {1, 4, 8, 16}) {
    k<MODE><<<148, 512>>>(out, 2000, w);
    k<MODE><<<148, 512>>>(out, 2000, w);
    long long h[16]; cudaMemcpy(h, out, 16 * 8, cudaMemcpyDeviceToHost);
    printf("%-34s %2d waiting warps: %4lld clk per wait\n", name, w, h[0]);
  }
  cudaFree(out);
}
int main() {
  run<0>("try_wait, all lanes");
  run<1>("try_wait, lane 0 + __syncwarp");
  run<2>("test_wait, all lanes");
  run<3>("test_wait, lane 0 + __syncwarp");
  run<4>("try_wait, lane 0 + vote");
  return 0;
}
