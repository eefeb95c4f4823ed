// Micro-benchmark: tcgen05.mma (kind::f16, M=128, K=16) issue rate per SM as a function of N, for
// A from shared memory (SS) and A from tensor memory (TS).  One CTA per SM, one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_bench tools/mma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../pangu_pytorch_b200/csrc/common.cuh"
using namespace pg;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(d), "r"(a), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

template <int N, int NACC, bool TS>
__global__ void __launch_bounds__(128, 1) bench(int reps, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (slot != 0u) __trap();
  constexpr uint32_t tmem = 0u;                  // whole TMEM: compile-time addresses keep the operands in uniform registers
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = make_idesc_f16(128, N, false);
    const uint64_t da = make_sdesc_sw128(smem_u32(smem));
    const uint64_t db = make_sdesc_sw128(smem_u32(smem) + 16384);
    long long t0 = 0;
    for (int pass = 0; pass < 2; ++pass) {       // pass 0 = warm-up
      t0 = clock64();
      for (int i = 0; i < reps / 8; ++i) {
        if (elect_one()) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            constexpr uint32_t dummy = 0;
            const uint32_t d = tmem + uint32_t(u % NACC) * uint32_t(N) + dummy;
            if (TS) umma_ts(d, tmem + 448 + 8 * (u & 3), db + uint64_t((u & 3) * 2), idesc, 1);
            else umma_f16_ss(d, da + uint64_t((u & 3) * 2), db + uint64_t((u & 3) * 2), idesc, 1);
          }
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, pass);
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

// MODE 0: converged warp + elect, RUNTIME ring index in the B descriptor and the accumulator address (values are
//         warp-uniform but live in ordinary registers); MODE 1: one divergent lane (if lane == 0), compile-time operands;
// MODE 2: converged + elect, runtime index, but the 4 MMAs of a k-block share one descriptor base (+ constant offsets)
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) bench_rt(int reps, int stages, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = make_idesc_f16(128, N, false);
    const uint32_t base = smem_u32(smem);
    long long t0 = 0;
    for (int pass = 0; pass < 2; ++pass) {
      t0 = clock64();
      if (MODE == 1) {
        if (threadIdx.x == 0) {
          const uint64_t da = make_sdesc_sw128(base), db = make_sdesc_sw128(base + 16384);
          for (int i = 0; i < reps / 4; ++i) {
#pragma unroll
            for (int u = 0; u < 4; ++u) umma_f16_ss(tmem, da + uint64_t(u * 2), db + uint64_t(u * 2), idesc, 1);
          }
          umma_commit(&bar);
        }
        __syncwarp();
      } else {
        int st = 0;
        for (int i = 0; i < reps / 4; ++i) {
          const uint64_t da = make_sdesc_sw128(base + st * 49152);
          const uint64_t db = make_sdesc_sw128(base + st * 49152 + 16384);
          const uint32_t d = tmem + uint32_t(i & 1) * uint32_t(N);
          if (elect_one()) {
#pragma unroll
            for (int u = 0; u < 4; ++u) umma_f16_ss(d, da + uint64_t(u * 2), db + uint64_t(u * 2), idesc, 1);
          }
          __syncwarp();
          if (++st == stages) st = 0;
        }
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
      }
      mbar_wait(&bar, pass);
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}
template <int N, int MODE>
void run_rt(long long* out) {
  const int reps = 4096;
  cudaFuncSetAttribute(bench_rt<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  bench_rt<N, MODE><<<148, 128, 160 * 1024>>>(reps, 3, out);
  long long c = 0;
  cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
  const double per = double(c) / reps;
  printf("SS N=%3d %s: %6.1f clk per MMA -> %5.0f FLOP/clk/SM (floor %3d)\n", N,
         MODE == 1 ? "single divergent lane, constant operands" : "converged+elect, RUNTIME stage index / accumulator", per,
         2.0 * 128 * N * 16 / per, N / 2);
}

template <int N, int NACC, bool TS>
void run(long long* out) {
  const int reps = 4096;
  cudaFuncSetAttribute(bench<N, NACC, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  bench<N, NACC, TS><<<148, 128, 100 * 1024>>>(reps, out);
  long long c = 0;
  cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  const double per = double(c) / reps;
  printf("%s N=%3d nacc=%d: %6.1f clk per MMA (M=128,K=16) -> %5.0f FLOP/clk/SM (floor N/2 = %3d clk)  %s\n", TS ? "TS" : "SS", N, NACC,
         per, 2.0 * 128 * N * 16 / per, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* out;
  cudaMalloc(&out, 8);
  run<32, 1, false>(out); run<32, 4, false>(out);
  run<64, 1, false>(out); run<64, 2, false>(out); run<64, 4, false>(out);
  run<128, 1, false>(out); run<128, 2, false>(out);
  run<144, 1, false>(out); run<192, 1, false>(out); run<192, 2, false>(out); run<256, 1, false>(out);
  run<32, 1, true>(out); run<32, 4, true>(out);
  run<64, 1, true>(out); run<64, 4, true>(out);
  run<128, 1, true>(out); run<192, 1, true>(out); run<192, 2, true>(out); run<256, 1, true>(out);
  run_rt<64, 0>(out); run_rt<192, 0>(out); run_rt<256, 0>(out);
  run_rt<64, 1>(out); run_rt<192, 1>(out); run_rt<256, 1>(out);
  return 0;
}
