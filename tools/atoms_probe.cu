// shared-memory atomic throughput probe: random vs conflict-free addressing, atomics vs plain RMW
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int WORDS = 16384;
template <int MODE>
__global__ void __launch_bounds__(512) k(uint32_t * out, int iters, uint32_t seed)
{
  extern __shared__ uint32_t cnt[];
  for (int i = threadIdx.x; i < WORDS; i += blockDim.x) cnt[i] = 0;
  __syncthreads();
  uint32_t x = seed * 2654435761u + threadIdx.x * 40503u + blockIdx.x * 9176u;
  int const lane = threadIdx.x & 31;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x = x * 1664525u + 1013904223u;
      uint32_t w = (x >> 10) & (WORDS - 1);
      if (MODE == 1 || MODE == 3) w = (w & ~31u) | lane;          // one lane per bank: conflict free
      uint32_t const inc = (x & 1) ? 0x10000u : 1u;
      if (MODE <= 1) atomicAdd(&cnt[w], inc);
      else { cnt[w] += inc; }                                       // plain RMW (racy, throughput only)
    }
  }
  __syncthreads();
  uint32_t s = 0;
  for (int i = threadIdx.x; i < WORDS; i += blockDim.x) s += cnt[i];
  if (s == 0x12345) out[0] = s;
}
template <int MODE> void run(const char * name)
{
  uint32_t * d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, WORDS * 4);
  int const blocks = 148 * 2, iters = 2000;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(a);
    k<MODE><<<blocks, 512, WORDS * 4>>>(d, iters, 7 + rep);
    cudaEventRecord(b); cudaEventSynchronize(b);
  }
  float ms; cudaEventElapsedTime(&ms, a, b);
  double const ops = double(blocks) * 512 * iters * 8;
  printf("%-28s %.3f ms  %.1f G lane-ops/s  = %.2f warp-instr/clk/SM\n", name, ms, ops / ms / 1e6,
         ops / 32 / (ms * 1e-3) / 148 / 1.965e9);
}
int main()
{
  run<0>("ATOMS random");
  run<1>("ATOMS conflict-free");
  run<2>("LDS+STS random");
  run<3>("LDS+STS conflict-free");
  return 0;
}
