// prints the observed semantics of the DPX/SIMD intrinsics on this device
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ uint32_t pk2(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
__global__ void k(const int * in, int * out)
{
  // runtime inputs so nothing is constant-folded
  uint32_t a = pk2(in[0], in[1]), b = pk2(in[2], in[3]);
  bool ph, pl;
  uint32_t m = __vibmax_s16x2(a, b, &ph, &pl);
  out[0] = (int)(int16_t)(m & 0xffff); out[1] = (int)(int16_t)(m >> 16); out[2] = pl; out[3] = ph;
  uint32_t m2 = __vibmin_s16x2(a, b, &ph, &pl);
  out[4] = (int)(int16_t)(m2 & 0xffff); out[5] = (int)(int16_t)(m2 >> 16); out[6] = pl; out[7] = ph;
}
int main()
{
  int cases[][4] = {{5, 9, 7, 9}, {-3, -10, -4, 2}, {7, 1, 5, 1}, {0, 0, 0, 0}, {100, -100, -100, 100}};
  int *din, *dout; cudaMalloc(&din, 16); cudaMalloc(&dout, 32);
  for (auto & c : cases) {
    cudaMemcpy(din, c, 16, cudaMemcpyHostToDevice);
    k<<<1, 1>>>(din, dout);
    int o[8]; cudaMemcpy(o, dout, 32, cudaMemcpyDeviceToHost);
    printf("a=(lo %d,hi %d) b=(lo %d,hi %d): max=(%d,%d) pl=%d ph=%d | min=(%d,%d) pl=%d ph=%d\n",
           c[0], c[1], c[2], c[3], o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
  }
  return 0;
}
