// pipe_probe.cu — issue-rate probe for the packed 16x2 integer instructions the aligner is made of
// (sm_100a).  For every candidate instruction (or mix) NCH independent dependency chains per thread
// run ITER times; the result is warp-instructions per clock per SM sub-partition (SMSP).  A second
// pass with a single chain gives the dependent-issue latency.  Used to decide which pipe (ALU / FMA)
// each operation of the recurrence should be issued on; see DESIGN.md.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/pipe_probe.cu -o build/pipe_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int ITER = 4096;

template <int OP>
__device__ __forceinline__ uint32_t op1(uint32_t a, uint32_t b, uint32_t c, uint32_t one)
{
  // b and c are OTHER chains' values of the previous iteration: nothing folds algebraically
  if (OP == 0) { return __vadd2(a, b); }                           // VIADD.16x2
  if (OP == 1) { return a + b; }                                   // IADD3 / IMAD.IADD (ptxas's choice)
  if (OP == 3) { return __vimax3_u16x2(a, b, c); }                 // VIMNMX3.U16x2
  if (OP == 4) { return __viaddmax_u16x2(a, b, c); }               // VIADDMNMX.U16x2
  if (OP == 5) { uint32_t d; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b)); return d; }  // IMAD
  if (OP == 6) { bool p, q; uint32_t m = __vibmax_u16x2(a, b, &p, &q); if (!p) { m += 1u; } if (!q) { m += 0x10000u; } return m; }
  if (OP == 7) { return static_cast<uint32_t>(max(static_cast<int>(a ^ c), static_cast<int>(b))); }  // LOP3 + IMNMX
  if (OP == 8) { return (a & b) ^ c; }                             // LOP3
  return a;
}

template <int OP, int NCH>
__global__ void probe_kernel(uint32_t * out, uint32_t seed, uint32_t one, long long * clocks)
{
  uint32_t a[NCH + 2];
#pragma unroll
  for (int k = 0; k < NCH + 2; k++) { a[k] = seed * (threadIdx.x + 1) + k * 0x00030005u; }
  long long const t0 = clock64();
  for (int it = 0; it < ITER; it++) {
    if (NCH > 1) {
      uint32_t n[NCH];
#pragma unroll
      for (int k = 0; k < NCH; k++) { n[k] = op1<OP>(a[k], a[(k + 1) % NCH], a[(k + 2) % NCH], one); }
#pragma unroll
      for (int k = 0; k < NCH; k++) { a[k] = n[k]; }
    } else {
      a[0] = op1<OP>(a[0], a[1], a[2], one);   // one dependent chain
    }
  }
  long long const t1 = clock64();
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k < NCH; k++) { r ^= a[k]; }
  if (r == 0x12345678u) { out[0] = r; }
  if (threadIdx.x == 0 && blockIdx.x == 0) { clocks[0] = t1 - t0; }
}

// the flag-free cell update in the two candidate formulations, 4 independent "rows" per thread:
//   MIX 0: everything on packed DPX instructions            (VIADD.16x2 + VIMNMX3 + 2 x (VIADD.16x2 + VIADDMNMX))
//   MIX 1: the three subtracts as IMAD (FMA pipe), the three max operations as DPX (ALU pipe)
template <int MIX>
__global__ void cell_kernel(uint32_t * out, uint32_t seed, uint32_t one, long long * clocks)
{
  constexpr int NR = 4;
  uint32_t H[NR], E[NR], F[NR];
#pragma unroll
  for (int k = 0; k < NR; k++) { H[k] = 0x80008000u + seed * (threadIdx.x + 1 + k); E[k] = H[k] - 0x00140014u; F[k] = H[k] - 0x00130013u; }
  uint32_t const S = 0x00020006u & seed, nS = 0u - S;
  uint32_t const QR = 0x00140014u, R = 0x00020002u;
  uint32_t const nQR = __vneg2(QR), nR = __vneg2(R);         // per-half negation (for the fused add+max)
  uint32_t const mQR = 0u - QR, mR = 0u - R;                 // 32-bit negation (for IMAD: a*1 + (-C))
  long long const t0 = clock64();
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int k = 0; k < NR; k++) {
      if (MIX == 0) {
        uint32_t const t = __vadd2(H[k], nS);
        uint32_t const h = __vimax3_u16x2(t, F[k], E[k]);
        F[k] = __viaddmax_u16x2(h, nQR, __vadd2(F[k], nR));
        E[k] = __viaddmax_u16x2(h, nQR, __vadd2(E[k], nR));
        H[k] = h;
      } else {
        uint32_t t, f, e;
        asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(H[k]), "r"(one), "r"(nS));
        uint32_t const h = __vimax3_u16x2(t, F[k], E[k]);
        asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(f) : "r"(F[k]), "r"(one), "r"(mR));
        asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(e) : "r"(E[k]), "r"(one), "r"(mR));
        F[k] = __viaddmax_u16x2(h, nQR, f);
        E[k] = __viaddmax_u16x2(h, nQR, e);
        H[k] = h;
        (void)mQR;
      }
    }
  }
  long long const t1 = clock64();
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k < NR; k++) { r ^= H[k] ^ E[k] ^ F[k]; }
  if (r == 0x12345678u) { out[0] = r; }
  if (threadIdx.x == 0 && blockIdx.x == 0) { clocks[0] = t1 - t0; }
}

template <class K>
static int run(const char * name, K kernel, int ops_per_iter, int warps_per_smsp, uint32_t * d_out, long long * d_clk)
{
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int const threads = 128 * warps_per_smsp;   // 4 SMSPs x warps_per_smsp warps
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  long long clk = 0;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    kernel<<<sms, threads>>>(d_out, 3u + rep, 1u, d_clk);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) { best = ms; CK(cudaMemcpy(&clk, d_clk, 8, cudaMemcpyDeviceToHost)); }
  }
  double const winstr = static_cast<double>(ops_per_iter) * ITER * warps_per_smsp;  // per SMSP
  std::printf("%-44s %2d warps/SMSP: %6.3f warp-instr/clk/SMSP  (%.1f clk per instr per warp)\n", name, warps_per_smsp,
              winstr / static_cast<double>(clk), static_cast<double>(clk) / (static_cast<double>(ops_per_iter) * ITER));
  return 0;
}

int main()
{
  uint32_t * d_out; long long * d_clk;
  CK(cudaMalloc(&d_out, 64)); CK(cudaMalloc(&d_clk, 64));
#define THRU(OP, NAME) run(NAME " x8 chains", probe_kernel<OP, 8>, 8, 4, d_out, d_clk); run(NAME " x1 chain (latency)", probe_kernel<OP, 1>, 1, 1, d_out, d_clk);
  THRU(0, "VIADD.16x2 (__vadd2)")
  THRU(1, "32-bit add (a+b)")
  THRU(3, "VIMNMX3.U16x2 (__vimax3_u16x2)")
  THRU(4, "VIADDMNMX.U16x2 (__viaddmax_u16x2)")
  THRU(5, "IMAD (mad.lo, runtime multiplier)")
  THRU(6, "VIMNMX.U16x2 + 2 predicated adds")
  THRU(7, "LOP3 + IMNMX s32 (2 instr)")
  THRU(8, "LOP3")
  for (int w : {1, 2, 3, 4, 6, 8}) {
    run("cell update, all-DPX (6 instr/row)", cell_kernel<0>, 6 * 4, w, d_out, d_clk);
    run("cell update, 3 DPX + 3 IMAD (6 instr/row)", cell_kernel<1>, 6 * 4, w, d_out, d_clk);
  }
  return 0;
}
