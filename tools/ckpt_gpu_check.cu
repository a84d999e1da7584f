// ROUND-2 GROUNDWORK — standalone GPU check of the experimental checkpoint path (not part of the product):
//   nw_ckpt_kernel<8> (forward pass without direction bits)  +  tb_ckpt_kernel (tile-recompute traceback)
// on the configs[1] shape, compared with the oracle on a sample and timed with CUDA events.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I oracle tools/ckpt_gpu_check.cu \
//        -Loracle -loracle -Xlinker -rpath=$PWD/oracle -o /tmp/ckpt_gpu_check && /tmp/ckpt_gpu_check [ntasks]
// It has been compiled, never run (no GPU was left in round 1): expect to debug it.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "../vsearch_b200/csrc/experimental/nw_ckpt.cuh"
#include "../vsearch_b200/csrc/experimental/tb_ckpt.h"
#include "oracle.h"

namespace vsg {
void Error::set(const std::string & m) { std::fprintf(stderr, "error: %s\n", m.c_str()); }
void count_launch(int) {}
}  // namespace vsg

using namespace vsg;

__global__ void tb_ckpt_kernel(const __grid_constant__ ScoreParams sp, DevSeqs qs, DevSeqs ts,
                               const FastTask * __restrict__ tasks, int ntasks, int R,
                               const uint2 * __restrict__ rowck, const uint2 * __restrict__ colck,
                               int32_t * __restrict__ stats)
{
  int const id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= 2 * ntasks) { return; }
  FastTask const tk = tasks[id >> 1];
  int const half = id & 1;
  int const out = half ? tk.out_hi : tk.out_lo;
  if (out < 0) { return; }
  uint32_t const t = half ? tk.thi : tk.tlo;
  ckpt::PairView pv;
  pv.rowck = reinterpret_cast<const ckpt::U2 *>(rowck + tk.dir_off);
  pv.colck = reinterpret_cast<const ckpt::U2 *>(colck + tk.bnd_off);
  pv.R = R; pv.half = half; pv.Q = qs.len[tk.q]; pv.D = ts.len[t];
  pv.q = qs.sym + qs.off[tk.q];
  pv.t = ts.sym + ts.off[t];
  ckpt::TbOut o;
  ckpt::traceback(sp, pv, o, [](char) {});
  int32_t * const st = stats + static_cast<size_t>(out) * VSG_STAT_WORDS;
  st[VSG_STAT_ALIGNED] = o.aligned; st[VSG_STAT_MATCHES] = o.matches; st[VSG_STAT_MISMATCHES] = o.mismatches;
  st[VSG_STAT_GAPS] = o.gaps; st[VSG_STAT_TRIM_LEFT] = o.trim_left; st[VSG_STAT_TRIM_RIGHT] = o.trim_right;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main(int argc, char ** argv)
{
  int const ntasks = argc > 1 ? std::atoi(argv[1]) : 16384;
  constexpr int R = 8;
  int const Q = 250, D = 1500;
  std::mt19937 rng(2024);
  const char acgt[4] = {'A', 'C', 'G', 'T'};
  const uint8_t code[4] = {1, 2, 4, 8};
  std::vector<char> qa(static_cast<size_t>(ntasks) * Q), ta(static_cast<size_t>(ntasks) * 2 * D);
  std::vector<uint8_t> qsym(qa.size()), tsym(ta.size());
  for (size_t x = 0; x < ta.size(); x++) { int const b = rng() & 3; ta[x] = acgt[b]; tsym[x] = code[b]; }
  for (int k = 0; k < ntasks; k++) {
    int const start = rng() % (D - Q + 1);
    for (int i = 0; i < Q; i++) {   // a 5 % mutated window of the first target, as in configs[1]
      size_t const src = static_cast<size_t>(2 * k) * D + start + i;
      int b = tsym[src] == 1 ? 0 : (tsym[src] == 2 ? 1 : (tsym[src] == 4 ? 2 : 3));
      if (rng() % 20 == 0) { b = rng() & 3; }
      qa[static_cast<size_t>(k) * Q + i] = acgt[b]; qsym[static_cast<size_t>(k) * Q + i] = code[b];
    }
  }
  std::vector<int64_t> qoff(ntasks), toff(2 * ntasks);
  std::vector<int32_t> qlen(ntasks, Q), tlen(2 * ntasks, D);
  for (int k = 0; k < ntasks; k++) { qoff[k] = static_cast<int64_t>(k) * Q; }
  for (int k = 0; k < 2 * ntasks; k++) { toff[k] = static_cast<int64_t>(k) * D; }

  ScoreParams sp{};
  int const go[6] = {1, 1, 18, 18, 1, 1}, ge[6] = {1, 1, 2, 2, 1, 1};
  for (int k = 0; k < 6; k++) { sp.go[k] = static_cast<int16_t>(go[k]); sp.ge[k] = static_cast<int16_t>(ge[k]); }
  sp.match = 2; sp.mismatch = -4; sp.n_mismatch = 0; sp.fallback = 0; sp.score_min = static_cast<int16_t>(-32768 + 20);
  for (int i = 0; i < 16; i++) {
    for (int j = 0; j < 16; j++) {
      bool const ai = __builtin_popcount(i) != 1, aj = __builtin_popcount(j) != 1;
      sp.S[i][j] = static_cast<int16_t>((ai || aj) ? 0 : (i == j ? 2 : -4));
    }
  }

  size_t const row_elems = static_cast<size_t>(D + 31) * 32;                 // uint2 per task
  size_t const col_elems = static_cast<size_t>(D / CKPT_KC + 2) * 32 * R;    // uint2 per task
  std::vector<FastTask> tasks(ntasks);
  for (int k = 0; k < ntasks; k++) {
    FastTask & t = tasks[k];
    t.q = k; t.tlo = 2 * k; t.thi = 2 * k + 1; t.out_lo = 2 * k; t.out_hi = 2 * k + 1; t.dmax = D;
    t.dir_off = static_cast<uint64_t>(k) * row_elems; t.bnd_off = static_cast<uint64_t>(k) * col_elems;
  }

  uint8_t *d_qsym, *d_tsym; int64_t *d_qoff, *d_toff; int32_t *d_qlen, *d_tlen, *d_stats; FastTask * d_tasks; uint2 *d_row, *d_col;
  CK(cudaMalloc(&d_qsym, qsym.size())); CK(cudaMalloc(&d_tsym, tsym.size()));
  CK(cudaMalloc(&d_qoff, qoff.size() * 8)); CK(cudaMalloc(&d_toff, toff.size() * 8));
  CK(cudaMalloc(&d_qlen, qlen.size() * 4)); CK(cudaMalloc(&d_tlen, tlen.size() * 4));
  CK(cudaMalloc(&d_stats, static_cast<size_t>(2 * ntasks) * VSG_STAT_WORDS * 4));
  CK(cudaMalloc(&d_tasks, tasks.size() * sizeof(FastTask)));
  CK(cudaMalloc(&d_row, row_elems * ntasks * sizeof(uint2))); CK(cudaMalloc(&d_col, col_elems * ntasks * sizeof(uint2)));
  CK(cudaMemcpy(d_qsym, qsym.data(), qsym.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_tsym, tsym.data(), tsym.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_qoff, qoff.data(), qoff.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_toff, toff.data(), toff.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_qlen, qlen.data(), qlen.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_tlen, tlen.data(), tlen.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_tasks, tasks.data(), tasks.size() * sizeof(FastTask), cudaMemcpyHostToDevice));
  CK(cudaMemset(d_stats, 0, static_cast<size_t>(2 * ntasks) * VSG_STAT_WORDS * 4));
  DevSeqs qs{d_qsym, d_qoff, d_qlen, ntasks}, ts{d_tsym, d_toff, d_tlen, 2 * ntasks};

  size_t const dyn = fast_dyn_smem(R, false);
  CK(cudaFuncSetAttribute(nw_ckpt_kernel<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dyn)));
  CK(cudaFuncSetAttribute(nw_ckpt_kernel<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dyn)));
  cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  float fwd_ms = 0, tb_ms = 0;
  // PLAIN variant: the same alignment problem under the shifted scoring S - 2, ge + 1 (every cell of anti-diagonal
  // i+j is lowered by i+j+2: identical direction bits, all scores <= 0, every add a plain 32-bit subtract)
  ScoreParams sp2 = sp;
  for (int i = 0; i < 16; i++) { for (int j = 0; j < 16; j++) { sp2.S[i][j] = static_cast<int16_t>(sp.S[i][j] - 2); } }
  for (int k = 0; k < 6; k++) { sp2.ge[k] = static_cast<int16_t>(sp.ge[k] + 1); }
  int bad_total = 0;
  for (int variant = 0; variant < 2; variant++) {
  ScoreParams const & spv = variant ? sp2 : sp;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    if (variant) { nw_ckpt_kernel<R, true><<<(ntasks + FAST_WARPS - 1) / FAST_WARPS, FAST_WARPS * 32, dyn>>>(spv, qs, ts, d_tasks, ntasks, d_row, d_col, d_stats); }
    else { nw_ckpt_kernel<R, false><<<(ntasks + FAST_WARPS - 1) / FAST_WARPS, FAST_WARPS * 32, dyn>>>(spv, qs, ts, d_tasks, ntasks, d_row, d_col, d_stats); }
    cudaEventRecord(e1);
    tb_ckpt_kernel<<<(2 * ntasks + 127) / 128, 128>>>(spv, qs, ts, d_tasks, ntasks, R, d_row, d_col, d_stats);
    cudaEventRecord(e2);
    CK(cudaEventSynchronize(e2));
    cudaEventElapsedTime(&fwd_ms, e0, e1); cudaEventElapsedTime(&tb_ms, e1, e2);
    double const cells = static_cast<double>(ntasks) * 2 * Q * D;
    std::printf("variant %s rep %d: forward %.2f ms (%.0f GCUPS), traceback %.2f ms\n", variant ? "plain-sub/shifted" : "all-dpx", rep, fwd_ms, cells / fwd_ms / 1e6, tb_ms);
  }
  CK(cudaGetLastError());

  std::vector<int32_t> stats(static_cast<size_t>(2 * ntasks) * VSG_STAT_WORDS);
  CK(cudaMemcpy(stats.data(), d_stats, stats.size() * 4, cudaMemcpyDeviceToHost));
  oracle_scoring sc; oracle_default_scoring(&sc);
  int bad = 0;
  int const nsample = std::min(2 * ntasks, 1024);
  for (int p = 0; p < nsample; p++) {
    int16_t os; uint16_t oa, om, omi, og; std::vector<char> cig(Q + D + 64);
    oracle_nw16(&sc, qa.data() + static_cast<size_t>(p / 2) * Q, Q, ta.data() + static_cast<size_t>(p) * D, D, &os, &oa, &om, &omi, &og, cig.data(), cig.size());
    int32_t const * st = stats.data() + static_cast<size_t>(p) * VSG_STAT_WORDS;
    int const got_score = st[VSG_STAT_SCORE] + (variant ? Q + D : 0);
    if (got_score != os || st[VSG_STAT_ALIGNED] != oa || st[VSG_STAT_MATCHES] != om || st[VSG_STAT_MISMATCHES] != omi || st[VSG_STAT_GAPS] != og) {
      if (++bad <= 5) { std::printf("MISMATCH pair %d: score %d/%d aligned %d/%d matches %d/%d gaps %d/%d\n", p, got_score, os, st[VSG_STAT_ALIGNED], oa, st[VSG_STAT_MATCHES], om, st[VSG_STAT_GAPS], og); }
    }
  }
  std::printf("%d of %d sampled pairs differ from the oracle\n", bad, nsample);
  bad_total += bad;
  }
  return bad_total != 0;
}
