// compile-only probe: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -cubin tools/ckpt_probe.cu
// then count the instructions of the steady loop in the SASS (see DESIGN.md, round-2 plan)
#include "../vsearch_b200/csrc/experimental/nw_ckpt.cuh"
void ckpt_probe_launch(const vsg::ScoreParams & sp, vsg::DevSeqs q, vsg::DevSeqs t, const vsg::FastTask * tasks, int n,
                       uint2 * rowck, uint2 * colck, int32_t * stats)
{
  vsg::nw_ckpt_kernel<8, true><<<(n + vsg::FAST_WARPS - 1) / vsg::FAST_WARPS, vsg::FAST_WARPS * 32, vsg::fast_dyn_smem(8, false)>>>(
      sp, q, t, tasks, n, rowck, colck, stats);
}
