// mask.cu — DUST soft-masking of sequence sets on the device.
//
// Replaces dust() / dust_all() (reference core/mask.cpp:79-188; called on every query by
// commands/usearch_global.cpp:386-389 and on the database by search_prep :579 — both ON by default).
// It is not part of the aligner/ranker path proper, but at several hundred thousand queries per second
// the reference's scalar dust() on the host would be the bottleneck, and the
// ranker honours the mask (lower-case symbols do not seed k-mers, unique.cpp:198-199).
//
// Semantics reproduced exactly: the whole sequence is upper-cased, then for windows of 64 symbols
// starting every 32 symbols the best-scoring sub-interval (word size 3; score 10*sum/j with integer
// division; strict ">" so the first (i, j) in scan order wins ties) is lower-cased if its score exceeds
// 20, and the window start skips ahead by 32 - b when the interval ends in the first half.
// One warp per sequence; lane l evaluates the interval starts i = l and i = l + 32 of a window with a
// private 64-entry count table in shared memory.
#include "vsg_internal.h"

namespace vsg {

constexpr int DUST_WARPS = 4;

__global__ void __launch_bounds__(DUST_WARPS * 32)
dust_kernel(uint8_t * __restrict__ sym, const int64_t * __restrict__ off, const int32_t * __restrict__ lens, int64_t n)
{
  __shared__ uint8_t counts[DUST_WARPS][32][64];
  __shared__ uint8_t words[DUST_WARPS][64];
  int const lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int64_t const w = static_cast<int64_t>(blockIdx.x) * DUST_WARPS + wib;
  if (w >= n) { return; }
  uint8_t * const s = sym + off[w];
  int const len = lens[w];
  // "convert sequence to upper case" (mask.cpp:131-139)
  for (int p = lane; p < len; p += 32) { s[p] &= 0x0f; }
  __syncwarp();
  uint8_t * const cnt = counts[wib][lane];
  uint8_t * const wd = words[wib];

  for (int i0 = 0; i0 < len; i0 += 32) {
    int const l = (len > i0 + 64) ? 64 : len - i0;
    int const l1 = l - 3 + 1 - 5;
    int bestv = 0, besti = 0, bestj = 0;
    if (l1 > 0) {
      // words[j] = the (up to) three symbols ending at j, 2 bits each, as wo() accumulates them
      for (int j = lane; j < l; j += 32) {
        unsigned word = 0;
        for (int t = (j >= 2 ? j - 2 : 0); t <= j; t++) {
          int const c = s[i0 + t] & 15;
          word = (word << 2) | ((c == 2) ? 1u : (c == 4) ? 2u : (c == 8) ? 3u : 0u);
        }
        wd[j] = static_cast<uint8_t>(word & 63u);
      }
      __syncwarp();
      for (int i = lane; i < l1; i += 32) {
        for (int k = 0; k < 64; k += 4) { *reinterpret_cast<uint32_t *>(cnt + k) = 0; }
        int sum = 0;
        for (int j = 2; j < l - i; j++) {
          int const word = wd[i + j];
          int const c = cnt[word];
          if (c != 0) {
            sum += c;
            int const v = 10 * sum / j;
            if (v > bestv) { bestv = v; besti = i; bestj = j; }
          }
          cnt[word] = static_cast<uint8_t>(c + 1);
        }
      }
      // first maximum in (i, j) scan order across lanes
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        int const ov = __shfl_xor_sync(0xffffffffu, bestv, d);
        int const oi = __shfl_xor_sync(0xffffffffu, besti, d);
        int const oj = __shfl_xor_sync(0xffffffffu, bestj, d);
        bool const take = (ov > bestv) || (ov == bestv && (oi < besti || (oi == besti && oj < bestj)));
        if (take) { bestv = ov; besti = oi; bestj = oj; }
      }
      __syncwarp();
    }
    if (bestv > 20) {
      int const a = besti, b = besti + bestj;
      for (int j = a + lane; j <= b; j += 32) { s[i0 + j] |= 0x10; }
      if (b < 32) { i0 += 32 - b; }
    }
    __syncwarp();
  }
}

}  // namespace vsg

using namespace vsg;

extern "C" int vsg_seqset_dust(vsg_ctx * c, vsg_seqset * s)
{
  if (c == nullptr || s == nullptr) { Error::set("vsg_seqset_dust: null argument"); return VSG_EINVAL; }
  VSG_CUDA_OK(cudaSetDevice(c->device));
  if (s->d.n == 0) { return VSG_OK; }
  int64_t const blocks = (s->d.n + DUST_WARPS - 1) / DUST_WARPS;
  dust_kernel<<<static_cast<unsigned>(blocks), DUST_WARPS * 32, 0, c->stream>>>(
      static_cast<uint8_t *>(s->b_sym.p), s->d.off, s->d.len, s->d.n);
  count_launch();
  VSG_CUDA_OK(cudaStreamSynchronize(c->stream));
  VSG_CUDA_OK(cudaGetLastError());
  return VSG_OK;
}

extern "C" int vsg_seqset_symbols(vsg_ctx * c, const vsg_seqset * s, uint8_t * out, int64_t cap)
{
  if (c == nullptr || s == nullptr || out == nullptr) { Error::set("vsg_seqset_symbols: null argument"); return VSG_EINVAL; }
  if (cap < s->total) { Error::set("vsg_seqset_symbols: buffer too small"); return VSG_ECAP; }
  VSG_CUDA_OK(cudaSetDevice(c->device));
  if (s->total > 0) {
    VSG_CUDA_OK(cudaMemcpyAsync(out, s->d.sym, static_cast<size_t>(s->total), cudaMemcpyDeviceToHost, c->stream));
    VSG_CUDA_OK(cudaStreamSynchronize(c->stream));
  }
  return VSG_OK;
}
