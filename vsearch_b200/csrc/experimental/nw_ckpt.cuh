// nw_ckpt.cuh — ROUND-2 GROUNDWORK, NOT PART OF THE PRODUCT: libvsg.so does not include this file and
// nothing launches it.  It exists so that the instruction count of a flag-free forward pass can be
// read off the SASS today (tools/ckpt_probe.cu) and the next round starts from compiling code.
// The algorithm it belongs to is pinned on the CPU by tools/proto_checkpoint_traceback.py.
//
// nw_ckpt_kernel<R>: the wavefront of nw_fast_kernel (one warp, one query, two targets as 16-bit
// halves, lane l owns rows [l*R, l*R+R), column s - l at step s) WITHOUT direction bits:
//   per row   t  = Hdiag + S                       |  h  = max3(t, F, E)        VIADD.16x2 + VIMNMX3.U16x2
//             F' = max(h - QRt, F - Rt)            |  E' = max(h - QRq, E - Rq) 2 x (VIADD + VIADDMNMX.U16x2)
// and, instead of 0.5 byte of directions per cell, the values a traceback needs to REGENERATE the
// directions of any R x KC tile:
//   row checkpoints   (H, F) leaving the lane's last row, every step         -> rowck[(s*32 + lane)]
//   column checkpoints (H, E) of the lane's R rows after every KC-th column  -> colck[((c+1)/KC*32 + lane)*R + r]
#pragma once

#include "../align_kernels.cuh"

namespace vsg {

constexpr int CKPT_KC = 32;   // columns between column checkpoints

template <int R, bool PLAIN>
__global__ void __launch_bounds__(FAST_WARPS * 32)
nw_ckpt_kernel(const __grid_constant__ ScoreParams sp, DevSeqs qs, DevSeqs ts,
               const FastTask * __restrict__ tasks, int ntasks,
               uint2 * __restrict__ rowck, uint2 * __restrict__ colck, int32_t * __restrict__ stats)
{
  static_assert(R <= 8, "profile variant only");
  constexpr int RQ = (R + 3) / 4;
  extern __shared__ uint4 prof_mem[];
  __shared__ uint4 ringA[FAST_WARPS][2 * RING];
  __shared__ uint32_t ringB[FAST_WARPS][2 * RING];

  int const lane = threadIdx.x & 31;
  int const wib = threadIdx.x >> 5;
  int const w = blockIdx.x * FAST_WARPS + wib;
  if (w >= ntasks) { return; }
  FastTask const tk = tasks[w];

  int const Q = qs.len[tk.q];
  uint8_t const * __restrict__ qsym = qs.sym + qs.off[tk.q];
  int const Dlo = ts.len[tk.tlo], Dhi = ts.len[tk.thi];
  uint8_t const * __restrict__ dlo_p = ts.sym + ts.off[tk.tlo];
  uint8_t const * __restrict__ dhi_p = ts.sym + ts.off[tk.thi];
  int const dmax = tk.dmax;
  int const nsteps = dmax + 31;

  int const QRqi = sp.go[Q_I] + sp.ge[Q_I], Rqi = sp.ge[Q_I];
  int const QRqr = sp.go[Q_R] + sp.ge[Q_R], Rqr = sp.ge[Q_R];
  int const QRti = sp.go[T_I] + sp.ge[T_I], Rti = sp.ge[T_I];
  int const QRtr = sp.go[T_R] + sp.ge[T_R], Rtr = sp.ge[T_R];
  int const gotl = sp.go[T_L], getl = sp.ge[T_L];
  int const goql = sp.go[Q_L], geql = sp.ge[Q_L];

  int const llast = (Q - 1) / R;
  int const rlast = (Q - 1) % R;
  int score_lo = 0, score_hi = 0;

  uint4 * const rA = ringA[wib];
  uint32_t * const rB = ringB[wib];
  uint32_t const rA_s = static_cast<uint32_t>(__cvta_generic_to_shared(rA));
  uint32_t const rB_s = static_cast<uint32_t>(__cvta_generic_to_shared(rB));
  uint4 * const myprof = prof_mem + static_cast<size_t>(wib) * 16 * RQ * 32 + lane;
  uint32_t const prof_s = static_cast<uint32_t>(__cvta_generic_to_shared(myprof));

  int const row0 = lane * R;
  // every quantity NEGATED where it is subtracted, so that add+max fuses (VIADDMNMX)
  uint32_t Hl[R], E[R], nQRq[R], nRq[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    int const i = row0 + r;
    bool const last = (i == Q - 1);
    nQRq[r] = pk1(-(last ? QRqr : QRqi));
    nRq[r] = PLAIN ? pk1(last ? Rqr : Rqi) : pk1(-(last ? Rqr : Rqi));   // PLAIN: positive, subtracted with a 32-bit SUB
    Hl[r] = BIAS2 - pk1(gotl + (i + 1) * getl);
    E[r] = __vadd2(Hl[r], nQRq[r]);
    asm volatile("" : "+r"(nQRq[r]), "+r"(nRq[r]));
  }
  {
    int code[RQ * 4];
#pragma unroll
    for (int r = 0; r < RQ * 4; r++) {
      int const i = row0 + r;
      code[r] = (r < R && i < Q) ? (1 << code_to_2bit(qsym[i] & 15)) : 1;
    }
    for (int tp = 0; tp < 16; tp++) {
      int const dlo = 1 << (tp & 3), dhi = 1 << (tp >> 2);
#pragma unroll
      for (int r4 = 0; r4 < RQ; r4++) {
        uint4 v;
        int const sg = PLAIN ? -1 : 1;   // PLAIN: scores are <= 0 (shifted scoring), stored negated and SUBtracted
        v.x = pk2(sg * sp.S[dlo][code[4 * r4 + 0]], sg * sp.S[dhi][code[4 * r4 + 0]]);
        v.y = pk2(sg * sp.S[dlo][code[4 * r4 + 1]], sg * sp.S[dhi][code[4 * r4 + 1]]);
        v.z = pk2(sg * sp.S[dlo][code[4 * r4 + 2]], sg * sp.S[dhi][code[4 * r4 + 2]]);
        v.w = pk2(sg * sp.S[dlo][code[4 * r4 + 3]], sg * sp.S[dhi][code[4 * r4 + 3]]);
        myprof[(tp * RQ + r4) * 32] = v;
      }
    }
  }
  uint32_t diag_in = (row0 == 0) ? BIAS2 : BIAS2 - pk1(gotl + row0 * getl);
  uint32_t Hout = BIAS2, Fout = BIAS2;
  uint2 * const myrow = rowck + tk.dir_off + lane;                               // + s * 32
  uint2 * const mycol = colck + tk.bnd_off + static_cast<size_t>(lane) * R;      // + block * 32 * R + r
  bool const capture = (lane == llast);

  auto step = [&](auto edge_tag, int c, uint32_t aA, uint32_t aB, int s) {
    constexpr bool EDGE = decltype(edge_tag)::value;
    uint32_t hin = __shfl_up_sync(0xffffffffu, Hout, 1);
    uint32_t fin = __shfl_up_sync(0xffffffffu, Fout, 1);
    if (!EDGE || (c >= 0 && c < dmax)) {
      uint4 const rec = lds128(aA);   // x: profile offset, y: -QRt, z: -Rt, w: H(-1,c)
      if (lane == 0) { hin = rec.w; fin = lds32(aB); }
      uint32_t t[R];
      uint32_t const pa = prof_s + rec.x;
#pragma unroll
      for (int r4 = 0; r4 < RQ; r4++) {
        uint4 const S4 = lds128(pa + r4 * 512u);
        uint32_t const Sv[4] = {S4.x, S4.y, S4.z, S4.w};
#pragma unroll
        for (int u = 0; u < 4; u++) {
          int const r = 4 * r4 + u;
          if (r < R) { t[r] = PLAIN ? ((r == 0 ? diag_in : Hl[r - 1]) - Sv[u]) : __vadd2(r == 0 ? diag_in : Hl[r - 1], Sv[u]); }
        }
      }
      uint32_t F = fin;
#pragma unroll
      for (int r = 0; r < R; r++) {
        uint32_t const h = __vimax3_u16x2(t[r], F, E[r]);
        Hl[r] = h;
        F = __viaddmax_u16x2(h, rec.y, PLAIN ? (F - rec.z) : __vadd2(F, rec.z));
        E[r] = __viaddmax_u16x2(h, nQRq[r], PLAIN ? (E[r] - nRq[r]) : __vadd2(E[r], nRq[r]));
      }
      Hout = Hl[R - 1];
      Fout = F;
      diag_in = hin;
      myrow[static_cast<size_t>(s) * 32] = make_uint2(Hout, Fout);
      if (((c + 1) & (CKPT_KC - 1)) == 0) {
        uint2 * const cp = mycol + static_cast<size_t>((c + 1) / CKPT_KC) * 32 * R;
#pragma unroll
        for (int r = 0; r < R; r++) { cp[r] = make_uint2(Hl[r], E[r]); }
      }
      if (EDGE && capture && (c == Dlo - 1 || c == Dhi - 1)) {
        uint32_t v = 0;
#pragma unroll
        for (int r = 0; r < R; r++) { if (r == rlast) { v = Hl[r]; } }
        if (c == Dlo - 1) { score_lo = static_cast<int>(v & 0xffffu) - static_cast<int>(BIAS); }
        if (c == Dhi - 1) { score_hi = static_cast<int>(v >> 16) - static_cast<int>(BIAS); }
      }
    }
  };

  int const cap_lo = Dlo - 1 + llast, cap_hi = Dhi - 1 + llast;
  for (int s0 = 0; s0 < nsteps; s0 += 32) {
    __syncwarp();
    int const cc = s0 + lane;
    if (cc < dmax) {
      int const a = (cc < Dlo) ? (dlo_p[cc] & 15) : 0;
      int const b = (cc < Dhi) ? (dhi_p[cc] & 15) : 0;
      uint4 rec;
      rec.x = static_cast<uint32_t>(code_to_2bit(a) + 4 * code_to_2bit(b)) * (RQ * 512u);
      rec.y = pk2(-(cc >= Dlo - 1 ? QRtr : QRti), -(cc >= Dhi - 1 ? QRtr : QRti));
      rec.z = PLAIN ? pk2(cc >= Dlo - 1 ? Rtr : Rti, cc >= Dhi - 1 ? Rtr : Rti) : pk2(-(cc >= Dlo - 1 ? Rtr : Rti), -(cc >= Dhi - 1 ? Rtr : Rti));
      rec.w = BIAS2 - pk1(goql + (cc + 1) * geql);
      uint32_t const fin0 = __vadd2(rec.w, rec.y);
      int const slot = cc & (RING - 1);
      rA[slot] = rec; rA[slot + RING] = rec;
      rB[slot] = fin0; rB[slot + RING] = fin0;
    }
    __syncwarp();
    uint32_t const slot0 = static_cast<uint32_t>(s0 - lane) & (RING - 1);
    uint32_t aA = rA_s + slot0 * 16u, aB = rB_s + slot0 * 4u;
    bool const steady = (s0 >= 32) && (s0 + 31 < dmax) &&
                        (static_cast<unsigned>(cap_lo - s0) >= 32u) && (static_cast<unsigned>(cap_hi - s0) >= 32u);
    if (steady) {
#pragma unroll 2
      for (int k = 0; k < 32; k++) { step(std::false_type{}, s0 - lane + k, aA + k * 16u, aB + k * 4u, s0 + k); }
    } else {
      int const kend = min(32, nsteps - s0);
      int c = s0 - lane;
      for (int k = 0; k < kend; k++, c++, aA += 16u, aB += 4u) { step(std::true_type{}, c, aA, aB, s0 + k); }
    }
  }
  if (lane == llast) {
    if (tk.out_lo >= 0) { stats[static_cast<size_t>(tk.out_lo) * VSG_STAT_WORDS + VSG_STAT_SCORE] = score_lo; }
    if (tk.out_hi >= 0) { stats[static_cast<size_t>(tk.out_hi) * VSG_STAT_WORDS + VSG_STAT_SCORE] = score_hi; }
  }
}

}  // namespace vsg
