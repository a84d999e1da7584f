// tb_ckpt.h — the traceback that goes with nw_ckpt_kernel (align_ckpt.cuh): backtrack16
// (align_simd.cpp:1132-1245) over direction bits that are REGENERATED, tile by tile, from the forward
// pass's checkpoints instead of being read from a stored direction matrix.
//
// A tile = the cells lane b (rows [b*R, b*R+R)) computes during one 32-step chunk k of the wavefront,
// i.e. columns [32k - b, 32k - b + 31] clipped to the matrix.  Its inputs are exactly what the kernel
// checkpointed: H/E of the lane's rows at the end of chunk k-1 (column checkpoint k), and H/F leaving
// lane b-1's last row at every step (row checkpoints); first-row / first-column tiles use the boundary
// formulas (align_simd.cpp:852-857, 1895-1910).  Regenerating a tile replays onestep
// (align_simd.cpp:765-780) in plain ints; the walk then follows backtrack16's priorities.  Only the part
// of a tile between its top-left corner and the cell the path enters it at is recomputed.
//
// Host/device: the same function runs inside traceback_ckpt kernels and, compiled for the CPU, in
// tools/ckpt_host_check.cpp, which checks it against the oracle (tests/test_ckpt_cpu.py) over
// checkpoints laid out exactly as the kernel writes them.
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#define VSG_CKPT_HD __host__ __device__ __forceinline__
#else
#define VSG_CKPT_HD inline
#endif
#ifdef __CUDA_ARCH__
#define VSG_CKPT_UNROLL _Pragma("unroll")
#else
#define VSG_CKPT_UNROLL
#endif

namespace vsg {
namespace ckpt {

#ifndef VSG_CK_CHUNK
#define VSG_CK_CHUNK 32
#endif
constexpr int CHUNK = VSG_CK_CHUNK;   // CK_CHUNK of align_ckpt.cuh
constexpr int RMAX = 16;    // rows per lane

struct U2 { uint32_t x, y; };  // layout of CUDA's uint2

// the layout functions of align_ckpt.cuh, restated for host compilation (static_asserted equal there)
VSG_CKPT_HD size_t row_index(int s, int l) { return (static_cast<size_t>(s >> 2) * 32 + l) * 4 + (s & 3); }
VSG_CKPT_HD size_t col_index(int k, int l, int r, int R) { return (static_cast<size_t>(k - 1) * R + r) * 32 + l; }

struct PairView {
  const U2 * rowck;   // the task's row checkpoints
  const U2 * colck;   // the task's column checkpoints
  int R, half, Q, D;
  int general;        // a symbol outside ACGT in either sequence: scores come from the 16x16 matrix
  const uint8_t * q;  // symbols, 4-bit code in the low nibble
  const uint8_t * t;
};

struct ckpt_true { static constexpr bool value = true; };
struct ckpt_false { static constexpr bool value = false; };

struct TbOut { int aligned, matches, mismatches, gaps, trim_left, trim_right; };  // VSG_STAT_* meanings

enum { CQ_L = 0, CT_L = 1, CQ_I = 2, CT_I = 3, CQ_R = 4, CT_R = 5 };

// Bits: storage of one regenerated tile — set(column, word, value) / get(column, word); word w of a column
// holds the 4-bit directions of rows 8w .. 8w+7.  The device keeps it in shared memory (one bank per thread),
// the host in a plain array.
struct HostBits {
  uint32_t w[CHUNK][RMAX / 8];
  void set(int bj, int k, uint32_t v) { w[bj][k] = v; }
  uint32_t get(int bj, int k) const { return w[bj][k]; }
};

// Rows: access to the row checkpoints of one lane for a tile's steps.  stage(l, s0, s1) announces the range
// [s0, s1] of steps about to be read (at most 34 of them: a tile's columns plus the diagonal input); the device
// version copies the sectors into shared memory with cp.async so that all of a tile's loads are in flight
// together instead of one dependent load per column; the host version reads memory directly.
struct HostRows {
  const U2 * rowck;
  int lane = 0;
  void stage(int l, int, int) { lane = l; }
  void wait() {}
  U2 get(int s) const { return rowck[row_index(s, lane)]; }
};

// max(a, b) and whether b > a strictly (the reference's compare of align_simd.cpp:765-780)
VSG_CKPT_HD int max_gt(int a, int b, bool & gt)
{
  gt = b > a;
  return b > a ? b : a;
}

// emit(op) receives the alignment's operations last to first ('M', 'I' = column consumed alone,
// 'D' = row consumed alone).  SP supplies S[16][16], go[6], ge[6], match, mismatch, n_mismatch — the SAME
// (shifted) scoring the forward kernel ran with: the direction bits do not depend on the shift.
// The regeneration runs on the checkpoints' own BIASED values (v + 0x8000 as plain ints): every comparison
// of a cell is between quantities carrying the same bias, so nothing has to be converted.
template <int RT, class SP, class Bits, class Rows, class Emit>
VSG_CKPT_HD void traceback(const SP & sp, const PairView & v, Bits & bits, Rows & rows, TbOut & out, Emit && emit)
{
  constexpr int B = 0x8000;
  int const R = v.R, Q = v.Q, D = v.D, sh = 16 * v.half;
  int const QRqi = sp.go[CQ_I] + sp.ge[CQ_I], Rqi = sp.ge[CQ_I], QRqr = sp.go[CQ_R] + sp.ge[CQ_R], Rqr = sp.ge[CQ_R];
  int const QRti = sp.go[CT_I] + sp.ge[CT_I], Rti = sp.ge[CT_I], QRtr = sp.go[CT_R] + sp.ge[CT_R], Rtr = sp.ge[CT_R];
  int const gotl = sp.go[CT_L], getl = sp.ge[CT_L], goql = sp.go[CQ_L], geql = sp.ge[CQ_L];
  int smatch = sp.match, smismatch = sp.mismatch;
#ifdef __CUDA_ARCH__
  asm volatile("" : "+r"(smatch), "+r"(smismatch));   // sign-extended once, not per cell
#endif
  auto half_of = [&](uint32_t w) { return static_cast<int>((w >> sh) & 0xffffu); };   // stays biased

  int i = Q - 1, j = D - 1;
  int b = i / R, i0 = b * R;
  char op = 0;
  int aligned = 0, matches = 0, mismatches = 0, gaps = 0;
  char last_run_op = 0; int last_run = 0; bool last_open = true;   // the run that ENDS the alignment
  char first_op = 0; int first_run = 0;                            // the run still open = the alignment's first
  auto push = [&](char nop) {
    aligned++;
    if (last_open) {
      if (last_run == 0 || nop == last_run_op) { last_run_op = nop; last_run++; }
      else { last_open = false; }
    }
    if (nop == first_op) { first_run++; } else { first_op = nop; first_run = 1; }
    emit(nop);
    op = nop;
  };

  while (i >= 0 && j >= 0) {
    // ---- regenerate the tile (lane b, chunk k) up to the cell (i, j) ----
    int const k = (j + b) / CHUNK;
    int const jlo = (CHUNK * k - b) > 0 ? (CHUNK * k - b) : 0;
    int const ni = i - i0 + 1, nj = j - jlo + 1;
    // the tile's row checkpoints: lane b-1 at steps (jlo-1)+(b-1) .. j+(b-1), the first one being the diagonal
    // input H(i0-1, jlo-1) of the tile's first cell
    if (b > 0) { rows.stage(b - 1, jlo + b - 2 > 0 ? jlo + b - 2 : 0, j + b - 1); }
    int hcol[RT], ecol[RT], qc[RT];
VSG_CKPT_UNROLL
    for (int a = 0; a < RT; a++) {
      hcol[a] = 0; ecol[a] = 0; qc[a] = 0;
      if (a < ni) {
        int const ii = i0 + a;
        qc[a] = v.q[ii] & 15;
        if (jlo == 0) {
          hcol[a] = B - (gotl + (ii + 1) * getl);                    // H(ii,-1)
          ecol[a] = hcol[a] - (ii == Q - 1 ? QRqr : QRqi);           // E(ii,0)
        } else {
          U2 const ck = v.colck[col_index(k, b, a, R)];
          hcol[a] = half_of(ck.x); ecol[a] = half_of(ck.y);
        }
      }
    }
    int t_raw = v.t[jlo];
    if (b > 0) { rows.wait(); }
    // H(i0-1, jlo-1): the diagonal input of the tile's first cell
    int hd;
    if (b == 0) { hd = jlo == 0 ? B : B - (goql + jlo * geql); }
    else if (jlo == 0) { hd = B - (gotl + i0 * getl); }
    else { hd = half_of(rows.get(jlo - 1 + b - 1).x); }
    // only the query's last row has other query-gap penalties, and it can only be the tile's last row
    int const alast = (i == Q - 1) ? ni - 1 : -1;
    {
      for (int bj = 0; bj < nj; bj++) {
        int const jj = jlo + bj;
        int const qrt = jj >= D - 1 ? QRtr : QRti, rt = jj >= D - 1 ? Rtr : Rti;
        int htop, f_in;
        if (b == 0) { htop = B - (goql + (jj + 1) * geql); f_in = htop - qrt; }
        else { U2 const ck = rows.get(jj + b - 1); htop = half_of(ck.x); f_in = half_of(ck.y); }
        int hdiag = hd;
        hd = htop;
        int const tc = t_raw & 15;
        if (bj + 1 < nj) { t_raw = v.t[jj + 1]; }   // next column's symbol: loaded one iteration before it is masked and used
        uint32_t w0 = 0, w1 = 0;
VSG_CKPT_UNROLL
        for (int a = 0; a < RT; a++) {
          if (a >= ni) { break; }
          int const S = v.general ? sp.S[tc][qc[a]] : (qc[a] == tc ? smatch : smismatch);
          int const t = hdiag + S;
          bool up, left, extup, extleft;
          int const m1 = max_gt(t, f_in, up);                          // up:      F > h
          int const e_in = ecol[a];
          int const h = max_gt(m1, e_in, left);                        // left:    E > h
          int const f = max_gt(h - qrt, f_in - rt, extup);             // extup:   F - R > H - QR
          bool const lastrow = (a == alast);
          int const e = max_gt(h - (lastrow ? QRqr : QRqi), e_in - (lastrow ? Rqr : Rqi), extleft);
          hdiag = hcol[a];
          hcol[a] = h;
          ecol[a] = e;
          f_in = f;
          uint32_t & w = (a < 8) ? w0 : w1;
          uint32_t const one = 1u << (4 * (a & 7));
          if (up) { w += one; }
          if (left) { w += 2u * one; }
          if (extup) { w += 4u * one; }
          if (extleft) { w += 8u * one; }
        }
        bits.set(bj, 0, w0);
        if (RT > 8) { bits.set(bj, 1, w1); }
      }
    }
    // ---- walk inside the tile (backtrack16's priorities, align_simd.cpp:1150-1210) ----
    while (i >= i0 && j >= jlo) {
      int const a = i - i0;
      uint32_t const d = (bits.get(j - jlo, RT > 8 ? (a >> 3) : 0) >> (4 * (a & 7))) & 15u;
      bool const ext_i = (op == 'I') && (d & 8u);
      bool const ext_d = !ext_i && (op == 'D') && (d & 4u);
      bool const open_i = !ext_i && !ext_d && (d & 2u);
      bool const open_d = !ext_i && !ext_d && !open_i && (d & 1u);
      bool const is_i = ext_i || open_i, is_d = ext_d || open_d;
      if ((open_i && op != 'I') || (open_d && op != 'D')) { gaps++; }
      if (!is_i && !is_d) {
        int const qa = v.q[i] & 15, cc = v.t[j] & 15;
        bool const hit = (qa & cc) != 0 && !(sp.n_mismatch && (qa == 15 || cc == 15));
        if (hit) { matches++; } else { mismatches++; }
      }
      if (!is_i) { i--; }
      if (!is_d) { j--; }
      push(is_i ? 'I' : (is_d ? 'D' : 'M'));
    }
    if (i < i0) { b--; i0 -= R; }
  }
  while (i >= 0) { if (op != 'D') { gaps++; } i--; push('D'); }
  while (j >= 0) { if (op != 'I') { gaps++; } j--; push('I'); }
  out.aligned = aligned; out.matches = matches; out.mismatches = mismatches; out.gaps = gaps;
  out.trim_left = first_op == 'D' ? first_run : (first_op == 'I' ? -first_run : 0);
  out.trim_right = last_run_op == 'D' ? last_run : (last_run_op == 'I' ? -last_run : 0);
}

}  // namespace ckpt
}  // namespace vsg
