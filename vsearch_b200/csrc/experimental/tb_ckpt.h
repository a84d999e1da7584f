// tb_ckpt.h — ROUND-2 GROUNDWORK, NOT PART OF THE PRODUCT (nothing in libvsg.so includes it).
// The traceback that goes with experimental/nw_ckpt.cuh: one thread (here: one plain function call)
// per pair regenerates the direction bits of the R x KC tiles its path crosses from the forward
// pass's checkpoints and walks them with backtrack16's priorities (align_simd.cpp:1132-1245).
// Written as host/device code so that the very same function is checked on the CPU against the
// oracle (tools/ckpt_host_check.cpp, tests/test_proto_ckpt_cpu.py) before a kernel wraps it.
//
// Checkpoint layout (what nw_ckpt_kernel writes for one task = one query x two targets, values are
// the kernel's biased halves v + 0x8000, low half = first target, high half = second):
//   rowck[s * 32 + l]                  = (H, F) leaving lane l's last row at step s (column s - l)
//   colck[(blk * 32 + l) * R + r]      = (H, E) of lane l's row r after column blk * KC - 1
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#define VSG_CKPT_HD __host__ __device__ __forceinline__
#else
#define VSG_CKPT_HD inline
#endif

namespace vsg {
namespace ckpt {

constexpr int KC = 32;    // columns per tile (CKPT_KC of nw_ckpt.cuh)
constexpr int RMAX = 8;   // rows per tile = rows per lane

struct U2 { uint32_t x, y; };  // layout of CUDA's uint2

struct PairView {
  const U2 * rowck;
  const U2 * colck;
  int R, half, Q, D;
  const uint8_t * q;  // symbols, 4-bit code in the low nibble
  const uint8_t * t;
};

struct TbOut { int aligned, matches, mismatches, gaps, trim_left, trim_right; };  // VSG_STAT_* meanings

enum { CQ_L = 0, CT_L = 1, CQ_I = 2, CT_I = 3, CQ_R = 4, CT_R = 5 };

template <class SP>
struct Walker {
  const SP & sp;
  const PairView & v;
  VSG_CKPT_HD Walker(const SP & s, const PairView & p) : sp(s), v(p) {}

  VSG_CKPT_HD int unbias(uint32_t w) const { return static_cast<int>((w >> (16 * v.half)) & 0xffffu) - 0x8000; }
  VSG_CKPT_HD int QRq(int i) const { return i == v.Q - 1 ? sp.go[CQ_R] + sp.ge[CQ_R] : sp.go[CQ_I] + sp.ge[CQ_I]; }
  VSG_CKPT_HD int Rq(int i) const { return i == v.Q - 1 ? sp.ge[CQ_R] : sp.ge[CQ_I]; }
  VSG_CKPT_HD int QRt(int j) const { return j >= v.D - 1 ? sp.go[CT_R] + sp.ge[CT_R] : sp.go[CT_I] + sp.ge[CT_I]; }
  VSG_CKPT_HD int Rt(int j) const { return j >= v.D - 1 ? sp.ge[CT_R] : sp.ge[CT_I]; }
  VSG_CKPT_HD int Hleft(int i) const { return i < 0 ? 0 : -(sp.go[CT_L] + (i + 1) * sp.ge[CT_L]); }   // H(i,-1)
  VSG_CKPT_HD int Htop(int j) const { return j < 0 ? 0 : -(sp.go[CQ_L] + (j + 1) * sp.ge[CQ_L]); }    // H(-1,j)

  // borders of the tile of lane b, column block c
  VSG_CKPT_HD int top(int b, int jj) const       // H(b*R - 1, jj), jj >= -1
  {
    if (b == 0) { return Htop(jj); }
    if (jj < 0) { return Hleft(b * v.R - 1); }
    return unbias(v.rowck[static_cast<size_t>(jj + b - 1) * 32 + (b - 1)].x);
  }
  VSG_CKPT_HD int ftop(int b, int jj) const      // F entering (b*R, jj)
  {
    if (b == 0) { return Htop(jj) - QRt(jj); }
    return unbias(v.rowck[static_cast<size_t>(jj + b - 1) * 32 + (b - 1)].y);
  }
  VSG_CKPT_HD int left(int b, int c, int ii) const   // H(ii, c*KC - 1)
  {
    if (c == 0) { return Hleft(ii); }
    return unbias(v.colck[(static_cast<size_t>(c) * 32 + b) * v.R + (ii - b * v.R)].x);
  }
  VSG_CKPT_HD int eleft(int b, int c, int ii) const  // E entering (ii, c*KC)
  {
    if (c == 0) { return Hleft(ii) - QRq(ii); }
    return unbias(v.colck[(static_cast<size_t>(c) * 32 + b) * v.R + (ii - b * v.R)].y);
  }
};

// emit(op) receives the alignment's operations last to first ('M', 'I' = column consumed alone,
// 'D' = row consumed alone, as in traceback_one)
template <class SP, class Emit>
VSG_CKPT_HD void traceback(const SP & sp, const PairView & v, TbOut & out, Emit && emit)
{
  Walker<SP> w(sp, v);
  int const R = v.R;
  int i = v.Q - 1, j = v.D - 1;
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

  uint8_t bits[RMAX][KC];
  int hcol[RMAX], ecol[RMAX];
  while (i >= 0 && j >= 0) {
    int const b = i / R, c = j / KC;
    int const i0 = b * R, j0 = c * KC;
    int const ni = i - i0 + 1, nj = j - j0 + 1;
    for (int a = 0; a < ni; a++) { hcol[a] = w.left(b, c, i0 + a); ecol[a] = w.eleft(b, c, i0 + a); }
    for (int bj = 0; bj < nj; bj++) {
      int const jj = j0 + bj;
      int hdiag = w.top(b, jj - 1);
      int f_in = w.ftop(b, jj);
      int const tc = v.t[jj] & 15;
      int const qrt = w.QRt(jj), rt = w.Rt(jj);
      for (int a = 0; a < ni; a++) {
        int const ii = i0 + a;
        int const t = hdiag + sp.S[tc][v.q[ii] & 15];
        int d = 0;
        if (f_in > t) { d |= 1; }
        int const m1 = t > f_in ? t : f_in;
        int const e_in = ecol[a];
        if (e_in > m1) { d |= 2; }
        int const h = m1 > e_in ? m1 : e_in;
        int const hf = h - qrt, f = f_in - rt;
        if (f > hf) { d |= 4; }
        int const he = h - w.QRq(ii), e = e_in - w.Rq(ii);
        if (e > he) { d |= 8; }
        bits[a][bj] = static_cast<uint8_t>(d);
        hdiag = hcol[a];
        hcol[a] = h;
        ecol[a] = e > he ? e : he;
        f_in = f > hf ? f : hf;
      }
    }
    while (i >= i0 && j >= j0) {
      int const d = bits[i - i0][j - j0];
      bool const ext_i = (op == 'I') && (d & 8);
      bool const ext_d = !ext_i && (op == 'D') && (d & 4);
      bool const open_i = !ext_i && !ext_d && (d & 2);
      bool const open_d = !ext_i && !ext_d && !open_i && (d & 1);
      bool const is_i = ext_i || open_i, is_d = ext_d || open_d;
      if ((open_i && op != 'I') || (open_d && op != 'D')) { gaps++; }
      if (!is_i && !is_d) {
        int const a = v.q[i] & 15, cc = v.t[j] & 15;
        bool const hit = (a & cc) != 0 && !(sp.n_mismatch && (a == 15 || cc == 15));
        if (hit) { matches++; } else { mismatches++; }
      }
      if (!is_i) { i--; }
      if (!is_d) { j--; }
      push(is_i ? 'I' : (is_d ? 'D' : 'M'));
    }
  }
  while (i >= 0) { if (op != 'D') { gaps++; } i--; push('D'); }
  while (j >= 0) { if (op != 'I') { gaps++; } j--; push('I'); }
  out.aligned = aligned; out.matches = matches; out.mismatches = mismatches; out.gaps = gaps;
  out.trim_left = first_op == 'D' ? first_run : (first_op == 'I' ? -first_run : 0);
  out.trim_right = last_run_op == 'D' ? last_run : (last_run_op == 'I' ? -last_run : 0);
}

}  // namespace ckpt
}  // namespace vsg
