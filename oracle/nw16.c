/* oracle/nw16.c — TEST INFRASTRUCTURE ONLY (the checker, never the product).
 *
 * Scalar restatement of ONE lane of the reference's 8-lane int16 aligner `search16`
 * (src/core/align_simd.cpp).  What is reproduced, with the reference lines followed:
 *   - clamping of the 14 scores/penalties to 16-bit cells and the "defer everything"
 *     flag                                             (:1264-1277, :1316-1373)
 *   - the 16x16 score matrix incl. n_mismatch          (:1319-1342)
 *   - special cases: forced fallback, empty query, empty / oversize target
 *                                                      (:1463-1539, :1867-1882, :130-134)
 *   - boundary seeding of a fresh lane                 (:1895-1910, :852-859, :885-887)
 *   - the cell update with its tie-breaking            (onestep, :752-781)
 *   - processing in blocks of 4 target columns, zero-symbol padding of the last
 *     block, right-end target penalties from column (D+3)%4 on
 *                                                      (:1735-1752, :1914-1925)
 *   - SATURATING signed 16-bit adds/subs everywhere    (:350-356)
 *   - per-block h_min/h_max (initialised to 0) and the sticky overflow flag
 *                                                      (:825-826, :1432-1444, :2029-2040)
 *   - top-row continuation between blocks              (:2043-2051)
 *   - traceback priorities, gap counting, run-length CIGAR (backtrack16, :1132-1245;
 *     pushop/finishop :1013-1049)
 * A lane's result does not depend on what the other 7 lanes hold, so one pair at a
 * time is an exact model of search16's per-target outputs.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

typedef int16_t cell;

static cell sat32(int32_t x) { return (cell)(x > 32767 ? 32767 : (x < -32768 ? -32768 : x)); }
static cell adds(cell a, cell b) { return sat32((int32_t)a + (int32_t)b); }
static cell subs(cell a, cell b) { return sat32((int32_t)a - (int32_t)b); }
static cell cmax(cell a, cell b) { return a > b ? a : b; }
static cell cmin(cell a, cell b) { return a < b ? a : b; }

void oracle_default_scoring(oracle_scoring * s)
{
  /* vsearch.h:450-461 defaults (match 2, mismatch -4, interior 20/2, terminal 2/1)
     after "open -= extension" (vsearch.cc:250-259) */
  int64_t const d[14] = {2, -4, 1, 1, 18, 18, 1, 1, 1, 1, 2, 2, 1, 1};
  memcpy(s->v, d, sizeof d);
  s->n_mismatch = 0;
}

int oracle_fits(uint64_t qlen, uint64_t dlen)
{
  return ((int64_t)(qlen + dlen) <= 65535LL) && ((int64_t)qlen * (int64_t)dlen <= 25000000LL);
}

static cell clamp_to_cell(int64_t v, int64_t limit, int * fallback)
{
  if (v > limit) { *fallback = 1; return (cell)limit; }
  if (v < -limit) { *fallback = 1; return (cell)(-limit); }
  return (cell)v;
}

typedef struct {
  cell match, mismatch;
  cell go_q_l, go_t_l, go_q_i, go_t_i, go_q_r, go_t_r;
  cell ge_q_l, ge_t_l, ge_q_i, ge_t_i, ge_q_r, ge_t_r;
  cell S[16][16];
  int n_mismatch, fallback;
} prep_t;

static void prep(const oracle_scoring * sc, prep_t * p)
{
  int fb = 0;
  int64_t const slim = 32767, plim = 32767 / 5; /* :1256-1257 */
  p->match = clamp_to_cell(sc->v[0], slim, &fb);
  p->mismatch = clamp_to_cell(sc->v[1], slim, &fb);
  p->go_q_l = clamp_to_cell(sc->v[2], plim, &fb);
  p->go_t_l = clamp_to_cell(sc->v[3], plim, &fb);
  p->go_q_i = clamp_to_cell(sc->v[4], plim, &fb);
  p->go_t_i = clamp_to_cell(sc->v[5], plim, &fb);
  p->go_q_r = clamp_to_cell(sc->v[6], plim, &fb);
  p->go_t_r = clamp_to_cell(sc->v[7], plim, &fb);
  p->ge_q_l = clamp_to_cell(sc->v[8], plim, &fb);
  p->ge_t_l = clamp_to_cell(sc->v[9], plim, &fb);
  p->ge_q_i = clamp_to_cell(sc->v[10], plim, &fb);
  p->ge_t_i = clamp_to_cell(sc->v[11], plim, &fb);
  p->ge_q_r = clamp_to_cell(sc->v[12], plim, &fb);
  p->ge_t_r = clamp_to_cell(sc->v[13], plim, &fb);
  p->n_mismatch = sc->n_mismatch != 0;
  p->fallback = fb;
  for (unsigned i = 0; i < 16; i++) {
    for (unsigned j = 0; j < 16; j++) {
      cell v;
      if (p->n_mismatch && (i == 15 || j == 15)) { v = p->mismatch; }
      else if (oracle_is_ambiguous_4bit(i) || oracle_is_ambiguous_4bit(j)) { v = 0; }
      else if (i == j) { v = p->match; }
      else { v = p->mismatch; }
      p->S[i][j] = v;
    }
  }
}

static void sentinel(int16_t * score, uint16_t * aligned, uint16_t * matches,
                     uint16_t * mismatches, uint16_t * gaps, char * cigar)
{
  *score = ORACLE_SENTINEL; *aligned = 0; *matches = 0; *mismatches = 0; *gaps = 0; cigar[0] = 0;
}

/* reversed run-length writer, as pushop/finishop build the string from its end */
typedef struct { char * end; char op; int count; } cigw;
static void flush_run(cigw * w)
{
  if (w->op != 0 && w->count != 0) {
    *--w->end = w->op;
    if (w->count > 1) {
      char buf[16];
      int const l = snprintf(buf, sizeof buf, "%d", w->count);
      w->end -= l;
      memcpy(w->end, buf, (size_t)l);
    }
  }
}
static void push(cigw * w, char op)
{
  if (op == w->op) { w->count++; return; }
  flush_run(w);
  w->op = op; w->count = 1;
}

int oracle_nw16(const oracle_scoring * sc,
                const char * q, int64_t qlen, const char * d, int64_t dlen,
                int16_t * score, uint16_t * aligned, uint16_t * matches,
                uint16_t * mismatches, uint16_t * gaps,
                char * cigar, size_t cigar_cap)
{
  prep_t P;
  prep(sc, &P);
  if (cigar_cap < (size_t)(qlen + dlen + 24)) { return -1; }

  if (P.fallback) { sentinel(score, aligned, matches, mismatches, gaps, cigar); return 0; }

  if (qlen == 0) { /* :1481-1539 */
    if (!oracle_fits(0, (uint64_t)dlen)) { sentinel(score, aligned, matches, mismatches, gaps, cigar); return 0; }
    *aligned = (uint16_t)dlen; *matches = 0; *mismatches = 0; *gaps = (uint16_t)dlen;
    if (dlen == 0) { *score = 0; cigar[0] = 0; }
    else {
      int64_t const a = -(int64_t)P.go_t_l - dlen * (int64_t)P.ge_t_l;
      int64_t const b = -(int64_t)P.go_t_r - dlen * (int64_t)P.ge_t_r;
      *score = (int16_t)(a > b ? a : b); /* plain narrowing, as the reference's static_cast<CELL> */
      snprintf(cigar, cigar_cap, "%lldI", (long long)dlen);
    }
    return 0;
  }

  if (dlen == 0 || !oracle_fits((uint64_t)qlen, (uint64_t)dlen)) { /* :1871-1881 */
    sentinel(score, aligned, matches, mismatches, gaps, cigar);
    return 0;
  }

  cell const QR_q_i = (cell)(P.go_q_i + P.ge_q_i), R_q_i = P.ge_q_i;
  cell const QR_q_r = (cell)(P.go_q_r + P.ge_q_r), R_q_r = P.ge_q_r;
  cell const QR_t_l = (cell)(P.go_t_l + P.ge_t_l), R_t_l = P.ge_t_l;
  cell const QR_t_i = (cell)(P.go_t_i + P.ge_t_i), R_t_i = P.ge_t_i;
  cell const QR_t_r = (cell)(P.go_t_r + P.ge_t_r), R_t_r = P.ge_t_r;
  cell const R_q_l = P.ge_q_l;

  int gpmax = 0; /* compute_score_min, :1432-1444 */
  {
    int const c[6] = {P.go_q_l + P.ge_q_l, P.go_q_i + P.ge_q_i, P.go_q_r + P.ge_q_r,
                      P.go_t_l + P.ge_t_l, P.go_t_i + P.ge_t_i, P.go_t_r + P.ge_t_r};
    for (int k = 0; k < 6; k++) { if (c[k] > gpmax) { gpmax = c[k]; } }
  }
  cell const score_min = (cell)(-32768 + gpmax);
  cell const score_max = 32767;

  cell * Hcol = (cell *)malloc(sizeof(cell) * (size_t)qlen);
  cell * Ecol = (cell *)malloc(sizeof(cell) * (size_t)qlen);
  uint8_t * dir = (uint8_t *)malloc((size_t)qlen * (size_t)dlen);
  uint8_t * qc = (uint8_t *)malloc((size_t)qlen);
  if (!Hcol || !Ecol || !dir || !qc) { abort(); }
  for (int64_t i = 0; i < qlen; i++) { qc[i] = oracle_map_4bit((unsigned char)q[i]); }

  /* fresh lane: H0..H3 = H(-1,-1..2), F0..F3 = H(-1,0..3)   (:1895-1910) */
  cell H[4], F[4], Sm[4] = {0, 0, 0, 0};
  H[0] = 0;
  H[1] = (cell)(-P.go_q_l - 1 * P.ge_q_l);
  H[2] = (cell)(-P.go_q_l - 2 * P.ge_q_l);
  H[3] = (cell)(-P.go_q_l - 3 * P.ge_q_l);
  F[0] = (cell)(-P.go_q_l - 1 * P.ge_q_l);
  F[1] = (cell)(-P.go_q_l - 2 * P.ge_q_l);
  F[2] = (cell)(-P.go_q_l - 3 * P.ge_q_l);
  F[3] = (cell)(-P.go_q_l - 4 * P.ge_q_l);

  int overflow = 0;
  int64_t const nblocks = (dlen + 3) / 4;
  for (int64_t b = 0; b < nblocks; b++) {
    int sym[4];
    cell QRt[4], Rt[4];
    int const ends_here = (4 * b + 4 >= dlen);
    for (int k = 0; k < 4; k++) {
      int64_t const j = 4 * b + k;
      sym[k] = j < dlen ? oracle_map_4bit((unsigned char)d[j]) : 0;
      int const right = ends_here && (k >= (int)((dlen + 3) % 4)); /* :1741-1751 */
      QRt[k] = right ? adds(QR_t_i, subs(QR_t_r, QR_t_i)) : QR_t_i;
      Rt[k] = right ? adds(R_t_i, subs(R_t_r, R_t_i)) : R_t_i;
    }
    cell h_min = 0, h_max = 0;
    cell h[4] = {H[0], H[1], H[2], H[3]};
    cell f[4], n[4] = {0, 0, 0, 0};
    for (int k = 0; k < 4; k++) { f[k] = subs(F[k], QRt[k]); }
    cell MQRtl = QR_t_l; /* M_QR_target_left of a freshly started lane */
    for (int64_t i = 0; i < qlen; i++) {
      int const last = (i == qlen - 1);
      cell h4 = 0, E;
      if (b == 0) { /* aligncolumns_first, :836-897 */
        if (!last) {
          h4 = subs(0, MQRtl);
          E = subs(subs(0, MQRtl), QR_q_i);
          MQRtl = adds(MQRtl, R_t_l);
        } else {
          E = subs(subs(0, MQRtl), QR_q_r);
        }
      } else { /* aligncolumns_rest, :959-998 */
        if (!last) { h4 = Hcol[i]; }
        E = Ecol[i];
      }
      cell const QRq = last ? QR_q_r : QR_q_i;
      cell const Rq = last ? R_q_r : R_q_i;
      for (int k = 0; k < 4; k++) { /* onestep, :765-780 */
        cell Hc = adds(h[k], P.S[sym[k]][qc[i]]);
        uint8_t bits = 0;
        if (f[k] > Hc) { bits |= 1; }
        Hc = cmax(Hc, f[k]);
        if (E > Hc) { bits |= 2; }
        Hc = cmax(Hc, E);
        h_min = cmin(h_min, Hc);
        h_max = cmax(h_max, Hc);
        n[k] = Hc;
        cell const HF = subs(Hc, QRt[k]);
        f[k] = subs(f[k], Rt[k]);
        if (f[k] > HF) { bits |= 4; }
        f[k] = cmax(f[k], HF);
        cell const HE = subs(Hc, QRq);
        E = subs(E, Rq);
        if (E > HE) { bits |= 8; }
        E = cmax(E, HE);
        int64_t const j = 4 * b + k;
        if (j < dlen) { dir[(size_t)i * (size_t)dlen + (size_t)j] = bits; }
      }
      Hcol[i] = n[3];
      Ecol[i] = E;
      h[0] = h4; h[1] = n[0]; h[2] = n[1]; h[3] = n[2];
    }
    for (int k = 0; k < 4; k++) { Sm[k] = n[k]; }
    if (h_min <= score_min || h_max >= score_max) { overflow = 1; }
    /* :2043-2051 */
    H[0] = subs(H[3], R_q_l); H[1] = subs(H[0], R_q_l); H[2] = subs(H[1], R_q_l); H[3] = subs(H[2], R_q_l);
    F[0] = subs(F[3], R_q_l); F[1] = subs(F[0], R_q_l); F[2] = subs(F[1], R_q_l); F[3] = subs(F[2], R_q_l);
  }

  if (overflow) {
    sentinel(score, aligned, matches, mismatches, gaps, cigar);
  } else {
    *score = Sm[(dlen + 3) % 4];
    /* backtrack16, :1132-1245 */
    uint16_t al = 0, ma = 0, mi = 0, ga = 0;
    int64_t i = qlen - 1, j = dlen - 1;
    char * const endp = cigar + qlen + dlen + 1;
    cigw w; w.end = endp; w.op = 0; w.count = 0;
    *--w.end = 0;
    char op = 0;
    while (i >= 0 && j >= 0) {
      al++;
      uint8_t const bits = dir[(size_t)i * (size_t)dlen + (size_t)j];
      if (op == 'I' && (bits & 8)) { j--; push(&w, 'I'); }
      else if (op == 'D' && (bits & 4)) { i--; push(&w, 'D'); }
      else if (bits & 2) { if (op != 'I') { ga++; } j--; push(&w, 'I'); op = 'I'; }
      else if (bits & 1) { if (op != 'D') { ga++; } i--; push(&w, 'D'); op = 'D'; }
      else {
        unsigned const a = qc[i], c = oracle_map_4bit((unsigned char)d[j]);
        if ((a & c) != 0) {
          if (P.n_mismatch && (a == 15 || c == 15)) { mi++; } else { ma++; }
        } else { mi++; }
        i--; j--; push(&w, 'M'); op = 'M';
      }
    }
    while (i >= 0) { al++; if (op != 'D') { ga++; } i--; push(&w, 'D'); op = 'D'; }
    while (j >= 0) { al++; if (op != 'I') { ga++; } j--; push(&w, 'I'); op = 'I'; }
    flush_run(&w);
    memmove(cigar, w.end, (size_t)(endp - w.end));
    *aligned = al; *matches = ma; *mismatches = mi; *gaps = ga;
  }
  free(Hcol); free(Ecol); free(dir); free(qc);
  return 0;
}
