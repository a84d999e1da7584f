"""Round-2 groundwork (prototype, CPU only, not part of the product): a forward pass that keeps NO
direction bits, only H/E at every KC-th column and H/F under every R-th row, and a traceback that
regenerates the direction bits of the R x KC tiles the path actually crosses.

Why: in nw_fast_kernel the eight predicated flag adds per cell pair cost as much as the recurrence
(DESIGN.md section 7).  Without them a cell pair is 6 packed instructions on sm_100a
(VIADDMNMX.U16x2 fuses add+max; checked with cuobjdump) instead of ~20, and the traceback recomputes
only (Q*KC + D*R) of the Q*D cells.  This script pins the algorithm — which boundary values are
needed, how the bits are regenerated, how the walk crosses tile edges — against the oracle, bit for
bit, before any CUDA is written.   python tools/proto_checkpoint_traceback.py [n_pairs]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import checkers  # noqa: E402

Q_L, T_L, Q_I, T_I, Q_R, T_R = range(6)


def code4(ch):
    return int(checkers.oracle().oracle_map_4bit(ch)) & 15


def score_matrix(pen, n_mismatch):
    match, mismatch = int(pen[0]), int(pen[1])
    amb = lambda x: bin(x).count("1") != 1
    S = [[0] * 16 for _ in range(16)]
    for i in range(16):
        for j in range(16):
            if n_mismatch and (i == 15 or j == 15):
                v = mismatch
            elif amb(i) or amb(j):
                v = 0
            elif i == j:
                v = match
            else:
                v = mismatch
            S[i][j] = v
    return S


class Problem:
    def __init__(self, q, t, pen, n_mismatch=0):
        self.q = [code4(c) for c in q]
        self.t = [code4(c) for c in t]
        self.Q, self.D = len(q), len(t)
        self.go = [int(x) for x in pen[2:8]]
        self.ge = [int(x) for x in pen[8:14]]
        self.S = score_matrix(pen, n_mismatch)
        self.n_mismatch = n_mismatch

    # penalties of the cell's row / column (align_simd.cpp:1741-1751: right-end values from the last one on)
    def QRq(self, i): return self.go[Q_R] + self.ge[Q_R] if i == self.Q - 1 else self.go[Q_I] + self.ge[Q_I]
    def Rq(self, i): return self.ge[Q_R] if i == self.Q - 1 else self.ge[Q_I]
    def QRt(self, j): return self.go[T_R] + self.ge[T_R] if j >= self.D - 1 else self.go[T_I] + self.ge[T_I]
    def Rt(self, j): return self.ge[T_R] if j >= self.D - 1 else self.ge[T_I]
    # matrix borders (align_simd.cpp:852-857, 1895-1901, 830-833)
    def Hleft(self, i): return 0 if i < 0 else -(self.go[T_L] + (i + 1) * self.ge[T_L])      # H(i,-1)
    def Htop(self, j): return 0 if j < 0 else -(self.go[Q_L] + (j + 1) * self.ge[Q_L])       # H(-1,j)

    def cell(self, i, j, hdiag, f_in, e_in):
        """one cell update; returns (h, f_out, e_out, bits): the kernel's recurrence and strict compares"""
        t = hdiag + self.S[self.t[j]][self.q[i]]
        bits = 0
        if f_in > t: bits |= 1
        m1 = max(t, f_in)
        if e_in > m1: bits |= 2
        h = max(m1, e_in)
        hf, f = h - self.QRt(j), f_in - self.Rt(j)
        if f > hf: bits |= 4
        he, e = h - self.QRq(i), e_in - self.Rq(i)
        if e > he: bits |= 8
        return h, max(hf, f), max(he, e), bits


def forward(p, R, KC):
    """score + checkpoints only.  rowck[b][j] = (H(bR-1, j), F entering (bR, j)); colck[c][i] =
    (H(i, cKC-1), E entering (i, cKC)); b >= 1, c >= 1 (the borders are analytic)."""
    Q, D = p.Q, p.D
    rowck = {b: [None] * D for b in range(1, (Q + R - 1) // R)}
    colck = {c: [None] * Q for c in range(1, (D + KC - 1) // KC)}
    hprev = [p.Hleft(i) for i in range(Q)]                    # H(i, j-1)
    e_in = [p.Hleft(i) - p.QRq(i) for i in range(Q)]          # E entering (i, j)
    score = None
    for j in range(D):
        if j % KC == 0 and j > 0:
            for i in range(Q):
                colck[j // KC][i] = (hprev[i], e_in[i])
        hdiag = p.Htop(j - 1)
        f_in = p.Htop(j) - p.QRt(j)
        habove = p.Htop(j)
        for i in range(Q):
            if i % R == 0 and i > 0:
                rowck[i // R][j] = (habove, f_in)
            h, f_out, e_out, _ = p.cell(i, j, hdiag, f_in, e_in[i])
            hdiag = hprev[i]
            hprev[i] = h
            e_in[i] = e_out
            f_in = f_out
            habove = h
        score = hprev[Q - 1]
    return score, rowck, colck


def traceback(p, R, KC, rowck, colck):
    """backtrack16's walk (align_simd.cpp:1132-1245) over regenerated tiles; returns ops (reversed order) and
    the number of cells recomputed"""
    Q, D = p.Q, p.D
    i, j = Q - 1, D - 1
    op = ""
    ops = []
    recomputed = 0
    while i >= 0 and j >= 0:
        b, c = i // R, j // KC
        i0, j0 = b * R, c * KC
        ni, nj = i - i0 + 1, j - j0 + 1
        # borders of the sub-rectangle rows [i0, i] x cols [j0, j]
        def top(jj):      # H(i0-1, jj), jj >= j0-1
            if b == 0: return p.Htop(jj)
            if jj < 0: return p.Hleft(i0 - 1)
            return rowck[b][jj][0]
        def ftop(jj):     # F entering (i0, jj)
            return p.Htop(jj) - p.QRt(jj) if b == 0 else rowck[b][jj][1]
        def left(ii):     # H(ii, j0-1)
            return p.Hleft(ii) if c == 0 else colck[c][ii][0]
        def eleft(ii):    # E entering (ii, j0)
            return p.Hleft(ii) - p.QRq(ii) if c == 0 else colck[c][ii][1]
        bits = [[0] * nj for _ in range(ni)]
        hcol = [left(i0 + a) for a in range(ni)]
        ecol = [eleft(i0 + a) for a in range(ni)]
        for bj in range(nj):
            jj = j0 + bj
            hdiag = top(jj - 1)
            f_in = ftop(jj)
            for a in range(ni):
                h, f_out, e_out, bt = p.cell(i0 + a, jj, hdiag, f_in, ecol[a])
                bits[a][bj] = bt
                hdiag = hcol[a]
                hcol[a] = h
                ecol[a] = e_out
                f_in = f_out
        recomputed += ni * nj
        while i >= i0 and j >= j0:
            d = bits[i - i0][j - j0]
            if op == "I" and (d & 8): j -= 1; nop = "I"
            elif op == "D" and (d & 4): i -= 1; nop = "D"
            elif d & 2: j -= 1; nop = "I"
            elif d & 1: i -= 1; nop = "D"
            else: i -= 1; j -= 1; nop = "M"
            ops.append(nop)
            op = nop
    while i >= 0:
        ops.append("D"); i -= 1
    while j >= 0:
        ops.append("I"); j -= 1
    return ops, recomputed


def cigar_of(ops_reversed):
    out = []
    ops = ops_reversed[::-1]
    k = 0
    while k < len(ops):
        m = k
        while m < len(ops) and ops[m] == ops[k]:
            m += 1
        out.append((str(m - k) if m - k > 1 else "") + ops[k])
        k = m
    return "".join(out)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rng = np.random.default_rng(7)
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    iupac = np.frombuffer(b"ACGTUNRYSWKMBDHVacgtn", dtype=np.uint8)
    bad = 0
    tot_cells = tot_re = 0
    for k in range(n):
        pen = checkers.DEFAULT_PEN.copy()
        if k % 3 == 2:
            pen = np.array([int(rng.integers(1, 5)), -int(rng.integers(1, 7))] + [int(rng.integers(0, 22)) for _ in range(6)]
                           + [int(rng.integers(0, 4)) for _ in range(6)], dtype=np.int64)
        A = iupac if k % 5 == 4 else alpha
        ql, dl = int(rng.integers(1, 90)), int(rng.integers(1, 200))
        if k % 2 == 0:   # related pair
            root = A[rng.integers(0, len(A), size=max(ql, dl))]
            q = root[:ql].copy(); t = root[:dl].copy()
            for s in (q, t):
                m = rng.random(len(s)) < 0.1
                s[m] = A[rng.integers(0, len(A), size=int(m.sum()))]
            q, t = q.tobytes(), t.tobytes()
        else:
            q = A[rng.integers(0, len(A), size=ql)].tobytes(); t = A[rng.integers(0, len(A), size=dl)].tobytes()
        R = int(rng.choice([1, 2, 3, 8])); KC = int(rng.choice([1, 4, 16, 32]))
        p = Problem(q, t, pen, n_mismatch=k % 7 == 6)
        score, rowck, colck = forward(p, R, KC)
        ops, re = traceback(p, R, KC, rowck, colck)
        want = checkers.oracle_nw16(q, t, pen, int(k % 7 == 6))
        got_cigar = cigar_of(ops)
        tot_cells += p.Q * p.D; tot_re += re
        if want[0] == 32767:
            continue   # the reference defers this pair (overflow flag); the exact kernel's business
        if score != want[0] or got_cigar != want[5] or len(ops) != want[1]:
            bad += 1
            if bad <= 5:
                print("MISMATCH", k, len(q), len(t), R, KC, score, want[0], got_cigar, want[5])
    print(f"{n} pairs, {bad} mismatches; cells recomputed in the traceback: {100.0 * tot_re / tot_cells:.1f} % of the matrix")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
