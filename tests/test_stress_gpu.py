"""Randomised stress of vsg_align_pairs against the oracle: ragged lengths from 1 to 2 600, mixed
pure-ACGT / IUPAC sequences (both kernel variants and their pairing rules), odd group sizes,
ungrouped pair lists, duplicate pairs, tiny direction budget (many chunks), several scoring sets."""
import os

import numpy as np
import pytest

import checkers
from vsearch_b200 import lib as vlib
from vsearch_b200 import synth

pytestmark = pytest.mark.gpu
IUPAC = b"ACGTACGTACGTNRYKMacgtn"


def rand_seq(rng, n, alphabet=b"ACGT"):
    a = np.frombuffer(alphabet, dtype=np.uint8)
    return a[rng.integers(0, a.shape[0], size=n)].tobytes()


def make_world(rng, nq, nt):
    lens = [1, 2, 5, 17, 31, 32, 33, 63, 64, 65, 100, 129, 200, 255, 256, 257, 300, 383, 400, 511, 512, 513, 640, 900, 1100, 1600, 2600]
    roots = {L: np.frombuffer(rand_seq(rng, L), dtype=np.uint8) for L in lens}
    def draw():
        L = lens[int(rng.integers(0, len(lens)))]
        kind = rng.random()
        if kind < 0.6:
            s = synth.mutate(rng, roots[L], float(rng.uniform(0, 0.25))).tobytes() or b"A"
        elif kind < 0.8:
            s = rand_seq(rng, L)
        else:
            s = rand_seq(rng, L, IUPAC)
        return s
    return [draw() for _ in range(nq)], [draw() for _ in range(nt)]


@pytest.fixture(params=["checkpoints", "direction-bits"])
def path(request):
    """both aligner paths: checkpoint kernels (threshold 0) and direction-bit kernels (threshold above any batch)"""
    old = os.environ.get("VSG_CKPT_MIN_PAIRS")
    os.environ["VSG_CKPT_MIN_PAIRS"] = "0" if request.param == "checkpoints" else "1000000000"
    yield request.param
    if old is None:
        os.environ.pop("VSG_CKPT_MIN_PAIRS", None)
    else:
        os.environ["VSG_CKPT_MIN_PAIRS"] = old


@pytest.mark.parametrize("seed,pen,budget", [
    (101, None, None),
    (102, [1, -2, 3, 3, 10, 10, 3, 3, 1, 1, 1, 1, 1, 1], "2"),
    (103, [5, -4, 0, 0, 12, 16, 0, 0, 0, 0, 3, 2, 0, 0], None),
])
def test_random_pairs(seed, pen, budget, path):
    rng = np.random.default_rng(seed)
    if budget:
        os.environ["VSG_DIR_BUDGET_MB"] = budget
    try:
        ctx = vlib.Context(0, pen=pen if pen is not None else vlib.DEFAULT_PEN)
    finally:
        os.environ.pop("VSG_DIR_BUDGET_MB", None)
    qseqs, tseqs = make_world(rng, 40, 60)
    pairs = []
    for qi in range(len(qseqs)):                       # grouped, odd and even group sizes
        for ti in rng.choice(len(tseqs), size=int(rng.integers(1, 8)), replace=False):
            pairs.append((qi, int(ti)))
    extra = [(int(rng.integers(0, len(qseqs))), int(rng.integers(0, len(tseqs)))) for _ in range(60)]
    pairs += extra + extra[:10]                        # ungrouped tail with duplicates
    qs = ctx.seqset(synth.SeqSet(qseqs)); ts = ctx.seqset(synth.SeqSet(tseqs))
    qi = np.array([p[0] for p in pairs], dtype=np.uint32); ti = np.array([p[1] for p in pairs], dtype=np.uint32)
    res = ctx.align_pairs(qs, ts, qi, ti, cigar=True)
    res2 = ctx.align_pairs(qs, ts, qi, ti, cigar=False)
    penarr = None if pen is None else np.array(pen, dtype=np.int64)
    cache = {}
    bad = []
    for k, (a, b) in enumerate(pairs):
        if (a, b) not in cache:
            cache[(a, b)] = checkers.oracle_nw16(qseqs[a], tseqs[b], penarr, 0)
        o = cache[(a, b)]
        g = (int(res.score[k]), int(res.aligned[k]), int(res.matches[k]), int(res.mismatches[k]), int(res.gaps[k]), res.cigars[k])
        g2 = (int(res2.score[k]), int(res2.aligned[k]), int(res2.matches[k]), int(res2.mismatches[k]), int(res2.gaps[k]))
        if o != g or g2 != g[:5]:
            bad.append((k, a, b, len(qseqs[a]), len(tseqs[b]), o[:5], g[:5], g2))
    assert not bad, f"{len(bad)} of {len(pairs)} differ: {bad[:4]}"
    assert res.fast_pairs > 0
    qs.close(); ts.close(); ctx.close()
