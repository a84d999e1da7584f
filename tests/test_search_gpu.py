"""GPU parity of the k-mer ranker (vsg_rank) and the whole search path (vsg_search_batch) against
the golden fixtures, the oracle and — when oracle/_ref travelled along — the unmodified reference."""
import json
import os

import numpy as np
import pytest

import checkers
from vsearch_b200 import lib as vlib
from vsearch_b200 import synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return json.load(open(os.path.join(G, name)))


@pytest.fixture(scope="module")
def ctx():
    c = vlib.Context(0)
    yield c
    c.close()


def rows_of(res, counts, q, max_results):
    out = []
    for j in range(int(counts[q])):
        r = res[q * max_results + j]
        out.append([r.target, r.id, r.matches, r.mismatches, r.gaps, r.alignment_length, r.accepted, r.strand])
    return out


def gpu_opts(id, maxaccepts, maxrejects, strand_both=0, mask_lower=0, k=8):
    o = vlib.default_search_opts()
    o.id = id; o.maxaccepts = maxaccepts; o.maxrejects = maxrejects; o.strand_both = strand_both
    o.mask_lower = mask_lower; o.wordlength = k
    return o


def test_rank_and_search_golden(ctx):
    g = load("rank_search_vectors.json")
    dbs = synth.SeqSet([d.encode() for d in g["db"]])
    qss = synth.SeqSet([q.encode() for q in g["queries"]])
    db = ctx.seqset(dbs); qs = ctx.seqset(qss)
    ix = ctx.index(db, 8, 0)
    nq = len(qss)
    for case in g["cases"]:
        th = case["tophits"]
        seqno, count, nc = ctx.rank(ix, qs, 0, nq, 12, th)
        for i in range(nq):
            assert [seqno[i, :nc[i]].tolist(), count[i, :nc[i]].tolist()] == case["topscores"][i], (case["id"], i)
        o = gpu_opts(case["id"], case["maxaccepts"], case["maxrejects"], case["strand_both"])
        res, counts, work = ctx.search(ix, db, qs, 0, nq, o, th)
        for i in range(nq):
            got = rows_of(res, counts, i, th)
            want = case["rows"][i]
            if case["strand_both"]:
                assert sorted(map(tuple, got)) == sorted(map(tuple, want)), (case["id"], i)
                assert [r[1] for r in got] == sorted([r[1] for r in got], reverse=True)
            else:
                assert got == want, (case["id"], i, got, want)
        assert work[0] > 0 and work[1] > 0
    ix.close(); db.close(); qs.close()


def test_reference_api_example_golden(ctx):
    g = load("search_api_example.json")
    p = g["params"]
    db = ctx.seqset(synth.SeqSet([s.encode() for s in g["ref_seqs"]]))
    qs = ctx.seqset(synth.SeqSet([s.encode() for s in g["query_seqs"]]))
    ix = ctx.index(db, p["wordlength"], 1)
    o = gpu_opts(p["id"], p["maxaccepts"], p["maxrejects"], mask_lower=1, k=p["wordlength"])
    res, counts, _ = ctx.search(ix, db, qs, 0, len(g["query_seqs"]), o, p["max_results"])
    got = []
    for i, ql in enumerate(g["query_labels"]):
        for r in rows_of(res, counts, i, p["max_results"]):
            got.append([ql, g["ref_labels"][r[0]], f"{r[1]:.1f}"])
    assert sorted(got) == sorted(g["expected_rows"])
    # and the full-precision rows the reference library returned when the fixture was made
    for i in range(len(g["query_labels"])):
        assert rows_of(res, counts, i, p["max_results"]) == g["full_rows"][i]
    ix.close(); db.close(); qs.close()


def test_multi_shard_database_vs_oracle(ctx):
    """> 32768 targets: several index shards, candidate-list overflow handling, ties on count"""
    rng = np.random.default_rng(31)
    roots = synth.random_seqs(rng, 40, 120)
    n = 70000
    pick = rng.integers(0, 40, size=n)
    seqs = []
    for i in range(n):
        s = roots[pick[i]].copy()
        pos = rng.integers(0, 120, size=3)
        s[pos] = synth.ACGT[rng.integers(0, 4, size=3)]
        seqs.append(s[: int(rng.integers(90, 121))].tobytes())
    dbs = synth.SeqSet(seqs)
    queries = [synth.mutate(rng, roots[i % 40], 0.03).tobytes() for i in range(24)]
    queries += [synth.random_seqs(rng, 1, 100)[0].tobytes(), b"ACGTACG"]   # no hit / shorter than k
    qss = synth.SeqSet(queries)
    db = ctx.seqset(dbs); qs = ctx.seqset(qss)
    ix = ctx.index(db, 8, 0)
    od = checkers.OracleDb(dbs)
    opts = checkers.search_opts(n, id=0.95, maxaccepts=2, maxrejects=8)
    seqno, count, nc = ctx.rank(ix, qs, 0, len(queries), opts.minwordmatches, opts.tophits)
    for i, q in enumerate(queries):
        s, c = od.topscores(q, opts)
        assert seqno[i, :nc[i]].tolist() == s.tolist() and count[i, :nc[i]].tolist() == c.tolist(), i
    o = gpu_opts(0.95, 2, 8)
    res, counts, work = ctx.search(ix, db, qs, 0, len(queries), o, opts.tophits)
    pairs = cells = 0
    for i, q in enumerate(queries):
        hits, p, cl = od.search(q, opts)
        pairs += p; cells += cl
        want = [[h.target, h.id, h.matches, h.mismatches, h.nwgaps, h.nwalignmentlength, h.accepted, h.strand]
                for h in hits]
        assert rows_of(res, counts, i, opts.tophits) == want, i
    assert (int(work[0]), int(work[1])) == (pairs, cells)   # same search16 workload as the reference's driver
    od.close(); ix.close(); db.close(); qs.close()


@pytest.mark.skipif(checkers.ref() is None, reason="oracle/_ref/libvsref.so not present")
def test_search_vs_compiled_reference(ctx):
    """config-2 shape in miniature: 250-nt windows of a random 1500-nt database, 5 % mutated"""
    dbs, qss, src = synth.config2_search(n_db=400, db_len=1500, n_q=120, q_len=250, div=0.05, seed=77)
    r = checkers.RefDb(dbs, id=0.9, maxaccepts=1, maxrejects=32)
    want = r.search(qss, max_results=r.tophits)
    th = r.tophits
    r.close()
    db = ctx.seqset(dbs); qs = ctx.seqset(qss)
    ix = ctx.index(db, 8, 0)
    res, counts, _ = ctx.search(ix, db, qs, 0, len(qss), gpu_opts(0.9, 1, 32), th)
    hit = 0
    for i in range(len(qss)):
        got = rows_of(res, counts, i, th)
        assert got == [list(t) for t in want[i]], i
        hit += bool(got) and got[0][0] == int(src[i])
    assert hit > 100
    ix.close(); db.close(); qs.close()
