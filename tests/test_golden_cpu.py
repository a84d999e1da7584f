"""The oracle against the committed golden fixtures (tests/golden/, generated from the unmodified
reference by tools/make_golden.py) — runs anywhere, no reference and no GPU needed."""
import json
import os

import numpy as np

import checkers
from vsearch_b200 import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return json.load(open(os.path.join(G, name)))


def test_nw16_known_answers():
    vec = load("nw16_vectors.json")["vectors"]
    assert len(vec) > 300
    for v in vec:
        pen = None if v["pen"] is None else np.array(v["pen"], dtype=np.int64)
        o = checkers.oracle_nw16(v["q"].encode("latin1"), v["t"].encode("latin1"), pen, v["nm"])
        assert list(o) == v["out"], v


def _hits_rows(hits):
    return [[h.target, h.id, h.matches, h.mismatches, h.nwgaps, h.nwalignmentlength, h.accepted, h.strand]
            for h in hits]


def test_topscores_and_search_known_answers():
    g = load("rank_search_vectors.json")
    db = synth.SeqSet([d.encode() for d in g["db"]])
    queries = [q.encode() for q in g["queries"]]
    o = checkers.OracleDb(db)
    for case in g["cases"]:
        opts = checkers.search_opts(len(db), id=case["id"], maxaccepts=case["maxaccepts"],
                                    maxrejects=case["maxrejects"])
        assert opts.tophits == case["tophits"]
        for i, q in enumerate(queries):
            s, c = o.topscores(q, opts)
            assert [s.tolist(), c.tolist()] == case["topscores"][i]
            hits, _, _ = o.search(q, opts)
            rows = _hits_rows(hits)
            if case["strand_both"]:
                rc = bytes(checkers.oracle().oracle_complement(b) for b in reversed(q))
                h2, _, _ = o.search(rc, opts, strand=1)
                rows = sorted(rows + _hits_rows(h2), key=lambda r: (-r[1], r[0]))
                want = sorted(case["rows"][i], key=lambda r: (-r[1], r[0]))
                assert sorted(map(tuple, rows)) == sorted(map(tuple, want))
            else:
                assert rows == case["rows"][i], (case["id"], i)
    o.close()


def test_reference_api_example_golden():
    """api_examples/example_search.cc part 1: rows of data/expected_search.tsv (id to 0.1 %)"""
    g = load("search_api_example.json")
    p = g["params"]
    db = synth.SeqSet([s.encode() for s in g["ref_seqs"]])
    o = checkers.OracleDb(db, k=p["wordlength"], mask_lower=1)
    opts = checkers.search_opts(len(db), id=p["id"], maxaccepts=p["maxaccepts"], maxrejects=p["maxrejects"],
                                k=p["wordlength"], mask_lower=1)
    got = []
    for ql, qs in zip(g["query_labels"], g["query_seqs"]):
        hits, _, _ = o.search(qs.encode(), opts)
        for h in hits[: p["max_results"]]:
            got.append([ql, g["ref_labels"][h.target], f"{h.id:.1f}"])
    assert sorted(got) == sorted(g["expected_rows"])
    o.close()
