"""Generate tests/golden/*.json from the UNMODIFIED reference (run in the build container only:
needs /root/reference and oracle/_ref/libvsref.so).  The fixtures travel to the GPU box; the
reference does not.

  search_api_example.json  inputs + expected rows of the reference's own golden test
                           (api_examples/example_search.cc:69-127, data/expected_search.tsv);
                           sequences are stored AFTER the reference's DUST pass (soft-masked,
                           lower case) because masking is outside the accelerated path
  nw16_vectors.json        search16 known answers (score, stats, CIGAR) incl. edge cases
  rank_search_vectors.json search_topscores lists and search_session_single rows
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import checkers  # noqa: E402
from vsearch_b200 import synth  # noqa: E402

REFDATA = "/root/reference/api_examples/data"
OUT = os.path.join(ROOT, "tests", "golden")


def read_fasta(path):
    labels, seqs = [], []
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            labels.append(line[1:]); seqs.append("")
        elif line:
            seqs[-1] += line
    return labels, seqs


def dust(seq: str) -> str:
    b = C.create_string_buffer(seq.encode())
    checkers.ref().vsref_dust(b, C.c_int(len(seq)))
    return b.value.decode()


def rand_seq(rng, n, alphabet=b"ACGT"):
    a = np.frombuffer(alphabet, dtype=np.uint8)
    return a[rng.integers(0, a.shape[0], size=n)].tobytes()


def main():
    os.makedirs(OUT, exist_ok=True)
    assert checkers.ref() is not None, "build oracle/_ref first (make -C oracle ref)"

    # 1. the reference's own search golden
    rl, rs = read_fasta(os.path.join(REFDATA, "chimera_ref.fasta"))
    ql, qs = read_fasta(os.path.join(REFDATA, "chimera_queries.fasta"))
    expected = [l.rstrip("\n").split("\t") for l in open(os.path.join(REFDATA, "expected_search.tsv"))]
    rs_m = [dust(s) for s in rs]
    qs_m = [dust(s) for s in qs]
    # cross-check: the reference library on these inputs reproduces its golden file
    db = synth.SeqSet([s.encode() for s in rs])
    r = checkers.RefDb(db, k=8, id=0.5, maxaccepts=3, maxrejects=16, dust=1)
    rows = r.search(synth.SeqSet([s.encode() for s in qs]), max_results=3)
    r.close()
    got = sorted(f"{ql[i]}\t{rl[t[0]]}\t{t[1]:.1f}" for i, rr in enumerate(rows) for t in rr)
    assert got == sorted("\t".join(e) for e in expected), "reference does not reproduce its own golden?"
    json.dump({"source": "api_examples/data/{chimera_ref,chimera_queries}.fasta + expected_search.tsv; "
                         "sequences after the reference's DUST pass",
               "params": {"wordlength": 8, "id": 0.5, "maxaccepts": 3, "maxrejects": 16, "max_results": 3},
               "ref_labels": rl, "ref_seqs": rs_m, "query_labels": ql, "query_seqs": qs_m,
               "expected_rows": expected,
               "full_rows": [[list(t) for t in rr] for rr in rows]},
              open(os.path.join(OUT, "search_api_example.json"), "w"), indent=0)

    # 2. search16 known answers
    rng = np.random.default_rng(20260922)
    iupac = b"ACGTUacgtuNnRYSWKMBDHVryswkmbdhvXx-"
    vec = []

    def add(q, ts, pen=None, nm=0):
        res = checkers.ref_search16(q, ts, pen, nm)
        for t, o in zip(ts, res):
            vec.append({"q": q.decode("latin1"), "t": t.decode("latin1"),
                        "pen": None if pen is None else [int(x) for x in pen], "nm": nm, "out": list(o)})

    for _ in range(40):
        L = int(rng.integers(1, 400))
        root = np.frombuffer(rand_seq(rng, L), dtype=np.uint8)
        q = synth.mutate(rng, root, 0.1).tobytes() or b"A"
        add(q, [synth.mutate(rng, root, float(rng.uniform(0, 0.35))).tobytes() for _ in range(4)]
            + [rand_seq(rng, int(rng.integers(1, 500)))])
    for nm in (0, 1):
        for _ in range(12):
            add(rand_seq(rng, int(rng.integers(1, 150)), iupac),
                [rand_seq(rng, int(rng.integers(1, 150)), iupac) for _ in range(4)], nm=nm)
    add(rand_seq(rng, 37), [b"", b"A", b"AC", b"ACG", b"ACGT", b"ACGTA"])
    add(b"", [b"", b"A", rand_seq(rng, 77)])
    add(b"A" * 50, [b"A" * 40, b"A" * 60, b"AT" * 25, b"T" * 50])
    add(b"ACAC" * 20, [b"CACA" * 20, b"AC" * 33, b"ACC" * 20])
    add(rand_seq(rng, 600), [rand_seq(rng, 700), rand_seq(rng, 1500)])       # multi-strip queries
    add(rand_seq(rng, 250), [rand_seq(rng, 1500), rand_seq(rng, 1499), rand_seq(rng, 1501)])
    pen_big = np.array([2, -4, 3000, 3000, 5000, 5000, 3000, 3000, 600, 600, 900, 900, 600, 600])
    for L in (10, 30, 60):
        add(rand_seq(rng, L), [rand_seq(rng, int(rng.integers(1, 2 * L))) for _ in range(6)], pen_big)
    pen_hi = np.array([3000, -3000, 1, 1, 18, 18, 1, 1, 1, 1, 2, 2, 1, 1])
    q = rand_seq(rng, 40)
    add(q, [q, q[:20], q + q], pen_hi)
    for _ in range(10):
        pen = np.array([int(rng.integers(1, 6)), -int(rng.integers(1, 8))]
                       + [int(rng.integers(0, 25)) for _ in range(6)] + [int(rng.integers(0, 5)) for _ in range(6)])
        root = np.frombuffer(rand_seq(rng, int(rng.integers(5, 200))), dtype=np.uint8)
        add(synth.mutate(rng, root, 0.15).tobytes() or b"C",
            [synth.mutate(rng, root, 0.25).tobytes() or b"G" for _ in range(4)], pen)
    json.dump({"source": "search16 (core/align_simd.cpp) through oracle/ref_shim.cpp", "vectors": vec},
              open(os.path.join(OUT, "nw16_vectors.json"), "w"))

    # 3. ranker + whole search
    roots = synth.random_seqs(rng, 10, 300)
    dbs = []
    for rr in range(10):
        for _ in range(7):
            dbs.append(synth.mutate(rng, roots[rr], float(rng.uniform(0, 0.2))).tobytes())
    dbs += [b"ACGT", b"N" * 50, dbs[0], dbs[1][:100], b"ACGTNNNNACGT" * 10, dbs[3].lower()]
    dbset = synth.SeqSet(dbs)
    queries = [synth.mutate(rng, roots[i % 10], 0.08).tobytes()[: int(rng.integers(60, 300))] for i in range(30)]
    queries += [b"ACGTACGTAC", b"ACG", synth.random_seqs(rng, 1, 200)[0].tobytes(), queries[0].lower()]
    cases = []
    for (idv, ma, mr, both) in ((0.9, 1, 32, 0), (0.5, 3, 16, 0), (0.97, 2, 4, 0), (0.8, 10, 10, 1)):
        r = checkers.RefDb(dbset, id=idv, maxaccepts=ma, maxrejects=mr, strand_both=both)
        tops = []
        for qq in queries:
            s1, c1 = r.topscores(qq)
            tops.append([s1.tolist(), c1.tolist()])
        rows = r.search(synth.SeqSet(queries), max_results=r.tophits)
        cases.append({"id": idv, "maxaccepts": ma, "maxrejects": mr, "strand_both": both, "tophits": r.tophits,
                      "topscores": tops, "rows": [[list(t) for t in rr] for rr in rows]})
        r.close()
    json.dump({"source": "search_topscores / search_session_single through oracle/ref_shim.cpp, masking none",
               "db": [d.decode() for d in dbs], "queries": [q.decode() for q in queries], "cases": cases},
              open(os.path.join(OUT, "rank_search_vectors.json"), "w"))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
