"""Seam 2 of the drop-in boundary on the GPU: the reference's library entry point search_batch()
(src/core/search.hpp:135-145) replaced by shim/search_batch_vsg.cpp + libvsg.so.

* the reference's OWN api_examples/example_search.cc, compiled unmodified and linked against the reference
  objects with both shims (oracle/Makefile: _ref/example_search_gpu), must reproduce the reference's golden
  api_examples/data/expected_search.tsv and pass its self-checks (batch == sequential, strand semantics);
* oracle/seam2_driver.cpp, linked once against the untouched reference and once against the shims, must print
  identical result records for option sets that exercise every filter of search_acceptable_unaligned /
  search_acceptable_aligned, both strands, the masking modes and the '*' gap penalties.
"""
import os
import subprocess

import numpy as np
import pytest

from vsearch_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")

needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "seam2_driver_gpu")),
                               reason="oracle/_ref (compiled reference + shims) not present")


@needs_ref
def test_reference_example_search_runs_on_the_gpu_shims():
    cwd = os.path.join(REF, "api_data")
    r = subprocess.run([os.path.join(REF, "example_search_gpu")], cwd=cwd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    want = sorted(open(os.path.join(cwd, "data", "expected_search.tsv")).read().splitlines())
    assert sorted(r.stdout.splitlines()) == want
    assert "PASS: batch search matches sequential search" in r.stderr
    assert "PASS: opt_strand=false (plus): RC query correctly found no hit" in r.stderr
    assert "PASS: opt_strand=true (both): RC query matched on minus strand" in r.stderr
    assert "FAIL" not in r.stderr


def _write_fasta(path, recs):
    with open(path, "w") as f:
        for head, seq in recs:
            f.write(f">{head}\n{seq}\n")


def _dataset(tmp_path):
    rng = np.random.default_rng(2025)
    comp = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")
    roots = synth.random_seqs(rng, 10, 420)
    low = b"AT" * 40 + b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA"          # DUST bait
    db = []
    for i in range(240):
        m = synth.mutate(rng, roots[i % 10], float(rng.uniform(0.0, 0.12))).tobytes()
        a, b = int(rng.integers(0, 40)), int(rng.integers(0, 40))
        s = m[a: len(m) - b]
        if i % 17 == 3:
            s = s[:150] + low + s[150:]
        if i % 29 == 5:
            s = s[:60] + b"NNNRYKM" + s[67:]                          # IUPAC -> the general kernel
        db.append([f"t{i};size={int(rng.integers(1, 60))}", s.decode()])
    queries = []
    for i in range(48):
        m = synth.mutate(rng, roots[i % 10], 0.04).tobytes()[int(rng.integers(0, 30)):]
        if i % 11 == 2:
            m = m[:100] + low + m[100:]
        if i % 4 == 1:
            m = m[::-1].translate(comp)                               # minus-strand queries
        if i % 9 == 4:
            m = m[:50].lower() + m[50:]                               # soft-masked input
        queries.append([f"q{i};size={int(rng.integers(1, 60))}", m.decode()])
    # identical sequences and shared labels for --self / --selfid / --idprefix / --idsuffix
    for k in range(6):
        head, seq = db[7 * k + 1]
        queries.append([head if k % 2 == 0 else f"dup{k};size=9", seq])
    for k in range(4):
        head, seq = db[5 * k + 2]
        queries.append([f"pre{k};size=3", seq[:12] + synth.mutate(rng, np.frombuffer(seq[12:-9].encode(), dtype=np.uint8), 0.05).tobytes().decode() + seq[-9:]])
    dbf, qf = str(tmp_path / "db.fasta"), str(tmp_path / "q.fasta")
    _write_fasta(dbf, db)
    _write_fasta(qf, queries)
    return dbf, qf


CASES = [
    ["id=0.8", "maxaccepts=3", "maxrejects=8"],                                         # defaults: dust on both sides
    ["id=0.8", "maxaccepts=2", "maxrejects=16", "strand=1"],
    ["id=0.85", "maxaccepts=4", "maxrejects=8", "qmask=none", "dbmask=none", "strand=1", "max_results=3"],
    ["id=0.8", "maxaccepts=3", "maxrejects=8", "self=1"],
    ["id=0.8", "maxaccepts=3", "maxrejects=8", "selfid=1", "strand=1"],
    ["id=0.7", "maxaccepts=5", "maxrejects=8", "idprefix=12", "idsuffix=9"],
    ["id=0.8", "maxaccepts=3", "maxrejects=8", "maxqsize=30", "mintsize=5"],
    ["id=0.8", "maxaccepts=3", "maxrejects=8", "minsizeratio=0.5", "maxsizeratio=3.0"],
    ["id=0.8", "maxaccepts=3", "maxrejects=8", "qmask=soft", "dbmask=soft"],
    ["id=0.8", "maxaccepts=3", "maxrejects=8", "qmask=soft", "hardmask=1", "dbmask=none"],
    ["id=0.8", "maxaccepts=3", "maxrejects=8", "maxsubs=20", "maxgaps=3", "mincols=300", "leftjust=0", "iddef=1"],
    ["id=0.8", "maxaccepts=2", "maxrejects=6", "infinite=qi,ti"],                       # '*' : every pair via the LMA callback
    ["id=0.8", "maxaccepts=2", "maxrejects=6", "infinite_ext=ql,tr", "strand=1"],
    ["id=0.8", "maxaccepts=0", "maxrejects=8"],                                         # library path: a zero limit finds nothing
    ["id=0.9", "maxaccepts=1", "maxrejects=32", "wordlength=7", "threads=3"],
    ["id=0.8", "maxaccepts=3", "maxrejects=8", "unoise_alpha=2.0"],                     # --cluster_unoise acceptance (searchcore.cpp:700-717)
    ["id=0.8", "maxaccepts=2", "maxrejects=16", "unoise_alpha=0.3", "strand=1"],
    ["id=0.9", "maxaccepts=2", "maxrejects=16", "wordlength=12"],                       # sparse index through the shim
]


@needs_ref
@pytest.mark.parametrize("case", CASES, ids=[" ".join(c) for c in CASES])
def test_search_batch_shim_equals_reference(tmp_path, case):
    dbf, qf = _dataset(tmp_path)
    outs = []
    for exe in ("seam2_driver_ref", "seam2_driver_gpu"):
        r = subprocess.run([os.path.join(REF, exe), dbf, qf] + case, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, (exe, r.stdout[-2000:], r.stderr[-2000:])
        outs.append(r.stdout.splitlines())
    assert outs[0] == outs[1], (len(outs[0]), len(outs[1]), [x for x in zip(outs[0], outs[1]) if x[0] != x[1]][:5])
    if "maxaccepts=0" not in case:
        assert len(outs[0]) > 5


# ---- the clustering half of seam 2: cluster_session_* / cluster_assign_* (src/core/cluster.hpp:78-118) ----
needs_cluster_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "seam2_cluster_driver_gpu")),
                                       reason="oracle/_ref (compiled reference + cluster shim) not present")


@needs_cluster_ref
def test_reference_example_cluster_runs_on_the_gpu_shims():
    """the reference's own api_examples/example_cluster.cc, unmodified, against shim/cluster_session_vsg.cpp"""
    cwd = os.path.join(REF, "api_data")
    r = subprocess.run([os.path.join(REF, "example_cluster_gpu")], cwd=cwd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    want = sorted(l for l in open(os.path.join(cwd, "data", "expected_cluster.uc")).read().splitlines() if l[:1] in "SH")
    got = sorted(l for l in r.stdout.splitlines() if l[:1] in "SH")
    assert got == want and len(got) > 0
    assert "PASS: batch cluster matches sequential" in r.stderr
    assert "FAIL" not in r.stderr


def _reads(tmp_path):
    rng = np.random.default_rng(77)
    roots = synth.random_seqs(rng, 40, 320)
    recs = []
    for i in range(1500):
        m = synth.mutate(rng, roots[int(rng.integers(0, 40))], float(rng.uniform(0.0, 0.06))).tobytes()
        a, b = int(rng.integers(0, 25)), int(rng.integers(0, 25))
        s = m[a: len(m) - b]
        if i % 41 == 7:
            s = s[:100] + b"ACACACACACACACACACACACACACACACACACAC" + s[100:]   # DUST bait
        if i % 97 == 11:
            s = s[:50] + b"NRY" + s[53:]                                        # IUPAC -> the general kernel
        recs.append([f"r{i};size={int(rng.integers(1, 200))}", s.decode()])
    path = str(tmp_path / "reads.fasta")
    _write_fasta(path, recs)
    return path


@needs_cluster_ref
@pytest.mark.parametrize("case", [
    ["id=0.97", "threads=1", "chunk=-1"],                      # cluster_assign_single, one by one
    ["id=0.97", "threads=8", "chunk=0"],                       # one cluster_assign_batch over everything
    ["id=0.95", "threads=32", "chunk=257", "maxrejects=16"],   # ranges that do not line up with the rounds
    ["id=0.9", "threads=128", "chunk=700", "qmask=none", "maxaccepts=2", "iddef=1"],
    ["id=0.97", "threads=4", "chunk=400", "minsl=0.9", "mid=0.9"],
    ["id=0.9", "threads=16", "chunk=300", "unoise_alpha=2.0"],   # --cluster_unoise acceptance: abundance skew instead of --id
    ["id=0.9", "threads=1", "chunk=0", "unoise_alpha=0.5", "maxaccepts=2"],
    ["id=0.9", "threads=8", "chunk=0", "maxaccepts=4", "maxrejects=16", "sizeorder=1"],   # the most abundant accepted centroid wins
])
def test_cluster_session_shim_equals_the_reference(tmp_path, case):
    reads = _reads(tmp_path)
    outs = []
    for exe in ("seam2_cluster_driver_ref", "seam2_cluster_driver_gpu"):
        r = subprocess.run([os.path.join(REF, exe), reads] + case, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, (exe, r.stdout[-2000:], r.stderr[-2000:])
        outs.append(r.stdout.splitlines())
    assert len(outs[0]) == 1500
    assert outs[0] == outs[1]
    ncent = sum(1 for l in outs[0] if l.split("\t")[2] == "1")
    assert 30 <= ncent < 1500
