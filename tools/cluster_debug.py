"""debug aid: first sequence (in processing order) whose cluster record differs between vsg_cluster_fast and the reference CLI"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vsearch_b200 import lib as vlib, synth
N = int(sys.argv[1]); T = int(sys.argv[2])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STOCK = os.path.join(ROOT, "oracle", "_ref", "vsearch")
rng = np.random.default_rng([3, 0])
nroots = max(50, N // 200)
roots = synth.random_seqs(rng, nroots, 300)
w = 1.0 / np.arange(1, nroots + 1); w /= w.sum()
reads = synth.mutate_batch(rng, roots[rng.choice(nroots, size=N, p=w)], 0.01)
labels = [f"a{i:08d}" for i in range(N)]
fa = "/tmp/dbg.fasta"
with open(fa, "wb") as f:
    for i in range(N):
        f.write(b">" + labels[i].encode() + b"\n" + reads.seq(i) + b"\n")
subprocess.run([STOCK, "--cluster_fast", fa, "--id", "0.97", "--threads", str(T), "--uc", "/tmp/dbg.uc", "--quiet"], check=True)
want = {}
for line in open("/tmp/dbg.uc"):
    f = line.rstrip("\n").split("\t")
    if f[0] == "S": want[f[8]] = ("S", int(f[1]), "*", "*")
    elif f[0] == "H": want[f[8]] = ("H", int(f[1]), f[3], f[9])
order = np.lexsort((np.arange(N), -reads.lens.astype(np.int64)))
ctx = vlib.Context(0)
ss = ctx.seqset(synth.SeqSet([reads.seq(int(i)) for i in order])); ss.dust()
o = vlib.default_search_opts(); o.id = 0.97; o.mask_lower = 1; o.maxrejects = 8   # --cluster_fast default (cli.cc:4163-4172)
res, ncl, work = vlib.cluster_fast(ctx, ss, o, T)
print("clusters", ncl, "ref", sum(1 for v in want.values() if v[0] == "S"))
nbad = 0
for k in range(N):
    lab = labels[order[k]]
    got = ("S", int(res["cluster"][k]), "*", "*") if res["centroid"][k] < 0 else ("H", int(res["cluster"][k]), f"{res['id'][k]:.1f}", labels[order[int(res['centroid'][k])]])
    if got != want[lab]:
        nbad += 1
        if nbad <= 5: print("pos", k, "round", k // T, "in-round", k % T, lab, "len", reads.lens[order[k]], "got", got, "want", want[lab])
print("mismatches", nbad)

# ---- dissect the first mismatch with the reference library (search against the centroids known at its round's start)
if nbad:
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import checkers
    first = next(k for k in range(N) if (("S", int(res["cluster"][k]), "*", "*") if res["centroid"][k] < 0 else ("H", int(res["cluster"][k]), f"{res['id'][k]:.1f}", labels[order[int(res['centroid'][k])]])) != want[labels[order[k]]])
    r0 = (first // T) * T
    cents = [k for k in range(r0) if res["centroid"][k] < 0]
    extras = [k for k in range(r0, first) if res["centroid"][k] < 0]
    print("first", first, "round start", r0, "centroids before round", len(cents), "extras before it in round", extras)
    seqs = [reads.seq(int(order[k])) for k in cents]
    dbs = synth.SeqSet(seqs)
    rdb = checkers.RefDb(dbs, id=0.97, maxaccepts=1, maxrejects=32, dust=1)
    q = reads.seq(int(order[first]))
    qss = synth.SeqSet([q])
    print("ref search rows:", rdb.search(qss, max_results=8))
    rs, rc = rdb.topscores(q)   # NOTE: the query is not dusted on this path
    rdb.close()
    cdb = ctx.seqset(dbs); cdb.dust(); cq = ctx.seqset(qss); cq.dust()
    cix = ctx.index(cdb, 8, 1)
    s_, c_, n_ = ctx.rank(cix, cq, 0, 1, 12, 41, mask_lower=1)
    print("gpu static rank:", list(zip(s_[0, :n_[0]].tolist(), c_[0, :n_[0]].tolist()))[:41])
    o2 = vlib.default_search_opts(); o2.id = 0.97; o2.mask_lower = 1
    rr, cc, ww = ctx.search(cix, cdb, cq, 0, 1, o2, 8)
    print("gpu static search:", [(rr[j].target, rr[j].id, rr[j].accepted) for j in range(cc[0])])
    tgt = int(res["centroid"][first])
    print("our accepted target pos", tgt, "index among centroids", cents.index(tgt) if tgt in cents else None)
    a = ctx.align_pairs(cq, cdb, np.zeros(len(cents), dtype=np.uint32), np.arange(len(cents), dtype=np.uint32))
    ids = 100.0 * a.matches / np.maximum(1, a.aligned.astype(np.int64) - a.trims.sum(axis=1))
    best = np.argsort(-ids)[:6]
    print("best ids over all centroids:", [(int(b), float(ids[b]), int(a.matches[b]), int(a.aligned[b]), a.trims[b].tolist()) for b in best])
    # the extras: shared k-mers (dusted sequences) and alignment with the query
    sym = ss.symbols(int(ss_total)) if False else None
    allsym = ss.symbols(int(np.sum(reads.lens)))
    offs = np.zeros(N + 1, dtype=np.int64); np.cumsum(reads.lens[order], out=offs[1:])
    def kms(k_):
        s = allsym[offs[k_]:offs[k_ + 1]]
        out = set()
        for p in range(len(s) - 7):
            w_ = s[p:p + 8]
            if all((int(x) & 15) in (1, 2, 4, 8) and not (int(x) & 16) for x in w_):
                v = 0
                for x in w_:
                    v = v * 4 + {1: 0, 2: 1, 4: 2, 8: 3}[int(x) & 15]
                out.add(v)
        return out
    kq = kms(first)
    for e in extras:
        ke = kms(e)
        al = ctx.align_pairs(ss, ss, np.array([first], dtype=np.uint32), np.array([e], dtype=np.uint32), cigar=True)
        print("extra", e, "len", reads.lens[order[e]], "shared", len(kq & ke), "query kmers", len(kq), "align: matches", int(al.matches[0]), "aligned", int(al.aligned[0]), "trims", al.trims[0].tolist(), al.cigars[0])
    print("lens of top candidates:", [(int(cents[t]), int(reads.lens[order[cents[t]]])) for t in s_[0, :12].tolist()])
    # reduced run: the centroids known so far + the query, through the CLI again
    red = "/tmp/dbg_red.fasta"
    with open(red, "wb") as f:
        for k_ in cents + extras + [first]:
            f.write(b">" + labels[order[k_]].encode() + b"\n" + reads.seq(int(order[k_])) + b"\n")
    for extra_args in ([], ["--qmask", "none"], ["--maxrejects", "64"], ["--maxaccepts", "2"]):
        subprocess.run([STOCK, "--cluster_fast", red, "--id", "0.97", "--threads", "1", "--uc", "/tmp/dbg_red.uc", "--quiet"] + extra_args, check=True)
        lines = [l.rstrip("\n") for l in open("/tmp/dbg_red.uc") if labels[order[first]] in l.split("\t")[8:9]]
        ns = sum(1 for l in open("/tmp/dbg_red.uc") if l.startswith("S"))
        print("reduced CLI", extra_args, "S records", ns, "->", lines)
