"""Stage-by-stage timing of the search path on the C2 workload (diagnostic, not the bench)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vsearch_b200 import lib as vlib, synth

NQ = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
C4 = "--c4" in sys.argv   # configs[3] shape: 1M x 1200 DB (31 shards), 150-nt queries
dbm = synth.config2_db(1_000_000, 1200, 4) if C4 else synth.config2_db(100_000, 1500, 2024)
ctx = vlib.Context(0)
db = ctx.seqset(synth.SeqSet.from_matrix(dbm))
t0 = time.time(); ix = ctx.index(db, 8, 0); print(f"index build {1e3*(time.time()-t0):.0f} ms")
qs_h, _ = synth.config2_query_batch(dbm, NQ, 150, 0.10, batch=1) if C4 else synth.config2_query_batch(dbm, NQ, batch=1)
qs = ctx.seqset(qs_h)
for rep in range(3):
    ctx.profile_reset(); t0 = time.time()
    seqno, count, nc = ctx.rank(ix, qs, 0, NQ, 12, 41)
    w = time.time() - t0; p = ctx.profile()
    print(f"rank  {NQ} queries: kernel {p.rank_ms:.1f} ms wall {1e3*w:.1f} ms  -> {NQ/p.rank_ms*1e3/1e6:.2f} Mq/s kernel; mean cands {nc.mean():.1f}")
qi = np.repeat(np.arange(NQ, dtype=np.uint32), 8)
ti = seqno[:, :8].reshape(-1).astype(np.uint32)
for rep in range(3):
    t0 = time.time()
    res = ctx.align_pairs(qs, db, qi, ti)
    w = time.time() - t0
    print(f"align {qi.shape[0]} pairs: fwd {res.fwd_ms:.1f} ms tb {res.tb_ms:.1f} ms wall {1e3*w:.1f} ms -> fwd {res.cells/res.fwd_ms/1e6:.0f} GCUPS")
if "--short" in sys.argv:
    sys.exit(0)
opts = vlib.default_search_opts(); opts.id = 0.9
for thr in (1, 2, 4, 8):
    os.environ["VSG_HOST_THREADS"] = str(thr)
    for rep in range(2):
        ctx.profile_reset(); t0 = time.time()
        r, counts, work = ctx.search(ix, db, qs, 0, NQ, opts, 1)
        w = time.time() - t0; p = ctx.profile()
    print(f"search threads={thr}: wall {1e3*w:.1f} ms ({NQ/w/1e3:.0f} kq/s, {work[1]/w/1e9:.0f} GCUPS) kernels fwd {p.fwd_ms:.0f} tb {p.traceback_ms:.0f} rank {p.rank_ms:.0f} ms")
