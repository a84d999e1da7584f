"""Sweep of the search driver's host knobs on the C2 workload (diagnostic, not the bench)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vsearch_b200 import lib as vlib, synth
NQ = 65536
dbm = synth.config2_db(100_000, 1500, 2024)
ctx = vlib.Context(0)
db = ctx.seqset(synth.SeqSet.from_matrix(dbm)); ix = ctx.index(db, 8, 0)
qs_h, _ = synth.config2_query_batch(dbm, NQ, batch=1); qs = ctx.seqset(qs_h)
opts = vlib.default_search_opts(); opts.id = 0.9
for sub in (2048, 4096, 8192):
    for thr in (6, 8, 12, 16):
        os.environ["VSG_HOST_THREADS"] = str(thr); os.environ["VSG_SUBBATCH"] = str(sub)
        best = 1e9
        for rep in range(5):
            t0 = time.time(); r, c, w = ctx.search(ix, db, qs, 0, NQ, opts, 1); best = min(best, time.time() - t0)
        print(f"subbatch {sub:6d} threads {thr:2d}: {1e3*best:.1f} ms  {w[1]/best/1e9:.0f} GCUPS  {NQ/best/1e3:.0f} kq/s", flush=True)
