import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["VSG_TRACE"] = "1"
import numpy as np
from vsearch_b200 import lib as vlib, synth
NQ = 65536
dbm = synth.config2_db(100_000, 1500, 2024)
ctx = vlib.Context(0)
db = ctx.seqset(synth.SeqSet.from_matrix(dbm)); ix = ctx.index(db, 8, 0)
qs_h, _ = synth.config2_query_batch(dbm, NQ, batch=1); qs = ctx.seqset(qs_h)
opts = vlib.default_search_opts(); opts.id = 0.9
for thr in (8,):
    os.environ["VSG_HOST_THREADS"] = str(thr)
    for rep in range(2):
        print(f"--- threads {thr} rep {rep}", file=sys.stderr, flush=True)
        t0 = time.time(); ctx.search(ix, db, qs, 0, NQ, opts, 1); print(f"wall {1e3*(time.time()-t0):.1f} ms", file=sys.stderr, flush=True)
