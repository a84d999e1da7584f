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
seqno, count, nc = ctx.rank(ix, qs, 0, NQ, 12, 41)
qi = np.repeat(np.arange(NQ, dtype=np.uint32), 8); ti = seqno[:, :8].reshape(-1).astype(np.uint32)
for rep in range(3):
    t0 = time.time(); res = ctx.align_pairs(qs, db, qi, ti); print(f"python wall {1e3*(time.time()-t0):.1f} ms", file=sys.stderr)
