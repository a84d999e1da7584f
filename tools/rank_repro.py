import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vsearch_b200 import lib as vlib, synth
NQ = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dbm = synth.config2_db(100_000, 1500, 2024)
ctx = vlib.Context(0)
db = ctx.seqset(synth.SeqSet.from_matrix(dbm))
ix = ctx.index(db, 8, 0)
qs_h, _ = synth.config2_query_batch(dbm, NQ, batch=1)
qs = ctx.seqset(qs_h)
ref = None
for rep in range(REPS):
    seqno, count, nc = ctx.rank(ix, qs, 0, NQ, 12, 41)
    if ref is None: ref = (seqno.copy(), count.copy())
    print(rep, "ok", bool((seqno == ref[0]).all() and (count == ref[1]).all()), flush=True)
