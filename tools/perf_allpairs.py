"""Quick throughput probe of the aligner on all-pairs workloads (not the bench contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vsearch_b200 import lib as vlib, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 200
reads = synth.config1_allpairs(n_reads=n, n_roots=20, length=L)
ctx = vlib.Context(0)
ss = ctx.seqset(reads)
qi, ti = np.triu_indices(n, k=1)
qi = qi.astype(np.uint32); ti = ti.astype(np.uint32)
for rep in range(3):
    t0 = time.time()
    res = ctx.align_pairs(ss, ss, qi, ti, cigar=False)
    dt = time.time() - t0
    print(f"n={n} L={L} pairs={qi.shape[0]} cells={res.cells:.3e} wall={dt*1e3:.1f}ms "
          f"fwd={res.fwd_ms:.2f}ms tb={res.tb_ms:.2f}ms  kernel GCUPS={res.cells/res.fwd_ms/1e6:.1f} "
          f"wall GCUPS={res.cells/dt/1e9:.1f} fast={res.fast_pairs} exact={res.exact_pairs}", flush=True)
