"""configs[2] shape through seam 1: the stock reference CLI versus the same CLI relinked with shim/align_simd_vsg.cpp
(oracle/_ref/vsearch_gpu) on --cluster_fast, C3-shaped reads (300 nt, 1 % divergence, Zipf-ish root choice).
Prints wall times and checks that the uc files agree at equal --threads.  Measurement for DESIGN.md (the cluster
round driver question); not part of the product."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vsearch_b200 import synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ROOTS = int(sys.argv[2]) if len(sys.argv) > 2 else max(50, N // 200)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STOCK = os.path.join(ROOT, "oracle", "_ref", "vsearch"); GPU = os.path.join(ROOT, "oracle", "_ref", "vsearch_gpu")
rng = np.random.default_rng(3)
roots = synth.random_seqs(rng, ROOTS, 300)
w = 1.0 / np.arange(1, ROOTS + 1); w /= w.sum()
pick = rng.choice(ROOTS, size=N, p=w)
reads = synth.mutate_batch(rng, roots[pick], 0.01)
fa = "/tmp/c3.fasta"; synth.write_fasta(fa, reads, "a")
def run(binary, threads, tag):
    uc = f"/tmp/c3_{tag}.uc"
    t0 = time.time()
    p = subprocess.run([binary, "--cluster_fast", fa, "--id", "0.97", "--uc", uc, "--threads", str(threads), "--quiet"], capture_output=True, text=True)
    dt = time.time() - t0
    assert p.returncode == 0, p.stderr[-1000:]
    return dt, sorted(open(uc).readlines())
nproc = os.cpu_count()
for thr in (8, nproc):
    tc, uc_c = run(STOCK, thr, f"cpu{thr}")
    print(f"stock  --threads {thr}: {tc:.2f} s  ({N / tc:.0f} reads/s)", flush=True)
tg, uc_g = run(GPU, 8, "gpu8")
tc8, uc_c8 = run(STOCK, 8, "cpu8b")
print(f"seam-1 --threads 8: {tg:.2f} s  ({N / tg:.0f} reads/s); identical uc at equal threads: {uc_g == uc_c8}; clusters {sum(1 for l in uc_c8 if l.startswith('C'))}")
