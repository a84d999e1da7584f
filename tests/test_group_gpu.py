"""vsg_group: several GPUs behind one process (database copied device to device, queries / all-pairs rows
sharded, no data-path collective).  Results must equal the single-context calls row for row; with one visible
GPU the same code runs as a group of one (and of two contexts on the same device, which exercises the
device-to-device copy and the sharding logic)."""
import os
import subprocess

import numpy as np
import pytest

from vsearch_b200 import lib as vlib
from vsearch_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def _ndev():
    import torch
    return torch.cuda.device_count()


def _rows(res, counts, nq, mr):
    return [[(r.target, r.id, r.matches, r.mismatches, r.gaps, r.alignment_length, r.accepted, r.strand)
             for r in (res[q * mr + j] for j in range(int(counts[q])))] for q in range(nq)]


@pytest.mark.parametrize("devices", [[0], [0, 0, 0], "all"])
def test_group_search_and_allpairs_equal_single_context(devices):
    if devices == "all":
        n = _ndev()
        if n < 2:
            pytest.skip("one GPU visible")
        devices = list(range(n))
    dbs, qss, _ = synth.config2_search(n_db=3000, db_len=900, n_q=1500, q_len=200, div=0.06, seed=11)
    # ragged query lengths so that equal-nucleotide sharding differs from equal counts
    rng = np.random.default_rng(3)
    qss = synth.SeqSet([qss.seq(i)[: int(rng.integers(60, 200))] for i in range(len(qss))])
    o = vlib.default_search_opts(); o.id = 0.9; o.maxaccepts = 2; o.maxrejects = 8; o.strand_both = 1
    mr = 3
    ctx = vlib.Context(0)
    db = ctx.seqset(dbs); q = ctx.seqset(qss); ix = ctx.index(db, 8, 0)
    res, counts, work = ctx.search(ix, db, q, 0, len(qss), o, mr)
    want = _rows(res, counts, len(qss), mr)
    g = vlib.Group(devices, dbs)
    st = g.stats()
    assert st["broadcast_bytes"] == (len(devices) - 1) * (int(dbs.lens.sum()) + 12 * len(dbs))
    gres, gcounts, gwork = g.search(qss, o, mr)
    assert _rows(gres, gcounts, len(qss), mr) == want
    assert gwork.tolist() == work.tolist()
    g.close()
    # all-pairs over a small read set
    reads = synth.config1_allpairs(n_reads=260, n_roots=6, length=180)
    oa = vlib.default_search_opts(); oa.id = 0.8
    ss = ctx.seqset(reads)
    h1, w1 = vlib.allpairs(ctx, ss, 0, len(reads), oa, 100000)
    g2 = vlib.Group(devices, reads)
    h2, w2 = g2.allpairs(oa, 100000)
    assert len(h1) > 500 and h1.tolist() == h2.tolist() and w1.tolist() == w2.tolist()
    g2.close()
    ss.close(); ix.close(); db.close(); q.close(); ctx.close()


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "seam2_driver_gpu")), reason="oracle/_ref not present")
def test_search_batch_shim_on_a_device_group(tmp_path):
    """the seam-2 drop-in with VSG_DEVICES: every visible GPU (or the same one three times) behind one search_batch call"""
    import test_seam2_gpu as t2
    dbf, qf = t2._dataset(tmp_path)
    n = _ndev()
    devs = ",".join(str(i) for i in range(n)) if n > 1 else "0,0,0"
    case = ["id=0.8", "maxaccepts=3", "maxrejects=8", "strand=1", "self=1", "minsizeratio=0.2"]
    outs = []
    for exe, env in (("seam2_driver_ref", {}), ("seam2_driver_gpu", {"VSG_DEVICES": devs})):
        r = subprocess.run([os.path.join(REF, exe), dbf, qf] + case, capture_output=True, text=True, timeout=900,
                           env=dict(os.environ, **env))
        assert r.returncode == 0, (exe, r.stdout[-2000:], r.stderr[-2000:])
        outs.append(r.stdout.splitlines())
    assert outs[0] == outs[1] and len(outs[0]) > 20
