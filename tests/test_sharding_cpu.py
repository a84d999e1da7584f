"""N > 1 host logic on CPU with the gloo backend (world_size 2): the database broadcast and the
query sharding / result gathering that bench.py does across GPUs, with the ORACLE standing in for
the device path (this is a test of the plumbing, not of the kernels)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import checkers
    from vsearch_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_db, L = 300, 400
    # rank 0 owns the database; everyone else receives the packed bytes + offsets + lengths
    cat = torch.zeros(n_db * L, dtype=torch.uint8)
    off = torch.zeros(n_db, dtype=torch.int64)
    ln = torch.zeros(n_db, dtype=torch.int32)
    dbm = synth.config2_db(n_db, L, seed=11)
    if rank == 0:
        cat.copy_(torch.from_numpy(dbm.reshape(-1))); off.copy_(torch.arange(n_db) * L); ln.fill_(L)
    for t in (cat, off, ln):
        dist.broadcast(t, 0)
    ss = synth.SeqSet.__new__(synth.SeqSet)
    ss.cat = np.concatenate([cat.numpy(), np.zeros(1, np.uint8)]); ss.offs = off.numpy(); ss.lens = ln.numpy()
    assert ss.seq(n_db - 1) == dbm[n_db - 1].tobytes()
    # contiguous query shards, as bench.py assigns batch (step*world + rank)
    nq = 40
    qs, src = synth.config2_query_batch(dbm, nq, q_len=120, div=0.05, seed=3, batch=0)
    lo, hi = rank * nq // world, (rank + 1) * nq // world
    od = checkers.OracleDb(ss)
    opts = checkers.search_opts(n_db, id=0.9)
    best = torch.full((nq,), -1, dtype=torch.int64)
    cells = torch.zeros(1, dtype=torch.int64)
    for q in range(lo, hi):
        hits, _, c = od.search(qs.seq(q), opts)
        best[q] = hits[0].target if hits else -1
        cells += c
    od.close()
    # results concatenate by query index; work adds up
    dist.all_reduce(best, op=dist.ReduceOp.MAX)
    dist.all_reduce(cells, op=dist.ReduceOp.SUM)
    if rank == 0:
        out.put((best.numpy().tolist(), int(cells.item()), src.tolist()))
    dist.destroy_process_group()


def test_two_rank_broadcast_and_query_sharding():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    best, cells, src = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert best == src          # every query finds the sequence it was cut from, on whichever rank
    assert cells > 0
