"""Deterministic synthetic FASTA-shaped workloads.

Alphabet uniform over ACGT; mutation operator ``mut(s, r)``: per base with probability
``r``: 80 % substitution (uniform base), 10 % deletion, 10 % insertion after the base.
Everything is generated with numpy's PCG64 from a fixed seed so that the CPU baseline,
the oracle and the GPU path all see byte-identical inputs.
"""
from __future__ import annotations

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_seqs(rng: np.random.Generator, n: int, length: int) -> np.ndarray:
    """n iid sequences of a fixed length as an (n, length) uint8 ASCII matrix."""
    return ACGT[rng.integers(0, 4, size=(n, length), dtype=np.uint8)]


def mutate(rng: np.random.Generator, s: np.ndarray, r: float) -> np.ndarray:
    """mut(s, r) of SURVEY §8(d) on one ASCII uint8 sequence."""
    n = s.shape[0]
    hit = rng.random(n) < r
    kind = rng.random(n)
    sub = hit & (kind < 0.8)
    dele = hit & (kind >= 0.8) & (kind < 0.9)
    ins = hit & (kind >= 0.9)
    out = s.copy()
    out[sub] = ACGT[rng.integers(0, 4, size=int(sub.sum()), dtype=np.uint8)]
    # build with insertions after the base, deletions dropped
    reps = np.ones(n, dtype=np.int64)
    reps[dele] = 0
    reps[ins] = 2
    idx = np.repeat(np.arange(n), reps)
    res = out[idx]
    # second copy of an inserted position becomes a random base
    second = np.zeros(idx.shape[0], dtype=bool)
    if idx.shape[0] > 1:
        second[1:] = idx[1:] == idx[:-1]
    res[second] = ACGT[rng.integers(0, 4, size=int(second.sum()), dtype=np.uint8)]
    return res


class SeqSet:
    """Concatenated ASCII sequences + offsets/lengths (the layout Database::add builds,
    reference src/core/db.cpp:170-226, minus headers)."""

    def __init__(self, seqs):
        self.lens = np.array([len(s) for s in seqs], dtype=np.int32)
        self.offs = np.zeros(len(seqs), dtype=np.int64)
        if len(seqs):
            np.cumsum(self.lens[:-1], out=self.offs[1:])
        total = int(self.lens.sum())
        self.cat = np.empty(total + 1, dtype=np.uint8)
        self.cat[total] = 0
        pos = 0
        for s in seqs:
            a = np.frombuffer(s, dtype=np.uint8) if isinstance(s, (bytes, bytearray)) else np.asarray(s, dtype=np.uint8)
            self.cat[pos:pos + a.shape[0]] = a
            pos += a.shape[0]

    def __len__(self):
        return int(self.lens.shape[0])

    def seq(self, i: int) -> bytes:
        o = int(self.offs[i])
        return self.cat[o:o + int(self.lens[i])].tobytes()

    @classmethod
    def from_matrix(cls, m: np.ndarray) -> "SeqSet":
        self = cls.__new__(cls)
        n, L = m.shape
        self.lens = np.full(n, L, dtype=np.int32)
        self.offs = np.arange(n, dtype=np.int64) * L
        self.cat = np.empty(n * L + 1, dtype=np.uint8)
        self.cat[:-1] = m.reshape(-1)
        self.cat[-1] = 0
        return self


def config1_allpairs(n_reads: int = 1000, n_roots: int = 20, length: int = 200,
                     div: float = 0.10, seed: int = 12345) -> SeqSet:
    """C1: reads i = mut(root[i mod n_roots], div)."""
    rng = np.random.default_rng(seed)
    roots = random_seqs(rng, n_roots, length)
    return SeqSet([mutate(rng, roots[i % n_roots], div) for i in range(n_reads)])


def config2_search(n_db: int = 100_000, db_len: int = 1500, n_q: int = 1_000_000,
                   q_len: int = 250, div: float = 0.05, seed: int = 2024):
    """C2: DB iid; query = mut(window of a uniformly chosen DB sequence, div).
    Returns (db SeqSet, queries SeqSet, source target of every query)."""
    rng = np.random.default_rng(seed)
    dbm = random_seqs(rng, n_db, db_len)
    src = rng.integers(0, n_db, size=n_q)
    start = rng.integers(0, db_len - q_len + 1, size=n_q)
    qs = [mutate(rng, dbm[src[i], start[i]:start[i] + q_len], div) for i in range(n_q)]
    return SeqSet.from_matrix(dbm), SeqSet(qs), src


def config5_allpairs(n_reads: int = 200_000, n_roots: int = 2000, length: int = 400,
                     div: float = 0.15, seed: int = 5) -> SeqSet:
    rng = np.random.default_rng(seed)
    roots = random_seqs(rng, n_roots, length)
    pick = rng.integers(0, n_roots, size=n_reads)
    return SeqSet([mutate(rng, roots[pick[i]], div) for i in range(n_reads)])


def write_fasta(path: str, ss: SeqSet, prefix: str) -> None:
    with open(path, "wb") as f:
        for i in range(len(ss)):
            f.write(b">" + prefix.encode() + str(i).encode() + b"\n" + ss.seq(i) + b"\n")


def mutate_batch(rng: np.random.Generator, m: np.ndarray, r: float) -> SeqSet:
    """mut(., r) applied to every row of an (n, L) ASCII matrix at once (vectorised; the draw
    order differs from `mutate`, so the two generators give different — equally distributed —
    sequences for the same seed)."""
    n, L = m.shape
    hit = rng.random((n, L)) < r
    kind = rng.random((n, L))
    sub = hit & (kind < 0.8)
    dele = hit & (kind >= 0.8) & (kind < 0.9)
    ins = hit & (kind >= 0.9)
    out = m.copy()
    out[sub] = ACGT[rng.integers(0, 4, size=int(sub.sum()), dtype=np.uint8)]
    reps = np.ones((n, L), dtype=np.int64)
    reps[dele] = 0
    reps[ins] = 2
    flat_reps = reps.reshape(-1)
    idx = np.repeat(np.arange(n * L), flat_reps)
    res = out.reshape(-1)[idx]
    second = np.zeros(idx.shape[0], dtype=bool)
    if idx.shape[0] > 1:
        second[1:] = idx[1:] == idx[:-1]
    res[second] = ACGT[rng.integers(0, 4, size=int(second.sum()), dtype=np.uint8)]
    ss = SeqSet.__new__(SeqSet)
    ss.lens = reps.sum(axis=1).astype(np.int32)
    ss.offs = np.zeros(n, dtype=np.int64)
    np.cumsum(ss.lens[:-1], out=ss.offs[1:])
    ss.cat = np.empty(res.shape[0] + 1, dtype=np.uint8)
    ss.cat[:-1] = res
    ss.cat[-1] = 0
    return ss


def config2_db(n_db: int = 100_000, db_len: int = 1500, seed: int = 2024) -> np.ndarray:
    """C2 database as an (n_db, db_len) ASCII matrix."""
    return random_seqs(np.random.default_rng(seed), n_db, db_len)


def config2_query_batch(dbm: np.ndarray, n_q: int, q_len: int = 250, div: float = 0.05,
                        seed: int = 2024, batch: int = 0):
    """Batch `batch` of the C2 query stream: windows of uniformly chosen DB sequences, mut(., div).
    Returns (SeqSet, source target per query)."""
    rng = np.random.default_rng([seed, 7919, batch])
    n_db, db_len = dbm.shape
    src = rng.integers(0, n_db, size=n_q)
    start = rng.integers(0, db_len - q_len + 1, size=n_q)
    cols = start[:, None] + np.arange(q_len)[None, :]
    win = dbm[src[:, None], cols]
    return mutate_batch(rng, win, div), src
