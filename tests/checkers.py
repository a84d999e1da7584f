"""ctypes loaders for the CHECKERS (oracle/liboracle.so and oracle/_ref/libvsref.so).

Test infrastructure only — the product never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

DEFAULT_PEN = np.array([2, -4, 1, 1, 18, 18, 1, 1, 1, 1, 2, 2, 1, 1], dtype=np.int64)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Scoring(C.Structure):
    _fields_ = [("v", C.c_int64 * 14), ("n_mismatch", C.c_int)]


def make_scoring(pen=None, n_mismatch=0):
    s = Scoring()
    pen = DEFAULT_PEN if pen is None else pen
    for i in range(14):
        s.v[i] = int(pen[i])
    s.n_mismatch = int(n_mismatch)
    return s


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
        _oracle = C.CDLL(path)
        _oracle.oracle_unique_kmers.restype = C.c_uint
        _oracle.oracle_index_build.restype = C.c_void_p
        _oracle.oracle_map_4bit.restype = C.c_ubyte
    return _oracle


_ref = None


def ref():
    """The unmodified reference behind a C ABI, or None when oracle/_ref was not built."""
    global _ref
    if _ref is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libvsref.so")
        if not os.path.exists(path):
            return None
        _ref = C.CDLL(path)
        _ref.vsref_db_create.restype = C.c_void_p
    return _ref


def oracle_nw16(q: bytes, d: bytes, pen=None, n_mismatch=0):
    lib = oracle()
    sc = make_scoring(pen, n_mismatch)
    score = C.c_int16(); al = C.c_uint16(); ma = C.c_uint16(); mi = C.c_uint16(); ga = C.c_uint16()
    cap = len(q) + len(d) + 64
    buf = C.create_string_buffer(cap)
    rc = lib.oracle_nw16(C.byref(sc), q, C.c_int64(len(q)), d, C.c_int64(len(d)),
                         C.byref(score), C.byref(al), C.byref(ma), C.byref(mi), C.byref(ga),
                         buf, C.c_size_t(cap))
    assert rc == 0
    return score.value, al.value, ma.value, mi.value, ga.value, buf.value.decode()


def ref_search16(q: bytes, targets, pen=None, n_mismatch=0):
    lib = ref()
    pen = np.ascontiguousarray(DEFAULT_PEN if pen is None else pen, dtype=np.int64)
    n = len(targets)
    lens = np.array([len(t) for t in targets], dtype=np.int32)
    offs = np.zeros(n, dtype=np.int64)
    if n:
        np.cumsum(lens[:-1], out=offs[1:])
    cat = b"".join(targets) + b"\0"
    scores = np.zeros(n, dtype=np.int16)
    al = np.zeros(n, dtype=np.uint16); ma = np.zeros(n, dtype=np.uint16)
    mi = np.zeros(n, dtype=np.uint16); ga = np.zeros(n, dtype=np.uint16)
    stride = len(q) + (int(lens.max()) if n else 0) + 64
    cig = C.create_string_buffer(stride * max(n, 1))
    rc = lib.vsref_search16(_p(pen, C.c_int64), C.c_int(n_mismatch), q, C.c_int(len(q)),
                            C.c_int(n), cat, _p(offs, C.c_int64), _p(lens, C.c_int),
                            _p(scores, C.c_int16), _p(al, C.c_uint16), _p(ma, C.c_uint16),
                            _p(mi, C.c_uint16), _p(ga, C.c_uint16), cig, C.c_int64(stride))
    assert rc == 0
    out = []
    raw = cig.raw
    for i in range(n):
        c = raw[i * stride:(i + 1) * stride].split(b"\0", 1)[0].decode()
        out.append((int(scores[i]), int(al[i]), int(ma[i]), int(mi[i]), int(ga[i]), c))
    return out


def oracle_unique_kmers(seq: bytes, k=8, mask_lower=0):
    out = np.zeros(max(len(seq), 1), dtype=np.uint32)
    n = oracle().oracle_unique_kmers(C.c_int(k), seq, C.c_int64(len(seq)), C.c_int(mask_lower),
                                     _p(out, C.c_uint32))
    return out[:n].copy()


def ref_unique_kmers(seq: bytes, k=8, mask_lower=0):
    out = np.zeros(max(len(seq), 1), dtype=np.uint32)
    n = ref().vsref_unique_count(C.c_int(k), seq, C.c_int(len(seq)), C.c_int(mask_lower),
                                 _p(out, C.c_uint32), C.c_int(out.shape[0]))
    return out[:n].copy()


class OracleHit(C.Structure):
    _fields_ = [("target", C.c_int), ("strand", C.c_int), ("count", C.c_uint),
                ("accepted", C.c_int), ("rejected", C.c_int), ("aligned", C.c_int), ("weak", C.c_int),
                ("nwscore", C.c_int), ("nwdiff", C.c_int), ("nwgaps", C.c_int), ("nwindels", C.c_int),
                ("nwalignmentlength", C.c_int), ("matches", C.c_int), ("mismatches", C.c_int),
                ("internal_alignmentlength", C.c_int), ("internal_gaps", C.c_int),
                ("internal_indels", C.c_int),
                ("trim_q_left", C.c_int), ("trim_q_right", C.c_int), ("trim_t_left", C.c_int),
                ("trim_t_right", C.c_int),
                ("id", C.c_double), ("id0", C.c_double), ("id1", C.c_double), ("id2", C.c_double),
                ("id3", C.c_double), ("id4", C.c_double), ("shortest", C.c_int), ("longest", C.c_int)]


class SearchOpts(C.Structure):
    _fields_ = [("id", C.c_double), ("weak_id", C.c_double), ("maxaccepts", C.c_int),
                ("maxrejects", C.c_int), ("minwordmatches", C.c_int), ("tophits", C.c_int),
                ("iddef", C.c_int), ("mask_lower", C.c_int)]


MINWORDMATCHES = [-1, -1, -1, 18, 17, 16, 15, 14, 12, 11, 10, 9, 8, 7, 5, 3]


def search_opts(n_db, id=0.9, maxaccepts=1, maxrejects=32, k=8, minwordmatches=-1, iddef=2,
                weak_id=10.0, mask_lower=0):
    """Effective options after the reference's fix-ups (vsearch.cc:186-276,
    usearch_global.cpp:598-614)."""
    o = SearchOpts()
    o.id = id
    o.weak_id = min(weak_id, id)
    o.maxaccepts = min(maxaccepts, n_db)
    o.maxrejects = min(maxrejects, n_db)
    o.minwordmatches = MINWORDMATCHES[k] if minwordmatches < 0 else minwordmatches
    o.tophits = min(o.maxaccepts + o.maxrejects + 8, n_db)
    o.iddef = iddef
    o.mask_lower = mask_lower
    return o


class OracleDb:
    def __init__(self, ss, k=8, mask_lower=0):
        self.ss = ss
        self.k = k
        self.h = oracle().oracle_index_build(C.c_int(k), C.c_int(len(ss)), _p(ss.cat, C.c_char),
                                             _p(ss.offs, C.c_int64), _p(ss.lens, C.c_int),
                                             C.c_int(mask_lower))

    def close(self):
        if self.h:
            oracle().oracle_index_free(C.c_void_p(self.h))
            self.h = None

    def topscores(self, q: bytes, opts):
        kmers = oracle_unique_kmers(q, self.k, opts.mask_lower)
        seqno = np.zeros(opts.tophits + 1, dtype=np.uint32)
        count = np.zeros(opts.tophits + 1, dtype=np.uint32)
        n = oracle().oracle_topscores(C.c_void_p(self.h), _p(self.ss.lens, C.c_int),
                                      _p(kmers, C.c_uint32), C.c_uint(kmers.shape[0]),
                                      C.c_int(opts.minwordmatches), C.c_int(opts.tophits),
                                      _p(seqno, C.c_uint32), _p(count, C.c_uint32))
        return seqno[:n].copy(), count[:n].copy()

    def search(self, q: bytes, opts, pen=None, strand=0):
        sc = make_scoring(pen)
        hits = (OracleHit * (opts.tophits + 1))()
        pairs = C.c_int64(); cells = C.c_int64()
        n = oracle().oracle_search_onequery(C.c_void_p(self.h), C.byref(sc), C.byref(opts),
                                            C.c_int(len(self.ss)), _p(self.ss.cat, C.c_char),
                                            _p(self.ss.offs, C.c_int64), _p(self.ss.lens, C.c_int),
                                            q, C.c_int(len(q)), C.c_int(strand),
                                            hits, C.c_int(opts.tophits + 1),
                                            C.byref(pairs), C.byref(cells))
        return [hits[i] for i in range(n)], pairs.value, cells.value


class RefDb:
    """Reference Database+Dbindex+session (one at a time per process)."""

    def __init__(self, ss, k=8, id=0.9, maxaccepts=1, maxrejects=32, minwordmatches=-1,
                 dust=0, strand_both=0, iddef=2):
        self.ss = ss
        self.h = ref().vsref_db_create(C.c_int(len(ss)), _p(ss.cat, C.c_char),
                                       _p(ss.offs, C.c_int64), _p(ss.lens, C.c_int),
                                       C.c_int(k), C.c_double(id), C.c_int(maxaccepts),
                                       C.c_int(maxrejects), C.c_int(minwordmatches), C.c_int(dust),
                                       C.c_int(strand_both), C.c_int(iddef))
        self.tophits = ref().vsref_db_tophits(C.c_void_p(self.h))

    def close(self):
        if self.h:
            ref().vsref_db_free(C.c_void_p(self.h))
            self.h = None

    def topscores(self, q: bytes):
        seqno = np.zeros(self.tophits + 1, dtype=np.uint32)
        count = np.zeros(self.tophits + 1, dtype=np.uint32)
        length = np.zeros(self.tophits + 1, dtype=np.uint32)
        n = ref().vsref_db_topscores(C.c_void_p(self.h), q, C.c_int(len(q)),
                                     _p(seqno, C.c_uint32), _p(count, C.c_uint32),
                                     _p(length, C.c_uint32))
        return seqno[:n].copy(), count[:n].copy()

    def search_rows(self, qs, max_results=8, threads=None):
        """the reference's own multi-threaded search_batch, every record kept: (counts, dict of flat arrays)"""
        nq = len(qs)
        threads = threads or (os.cpu_count() or 1)
        counts = np.zeros(nq, dtype=np.int32)
        m = nq * max_results
        a = {k: np.zeros(m, dtype=np.int32) for k in ("target", "matches", "mismatches", "gaps", "alnlen", "accepted", "strand")}
        a["id"] = np.zeros(m, dtype=np.float64)
        ref().vsref_db_search_batch_rows(C.c_void_p(self.h), C.c_int(nq), _p(qs.cat, C.c_char), _p(qs.offs, C.c_int64),
                                         _p(qs.lens, C.c_int), C.c_int(threads), C.c_int(max_results), _p(counts, C.c_int),
                                         _p(a["target"], C.c_int), _p(a["id"], C.c_double), _p(a["matches"], C.c_int),
                                         _p(a["mismatches"], C.c_int), _p(a["gaps"], C.c_int), _p(a["alnlen"], C.c_int),
                                         _p(a["accepted"], C.c_int), _p(a["strand"], C.c_int))
        return counts, a

    def search(self, qs, max_results=8):
        nq = len(qs)
        counts = np.zeros(nq, dtype=np.int32)
        m = nq * max_results
        target = np.zeros(m, dtype=np.int32); idv = np.zeros(m, dtype=np.float64)
        ma = np.zeros(m, dtype=np.int32); mi = np.zeros(m, dtype=np.int32)
        ga = np.zeros(m, dtype=np.int32); al = np.zeros(m, dtype=np.int32)
        acc = np.zeros(m, dtype=np.int32); st = np.zeros(m, dtype=np.int32)
        ref().vsref_db_search(C.c_void_p(self.h), C.c_int(nq), _p(qs.cat, C.c_char),
                              _p(qs.offs, C.c_int64), _p(qs.lens, C.c_int), C.c_int(max_results),
                              _p(counts, C.c_int), _p(target, C.c_int), _p(idv, C.c_double),
                              _p(ma, C.c_int), _p(mi, C.c_int), _p(ga, C.c_int), _p(al, C.c_int),
                              _p(acc, C.c_int), _p(st, C.c_int))
        out = []
        for q in range(nq):
            rows = []
            for j in range(counts[q]):
                o = q * max_results + j
                rows.append((int(target[o]), float(idv[o]), int(ma[o]), int(mi[o]), int(ga[o]),
                             int(al[o]), int(acc[o]), int(st[o])))
            out.append(rows)
        return out
