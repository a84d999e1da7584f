"""The C-ABI library loads and exports every symbol include/vsg.h declares; without a GPU it
refuses to work instead of silently falling back (no compute calls here)."""
import ctypes as C
import os

import pytest

from vsearch_b200 import lib as vlib


def test_library_exports_every_declared_symbol():
    names = vlib.declared_symbols()
    assert {"vsg_ctx_create", "vsg_seqset_create", "vsg_align_pairs", "vsg_index_create", "vsg_rank",
            "vsg_search_batch"} <= set(names)
    lib = vlib.load()
    for n in names:
        assert hasattr(lib, n), n
    assert b"sm_100a" in lib.vsg_version()


def test_header_mentions_reference_interfaces():
    text = open(vlib.HEADER).read()
    for ref in ("core/align_simd.cpp", "core/searchcore.cpp", "core/dbindex", "core/search.hpp"):
        assert ref in text


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(vlib.VsgError, match="no CUDA device"):
        vlib.Context(0)


def test_product_does_not_import_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for d, _, files in os.walk(os.path.join(root, "vsearch_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(d, f)).read()
                assert "liboracle" not in text and "oracle/" not in text and "checkers" not in text, f


def test_allpairs_partition_balances_the_triangle():
    """pure host arithmetic behind the multi-GPU row sharding of --allpairs_global (no device needed)"""
    import numpy as np
    lib = vlib.load()
    rng = np.random.default_rng(3)
    lens = rng.integers(50, 450, size=5000).astype(np.int32)
    for nparts in (1, 2, 3, 8):
        b = np.zeros(nparts + 1, dtype=np.int64)
        rc = lib.vsg_allpairs_partition(lens.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int64(len(lens)), C.c_int(nparts),
                                        b.ctypes.data_as(C.POINTER(C.c_int64)))
        assert rc == 0 and b[0] == 0 and b[-1] == len(lens) and np.all(np.diff(b) >= 0)
        suffix = np.concatenate([np.cumsum(lens[::-1].astype(np.float64))[::-1][1:], [0.0]])
        work = lens * suffix
        parts = np.array([work[b[p]:b[p + 1]].sum() for p in range(nparts)])
        assert parts.max() <= 1.02 * parts.mean() + work.max()
        if nparts > 1:   # equal row counts would be badly skewed
            eq = np.array([w.sum() for w in np.array_split(work, nparts)])
            assert eq.max() / eq.mean() > parts.max() / parts.mean()
