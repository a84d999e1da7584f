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
