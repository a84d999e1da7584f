"""vsearch_b200 — B200-native implementation of the vsearch search16 / k-mer-ranker hot path.

The product is the C-ABI library ``vsearch_b200/csrc/libvsg.so`` (CUDA, sm_100a) declared in
``include/vsg.h``; this package is only the thin Python loader used by tests and bench.py.
"""
