import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)


# vsg_align_pairs routes calls of fewer than VSG_CKPT_MIN_PAIRS pairs (default 2048) through the direction-bit
# kernels, which have the lower latency; the tests' batches are small, so they pin the threshold to 0 to run the
# checkpoint kernels (the product's main path) and switch to the other path explicitly where both are compared.
os.environ.setdefault("VSG_CKPT_MIN_PAIRS", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs the compiled reference in oracle/_ref")
