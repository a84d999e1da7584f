"""The traceback of the checkpoint aligner (vsearch_b200/csrc/tb_ckpt.h: direction bits regenerated tile by tile
from the forward pass's H/E/F checkpoints) compiled for the HOST and checked against the oracle over checkpoints
in the device layout, written by a scalar model of nw_ckpt_kernel under its shifted scoring
(tools/ckpt_host_check.cpp).  CPU only: it pins the algorithm and the layout, the GPU tests pin the kernels."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_device_checkpoint_traceback_matches_oracle(tmp_path):
    import checkers
    checkers.oracle()   # builds oracle/liboracle.so if needed
    exe = str(tmp_path / "ckpt_host_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "oracle"),
                           os.path.join(ROOT, "tools", "ckpt_host_check.cpp"), "-L", os.path.join(ROOT, "oracle"),
                           "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-o", exe])
    r = subprocess.run([exe, "500"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and " 0 mismatches" in r.stdout, r.stdout + r.stderr
