"""The round-2 groundwork prototype (tools/proto_checkpoint_traceback.py: forward pass without direction
bits, traceback regenerating them tile by tile from H/E/F checkpoints) stays bit-identical with the
oracle.  CPU only; nothing of the product depends on it yet."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_checkpoint_traceback_prototype_matches_oracle():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "proto_checkpoint_traceback.py"), "80"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 mismatches" in r.stdout, r.stdout


def test_host_device_checkpoint_traceback_matches_oracle(tmp_path):
    """vsearch_b200/csrc/experimental/tb_ckpt.h, compiled for the host, over checkpoints in the device layout"""
    import checkers
    checkers.oracle()   # builds oracle/liboracle.so if needed
    exe = str(tmp_path / "ckpt_host_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "oracle"),
                           os.path.join(ROOT, "tools", "ckpt_host_check.cpp"), "-L", os.path.join(ROOT, "oracle"),
                           "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-o", exe])
    r = subprocess.run([exe, "300"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and " 0 mismatches" in r.stdout, r.stdout + r.stderr
