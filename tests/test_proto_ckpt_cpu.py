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
