"""A short run of scripts/fuzz_parity.py inside the GPU suite: random small graphs with tie-heavy data, duplicate list
entries and random k / ef / PQ shapes; exact traversal, ADC traversal + rerank (with statistics = visited bitset,
without = bitset-free, with a rerank limit) and two-level search against the oracle.  Longer sweeps: run the script
itself under gpurun with SEED / ROUNDS / BUDGET_S (round 1: five seeds, ~6000 checks, no mismatch)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_fuzz_parity_short(gpu_lib):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SEED="11", ROUNDS="14", BUDGET_S="40")
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "fuzz_parity.py")], env=env, capture_output=True, text=True, timeout=300)
    tail = "\n".join(r.stdout.strip().split("\n")[-8:])
    assert r.returncode == 0, tail + r.stderr[-500:]
    assert "0 mismatches" in tail, tail
