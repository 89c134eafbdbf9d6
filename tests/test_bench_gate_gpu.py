"""bench.py verifies its first step against the CPU reference before it times anything (BASELINE.md section 4 item 5,
VERDICT r1 item 3): a healthy run prints the JSON line with parity_gate.mismatching_words == 0, a run whose root table
has one flipped bit exits non-zero without a result line."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench(*extra):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3", "--cpu-sample", "64", *extra],
                          capture_output=True, text=True, timeout=600)


def test_bench_parity_gate_passes_and_reports():
    r = _bench()
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["parity_gate"]["mismatching_words"] == 0 and d["parity_gate"]["checked_polynomials"] == 64
    assert d["roofline"]["frac"] > 0 and d["cpu_baseline"]["value"] > 0 and d["gpu_launches"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] >= 1024 * 65536 * 8


def test_bench_parity_gate_catches_a_wrong_twiddle():
    r = _bench("--corrupt-twiddle")
    assert r.returncode != 0
    assert "PARITY GATE FAILED" in r.stderr
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
