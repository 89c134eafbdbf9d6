"""A short run of the randomised differential test (tools/fuzz_parity.py: random width / ring / batch / modulus / direction /
signed / RNS cases through the C ABI, every output word against the oracle) inside the GPU suite; the long runs are kept under
profiles/r2_fuzz_parity.jsonl."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_randomised_parity_short_run():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_parity.py"), "12", "7"], capture_output=True, text=True,
                         cwd=ROOT, timeout=300)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert lines, out.stderr[-2000:]
    summary = json.loads(lines[-1])
    assert out.returncode == 0 and summary["mismatches"] == 0, "\n".join(lines[:10])
    assert summary["fuzz_cases"] >= 50
