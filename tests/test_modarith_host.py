"""Host-side exactness check of the integer helpers the kernels use on the device as well (modarith.cuh is
__host__ __device__ for them): the base-2^32 long division behind the RNS kernels' per-segment reciprocal against the
compiler's 128-bit division.  Compiles tests/modarith_probe.cu with nvcc (no GPU needed)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_long_division_helpers_match_128_bit_division(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "modarith_probe")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-w", "-I", os.path.join(ROOT, "gpu_ntt_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
                    "-o", exe, os.path.join(ROOT, "tests", "modarith_probe.cu")], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "bad=0" in out.stdout
