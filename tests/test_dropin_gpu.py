"""The reference's OWN example programs (example/ntt_merge/*.cu, example/ntt_4step/*.cu), compiled UNCHANGED against
include/gpuntt/ and linked with gpu_ntt_b200/lib/libntt-1.0.a by gpu_ntt_b200/build_cxx.sh (binaries in
tests/_dropin/, built in the container where /root/reference exists; they travel to the GPU box).  They are the
reference's test-suite: random inputs, GPU result compared with its CPU class, "All Correct" printed."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
DROPIN = os.path.join(HERE, "_dropin")


def run(exe, *args, timeout=600):
    path = os.path.join(DROPIN, exe)
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (run gpu_ntt_b200/build_cxx.sh where the reference tree is mounted)")
    r = subprocess.run([path, *map(str, args)], capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


@pytest.mark.parametrize("logn,batch", [(12, 1), (16, 4), (10, 8), (17, 2), (5, 3)])
def test_reference_gpu_merge_ntt_example(logn, batch):
    out = run("gpu_merge_ntt_examples", logn, batch)
    assert out.count("All Correct for PerPolynomial NTT.") == 2 and "All Correct for PerCoefficient NTT." in out
    assert "Error" not in out


@pytest.mark.parametrize("logn,batch", [(12, 1), (16, 4), (11, 8), (17, 2)])
def test_reference_gpu_merge_intt_example(logn, batch):
    out = run("gpu_merge_intt_examples", logn, batch)
    assert out.count("All Correct") == 3 and "Error" not in out


def test_reference_cpu_merge_example_config_c1():
    """BASELINE config C1: cpu_merge_ntt_examples 12 1 (our NTTCPU vs schoolbook)."""
    assert "All Correct." in run("cpu_merge_ntt_examples", 12, 1)


@pytest.mark.parametrize("logn,batch", [(12, 2), (16, 2), (20, 1)])
def test_reference_gpu_4step_examples(logn, batch):
    assert "All Correct." in run("gpu_4step_ntt_examples", logn, batch)
    assert "All Correct." in run("gpu_4step_intt_examples", logn, batch)


def test_reference_cpu_4step_example():
    assert "All Correct." in run("cpu_4step_ntt_examples", 12, 1)
