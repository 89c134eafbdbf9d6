"""The CMake package (project GPUNTT, target GPUNTT::ntt, archive libntt-1.0.a, headers under
include/GPUNTT-1.0, package config in lib/cmake/GPUNTT-1.0).  Always: the project CONFIGURES with
-DGPUNTT_BUILD_EXAMPLES=ON (the round-1 tree failed there: examples/ was empty).  With GPUNTT_TEST_CMAKE_INSTALL=1
(several minutes of nvcc): build, `cmake --install`, then configure + build + run tests/cmake_consumer against the
installed package with find_package(GPUNTT CONFIG) -- the outcome of that run is kept in
profiles/r2_cmake_install_check.txt."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CMAKE = shutil.which("cmake")
pytestmark = pytest.mark.skipif(CMAKE is None or shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"),
                                reason="needs cmake and nvcc")
ENV = dict(os.environ, CUDACXX=os.environ.get("CUDACXX", "/usr/local/cuda/bin/nvcc"))


def test_configures_with_examples(tmp_path):
    b = tmp_path / "build"
    r = subprocess.run([CMAKE, "-S", ROOT, "-B", str(b), "-DGPUNTT_BUILD_EXAMPLES=ON"], capture_output=True, text=True, env=ENV)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    help_ = subprocess.run([CMAKE, "--build", str(b), "--target", "help"], capture_output=True, text=True, env=ENV).stdout
    for target in ("ntt", "gpu_merge_examples", "gpu_4step_examples"):
        assert target in help_, help_[-2000:]


@pytest.mark.skipif(os.environ.get("GPUNTT_TEST_CMAKE_INSTALL") != "1", reason="several minutes of nvcc: set GPUNTT_TEST_CMAKE_INSTALL=1")
def test_install_and_find_package_consumer(tmp_path):
    b, prefix, cb = tmp_path / "build", tmp_path / "prefix", tmp_path / "consumer_build"
    subprocess.check_call([CMAKE, "-S", ROOT, "-B", str(b), f"-DCMAKE_INSTALL_PREFIX={prefix}"], env=ENV)
    subprocess.check_call([CMAKE, "--build", str(b), "-j", str(os.cpu_count() or 4)], env=ENV)
    subprocess.check_call([CMAKE, "--install", str(b)], env=ENV)
    assert (prefix / "lib" / "libntt-1.0.a").exists() or (prefix / "lib64" / "libntt-1.0.a").exists()
    assert (prefix / "include" / "GPUNTT-1.0" / "gpuntt" / "ntt_merge" / "ntt.cuh").exists()
    subprocess.check_call([CMAKE, "-S", os.path.join(ROOT, "tests", "cmake_consumer"), "-B", str(cb), f"-DCMAKE_PREFIX_PATH={prefix}"], env=ENV)
    subprocess.check_call([CMAKE, "--build", str(cb)], env=ENV)
    out = subprocess.run([str(cb / "consumer")], capture_output=True, text=True)      # CPU-only part of the program
    assert out.returncode == 0 and "modulus 576460756061519873" in out.stdout, out.stdout + out.stderr
