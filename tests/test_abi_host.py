"""CPU-only: the C++ link surface is the reference's.  tests/abi_probe.cu is compiled (host only, g++) against
the reference's headers and against include/gpuntt: the struct layouts printed must be identical, and every
symbol the reference-header build leaves undefined (the mangled GPU_NTT<...> etc. a caller compiled against
GPU-NTT needs) must be defined by gpu_ntt_b200/lib/libntt-1.0.a.  Needs /root/reference (build container)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_INC = "/root/reference/src/include"
CUDA_INC = "/usr/local/cuda/include"
LIB = os.path.join(ROOT, "gpu_ntt_b200", "lib", "libntt-1.0.a")
PROBE = os.path.join(ROOT, "tests", "abi_probe.cu")

pytestmark = pytest.mark.skipif(not os.path.isdir(REF_INC) or not os.path.isdir(CUDA_INC),
                                reason="reference headers / CUDA headers not present")


def compile_probe(tmp_path, inc, tag):
    obj = str(tmp_path / f"probe_{tag}.o")
    subprocess.check_call(["g++", "-std=c++17", "-w", "-x", "c++", "-c", PROBE, "-I", inc, "-I", CUDA_INC, "-o", obj])
    return obj


def test_struct_layouts_match_the_reference(tmp_path):
    outs = []
    for inc, tag in ((REF_INC, "ref"), (os.path.join(ROOT, "include"), "ours")):
        obj = compile_probe(tmp_path, inc, tag)
        exe = str(tmp_path / f"probe_{tag}")
        # layout printing needs no library: link with the symbols left unresolved
        subprocess.check_call(["g++", obj, "-o", exe, "-Wl,--unresolved-symbols=ignore-all"])
        outs.append(subprocess.run([exe], capture_output=True, text=True, check=True).stdout)
    assert outs[0] == outs[1] and "u64 Modulus 24 8 | cfg 40 rns 40 c4 24 r4 24" in outs[0]


def test_library_defines_every_symbol_a_reference_caller_needs(tmp_path):
    if not os.path.exists(LIB):
        subprocess.check_call(["bash", os.path.join(ROOT, "gpu_ntt_b200", "build_cxx.sh")])
    obj = compile_probe(tmp_path, REF_INC, "ref")
    und = subprocess.run(["nm", "-u", obj], capture_output=True, text=True, check=True).stdout.split("\n")
    need = {l.split()[-1] for l in und if l.strip() and ("gpuntt" in l or "Modulus" in l)}
    assert len(need) >= 60, need
    defined = subprocess.run(["nm", "--defined-only", LIB], capture_output=True, text=True, check=True).stdout
    have = {l.split()[-1] for l in defined.split("\n") if len(l.split()) >= 3}
    missing = sorted(s for s in need if s not in have)
    assert not missing, "symbols a GPU-NTT caller links against but libntt-1.0.a lacks:\n" + "\n".join(missing)
