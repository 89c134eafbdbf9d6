"""CPU-only checks of the host logic: the C-ABI library loads and exports every symbol the header
declares (no compute calls), and the launch plans it reports compute the reference transform when
their index algebra is emulated in Python against the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gpu_ntt_b200 as G
from gpu_ntt_b200 import capi
from oracle import oracle as O
from tests.plan_emulator import bank_conflicts, emulate, parse_plan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "gpuntt_b200.h")).read()
    names = set(re.findall(r"\b(gpuntt_b200_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 10
    L = capi.lib()
    for nme in names:
        assert hasattr(L, nme), f"{nme} declared in include/gpuntt_b200.h but not exported"
    assert L.gpuntt_b200_version() == 100


def test_plain_c_caller_compiles_links_and_runs(tmp_path):
    """include/gpuntt_b200.h is valid C11 (-pedantic) and the shared library links into a C program: what a cgo / JNI / N-API
    binding of INTEGRATION.md builds on.  tests/c_caller.c only makes calls that return before any CUDA work."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    capi.lib()
    libdir = os.path.join(ROOT, "gpu_ntt_b200", "lib")
    exe = str(tmp_path / "c_caller")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_caller.c"), "-o", exe, "-L", libdir, "-lgpuntt_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "c caller ok" in out.stdout, out.stdout + out.stderr


def test_argument_validation_without_gpu():
    """Error behaviour mirrors the reference's exceptions (ntt.cu:2088-2091, 2253) as status codes."""
    with pytest.raises(G.GpuNttError) as e:
        capi.merge_ntt(in_ptr=16, out_ptr=16, table_ptr=16, n_power=0, batch=1, modulus=17)
    assert e.value.status == capi.ERR_N_POWER and "Invalid n_power range!" in e.value.message
    with pytest.raises(G.GpuNttError) as e:
        capi.merge_ntt(in_ptr=16, out_ptr=16, table_ptr=16, n_power=29, batch=1, modulus=17)
    assert e.value.status == capi.ERR_N_POWER
    with pytest.raises(G.GpuNttError) as e:
        capi.merge_ntt(in_ptr=16, out_ptr=16, table_ptr=16, n_power=4, batch=1, modulus=17, layout=7)
    assert e.value.status == capi.ERR_LAYOUT and "Invalid ntt_layout!" in e.value.message
    with pytest.raises(G.GpuNttError) as e:
        capi.merge_ntt(in_ptr=0, out_ptr=16, table_ptr=16, n_power=4, batch=1, modulus=17)
    assert e.value.status == capi.ERR_ARGUMENT
    # empty batch is a no-op, exactly like launching a zero-sized grid would be
    capi.merge_ntt(in_ptr=0, out_ptr=0, table_ptr=0, n_power=4, batch=0, modulus=17)


@pytest.mark.parametrize("bits", [64, 32])
def test_plans_are_well_formed(bits):
    for n in range(1, 29):
        ps = parse_plan(capi.describe_plan(n, bits))
        assert sum(p["d"] for p in ps) == n
        lo_expect = n
        for p in ps:  # passes go from the high bits down, covering every bit exactly once
            lo_expect -= p["d"]
            assert p["lo"] == lo_expect
            assert sum(p["rounds"]) == p["d"] and all(1 <= r <= (5 if bits == 32 else 4) for r in p["rounds"])
            assert (bits // 8) << p["k"] <= 64 * 1024
            assert p["k"] == p["d"] + p["c"] if p["lo"] else p["k"] >= p["d"]
            assert p["c"] <= p["lo"]
        assert ps[-1]["lo"] == 0 and ps[-1]["c"] == 0


@pytest.mark.parametrize("bits", [64, 32])
def test_plans_are_bank_conflict_free(bits):
    for n in (5, 10, 12, 13, 14, 16, 17, 20, 24, 28):
        worst = bank_conflicts(parse_plan(capi.describe_plan(n, bits)), bits)
        assert max(worst.values()) == 1, (n, worst)


@pytest.mark.parametrize("bits,n,poly,batch", [(64, 3, 1, 5), (64, 11, 0, 3), (64, 12, 1, 1), (64, 14, 1, 1),
                                               (32, 5, 0, 70), (32, 13, 1, 1), (32, 15, 0, 1), (64, 14, 0, 2)])
def test_plan_emulation_matches_oracle(bits, n, poly, batch):
    P = O.merge_params(n, poly, bits)
    x = O.example_input(P.modulus, batch << n, seed=n)
    passes = parse_plan(capi.describe_plan(n, bits))
    plus = 1 if poly == O.X_N_plus else 0
    y = emulate([int(v) for v in x], n, P.modulus, [int(v) for v in P.fwd_br], plus, passes, False)
    assert y == [int(v) for v in O.merge_ntt(x, P)]
    z = emulate(y, n, P.modulus, [int(v) for v in P.inv_br], plus, passes, True, n_inv=P.n_inv)
    assert z == [int(v) for v in x]
