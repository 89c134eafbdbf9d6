#!/usr/bin/env python
"""Device-time of the BASELINE.json configurations other than the bench.py headline (C3, C4) and a sweep over ring
sizes, through the C ABI, CUDA events, data resident in HBM.  Timing only (parity is tests/ -m gpu): the 4-step
tables are random residues, the merge tables come from NTTParameters.  One JSON object per line.

    python tools/perf_configs.py [--quick]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from gpu_ntt_b200 import capi  # noqa: E402
from gpu_ntt_b200.params import NTTParameters, X_N_minus, X_N_plus  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def time_ms(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def dev(a, bits):
    if bits == 64:
        return torch.from_numpy(np.ascontiguousarray(a).astype(np.uint64).view(np.int64)).cuda()
    return torch.from_numpy(np.ascontiguousarray(a).astype(np.uint32).view(np.int32)).cuda()


def merge_case(name, logn, batch, bits, poly, iters, inverse=False, both=False):
    P = NTTParameters(logn, poly, bits)
    p = P.modulus
    tab = dev(P.gpu_root_of_unity_table_generator(P.forward_root_of_unity_table), bits)
    itab = dev(P.gpu_root_of_unity_table_generator(P.inverse_root_of_unity_table), bits)
    dt = torch.int64 if bits == 64 else torch.int32
    x = torch.randint(0, p, (batch, 1 << logn), dtype=dt, device="cuda")

    def fwd():
        capi.ntt(x, tab, p, logn, poly)

    def inv():
        capi.intt(x, itab, p, P.n_inv, logn, poly)

    def rt():
        fwd()
        inv()
    fn = rt if both else (inv if inverse else fwd)
    ms = time_ms(fn, iters)
    lib = capi.lib()
    lib.gpuntt_b200_set_profiling(1)
    capi.profile_read()
    fn()
    launches = [(k, round(m, 4)) for k, m in capi.profile_read()]
    lib.gpuntt_b200_set_profiling(0)
    ntts = batch * (2 if both else 1)
    bytes_alg = 2 * (1 << logn) * (bits // 8) * ntts
    gbs = bytes_alg / (ms * 1e-3) / 1e9
    out = {"case": name, "logn": logn, "batch": batch, "bits": bits, "ring": "X^N-1" if poly == X_N_minus else "X^N+1",
           "op": "fwd+inv" if both else ("inv" if inverse else "fwd"), "ms": round(ms, 4),
           "ntt_per_s": round(ntts / (ms * 1e-3), 1), "alg_GBps": round(gbs, 1), "frac_hbm": round(gbs / peak(), 4),
           "launches_kind_ms": launches, "plan": capi.describe_plan(logn, bits).strip()}
    print(json.dumps(out), flush=True)
    del x
    torch.cuda.empty_cache()


def rns_case(name, logn, batch, mod_count, iters, top_log2=59):
    """RNS forward (GPU_NTT RNS overload): mod_count NTT-friendly primes below 2^top_log2, random tables (timing only)."""
    sys.path.insert(0, os.path.join(ROOT))
    from tests.test_merge_gpu import rns_primes
    primes = [p for p, _ in rns_primes(64, logn, mod_count, top_log2)]
    rng = np.random.default_rng(2)
    n = 1 << logn
    tab = np.concatenate([rng.integers(1, p, n, dtype=np.uint64) for p in primes])
    mods = np.array([[p, p.bit_length(), 0] for p in primes], dtype=np.uint64).ravel()
    d_tab, d_mods = dev(tab, 64), dev(mods, 64)
    x = torch.randint(0, min(primes), (batch, n), dtype=torch.int64, device="cuda")
    res = {}
    for label, force in (("tuned", 0), ("generic", 1)):
        capi.lib().gpuntt_b200_force_generic_path(force)

        def fn():
            capi.merge_ntt(in_ptr=x.data_ptr(), out_ptr=x.data_ptr(), table_ptr=d_tab.data_ptr(), n_power=logn, batch=batch,
                           element_bits=64, direction=capi.FORWARD, reduction_poly=X_N_plus, mod_count=mod_count,
                           modulus_dev=d_mods.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
        res[label] = round(time_ms(fn, iters), 4)
    capi.lib().gpuntt_b200_force_generic_path(0)
    gbs = 2 * n * 8 * batch / (res["tuned"] * 1e-3) / 1e9
    print(json.dumps({"case": name, "logn": logn, "batch": batch, "mod_count": mod_count, "prime_bits": top_log2, "ms": res["tuned"],
                      "ms_generic_kernel": res["generic"], "ntt_per_s": round(batch / (res["tuned"] * 1e-3), 1),
                      "alg_GBps": round(gbs, 1), "frac_hbm": round(gbs / peak(), 4)}), flush=True)


def fourstep_case(name, logn, batch, iters, contract, inverse=False):
    n1, n2 = capi.fourstep_shape(logn)
    p = 576460753175838721 if logn == 24 else 576460752303415297
    rng = np.random.default_rng(1)
    t1 = dev(rng.integers(1, p, n1 // 2, dtype=np.uint64), 64)
    t2 = dev(rng.integers(1, p, n2 // 2, dtype=np.uint64), 64)
    t2[0] = 1
    t1[0] = 1
    w = torch.randint(0, p, (1 << logn,), dtype=torch.int64, device="cuda")
    x = torch.randint(0, p, (batch, 1 << logn), dtype=torch.int64, device="cuda")
    out = torch.empty_like(x) if contract == capi.FOURSTEP_REFERENCE else None

    def fn():
        if inverse:
            capi.fourstep_ntt(x, t1, t2, w, p, logn, direction=capi.INVERSE, mod_inverse=12345, io_contract=contract, out=out)
        else:
            capi.fourstep_ntt(x, t1, t2, w, p, logn, io_contract=contract, out=out)
    ms = time_ms(fn, iters, warm=2)
    lib = capi.lib()
    lib.gpuntt_b200_set_profiling(1)
    capi.profile_read()
    fn()
    launches = [(k, round(m, 4)) for k, m in capi.profile_read()]   # kind 9 = transpose, 0 = twiddle prep, k = k-th pass
    lib.gpuntt_b200_set_profiling(0)
    bytes_alg = 2 * (1 << logn) * 8 * batch
    gbs = bytes_alg / (ms * 1e-3) / 1e9
    print(json.dumps({"case": name, "logn": logn, "batch": batch, "bits": 64, "n1": n1, "n2": n2,
                      "contract": "fused" if contract == capi.FOURSTEP_FUSED else "reference", "op": "inv" if inverse else "fwd", "ms": round(ms, 4),
                      "ntt_per_s": round(batch / (ms * 1e-3), 2), "alg_GBps": round(gbs, 1),
                      "frac_hbm": round(gbs / peak(), 4), "launches_kind_ms": launches}), flush=True)
    del x, w, out
    torch.cuda.empty_cache()


def percoef_case(name, logh, w, iters, bits=64):
    """NTTLayout::PerCoefficient: one [2^logh][w] matrix, every column a transform; tuned strided passes vs the generic kernel"""
    P = NTTParameters(logh, X_N_plus, bits)
    tab = dev(P.gpu_root_of_unity_table_generator(P.forward_root_of_unity_table), bits)
    x = torch.randint(0, P.modulus, (1 << logh, w), dtype=torch.int64 if bits == 64 else torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream

    def fn():
        capi.merge_ntt(in_ptr=x.data_ptr(), out_ptr=x.data_ptr(), table_ptr=tab.data_ptr(), n_power=logh, batch=w, element_bits=bits,
                       direction=capi.FORWARD, reduction_poly=X_N_plus, layout=capi.PerCoefficient, modulus=P.modulus, stream=s)
    res = {}
    for tag, generic in (("tuned", 0), ("generic", 1)):
        capi.lib().gpuntt_b200_force_generic_path(generic)
        res[tag] = time_ms(fn, iters)
        if not generic:
            launches = capi.lib().gpuntt_b200_last_launch_count()
    capi.lib().gpuntt_b200_force_generic_path(0)
    bytes_alg = 2 * (1 << logh) * (bits // 8) * w
    gbs = bytes_alg / (res["tuned"] * 1e-3) / 1e9
    print(json.dumps({"case": name, "logn": logh, "batch": w, "bits": bits, "layout": "PerCoefficient", "ms": round(res["tuned"], 4),
                      "ms_generic_kernel": round(res["generic"], 4), "launches": launches, "ntt_per_s": round(w / (res["tuned"] * 1e-3), 1),
                      "alg_GBps": round(gbs, 1), "frac_hbm": round(gbs / peak(), 4)}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--percoeff-only", action="store_true")
    args = ap.parse_args()
    it = 5 if args.quick else 20
    capi.lib()
    if args.percoeff_only:
        percoef_case("PerCoefficient 2^9 x 131072 (X^N+1)", 9, 131072, it)
        percoef_case("PerCoefficient Data32 2^9 x 262144 (X^N+1)", 9, 262144, it, bits=32)
        percoef_case("PerCoefficient Data32 2^8 x 524288 (X^N+1)", 8, 524288, it, bits=32)
        percoef_case("PerCoefficient Data32 2^9 x 1024 (the reference example's shape)", 9, 1024, it, bits=32)
        return
    merge_case("C2 fwd", 16, 1024, 64, X_N_minus, it)
    merge_case("C2 inv", 16, 1024, 64, X_N_minus, it, inverse=True)
    merge_case("C2 negacyclic fwd", 16, 1024, 64, X_N_plus, it)
    merge_case("C3 fwd+inv", 14, 4096, 32, X_N_minus, it, both=True)
    merge_case("C3 fwd", 14, 4096, 32, X_N_minus, it)
    rns_case("RNS fwd 4 x 59-bit primes", 16, 1024, 4, it)
    rns_case("RNS fwd 4 x 61-bit primes (exact kernels)", 16, 1024, 4, it, top_log2=61)
    fourstep_case("C4 4-step fused", 24, 16, max(2, it // 4), capi.FOURSTEP_FUSED)
    fourstep_case("C4 4-step reference contract", 24, 16, max(2, it // 4), capi.FOURSTEP_REFERENCE)
    fourstep_case("C4 4-step inverse fused", 24, 16, max(2, it // 4), capi.FOURSTEP_FUSED, inverse=True)
    fourstep_case("4-step fused logN=20", 20, 64, max(2, it // 4), capi.FOURSTEP_FUSED)
    percoef_case("PerCoefficient 2^9 x 131072 (X^N+1)", 9, 131072, it)
    percoef_case("PerCoefficient 2^8 x 262144 (X^N+1)", 8, 262144, it)
    percoef_case("PerCoefficient 2^9 x 1024 (the reference example's shape)", 9, 1024, it)
    percoef_case("PerCoefficient Data32 2^9 x 262144 (X^N+1)", 9, 262144, it, bits=32)
    percoef_case("PerCoefficient Data32 2^8 x 524288 (X^N+1)", 8, 524288, it, bits=32)
    if not args.quick:
        for logn in (8, 10, 11, 12, 13, 14, 15, 17, 18, 20, 22, 24):
            merge_case(f"u64 logN={logn}", logn, max(1, (1 << 26) >> logn), 64, X_N_minus, it)
        for logn in (10, 12, 16):
            merge_case(f"u32 logN={logn}", logn, max(1, (1 << 27) >> logn), 32, X_N_minus, it)


if __name__ == "__main__":
    main()
