#!/usr/bin/env python
"""Randomised differential test of the Merge-NTT entry point against the oracle (test infrastructure; run on a GPU box):
random (width, ring size, ring type, batch, modulus anywhere in the accepted range, direction, in / out of place, signed I/O,
tuned / generic kernels, single-launch knob; three cases in ten through the RNS overloads with 1..5 random moduli, a
share through GPU_4STEP_NTT in both I/O contracts and directions, one in ten in NTTLayout::PerCoefficient) with extreme
inputs mixed in; every output word is compared with NTTCPU's
restatement.  Usage: fuzz_parity.py [seconds] [seed] [4-step share] [PerCoefficient share].  Prints one JSON line per mismatch and a summary line; exit code 1 on
any mismatch."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gpu_ntt_b200 import capi  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.gpu_util import to_dev, to_host, to_host_signed  # noqa: E402
from tests.test_moduli_gpu import custom_params, ntt_prime_below  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
fourstep_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1     # share of the cases that go through GPU_4STEP_NTT
percoeff_share = float(sys.argv[4]) if len(sys.argv) > 4 else 0.1     # ... in NTTLayout::PerCoefficient
rng = np.random.default_rng(seed)
lib = capi.lib()
params_cache = {}


def params(logn, poly, p):
    key = (logn, poly, p)
    if key not in params_cache:
        if len(params_cache) > 64:
            params_cache.clear()
        params_cache[key] = custom_params(logn, poly, p)
    return params_cache[key]


def pick_case():
    bits = int(rng.choice([32, 64]))
    logn = int(rng.choice([int(rng.integers(1, 13)), int(rng.integers(7, 18)), int(rng.integers(12, 21))]))
    poly = int(rng.choice([O.X_N_minus, O.X_N_plus]))
    two_n = 2 << logn
    max_bits = 30 if bits == 32 else 62
    lo_bits = logn + 3
    # modulus: a prime p = 1 (mod 2N) below a random limit; half of the draws sit next to a policy boundary
    if rng.random() < 0.5:
        edges = [1 << 29, (1 << 29) + (1 << 22), (1 << 30) - 1] if bits == 32 else \
            [1 << 36, (1 << 36) + (1 << 31), 1 << 40, (1 << 40) + (1 << 34), (1 << 60) - (1 << 31), (1 << 60) - (1 << 31) + (1 << 40),
             (1 << 60) + (1 << 58), (1 << 60) + (1 << 58) + (1 << 45), (1 << 62) - 1]
        limit = int(edges[int(rng.integers(0, len(edges)))])
    else:
        b = int(rng.integers(lo_bits, max_bits + 1))
        limit = (1 << b) - int(rng.integers(0, 1 << (b - 2)))
    if limit <= two_n + 1:
        return None
    try:
        p = ntt_prime_below(limit, two_n)
    except ValueError:
        return None
    cap = max(1, (1 << 21) >> logn)
    chunk = max(1, ((1 << 11) if bits == 64 else (1 << 12)) >> logn)
    batch = int(rng.choice([1, 2, 3, int(rng.integers(1, cap + 1)), min(cap, chunk * int(rng.integers(1, 9))),
                            min(cap, chunk * int(rng.integers(1, 9)) + 1), max(1, min(cap, chunk - 1))]))
    signed = bool(rng.random() < 0.25)     # (the reference's signed overloads are out of place: T* in, TU* out)
    return dict(bits=bits, logn=logn, poly=poly, p=p, batch=batch, inverse=bool(rng.integers(0, 2)),
                inplace=bool(rng.integers(0, 2)) and not signed, signed=signed, generic=bool(rng.random() < 0.15), fused=int(rng.choice([1, 1, 2, 0])),
                extreme=int(rng.integers(0, 4)))


def run_case(c):
    bits, logn, poly, p, batch = c["bits"], c["logn"], c["poly"], c["p"], c["batch"]
    P = params(logn, poly, p)
    n = 1 << logn
    x = rng.integers(0, p, size=(batch, n), dtype=np.uint64)
    if c["extreme"] == 1:
        x[0, :] = p - 1
    elif c["extreme"] == 2:
        x[-1, ::2] = p - 1
        x[-1, 1::2] = 0
    elif c["extreme"] == 3:
        x[:, : min(n, 8)] = p - 1
    lib.gpuntt_b200_force_generic_path(1 if c["generic"] else 0)
    capi.tune(capi.TUNE_FUSED_PASSES, c["fused"])
    st = torch.cuda.current_stream().cuda_stream
    try:
        if not c["inverse"]:
            want = O.merge_ntt(x, P)
            if c["signed"]:
                # signed input: residues above p/2 are handed over as negative numbers (test_merge_ntt.cu:184-341)
                sx = O.centered(x, p)
                d = torch.from_numpy(sx if bits == 64 else sx.astype(np.int32)).cuda()
            else:
                d = to_dev(x, bits)
            out = d if c["inplace"] else torch.zeros_like(d)
            capi.merge_ntt(in_ptr=d.data_ptr(), out_ptr=out.data_ptr(), table_ptr=to_dev(P.fwd_br, bits).data_ptr(), n_power=logn,
                           batch=batch, element_bits=bits, direction=capi.FORWARD, reduction_poly=poly, modulus=p,
                           is_signed=c["signed"], stream=st)
            torch.cuda.synchronize()
            got = to_host(out, bits).reshape(batch, n)
            return bool((got == want).all())
        want = O.merge_intt(x, P)
        d = to_dev(x, bits)
        out = d if c["inplace"] else torch.zeros_like(d)
        capi.merge_ntt(in_ptr=d.data_ptr(), out_ptr=out.data_ptr(), table_ptr=to_dev(P.inv_br, bits).data_ptr(), n_power=logn,
                       batch=batch, element_bits=bits, direction=capi.INVERSE, reduction_poly=poly, modulus=p, mod_inverse=P.n_inv,
                       is_signed=c["signed"], stream=st)
        torch.cuda.synchronize()
        if c["signed"]:
            got = to_host_signed(out).reshape(batch, n)
            return bool((got == O.centered(want, p).reshape(batch, n)).all())
        return bool((to_host(out, bits).reshape(batch, n) == want).all())
    finally:
        lib.gpuntt_b200_force_generic_path(0)
        capi.tune(capi.TUNE_FUSED_PASSES, 1)


def random_prime(bits, two_n, logn):
    max_bits = 30 if bits == 32 else 62
    for _ in range(20):
        b = int(rng.integers(logn + 3, max_bits + 1))
        limit = (1 << b) - int(rng.integers(0, 1 << (b - 2)))
        if limit <= two_n + 1:
            continue
        try:
            return ntt_prime_below(limit, two_n)
        except ValueError:
            continue
    return None


def pick_rns_case():
    bits = int(rng.choice([32, 64]))
    logn = int(rng.choice([int(rng.integers(2, 12)), int(rng.integers(10, 18))]))
    mc = int(rng.integers(1, 6))
    primes = []
    for _ in range(mc):
        p = random_prime(bits, 2 << logn, logn)
        if p is None or p in primes:
            return None
        primes.append(p)
    cap = max(mc, (1 << 20) >> logn)
    per_slot = int(rng.integers(1, max(2, cap // mc + 1)))
    batch = per_slot * mc if rng.random() < 0.8 else int(rng.integers(1, cap + 1))    # (not a multiple of mod_count: generic kernel)
    return dict(rns=True, bits=bits, logn=logn, poly=int(rng.choice([O.X_N_minus, O.X_N_plus])), primes=primes, batch=batch,
                inverse=bool(rng.integers(0, 2)), inplace=bool(rng.integers(0, 2)), signed=False, generic=bool(rng.random() < 0.1),
                fused=int(rng.choice([1, 2, 0])))


def run_rns_case(c):
    """RNS overloads: polynomial b uses modulus[b % mod_count], table slice (b % mod_count) << n_power, mod_inverse[b % mod_count]
    (ntt.cu:613-619, 672-673, 1225-1226 of the reference)."""
    bits, logn, poly, primes, batch = c["bits"], c["logn"], c["poly"], c["primes"], c["batch"]
    mc, n = len(primes), 1 << logn
    Ps = [params(logn, poly, p) for p in primes]
    tab = np.zeros(mc << logn, dtype=np.uint64)
    mods = np.zeros((mc, 3), dtype=np.uint64)
    ninvs = np.zeros(mc, dtype=np.uint64)
    for m, P in enumerate(Ps):
        t = P.inv_br if c["inverse"] else P.fwd_br
        tab[m << logn:(m << logn) + t.size] = t
        bit, mu = O.modulus(P.modulus, bits)
        mods[m] = (P.modulus, bit, mu)
        ninvs[m] = P.n_inv
    x = np.stack([rng.integers(0, primes[b % mc], size=n, dtype=np.uint64) for b in range(batch)])
    x[0, :] = primes[0] - 1
    fn = O.merge_intt if c["inverse"] else O.merge_ntt
    want = np.stack([fn(x[b], Ps[b % mc]).reshape(n) for b in range(batch)])
    lib.gpuntt_b200_force_generic_path(1 if c["generic"] else 0)
    capi.tune(capi.TUNE_FUSED_PASSES, c["fused"])
    try:
        d = to_dev(x, bits)
        out = d if c["inplace"] else torch.zeros_like(d)
        mods_d, ninv_d, tab_d = to_dev(mods.ravel(), bits), to_dev(ninvs, bits), to_dev(tab, bits)
        capi.merge_ntt(in_ptr=d.data_ptr(), out_ptr=out.data_ptr(), table_ptr=tab_d.data_ptr(), n_power=logn, batch=batch,
                       element_bits=bits, direction=capi.INVERSE if c["inverse"] else capi.FORWARD, reduction_poly=poly, mod_count=mc,
                       modulus_dev=mods_d.data_ptr(), mod_inverse_dev=ninv_d.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        return bool((to_host(out, bits).reshape(batch, n) == want).all())
    finally:
        lib.gpuntt_b200_force_generic_path(0)
        capi.tune(capi.TUNE_FUSED_PASSES, 1)


def pick_percoeff_case():
    bits = int(rng.choice([32, 64, 64]))
    logn = int(rng.integers(1, 10))
    p = None
    if bits == 64 and rng.random() < 0.4:
        edges = [1 << 36, 1 << 40, (1 << 40) + (1 << 34), (1 << 60) - (1 << 31), (1 << 60) - (1 << 31) + (1 << 40), (1 << 60) + (1 << 58),
                 (1 << 60) + (1 << 58) + (1 << 45), (1 << 62) - 1]
        p = ntt_prime_below(int(edges[int(rng.integers(0, len(edges)))]), 2 << logn)
    else:
        p = random_prime(bits, 2 << logn, logn)
    if p is None:
        return None
    w = 1 << int(rng.integers(0, 21 - logn))       # the reference's limit for this layout: a power-of-two batch (ntt.cu:2235)
    signed = bool(rng.random() < 0.2)
    return dict(percoeff=True, bits=bits, logn=logn, poly=int(rng.choice([O.X_N_minus, O.X_N_plus])), p=p, batch=w,
                inverse=bool(rng.integers(0, 2)), inplace=bool(rng.integers(0, 2)) and not signed, signed=signed,
                generic=bool(rng.random() < 0.15))


def run_percoeff_case(c):
    """NTTLayout::PerCoefficient (ForwardCoreTranspose / InverseCoreTranspose, ntt.cu:1554-2074): one [N][batch] matrix whose columns
    are the transforms."""
    bits, logn, poly, p, w = c["bits"], c["logn"], c["poly"], c["p"], c["batch"]
    P = params(logn, poly, p)
    h = 1 << logn
    x = rng.integers(0, p, size=(h, w), dtype=np.uint64)
    x[:, 0] = p - 1
    fn = O.merge_intt if c["inverse"] else O.merge_ntt
    want = fn(np.ascontiguousarray(x.T), P).reshape(w, h).T
    lib.gpuntt_b200_force_generic_path(1 if c["generic"] else 0)
    try:
        if c["signed"] and not c["inverse"]:
            sx = O.centered(x, p)
            d = torch.from_numpy(np.ascontiguousarray(sx if bits == 64 else sx.astype(np.int32))).cuda()
        else:
            d = to_dev(x, bits)
        out = d if c["inplace"] else torch.zeros_like(d)
        capi.merge_ntt(in_ptr=d.data_ptr(), out_ptr=out.data_ptr(), table_ptr=to_dev(P.inv_br if c["inverse"] else P.fwd_br, bits).data_ptr(),
                       n_power=logn, batch=w, element_bits=bits, direction=capi.INVERSE if c["inverse"] else capi.FORWARD,
                       reduction_poly=poly, layout=capi.PerCoefficient, modulus=p, mod_inverse=P.n_inv if c["inverse"] else 0,
                       is_signed=c["signed"], stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        if c["signed"] and c["inverse"]:
            return bool((to_host_signed(out).reshape(h, w) == O.centered(want, p).reshape(h, w)).all())
        return bool((to_host(out, bits).reshape(h, w) == want).all())
    finally:
        lib.gpuntt_b200_force_generic_path(0)


fs_cache = {}


def pick_4step_case():
    bits = int(rng.choice([32, 64, 64]))
    logn = int(rng.choice([int(rng.integers(12, 18)), int(rng.integers(12, 22))]))
    n = 1 << logn
    max_bits = 30 if bits == 32 else 62
    if rng.random() < 0.4 and bits == 64:
        edges = [1 << 40, (1 << 40) + (1 << 34), (1 << 60) - (1 << 31), (1 << 60) - (1 << 31) + (1 << 40), (1 << 60) + (1 << 58),
                 (1 << 60) + (1 << 58) + (1 << 45), (1 << 62) - 1]
        limit = int(edges[int(rng.integers(0, len(edges)))])
    else:
        b = int(rng.integers(logn + 3, max_bits + 1))
        limit = (1 << b) - int(rng.integers(0, 1 << (b - 2)))
    try:
        p = ntt_prime_below(limit, n)
    except ValueError:
        return None
    batch = int(rng.integers(1, max(2, min(9, ((1 << 22) >> logn) + 1))))
    fused = bool(rng.integers(0, 2))
    return dict(fourstep=True, bits=bits, logn=logn, p=p, batch=batch, inverse=bool(rng.integers(0, 2)), fused_contract=fused,
                inplace=fused and bool(rng.integers(0, 2)), signed=False, generic=bool(rng.random() < 0.1),
                transposed_knob=int(rng.choice([1, 1, 0])))


def run_4step_case(c):
    """GPU_4STEP_NTT in both I/O contracts (the reference's, between the caller's transposes, and the fused natural-order one)
    against NTT_4STEP_CPU (test_4step_ntt.cu:147-166, test_4step_intt.cu:82-84, 155-166)."""
    from tests.test_4step_gpu import custom_fourstep_params, tables, transposed
    bits, logn, p, batch = c["bits"], c["logn"], c["p"], c["batch"]
    if (logn, p) not in fs_cache:
        if len(fs_cache) > 8:
            fs_cache.clear()
        fs_cache[(logn, p)] = custom_fourstep_params(logn, p)
    P = fs_cache[(logn, p)]
    n = P.n
    x = rng.integers(0, p, size=(batch, n), dtype=np.uint64)
    x[0, ::3] = p - 1
    inv = c["inverse"]
    want = O.fourstep_intt(x, P) if inv else O.fourstep_ntt(x, P)
    t1, t2, W = tables(P, bits, inv)
    lib.gpuntt_b200_force_generic_path(1 if c["generic"] else 0)
    capi.tune(3, c["transposed_knob"])      # knob 3: transposing stores inside the passes / transpose kernels
    try:
        if c["fused_contract"]:
            d = to_dev(x, bits)
            out = d if c["inplace"] else torch.zeros_like(d)
            capi.fourstep_ntt(d.view(batch, n), t1, t2, W, p, logn, direction=capi.INVERSE if inv else capi.FORWARD,
                              mod_inverse=P.n_inv if inv else 0, out=out.view(batch, n))
            torch.cuda.synchronize()
            return bool((to_host(out, bits).reshape(batch, n) == want).all())
        src = O.fourstep_intt_first_transpose(x, P) if inv else transposed(x, P.n1, P.n2)
        d = to_dev(src, bits)
        r = torch.zeros_like(d)
        capi.fourstep_ntt(d.view(batch, n), t1, t2, W, p, logn, direction=capi.INVERSE if inv else capi.FORWARD,
                          mod_inverse=P.n_inv if inv else 0, io_contract=capi.FOURSTEP_REFERENCE, out=r.view(batch, n))
        torch.cuda.synchronize()
        return bool((transposed(to_host(r, bits), P.n1, P.n2).reshape(batch, n) == want).all())
    finally:
        lib.gpuntt_b200_force_generic_path(0)
        capi.tune(3, 1)


def main():
    t0 = time.time()
    done = bad = 0
    kinds = {}
    while time.time() - t0 < budget:
        u = rng.random()
        rns, fs, pc = u < 0.3, (u >= 0.3 and u < 0.3 + fourstep_share), u >= 1.0 - percoeff_share
        c = pick_rns_case() if rns else pick_4step_case() if fs else pick_percoeff_case() if pc else pick_case()
        if c is None:
            continue
        try:
            ok = run_rns_case(c) if rns else run_4step_case(c) if fs else run_percoeff_case(c) if pc else run_case(c)
        except Exception as e:  # an error code from the library is a finding as well
            ok = False
            c["error"] = repr(e)[:200]
        done += 1
        k = (c["bits"], "inv" if c["inverse"] else "fwd", "rns" if rns else "4step" if fs else "percoeff" if pc else "signed" if c["signed"] else "unsigned",
             "generic" if c["generic"] else "tuned")
        kinds[k] = kinds.get(k, 0) + 1
        if not ok:
            bad += 1
            print(json.dumps({"mismatch": c}), flush=True)
    print(json.dumps({"fuzz_cases": done, "mismatches": bad, "seconds": round(time.time() - t0, 1), "seed": seed,
                      "by_kind": {"/".join(map(str, k)): v for k, v in sorted(kinds.items())}}))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
