#!/usr/bin/env python
"""One call of every kernel family of the library at a saturating size (run under ncu by tools/gpu_full_pass.sh: one metrics row per
kernel for profiles/r2_v4_kernel_families_ncu.txt).  Random tables where parity does not matter (timing / counters only)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gpu_ntt_b200 import capi  # noqa: E402
from gpu_ntt_b200.params import NTTParameters, X_N_minus, X_N_plus  # noqa: E402
from perf_configs import dev  # noqa: E402

lib = capi.lib()


def merge(logn, batch, bits, inverse=False, fused=1, generic=0, poly=X_N_minus):
    P = NTTParameters(logn, poly, bits)
    tab = dev(P.gpu_root_of_unity_table_generator(P.inverse_root_of_unity_table if inverse else P.forward_root_of_unity_table), bits)
    x = torch.randint(0, P.modulus, (batch, 1 << logn), dtype=torch.int64 if bits == 64 else torch.int32, device="cuda")
    capi.tune(capi.TUNE_FUSED_PASSES, fused)
    lib.gpuntt_b200_force_generic_path(generic)
    if inverse:
        capi.intt(x, tab, P.modulus, P.n_inv, logn, poly)
    else:
        capi.ntt(x, tab, P.modulus, logn, poly)
    torch.cuda.synchronize()
    capi.tune(capi.TUNE_FUSED_PASSES, 1)
    lib.gpuntt_b200_force_generic_path(0)


def rns(logn, batch, mc, fused, top=59):
    from tests.test_merge_gpu import rns_primes
    primes = [p for p, _ in rns_primes(64, logn, mc, top)]
    rng = np.random.default_rng(2)
    n = 1 << logn
    tab = dev(np.concatenate([rng.integers(1, p, n, dtype=np.uint64) for p in primes]), 64)
    mods = dev(np.array([[p, p.bit_length(), 0] for p in primes], dtype=np.uint64).ravel(), 64)
    x = torch.randint(0, min(primes), (batch, n), dtype=torch.int64, device="cuda")
    capi.tune(capi.TUNE_FUSED_PASSES, fused)
    capi.merge_ntt(in_ptr=x.data_ptr(), out_ptr=x.data_ptr(), table_ptr=tab.data_ptr(), n_power=logn, batch=batch, element_bits=64,
                   direction=capi.FORWARD, reduction_poly=X_N_plus, mod_count=mc, modulus_dev=mods.data_ptr(),
                   stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    capi.tune(capi.TUNE_FUSED_PASSES, 1)


def fourstep(logn, batch, contract, inverse=False, resident=1):
    n1, n2 = capi.fourstep_shape(logn)
    p = 576460753175838721 if logn == 24 else 576460752303415297
    rng = np.random.default_rng(1)
    t1 = dev(rng.integers(1, p, n1 // 2, dtype=np.uint64), 64)
    t2 = dev(rng.integers(1, p, n2 // 2, dtype=np.uint64), 64)
    t1[0] = 1
    t2[0] = 1
    w = torch.randint(0, p, (1 << logn,), dtype=torch.int64, device="cuda")
    x = torch.randint(0, p, (batch, 1 << logn), dtype=torch.int64, device="cuda")
    out = torch.empty_like(x) if contract == capi.FOURSTEP_REFERENCE else None
    capi.tune(5, resident)
    capi.fourstep_ntt(x, t1, t2, w, p, logn, direction=capi.INVERSE if inverse else capi.FORWARD, mod_inverse=12345 if inverse else 0,
                      io_contract=contract, out=out)
    torch.cuda.synchronize()
    capi.tune(5, 1)


only = set(os.environ.get("GPUNTT_FAMILIES_ONLY", "").split(",")) - {""}  # e.g. GPUNTT_FAMILIES_ONLY=23,24: just those cases


def case(label, fn):
    if only and label.split()[0] not in only:
        return
    print(label)
    fn()


case("1 C2 forward, one launch per pass", lambda: merge(16, 1024, 64, fused=0))
case("2 C2 inverse, one launch per pass", lambda: merge(16, 1024, 64, inverse=True, fused=0))
case("3 C2 forward, single-launch kernel", lambda: merge(16, 1024, 64, fused=2))
case("4 C3 forward, one launch per pass (32-bit shapes)", lambda: merge(14, 4096, 32, fused=0))
case("5 C3 inverse, one launch per pass", lambda: merge(14, 4096, 32, inverse=True, fused=0))
case("6 C3 forward, single-launch kernel", lambda: merge(14, 4096, 32, fused=2))
case("7 C3 inverse, single-launch kernel", lambda: merge(14, 4096, 32, inverse=True, fused=2))
case("8 small ring 64-bit 2^10 x 65536 (fast_small, three rounds)", lambda: merge(10, 65536, 64))
case("9 small ring 32-bit 2^12 x 32768", lambda: merge(12, 32768, 32))
case("10 three-pass ring 64-bit 2^20 x 64", lambda: merge(20, 64, 64))
case("11 RNS forward 4 x 59-bit, 2^16 x 1024 (fast_pass_dual_kernel)", lambda: rns(16, 1024, 4, 0))
case("12 RNS forward 4 x 59-bit, 2^14 x 32 (fused2_rns_kernel)", lambda: rns(14, 32, 4, 2))
case("13 generic path 64-bit 2^16 x 256 (twiddle_prep_kernel + merge_pass_kernel x 2)", lambda: merge(16, 256, 64, generic=1))
case("14 generic path 32-bit 2^14 x 1024", lambda: merge(14, 1024, 32, generic=1))
case("15 C4 fused contract forward (w_pairs_kernel, wcol_kernel strided + transposing store, two strided row passes)", lambda: fourstep(24, 16, capi.FOURSTEP_FUSED))
case("16 C4 fused forward with per-tile pairs (fast_pass_kernel WMUL + TS)", lambda: fourstep(24, 16, capi.FOURSTEP_FUSED, resident=0))
case("17 C4 reference contract forward (w_pairs_tile_kernel, wcol_kernel contiguous, strided row pass, strided row pass + transposing store)", lambda: fourstep(24, 16, capi.FOURSTEP_REFERENCE))
case("18 C4 fused inverse (pairs, strided n1 pass + transposing store, wcol_kernel inverse product pass, last strided pass)", lambda: fourstep(24, 16, capi.FOURSTEP_FUSED, inverse=True))
case("19 C4 reference contract inverse (pairs, contiguous n1 pass, wcol_kernel inverse product pass + transposing store, last pass along the rows)", lambda: fourstep(24, 16, capi.FOURSTEP_REFERENCE, inverse=True))
capi.tune(3, 0)
case("20 C4 reference contract forward with knob 4STEP_TRANSPOSED = 0 (transpose_kernel, column pass, row passes)", lambda: fourstep(24, 16, capi.FOURSTEP_REFERENCE))
capi.tune(3, 1)
case("21 one-tile ring, 32-bit 2^13 x 16384 inverse (whole transform in a tile)", lambda: merge(13, 16384, 32, inverse=True))
case("22 small-tile single-launch kernel, 64-bit 2^13 x 8 forward (1024-element tiles)", lambda: merge(13, 8, 64))
case("23 one polynomial of a three-pass ring, 64-bit 2^24 x 1 (contiguous pass on 2048-element single-polynomial tiles)", lambda: merge(24, 1, 64))
case("24 one polynomial of a three-pass ring, 32-bit 2^24 x 1 (contiguous pass on 4096-element single-polynomial tiles)", lambda: merge(24, 1, 32))
