"""Modulus-range coverage of the 64-bit paths: the engine picks an arithmetic policy from the size of p (exact,
lazy "fast", F60 -- gpu_ntt_b200/csrc/modarith.cuh), so every policy boundary gets a prime on each side, on the tuned
2^16 kernels and on the generic pass kernel, forward and inverse, against the oracle (the reference's NTTCPU accepts
any modulus below 2^62, modular_arith.cuh:66-67)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
from gpu_ntt_b200 import capi  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.gpu_util import to_dev, to_host  # noqa: E402

pytestmark = pytest.mark.gpu


def _is_prime(n: int) -> bool:
    if n < 2:
        return False
    for q in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % q == 0:
            return n == q
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def ntt_prime_below(limit: int, two_n: int) -> int:
    """largest prime p < limit with p = 1 (mod two_n)"""
    k = (limit - 2) // two_n
    while k > 0:
        p = k * two_n + 1
        if _is_prime(p):
            return p
        k -= 1
    raise ValueError("no prime")


def custom_params(logn: int, poly: int, p: int) -> O.MergeParams:
    """MergeParams for an arbitrary NTT-friendly prime (what NTTParameters(LOGN, NTTFactors, poly) builds)."""
    two_n = 2 << logn
    assert (p - 1) % two_n == 0
    psi = 0
    for g in range(2, 1000):
        c = pow(g, (p - 1) // two_n, p)
        if pow(c, two_n // 2, p) == p - 1:      # order exactly 2N
            psi = c
            break
    assert psi
    omega = psi * psi % p
    minus = poly == O.X_N_minus
    root = omega if minus else psi
    n = 1 << logn
    P = O.MergeParams(logn, poly, 64, p, omega, psi, pow(n, p - 2, p), root, pow(root, p - 2, p), n >> 1 if minus else n, n)
    P.fwd = np.empty(P.root_size, dtype=np.uint64)
    P.inv = np.empty(P.root_size, dtype=np.uint64)
    O.lib().ora_power_table(P.root, P.modulus, P.root_size, P.fwd)
    O.lib().ora_power_table(P.inv_root, P.modulus, P.root_size, P.inv)
    P.fwd_br = O.bitrev_table(P.fwd)
    P.inv_br = O.bitrev_table(P.inv)
    return P


LIMITS = [1 << 29, 1 << 33, (1 << 36) - 1, (1 << 36) + (1 << 30), (1 << 40) - 1, (1 << 40) + (1 << 33), 1 << 50,
          (1 << 60) - (1 << 31), (1 << 60) + (1 << 25), (1 << 60) + (1 << 58), (1 << 60) + (1 << 58) + (1 << 40), (1 << 62) - 1]


@pytest.mark.parametrize("limit", LIMITS)
@pytest.mark.parametrize("logn,poly", [(16, O.X_N_minus), (16, O.X_N_plus), (13, O.X_N_minus), (10, O.X_N_plus), (18, O.X_N_minus)])
def test_modulus_ranges(limit, logn, poly):
    p = ntt_prime_below(limit, 2 << logn)
    P = custom_params(logn, poly, p)
    batch = 3 if logn > 11 else 4     # (small rings take the one-launch kernel when batch * N fills whole 2048-element chunks)
    rng = np.random.default_rng(limit % 1000003 + logn)
    x = rng.integers(0, p, size=(batch, 1 << logn), dtype=np.uint64)
    x[0, :4] = (p - 1, 0, p - 1, 1)           # extremes
    x[1, :] = p - 1
    want = O.merge_ntt(x, P)
    d = to_dev(x, 64)
    tab, itab = to_dev(P.fwd_br, 64), to_dev(P.inv_br, 64)
    capi.lib().gpuntt_b200_set_profiling(1)      # launch kinds: 0 = the generic path's twiddle-prep kernel
    capi.profile_read()
    capi.ntt(d, tab, p, logn, poly)
    torch.cuda.synchronize()
    # every modulus the reference accepts (p < 2^62) stays on the tuned kernels: exact policy where the lazy ones do not apply
    assert 0 not in [k for k, _ in capi.profile_read()], f"p={p} fell back to the generic kernel"
    assert (to_host(d, 64).reshape(batch, -1) == want).all(), f"forward mismatch p={p}"
    capi.intt(d, itab, p, P.n_inv, logn, poly)
    torch.cuda.synchronize()
    kinds = [k for k, _ in capi.profile_read()]
    capi.lib().gpuntt_b200_set_profiling(0)
    assert 0 not in kinds, f"p={p} fell back to the generic kernel (inverse)"
    assert (to_host(d, 64).reshape(batch, -1) == x).all(), f"inverse mismatch p={p}"
    # the generic pass kernel with the same modulus
    capi.lib().gpuntt_b200_force_generic_path(1)
    try:
        d = to_dev(x, 64)
        capi.ntt(d, tab, p, logn, poly)
        torch.cuda.synchronize()
        assert (to_host(d, 64).reshape(batch, -1) == want).all(), f"generic forward mismatch p={p}"
        capi.intt(d, itab, p, P.n_inv, logn, poly)
        torch.cuda.synchronize()
        assert (to_host(d, 64).reshape(batch, -1) == x).all(), f"generic inverse mismatch p={p}"
    finally:
        capi.lib().gpuntt_b200_force_generic_path(0)


# 32-bit data: the forward kernels are lazy (ModL32) up to p = 2^29 and exact above, the inverse kernels exact throughout; the
# reference takes Data32 moduli up to 30 bits (modular_arith.cuh:66).  A prime on each side of 2^29, the top of the 30-bit range,
# and the small primes lattice schemes use (12289 = 3 * 2^12 + 1, 7681, 40961 where the ring allows them).
LIMITS32 = [1 << 13, 1 << 14, 1 << 17, 1 << 24, 1 << 29, (1 << 29) + (1 << 20), (1 << 29) + (1 << 28), (1 << 30) - 1]


@pytest.mark.parametrize("limit", LIMITS32)
@pytest.mark.parametrize("logn,poly,batch", [(8, O.X_N_plus, 48), (10, O.X_N_minus, 5), (11, O.X_N_plus, 8), (12, O.X_N_minus, 4),
                                             (13, O.X_N_plus, 3), (14, O.X_N_minus, 6), (16, O.X_N_plus, 2), (19, O.X_N_minus, 2),
                                             (20, O.X_N_plus, 1)])
def test_modulus_ranges_32bit(limit, logn, poly, batch):
    two_n = 2 << logn
    if limit <= two_n:
        pytest.skip("no prime p = 1 (mod 2N) below this limit")
    try:
        p = ntt_prime_below(limit, two_n)
    except ValueError:
        pytest.skip("no prime p = 1 (mod 2N) below this limit")
    P = custom_params(logn, poly, p)
    rng = np.random.default_rng(limit % 1000003 + logn)
    x = rng.integers(0, p, size=(batch, 1 << logn), dtype=np.uint64)
    x[0, :4] = (p - 1, 0, p - 1, 1)           # extremes
    x[-1, :] = p - 1
    want = O.merge_ntt(x, P)
    tab, itab = to_dev(P.fwd_br, 32), to_dev(P.inv_br, 32)
    for generic in (0, 1):
        capi.lib().gpuntt_b200_force_generic_path(generic)
        try:
            for inplace in (True, False):
                d = to_dev(x, 32)
                out = d if inplace else torch.zeros_like(d)
                capi.ntt(d, tab, p, logn, poly, out=out)
                torch.cuda.synchronize()
                assert (to_host(out, 32).reshape(batch, -1) == want).all(), f"forward mismatch p={p} generic={generic} inplace={inplace}"
                back = out if inplace else torch.zeros_like(d)
                capi.intt(out, itab, p, P.n_inv, logn, poly, out=back)
                torch.cuda.synchronize()
                assert (to_host(back, 32).reshape(batch, -1) == x).all(), f"inverse mismatch p={p} generic={generic} inplace={inplace}"
        finally:
            capi.lib().gpuntt_b200_force_generic_path(0)
