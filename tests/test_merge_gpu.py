"""GPU parity tests of the Merge-NTT path: every call goes through the C ABI
(include/gpuntt_b200.h via gpu_ntt_b200.capi) and is compared bit-exactly with the oracle on the
same seeded inputs, mirroring the reference's gpu_merge_ntt_examples / gpu_merge_intt_examples
(example/ntt_merge/test_merge_ntt.cu, test_merge_intt.cu)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from gpu_ntt_b200 import capi  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.gpu_util import to_dev, to_host, to_host_signed  # noqa: E402

pytestmark = pytest.mark.gpu


def run_fwd(x, P, bits, poly, inplace=True):
    n = P.logn
    d = to_dev(x, bits)
    tab = to_dev(P.fwd_br, bits)
    out = d if inplace else torch.zeros_like(d)
    capi.ntt(d.view(-1, P.n), tab, P.modulus, n, poly, out=out.view(-1, P.n))
    torch.cuda.synchronize()
    if not inplace:
        assert (to_host(d, bits) == x).all(), "out-of-place call modified its input"
    return to_host(out, bits)


def run_inv(y, P, bits, poly, inplace=True):
    d = to_dev(y, bits)
    tab = to_dev(P.inv_br, bits)
    out = d if inplace else torch.zeros_like(d)
    capi.intt(d.view(-1, P.n), tab, P.modulus, P.n_inv, P.logn, poly, out=out.view(-1, P.n))
    torch.cuda.synchronize()
    return to_host(out, bits)


@pytest.mark.parametrize("bits", [64, 32])
@pytest.mark.parametrize("poly", [O.X_N_minus, O.X_N_plus])
@pytest.mark.parametrize("logn,batch", [(1, 1), (1, 7), (2, 3), (3, 1000), (4, 1), (5, 33), (7, 5), (9, 9),
                                        (10, 3), (11, 2), (12, 5), (13, 3), (14, 2), (15, 1), (16, 3), (17, 1)])
def test_forward_and_inverse_match_oracle(bits, poly, logn, batch):
    P = O.merge_params(logn, poly, bits)
    x = O.example_input(P.modulus, batch << logn, seed=logn + batch)
    want = O.merge_ntt(x, P)
    got = run_fwd(x, P, bits, poly, inplace=True)
    assert (got == want).all()
    got2 = run_fwd(x, P, bits, poly, inplace=False)
    assert (got2 == want).all()
    back = run_inv(want, P, bits, poly, inplace=True)
    assert (back == x).all()
    back2 = run_inv(want, P, bits, poly, inplace=False)   # the reference's single-modulus out-of-place INTT is
    assert (back2 == x).all()                            # broken for logn >= 11 (ntt.cu:2367-2390); ours is not
    winv = O.merge_intt(x, P)
    assert (run_inv(x, P, bits, poly) == winv).all()


@pytest.mark.parametrize("bits,logn", [(64, 18), (64, 20), (32, 19), (64, 23), (32, 24)])
def test_large_rings(bits, logn):
    """2- and 3-pass plans (the reference switches kernels at logn 17 and 25, ntt.cuh:637-697)."""
    P = O.merge_params(logn, O.X_N_minus, bits)
    x = O.example_input(P.modulus, 1 << logn, seed=1)
    want = O.merge_ntt(x, P)
    assert (run_fwd(x, P, bits, O.X_N_minus) == want).all()
    assert (run_inv(want, P, bits, O.X_N_minus) == x).all()


@pytest.mark.parametrize("bits,logn", [(64, 25), (64, 26), (64, 27), (32, 25), (32, 26)])
def test_large_rings_above_2_24(bits, logn):
    """logN 25..27 (ForwardCore_/InverseCore_ of the reference, ntt.cu:763-1084, 1320-1552; plans ntt.cuh:669-697):
    forward against the oracle, word for word, and the inverse round trip.  (2^27 x 8 B = 1 GiB per buffer; the CPU oracle
    needs about a minute at 2^27, so one polynomial per size.)"""
    P = O.merge_params(logn, O.X_N_minus, bits)
    x = O.example_input(P.modulus, 1 << logn, seed=logn)
    want = O.merge_ntt(x, P)
    d = to_dev(x, bits)
    tab = to_dev(P.fwd_br, bits)
    capi.ntt(d.view(1, -1), tab, P.modulus, logn, O.X_N_minus)
    torch.cuda.synchronize()
    got = to_host(d, bits)
    assert (got == want).all(), f"{int((got != want).sum())} words differ"
    del tab
    itab = to_dev(P.inv_br, bits)
    capi.intt(d.view(1, -1), itab, P.modulus, P.n_inv, logn, O.X_N_minus)
    torch.cuda.synchronize()
    assert (to_host(d, bits) == x).all()


@pytest.mark.parametrize("bits,logn,batch,poly", [(64, 25, 2, O.X_N_plus), (64, 26, 1, O.X_N_plus), (32, 27, 1, O.X_N_minus)])
def test_large_rings_tuned_four_pass_plans(bits, logn, batch, poly):
    """Rings above 2^24 (64-bit) / 2^26 (32-bit) take the tuned kernels as three strided passes + the contiguous pass: launch
    count 4, every word against the oracle, inverse round trip (negacyclic rings and two polynomials per call as well; the
    X^N-1 single-polynomial cases are in test_large_rings_above_2_24)."""
    P = O.merge_params(logn, poly, bits)
    x = O.example_input(P.modulus, batch << logn, seed=logn)
    want = threaded_oracle(O.merge_ntt, x, P)
    d = to_dev(x, bits)
    tab = to_dev(P.fwd_br, bits)
    capi.ntt(d.view(batch, -1), tab, P.modulus, logn, poly)
    assert capi.lib().gpuntt_b200_last_launch_count() == 4
    torch.cuda.synchronize()
    got = to_host(d, bits)
    assert (got == want).all(), f"{int((got != want).sum())} words differ"
    del tab
    itab = to_dev(P.inv_br, bits)
    capi.intt(d.view(batch, -1), itab, P.modulus, P.n_inv, logn, poly)
    assert capi.lib().gpuntt_b200_last_launch_count() == 4
    torch.cuda.synchronize()
    assert (to_host(d, bits) == x).all()


def _sparse_closed_form(P, logn, nz, ks):
    """X_k = sum_j x_j w^(j k) for the sparse input nz = {j: x_j}; NTTCPU::ntt leaves X_k at the bit-reversed position of k"""
    p, w = P.modulus, P.root
    out = {}
    for k in ks:
        pos = int(format(k, f"0{logn}b")[::-1], 2)
        out[pos] = sum(v * pow(w, j * k, p) for j, v in nz.items()) % p
    return out


def test_ring_2_28_properties():
    """N = 2^28, the largest ring the reference accepts (2 GiB per polynomial; the CPU oracle would need minutes): a sparse
    input whose transform is known in closed form, sampled at 2000 outputs (the closed form itself is first checked against
    the oracle at N = 2^10), the tuned plan's launch count, and the inverse round trip of a dense random input."""
    bits = 64
    small = O.merge_params(10, O.X_N_minus, bits)
    nz = {0: 5, 1: 7, 77: 123456789, 600: small.modulus - 2}
    xs = np.zeros(1 << 10, dtype=np.uint64)
    for j, v in nz.items():
        xs[j] = v
    ys = O.merge_ntt(xs, small)
    for pos, v in _sparse_closed_form(small, 10, nz, range(0, 1024, 37)).items():
        assert int(ys[pos]) == v
    logn = 28
    n = 1 << logn
    P = O.merge_params(logn, O.X_N_minus, bits)
    p = P.modulus
    nz = {0: 3, 1: 11, 12345: p - 1, (n >> 1) + 7: 987654321, n - 1: 42}
    d = torch.zeros(n, dtype=torch.int64, device="cuda")
    for j, v in nz.items():
        d[j] = v if v < (1 << 63) else v - (1 << 64)
    tab = to_dev(P.fwd_br, bits)
    capi.ntt(d.view(1, -1), tab, p, logn, O.X_N_minus)
    assert capi.lib().gpuntt_b200_last_launch_count() == 4
    torch.cuda.synchronize()
    rng = np.random.default_rng(28)
    ks = [0, 1, n - 1, n >> 1] + [int(k) for k in rng.integers(0, n, 2000)]
    want = _sparse_closed_form(P, logn, nz, ks)
    idx = torch.tensor(sorted(want), dtype=torch.int64, device="cuda")
    got = d[idx].cpu().numpy().view(np.uint64)
    assert [int(v) for v in got] == [want[k] for k in sorted(want)]
    del tab, d
    x = torch.randint(0, p, (n,), dtype=torch.int64, device="cuda")
    y = x.clone()
    tab = to_dev(P.fwd_br, bits)
    capi.ntt(y.view(1, -1), tab, p, logn, O.X_N_minus)
    del tab
    itab = to_dev(P.inv_br, bits)
    capi.intt(y.view(1, -1), itab, p, P.n_inv, logn, O.X_N_minus)
    assert capi.lib().gpuntt_b200_last_launch_count() == 4
    torch.cuda.synchronize()
    assert bool((x == y).all())


@pytest.mark.parametrize("logn,batch,poly", [(18, 3, O.X_N_plus), (19, 2, O.X_N_minus), (21, 2, O.X_N_plus),
                                             (22, 1, O.X_N_minus), (24, 1, O.X_N_plus), (24, 2, O.X_N_minus)])
def test_large_rings_tuned_three_pass_plans(logn, batch, poly):
    """64-bit rings of 2^17..2^24 take the tuned kernels as two strided passes + the contiguous pass (ranges of
    twiddles per block of matrix rows); several polynomials per call, both ring types, in place and out of place."""
    P = O.merge_params(logn, poly, 64)
    x = O.example_input(P.modulus, batch << logn, seed=logn)
    want = O.merge_ntt(x, P)
    assert (run_fwd(x, P, 64, poly) == want).all()
    assert (run_fwd(x, P, 64, poly, inplace=False) == want).all()
    assert (run_inv(want, P, 64, poly) == x).all()
    assert (run_inv(want, P, 64, poly, inplace=False) == x).all()


def test_golden_vectors_through_c_abi(golden):
    """Outputs on the example drivers' seed-0 input equal the hashes recorded from the reference."""
    for g in golden["merge"]:
        if g["logn"] > 16:
            continue
        bits, poly, logn = g["width"], g["poly"], g["logn"]
        P = O.merge_params(logn, poly, bits)
        x = O.example_input(P.modulus, g["batch"] << logn)
        y = run_fwd(x, P, bits, poly)
        assert str(O.fold_hash(y)) == g["ntt_hash"], g
        z = run_inv(x, P, bits, poly)
        assert str(O.fold_hash(z)) == g["intt_hash"], g


@pytest.mark.parametrize("bits", [64, 32])
@pytest.mark.parametrize("logn,batch", [(10, 8), (13, 5), (14, 64), (16, 3), (17, 2)])
def test_signed_variants_take_the_tuned_kernels(bits, logn, batch):
    """Data32s/Data64s on the tuned kernels: the first forward round fixes negative inputs up as it loads, the last
    inverse round centres its outputs (ntt.cu:481-489, 1178-1186 of the reference) -- no generic-kernel launches
    (those would be a twiddle-prep kernel plus the passes)."""
    P = O.merge_params(logn, O.X_N_plus, bits)
    p = P.modulus
    rng = np.random.default_rng(logn * 7 + bits)
    mag = rng.integers(0, p // 2, size=batch << logn, dtype=np.int64)
    sx = np.where(rng.integers(0, 2, size=mag.size) == 1, -mag, mag).astype(np.int64)
    sx[:3] = (-(p // 2) + 1, -1, p // 2 - 1)
    want = O.merge_ntt(O.reduce_signed(sx, p), P)
    d = torch.from_numpy(sx if bits == 64 else sx.astype(np.int32)).cuda()
    out = torch.zeros_like(d)
    tab, itab = to_dev(P.fwd_br, bits), to_dev(P.inv_br, bits)
    st = torch.cuda.current_stream().cuda_stream
    capi.lib().gpuntt_b200_set_profiling(1)      # launch kinds: 0 = the generic path's twiddle-prep kernel
    capi.profile_read()
    capi.merge_ntt(in_ptr=d.data_ptr(), out_ptr=out.data_ptr(), table_ptr=tab.data_ptr(), n_power=logn, batch=batch,
                   element_bits=bits, direction=capi.FORWARD, reduction_poly=O.X_N_plus, modulus=p, is_signed=True, stream=st)
    torch.cuda.synchronize()
    assert 0 not in [k for k, _ in capi.profile_read()], "signed forward fell back to the generic kernel"
    assert (to_host(out, bits) == want).all()
    back = torch.zeros_like(d)
    capi.merge_ntt(in_ptr=out.data_ptr(), out_ptr=back.data_ptr(), table_ptr=itab.data_ptr(), n_power=logn, batch=batch,
                   element_bits=bits, direction=capi.INVERSE, reduction_poly=O.X_N_plus, modulus=p, mod_inverse=P.n_inv,
                   is_signed=True, stream=st)
    torch.cuda.synchronize()
    kinds = [k for k, _ in capi.profile_read()]
    capi.lib().gpuntt_b200_set_profiling(0)
    assert 0 not in kinds, "signed inverse fell back to the generic kernel"
    assert (to_host_signed(back) == sx).all()


@pytest.mark.parametrize("bits", [64, 32])
@pytest.mark.parametrize("logn", [4, 12, 16])
def test_signed_variants(bits, logn):
    """Data32s/Data64s: signed input on forward (test_merge_ntt.cu:184-341), centred output on inverse."""
    P = O.merge_params(logn, O.X_N_plus, bits)
    p = P.modulus
    rng = np.random.RandomState(logn)
    mag = rng.randint(0, 2**31, size=3 << logn).astype(np.int64) * (1 if bits == 32 else 2**20) % (p // 2)
    sx = np.where(rng.randint(0, 2, size=mag.size) == 1, -mag, mag).astype(np.int64)
    want = O.merge_ntt(O.reduce_signed(sx, p), P)
    d = torch.from_numpy(sx if bits == 64 else sx.astype(np.int32)).cuda()
    out = torch.zeros_like(d)
    tab = to_dev(P.fwd_br, bits)
    capi.merge_ntt(in_ptr=d.data_ptr(), out_ptr=out.data_ptr(), table_ptr=tab.data_ptr(), n_power=logn, batch=3,
                   element_bits=bits, direction=capi.FORWARD, reduction_poly=O.X_N_plus, modulus=p, is_signed=True,
                   stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert (to_host(out, bits) == want).all()
    # inverse with centred signed output
    itab = to_dev(P.inv_br, bits)
    back = torch.zeros_like(d)
    capi.merge_ntt(in_ptr=out.data_ptr(), out_ptr=back.data_ptr(), table_ptr=itab.data_ptr(), n_power=logn, batch=3,
                   element_bits=bits, direction=capi.INVERSE, reduction_poly=O.X_N_plus, modulus=p,
                   mod_inverse=P.n_inv, is_signed=True, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert (to_host_signed(back) == O.centered(O.merge_intt(want, P), p)).all()
    assert (to_host_signed(back) == sx).all()


def _is_prime(n):
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


def rns_primes(bits, logn, count, top_log2=None):
    m = 1 << (logn + 1)
    top = (1 << (top_log2 or (61 if bits == 64 else 29))) // m
    out = []
    k = top
    while len(out) < count:
        p = k * m + 1
        if _is_prime(p):
            for g in range(2, 200):
                psi = pow(g, (p - 1) // m, p)
                if pow(psi, m // 2, p) == p - 1:
                    out.append((p, psi))
                    break
        k -= 1
    return out


@pytest.mark.parametrize("bits", [64, 32])
@pytest.mark.parametrize("logn,batch,mod_count", [(3, 10, 3), (10, 7, 2), (12, 6, 3), (14, 5, 2), (16, 4, 4)])
def test_rns_form(bits, logn, batch, mod_count):
    """RNS overloads: polynomial b uses modulus[b % mod_count], table slice (b % mod_count) << n_power,
    mod_inverse[b % mod_count] (ntt.cu:613-619, 672-673, 1225-1226).  61-/29-bit primes exercise the top
    of the supported modulus range."""
    _rns_roundtrip(bits, logn, batch, mod_count, rns_primes(bits, logn, mod_count))


@pytest.mark.parametrize("bits,logn,batch,mod_count,tops", [
    (64, 12, 6, 3, (59, 59, 59)),        # every modulus allows the lazy policies -> F60 / lazy inverse kernels
    (64, 16, 8, 4, (59, 58, 50, 45)),
    (64, 14, 4, 2, (59, 61)),            # one modulus outside -> the exact-policy kernels do the passes
    (64, 16, 6, 2, (61, 61)),
    (64, 13, 9, 3, (59, 59, 59)),        # odd number of polynomials per slot (ragged last tile)
    (32, 14, 6, 3, (29, 29, 29)),
    (32, 16, 4, 2, (29, 25)),
    (64, 17, 4, 2, (59, 58)),            # three-pass plans: two strided passes with several twiddle ranges per slot
    (64, 18, 3, 3, (59, 61, 50)),
    (32, 19, 4, 2, (29, 28)),
])
@pytest.mark.parametrize("fused", [2, 0])
def test_rns_form_tuned_kernels(bits, logn, batch, mod_count, tops, fused):
    """RNS overloads on the tuned kernels (two-pass ring sizes, batch a multiple of mod_count): per-slot modulus,
    table slice and N^-1 are read per segment on the device, and for 64-bit data every kernel holds the lazy and
    the exact arithmetic body and picks one from the moduli it finds.  fused = 2: the single-launch kernel (the CTAs
    are partitioned among the modulus slots, each picks its policy from its own modulus); 0: one launch per pass."""
    primes = [rns_primes(bits, logn, 1 + i, t)[i] for i, t in enumerate(tops)]
    assert len({p for p, _ in primes}) == mod_count
    try:
        capi.tune(capi.TUNE_FUSED_PASSES, fused)
        _rns_roundtrip(bits, logn, batch, mod_count, primes)
    finally:
        capi.tune(capi.TUNE_FUSED_PASSES, 1)
    # the inverse call was the last one: one launch per pass (64-bit: the dual kernel picks the lazy or the exact body
    # from the modulus array on the device)
    assert capi.lib().gpuntt_b200_last_launch_count() == ((1 if fused else 2) if logn <= (16 if bits == 64 else 18) else 3)


@pytest.mark.parametrize("bits,logn,batch,mod_count,tops", [
    (64, 7, 6, 3, (59, 59, 59)),         # 32 polynomials of a slot per tile, two of them present
    (64, 8, 200, 4, (59, 58, 50, 45)),   # several tiles per slot, ragged last one
    (64, 9, 10, 2, (59, 61)),            # one modulus outside the lazy range: the exact body
    (64, 10, 9, 3, (61, 61, 61)),
    (64, 11, 12, 4, (59, 59, 59, 59)),
    (64, 11, 2, 2, (59, 60)),            # one polynomial per slot: half-empty tiles
    (32, 8, 96, 3, (29, 29, 29)),
    (32, 10, 14, 2, (29, 25)),
    (32, 12, 6, 3, (29, 28, 27)),
])
def test_rns_small_rings_tuned_kernels(bits, logn, batch, mod_count, tops):
    """RNS overloads on rings of 2^7..2^11 (64-bit) / 2^8..2^12 (32-bit): ONE launch whose tiles hold whole polynomials of one
    modulus slot (4-D tensor map {row, rows, slot, polynomial within the slot}); forward and inverse against the oracle."""
    primes = [rns_primes(bits, logn, 1 + i, t)[i] for i, t in enumerate(tops)]
    assert len({p for p, _ in primes}) == mod_count
    _rns_roundtrip(bits, logn, batch, mod_count, primes)
    assert capi.lib().gpuntt_b200_last_launch_count() == 1


@pytest.mark.parametrize("bits,logn,batch,mod_count,tops", [(64, 12, 1202, 3, (59, 61, 50)), (32, 14, 301, 2, (29, 27)),
                                                            (64, 10, 4099, 4, (59, 59, 58, 45))])
def test_rns_batch_not_a_multiple_of_mod_count_is_split(bits, logn, batch, mod_count, tops):
    """From 2^22 elements up, an RNS batch that is not a multiple of mod_count runs its whole rounds of slots on the tuned kernels
    and only the last (fewer than mod_count) polynomials on the generic kernel; below, the generic kernel takes the whole call."""
    primes = [rns_primes(bits, logn, 1 + i, t)[i] for i, t in enumerate(tops)]
    _rns_roundtrip(bits, logn, batch, mod_count, primes)
    split_launches = capi.lib().gpuntt_b200_last_launch_count()
    _rns_roundtrip(bits, logn, batch % mod_count, mod_count, primes)
    tail_launches = capi.lib().gpuntt_b200_last_launch_count()
    _rns_roundtrip(bits, logn, batch - batch % mod_count, mod_count, primes)
    assert split_launches == capi.lib().gpuntt_b200_last_launch_count() + tail_launches


def _rns_roundtrip(bits, logn, batch, mod_count, primes):
    n = 1 << logn
    fwd_tab = np.zeros(mod_count << logn, dtype=np.uint64)
    inv_tab = np.zeros(mod_count << logn, dtype=np.uint64)
    mods = np.zeros((mod_count, 3), dtype=np.uint64)
    ninvs = np.zeros(mod_count, dtype=np.uint64)
    params = []
    for m, (p, psi) in enumerate(primes):
        fwd = np.array([pow(psi, i, p) for i in range(n)], dtype=np.uint64)
        ipsi = pow(psi, p - 2, p)
        inv = np.array([pow(ipsi, i, p) for i in range(n)], dtype=np.uint64)
        fwd_tab[m << logn:(m + 1) << logn] = O.bitrev_table(fwd)
        inv_tab[m << logn:(m + 1) << logn] = O.bitrev_table(inv)
        bit, mu = O.modulus(p, bits)
        mods[m] = (p, bit, mu)
        ninvs[m] = pow(n, p - 2, p)
        P = O.MergeParams(logn, O.X_N_plus, bits, p, 0, psi, int(ninvs[m]), psi, ipsi, n, n)
        P.fwd, P.inv = fwd, inv
        params.append(P)
    rng = np.random.RandomState(5)
    x = np.zeros((batch, n), dtype=np.uint64)
    for b in range(batch):
        p = primes[b % mod_count][0]
        x[b] = (rng.randint(0, 2**31, size=n).astype(np.uint64) * np.uint64(2**31) +
                rng.randint(0, 2**31, size=n).astype(np.uint64)) % np.uint64(p)
    want = np.stack([O.merge_ntt(x[b], params[b % mod_count]) for b in range(batch)])
    d = to_dev(x, bits)
    s = torch.cuda.current_stream().cuda_stream
    mods_d = to_dev(mods.ravel(), bits)
    ninv_d = to_dev(ninvs, bits)
    capi.merge_ntt(in_ptr=d.data_ptr(), out_ptr=d.data_ptr(), table_ptr=to_dev(fwd_tab, bits).data_ptr(),
                   n_power=logn, batch=batch, element_bits=bits, direction=capi.FORWARD,
                   reduction_poly=O.X_N_plus, mod_count=mod_count, modulus_dev=mods_d.data_ptr(), stream=s)
    torch.cuda.synchronize()
    assert (to_host(d, bits) == want).all()
    capi.merge_ntt(in_ptr=d.data_ptr(), out_ptr=d.data_ptr(), table_ptr=to_dev(inv_tab, bits).data_ptr(),
                   n_power=logn, batch=batch, element_bits=bits, direction=capi.INVERSE,
                   reduction_poly=O.X_N_plus, mod_count=mod_count, modulus_dev=mods_d.data_ptr(),
                   mod_inverse_dev=ninv_d.data_ptr(), stream=s)
    torch.cuda.synchronize()
    assert (to_host(d, bits) == x).all()


def threaded_oracle(fn, x, P, threads=None):
    """fn(row, P) for every polynomial of x, sharded over host threads (ctypes releases the GIL inside the C oracle)."""
    import concurrent.futures as cf
    import os
    threads = threads or min(32, os.cpu_count() or 1)
    rows = x.reshape(-1, P.n)
    out = np.empty_like(rows)

    def work(idx):
        for r in idx:
            out[r] = fn(rows[r], P)
    with cf.ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, [list(range(i, rows.shape[0], threads)) for i in range(threads)]))
    return out.reshape(x.shape)


@pytest.mark.parametrize("fused", [1, 2, 0])
def test_config_c2_full_size_every_polynomial(fused):
    """BASELINE config 2 exactly as SURVEY 8(d) prescribes it: NTTParameters<Data64>(16, X_N_minus), the example drivers'
    seed-0 mt19937 stream, batch 1024, in place -- and EVERY one of the 1024 x 65536 output words against the oracle
    (NTTCPU::ntt restated; 16 host threads need well under a second for it).  Run on the default kernel choice, on the
    single-launch kernel and on one launch per pass.  Then the inverse restores every input word."""
    logn, batch, bits = 16, 1024, 64
    P = O.merge_params(logn, O.X_N_minus, bits)
    p = P.modulus
    x = O.example_input(p, batch << logn, seed=0).reshape(batch, P.n)
    want = threaded_oracle(O.merge_ntt, x, P)
    tab, itab = to_dev(P.fwd_br, bits), to_dev(P.inv_br, bits)
    d = to_dev(x, bits)
    try:
        capi.tune(capi.TUNE_FUSED_PASSES, fused)
        capi.ntt(d, tab, p, logn, O.X_N_minus)
        torch.cuda.synchronize()
        assert capi.lib().gpuntt_b200_last_launch_count() == (1 if fused == 2 else 2)
        got = to_host(d, bits)
        assert (got == want).all(), f"{int((got != want).sum())} words differ"
        if fused == 1:
            assert O.fold_hash(got[0]) == 2501385232060115022    # SURVEY 8(c) KAT of the first polynomial
        capi.intt(d, itab, p, P.n_inv, logn, O.X_N_minus)
        torch.cuda.synchronize()
        assert (to_host(d, bits) == x).all()
    finally:
        capi.tune(capi.TUNE_FUSED_PASSES, 1)


@pytest.mark.parametrize("fused", [1, 0])
def test_config_c3_full_size_every_polynomial(fused):
    """BASELINE config 3 as SURVEY 8(d) prescribes it: NTTParameters<Data32>(14, X_N_minus) (p = 469762049), seed-0 stream,
    batch 4096: forward == NTTCPU::ntt for every polynomial, inverse(forward(x)) == x, bit exact; the forward transform
    is ONE launch on the default path."""
    logn, batch, bits = 14, 4096, 32
    P = O.merge_params(logn, O.X_N_minus, bits)
    p = P.modulus
    assert p == 469762049
    x = O.example_input(p, batch << logn, seed=0).reshape(batch, P.n)
    want = threaded_oracle(O.merge_ntt, x, P)
    tab, itab = to_dev(P.fwd_br, bits), to_dev(P.inv_br, bits)
    d = to_dev(x, bits)
    try:
        capi.tune(capi.TUNE_FUSED_PASSES, fused)
        capi.ntt(d, tab, p, logn, O.X_N_minus)
        torch.cuda.synchronize()
        assert capi.lib().gpuntt_b200_last_launch_count() == (1 if fused else 2)
        got = to_host(d, bits)
        assert (got == want).all(), f"{int((got != want).sum())} words differ"
        capi.intt(d, itab, p, P.n_inv, logn, O.X_N_minus)
        torch.cuda.synchronize()
        assert capi.lib().gpuntt_b200_last_launch_count() == (1 if fused else 2)
        assert (to_host(d, bits) == x).all()
    finally:
        capi.tune(capi.TUNE_FUSED_PASSES, 1)


def test_config_c2_linearity_and_range():
    """Size-independent properties at the C2 size on device-generated data: NTT(a+b) == NTT(a)+NTT(b), canonical outputs."""
    logn, batch, bits = 16, 1024, 64
    P = O.merge_params(logn, O.X_N_minus, bits)
    p = P.modulus
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randint(0, p, (batch, P.n), dtype=torch.int64, device="cuda", generator=g)
    x0 = x.clone()
    tab = to_dev(P.fwd_br, bits)
    capi.ntt(x, tab, p, logn, O.X_N_minus)
    s = x0[:512] + x0[512:]
    s = torch.where(s >= p, s - p, s)
    capi.ntt(s, tab, p, logn, O.X_N_minus)
    t = x[:512] + x[512:]
    t = torch.where(t >= p, t - p, t)
    assert torch.equal(s, t)
    assert int(x.max()) < p and int(x.min()) >= 0


def test_host_buffer_entry_point():
    """gpuntt_b200_merge_ntt_host: host in/out, copies inside (the e2e path of bench.py)."""
    import ctypes as C
    logn, batch = 12, 8
    P = O.merge_params(logn, O.X_N_minus, 64)
    x = O.example_input(P.modulus, batch << logn, seed=9)
    out = np.zeros_like(x)
    d = capi.MergeDesc(64, 0, capi.FORWARD, logn, capi.PerPolynomial, O.X_N_minus, batch, 0,
                       x.ctypes.data, out.ctypes.data, None, P.modulus, 0, None, None, None)
    capi.check(capi.lib().gpuntt_b200_merge_ntt_host(C.byref(d), P.fwd_br.ctypes.data, P.fwd_br.size))
    assert (out == O.merge_ntt(x, P)).all()
    assert capi.lib().gpuntt_b200_last_launch_count() >= 1


def test_host_buffer_entry_point_chunked_pipeline():
    """Same entry point at a size that is cut into several chunks (64 polynomials of 2^16 each) with a ragged
    last chunk: H2D / transform / D2H of different chunks overlap on the engine's three streams."""
    import ctypes as C
    logn, batch = 16, 150
    P = O.merge_params(logn, O.X_N_minus, 64)
    x = O.example_input(P.modulus, batch << logn, seed=21)
    out = np.zeros_like(x)
    d = capi.MergeDesc(64, 0, capi.FORWARD, logn, capi.PerPolynomial, O.X_N_minus, batch, 0,
                       x.ctypes.data, out.ctypes.data, None, P.modulus, 0, None, None, None)
    capi.check(capi.lib().gpuntt_b200_merge_ntt_host(C.byref(d), P.fwd_br.ctypes.data, P.fwd_br.size))
    assert (out == O.merge_ntt(x, P)).all()
    assert capi.lib().gpuntt_b200_last_launch_count() in (3, 6)  # 3 chunks x (one fused launch | two passes)


def test_streams_and_no_sync():
    """Calls only enqueue on cfg.stream: two streams, two different transforms, results both right."""
    P = O.merge_params(13, O.X_N_minus, 64)
    xs = [O.example_input(P.modulus, 4 << 13, seed=s) for s in (1, 2)]
    tab = to_dev(P.fwd_br, 64)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    ds = [to_dev(x, 64) for x in xs]
    torch.cuda.synchronize()
    for st, d in zip(streams, ds):
        capi.ntt(d.view(-1, P.n), tab, P.modulus, 13, O.X_N_minus, stream=st)
    torch.cuda.synchronize()
    for x, d in zip(xs, ds):
        assert (to_host(d, 64) == O.merge_ntt(x, P)).all()


@pytest.mark.parametrize("poly", [O.X_N_minus, O.X_N_plus])
@pytest.mark.parametrize("batch", [1, 2, 5, 64, 301])
def test_fast_path_and_generic_path_agree_with_oracle(poly, batch):
    """logn=16 Data64 takes the persistent TMA kernels (merge_fast.cu); the same call with the fast path
    disabled takes the generic pass kernel.  Both must equal the oracle bit for bit (odd batches leave a
    half-empty last tile; 301 polynomials make every CTA cross a range boundary)."""
    logn, bits = 16, 64
    P = O.merge_params(logn, poly, bits)
    x = O.example_input(P.modulus, batch << logn, seed=batch)
    idx = sorted(set([0, batch - 1, batch // 2]))
    want = {b: O.merge_ntt(x.reshape(batch, -1)[b], P) for b in idx}
    for force in (0, 1):
        capi.lib().gpuntt_b200_force_generic_path(force)
        try:
            y = run_fwd(x, P, bits, poly, inplace=(force == 0)).reshape(batch, -1)
            for b in idx:
                assert (y[b] == want[b]).all(), (force, b)
            back = run_inv(y.ravel(), P, bits, poly, inplace=(force == 1))
            assert (back == x).all(), force
        finally:
            capi.lib().gpuntt_b200_force_generic_path(0)


def test_fast_path_large_modulus_uses_exact_arithmetic():
    """A 61-bit prime is above the fast-arithmetic limit: the persistent kernels run with exact quotients."""
    logn = 16
    (p, psi), = rns_primes(64, logn, 1)
    assert p.bit_length() == 61
    n = 1 << logn
    omega = psi * psi % p
    fwd = np.array([pow(omega, i, p) for i in range(n // 2)], dtype=np.uint64)
    inv = np.array([pow(pow(omega, p - 2, p), i, p) for i in range(n // 2)], dtype=np.uint64)
    P = O.MergeParams(logn, O.X_N_minus, 64, p, omega, psi, pow(n, p - 2, p), omega, pow(omega, p - 2, p), n // 2, n)
    P.fwd, P.inv, P.fwd_br, P.inv_br = fwd, inv, O.bitrev_table(fwd), O.bitrev_table(inv)
    rng = np.random.RandomState(3)
    x = (rng.randint(0, 2**31, size=2 * n).astype(np.uint64) * np.uint64(2**31) +
         rng.randint(0, 2**31, size=2 * n).astype(np.uint64)) % np.uint64(p)
    y = run_fwd(x, P, 64, O.X_N_minus)
    assert (y == O.merge_ntt(x, P)).all()
    assert (run_inv(y, P, 64, O.X_N_minus) == x).all()


@pytest.mark.parametrize("bits", [64, 32])
@pytest.mark.parametrize("logn,batch,mod_count", [(10, 6, 2), (12, 8, 4), (16, 5, 3), (16, 6, 3)])
def test_rns_ordered_entry_points(bits, logn, batch, mod_count):
    """GPU_NTT_Modulus_Ordered: polynomial b uses modulus / table slice / n^-1 number order[b % mod_count]
    (ntt.cu:3117-3118); GPU_NTT_Poly_Ordered: the b-th transform runs in polynomial slot order[b] with modulus
    b % mod_count (ntt.cu:3796-3798)."""
    n = 1 << logn
    total_primes = mod_count + 2
    primes = rns_primes(bits, logn, total_primes)
    fwd_tab = np.zeros(total_primes << logn, dtype=np.uint64)
    inv_tab = np.zeros(total_primes << logn, dtype=np.uint64)
    mods = np.zeros((total_primes, 3), dtype=np.uint64)
    ninvs = np.zeros(total_primes, dtype=np.uint64)
    params = []
    for m, (p, psi) in enumerate(primes):
        fwd = np.array([pow(psi, i, p) for i in range(n)], dtype=np.uint64)
        ipsi = pow(psi, p - 2, p)
        inv = np.array([pow(ipsi, i, p) for i in range(n)], dtype=np.uint64)
        fwd_tab[m << logn:(m + 1) << logn] = O.bitrev_table(fwd)
        inv_tab[m << logn:(m + 1) << logn] = O.bitrev_table(inv)
        bit, mu = O.modulus(p, bits)
        mods[m] = (p, bit, mu)
        ninvs[m] = pow(n, p - 2, p)
        P = O.MergeParams(logn, O.X_N_plus, bits, p, 0, psi, int(ninvs[m]), psi, ipsi, n, n)
        P.fwd, P.inv = fwd, inv
        params.append(P)
    rng = np.random.RandomState(11)
    s = torch.cuda.current_stream().cuda_stream
    mods_d, ninv_d = to_dev(mods.ravel(), bits), to_dev(ninvs, bits)
    ftab, itab = to_dev(fwd_tab, bits), to_dev(inv_tab, bits)

    def rand_poly(p):
        return (rng.randint(0, 2**31, size=n).astype(np.uint64) * np.uint64(2**31) +
                rng.randint(0, 2**31, size=n).astype(np.uint64)) % np.uint64(p)

    # --- modulus ordered
    order = rng.permutation(total_primes)[:mod_count].astype(np.int32)
    x = np.stack([rand_poly(primes[order[b % mod_count]][0]) for b in range(batch)])
    want = np.stack([O.merge_ntt(x[b], params[order[b % mod_count]]) for b in range(batch)])
    order_d = torch.from_numpy(order).cuda()
    d = to_dev(x, bits)
    out = torch.zeros_like(d)
    capi.merge_ntt(in_ptr=d.data_ptr(), out_ptr=out.data_ptr(), table_ptr=ftab.data_ptr(), n_power=logn, batch=batch,
                   element_bits=bits, direction=capi.FORWARD, reduction_poly=O.X_N_plus, mod_count=mod_count,
                   modulus_dev=mods_d.data_ptr(), stream=s, modulus_order_dev=order_d.data_ptr())
    torch.cuda.synchronize()
    assert (to_host(out, bits) == want).all()
    capi.merge_ntt(in_ptr=out.data_ptr(), out_ptr=out.data_ptr(), table_ptr=itab.data_ptr(), n_power=logn, batch=batch,
                   element_bits=bits, direction=capi.INVERSE, reduction_poly=O.X_N_plus, mod_count=mod_count,
                   modulus_dev=mods_d.data_ptr(), mod_inverse_dev=ninv_d.data_ptr(), stream=s,
                   modulus_order_dev=order_d.data_ptr())
    torch.cuda.synchronize()
    assert (to_host(out, bits) == x).all()

    # --- poly ordered: `slots` polynomial slots, `batch` of them transformed in the order given
    slots = batch + 3
    porder = rng.permutation(slots)[:batch].astype(np.int32)
    buf = np.stack([rand_poly(primes[0][0] if True else 0) for _ in range(slots)])
    for b in range(batch):
        buf[porder[b]] = rand_poly(primes[b % mod_count][0])
    want = buf.copy()
    for b in range(batch):
        want[porder[b]] = O.merge_ntt(buf[porder[b]], params[b % mod_count])
    porder_d = torch.from_numpy(porder).cuda()
    d = to_dev(buf, bits)
    capi.merge_ntt(in_ptr=d.data_ptr(), out_ptr=d.data_ptr(), table_ptr=ftab.data_ptr(), n_power=logn, batch=batch,
                   element_bits=bits, direction=capi.FORWARD, reduction_poly=O.X_N_plus, mod_count=mod_count,
                   modulus_dev=mods_d.data_ptr(), stream=s, poly_order_dev=porder_d.data_ptr())
    torch.cuda.synchronize()
    assert (to_host(d, bits).reshape(slots, n) == want).all()
    capi.merge_ntt(in_ptr=d.data_ptr(), out_ptr=d.data_ptr(), table_ptr=itab.data_ptr(), n_power=logn, batch=batch,
                   element_bits=bits, direction=capi.INVERSE, reduction_poly=O.X_N_plus, mod_count=mod_count,
                   modulus_dev=mods_d.data_ptr(), mod_inverse_dev=ninv_d.data_ptr(), stream=s,
                   poly_order_dev=porder_d.data_ptr())
    torch.cuda.synchronize()
    assert (to_host(d, bits).reshape(slots, n) == buf).all()
    if batch % mod_count == 0 and logn >= (12 if bits == 64 else 14):
        # two-pass ring, whole slots: the tuned RNS kernels (policy kernel + lazy/exact kernel per pass for 64-bit)
        assert capi.lib().gpuntt_b200_last_launch_count() == 2


@pytest.mark.parametrize("bits", [64, 32])
@pytest.mark.parametrize("poly", [O.X_N_plus, O.X_N_minus])
@pytest.mark.parametrize("logh,w", [(9, 1024), (7, 64), (5, 2), (3, 8), (1, 4), (9, 1), (8, 16)])
def test_per_coefficient_layout(bits, poly, logh, w):
    """NTTLayout::PerCoefficient: the buffer is an H x W row-major matrix and every column is a transform of
    length H = 2^n_power, batch_size = W (example/ntt_merge/test_merge_ntt.cu:343-474: compared with PerPolynomial
    on the transposed matrix; H = 512 x W = 1024 there, H = 128 in test_merge_intt.cu)."""
    h = 1 << logh
    P = O.merge_params(logh, poly, bits)
    x = O.example_input(P.modulus, h * w, seed=logh * 7 + w).reshape(h, w)
    want = O.merge_ntt(np.ascontiguousarray(x.T), P).reshape(w, h).T      # column j of the result = NTT(column j)
    s = torch.cuda.current_stream().cuda_stream
    d = to_dev(x, bits)
    tab, itab = to_dev(P.fwd_br, bits), to_dev(P.inv_br, bits)
    capi.merge_ntt(in_ptr=d.data_ptr(), out_ptr=d.data_ptr(), table_ptr=tab.data_ptr(), n_power=logh, batch=w,
                   element_bits=bits, direction=capi.FORWARD, reduction_poly=poly, layout=capi.PerCoefficient,
                   modulus=P.modulus, stream=s)
    torch.cuda.synchronize()
    assert (to_host(d, bits).reshape(h, w) == want).all()
    out = torch.zeros_like(d)
    capi.merge_ntt(in_ptr=d.data_ptr(), out_ptr=out.data_ptr(), table_ptr=itab.data_ptr(), n_power=logh, batch=w,
                   element_bits=bits, direction=capi.INVERSE, reduction_poly=poly, layout=capi.PerCoefficient,
                   modulus=P.modulus, mod_inverse=P.n_inv, stream=s)
    torch.cuda.synchronize()
    assert (to_host(out, bits).reshape(h, w) == x).all()


@pytest.mark.parametrize("poly", [O.X_N_plus, O.X_N_minus])
@pytest.mark.parametrize("bits,logh,w,launches", [(64, 4, 256, 1), (64, 5, 128, 1), (64, 6, 64, 1), (64, 7, 4096, 1), (64, 8, 16, 1),
                                                  (64, 8, 2048, 1), (64, 9, 256, 2), (64, 9, 4096, 2),
                                                  (32, 3, 1024, 1), (32, 4, 512, 1), (32, 5, 256, 1), (32, 6, 128, 1), (32, 7, 4096, 1),
                                                  (32, 8, 32, 1), (32, 8, 2048, 1), (32, 9, 512, 2), (32, 9, 4096, 2),
                                                  (32, -7, 4096, 1), (32, -9, 1024, 2)])
def test_per_coefficient_on_tuned_kernels(poly, bits, logh, w, launches):
    """PerCoefficient calls whose batch is at least one tile wide run as strided passes of the tuned kernels (one pass up to H = 256,
    two for H = 512): launch count asserted, every word against the oracle, in place and out of place, and the generic kernel gives
    the same words.  (32-bit, negative logh: a 30-bit prime, i.e. the exact-policy forward kernels instead of the lazy ones.)"""
    if logh < 0:
        from tests.test_moduli_gpu import custom_params, ntt_prime_below
        logh = -logh
        P = custom_params(logh, poly, ntt_prime_below((1 << 30) - 1, 2 << logh))
    else:
        P = O.merge_params(logh, poly, bits)
    h = 1 << logh
    x = O.example_input(P.modulus, h * w, seed=logh * 3 + w).reshape(h, w)
    want = O.merge_ntt(np.ascontiguousarray(x.T), P).reshape(w, h).T
    s = torch.cuda.current_stream().cuda_stream
    tab, itab = to_dev(P.fwd_br, bits), to_dev(P.inv_br, bits)
    for generic in (0, 1):
        capi.lib().gpuntt_b200_force_generic_path(generic)
        try:
            d = to_dev(x, bits)
            out = torch.zeros_like(d)
            capi.merge_ntt(in_ptr=d.data_ptr(), out_ptr=out.data_ptr(), table_ptr=tab.data_ptr(), n_power=logh, batch=w, element_bits=bits,
                           direction=capi.FORWARD, reduction_poly=poly, layout=capi.PerCoefficient, modulus=P.modulus, stream=s)
            if not generic:
                assert capi.lib().gpuntt_b200_last_launch_count() == launches
            torch.cuda.synchronize()
            assert (to_host(out, bits).reshape(h, w) == want).all(), generic
            assert (to_host(d, bits).reshape(h, w) == x).all(), "out-of-place call modified its input"
            capi.merge_ntt(in_ptr=out.data_ptr(), out_ptr=out.data_ptr(), table_ptr=itab.data_ptr(), n_power=logh, batch=w, element_bits=bits,
                           direction=capi.INVERSE, reduction_poly=poly, layout=capi.PerCoefficient, modulus=P.modulus, mod_inverse=P.n_inv,
                           stream=s)
            if not generic:
                assert capi.lib().gpuntt_b200_last_launch_count() == launches
            torch.cuda.synchronize()
            assert (to_host(out, bits).reshape(h, w) == x).all(), generic
        finally:
            capi.lib().gpuntt_b200_force_generic_path(0)


def test_per_coefficient_rns_and_limits():
    """RNS PerCoefficient: column j uses modulus[j % mod_count] (ntt.cu:1737-1741); n_power > 9 is rejected like the
    reference (ntt.cu:2230-2233)."""
    bits, logh, w, mod_count = 64, 6, 8, 2
    h = 1 << logh
    primes = rns_primes(bits, logh, mod_count)
    fwd_tab = np.zeros(mod_count << logh, dtype=np.uint64)
    mods = np.zeros((mod_count, 3), dtype=np.uint64)
    params = []
    for m, (p, psi) in enumerate(primes):
        fwd = np.array([pow(psi, i, p) for i in range(h)], dtype=np.uint64)
        fwd_tab[m << logh:(m + 1) << logh] = O.bitrev_table(fwd)
        bit, mu = O.modulus(p, bits)
        mods[m] = (p, bit, mu)
        P = O.MergeParams(logh, O.X_N_plus, bits, p, 0, psi, pow(h, p - 2, p), psi, pow(psi, p - 2, p), h, h)
        P.fwd = fwd
        params.append(P)
    rng = np.random.RandomState(3)
    x = np.zeros((h, w), dtype=np.uint64)
    for j in range(w):
        x[:, j] = rng.randint(0, 2**31, size=h).astype(np.uint64) % np.uint64(primes[j % mod_count][0])
    want = np.stack([O.merge_ntt(np.ascontiguousarray(x[:, j]), params[j % mod_count]) for j in range(w)], axis=1)
    d = to_dev(x, bits)
    mods_d = to_dev(mods.ravel(), bits)
    s = torch.cuda.current_stream().cuda_stream
    capi.merge_ntt(in_ptr=d.data_ptr(), out_ptr=d.data_ptr(), table_ptr=to_dev(fwd_tab, bits).data_ptr(), n_power=logh,
                   batch=w, element_bits=bits, direction=capi.FORWARD, reduction_poly=O.X_N_plus,
                   layout=capi.PerCoefficient, mod_count=mod_count, modulus_dev=mods_d.data_ptr(), stream=s)
    torch.cuda.synchronize()
    assert (to_host(d, bits).reshape(h, w) == want).all()
    with pytest.raises(capi.GpuNttError) as ei:
        capi.merge_ntt(in_ptr=d.data_ptr(), out_ptr=d.data_ptr(), table_ptr=d.data_ptr(), n_power=10, batch=4,
                       layout=capi.PerCoefficient, modulus=primes[0][0], stream=s)
    assert ei.value.status == capi.ERR_N_POWER


@pytest.mark.parametrize("bits,logn,batch", [(64, 16, 4), (32, 14, 3), (64, 20, 1)])
def test_cuda_graph_capture_and_replay(bits, logn, batch):
    """The entry points only enqueue work on cfg.stream (no synchronisation, no allocation once the cached scratch
    exists), so a caller may capture them into a CUDA graph: capture forward + inverse, replay on new data."""
    P = O.merge_params(logn, O.X_N_minus, bits)
    tab, itab = to_dev(P.fwd_br, bits), to_dev(P.inv_br, bits)
    x = O.example_input(P.modulus, batch << logn, seed=77)
    d = to_dev(x, bits).view(batch, -1)
    fwd = torch.empty_like(d)
    back = torch.empty_like(d)
    capi.ntt(d, tab, P.modulus, logn, O.X_N_minus, out=fwd)          # warm-up: scratch buffers, function attributes
    capi.intt(fwd, itab, P.modulus, P.n_inv, logn, O.X_N_minus, out=back)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        capi.ntt(d, tab, P.modulus, logn, O.X_N_minus, out=fwd)
        capi.intt(fwd, itab, P.modulus, P.n_inv, logn, O.X_N_minus, out=back)
    y = O.example_input(P.modulus, batch << logn, seed=78)
    d.copy_(to_dev(y, bits).view(batch, -1))
    fwd.zero_()
    back.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert (to_host(fwd, bits).ravel() == O.merge_ntt(y, P).ravel()).all()
    assert (to_host(back, bits).ravel() == y).all()


@pytest.mark.parametrize("bits", [64, 32])
@pytest.mark.parametrize("poly", [O.X_N_minus, O.X_N_plus])
@pytest.mark.parametrize("logn", [7, 8, 9, 10, 11, 12])
def test_small_rings_single_pass_tuned_kernels(bits, poly, logn):
    """Rings of 2^7..2^11 (64-bit) / 2^8..2^12 (32-bit) whose batch fills whole 2048/4096-element chunks run as ONE launch of
    the tuned kernel (whole transforms inside a tile, two or three register rounds; replaces the reference's LowRing
    kernels and one-launch plans, ntt.cu:11-433).  Batches: one chunk (half-empty tile), an odd number of chunks, and enough
    chunks that every CTA of the persistent grid gets several tiles.  The same calls on the generic kernel agree."""
    small = (7 <= logn <= 11) if bits == 64 else (8 <= logn <= 12)
    if not small:
        pytest.skip("not a small-ring size for this width")
    chunk = 2048 if bits == 64 else 4096
    P = O.merge_params(logn, poly, bits)
    for chunks in (1, 5, 1500):
        batch = (chunks * chunk) >> logn
        x = O.example_input(P.modulus, batch << logn, seed=logn + chunks)
        want = O.merge_ntt(x, P)
        for inplace in (True, False):
            got = run_fwd(x, P, bits, poly, inplace=inplace)
            assert capi.lib().gpuntt_b200_last_launch_count() == 1, "forward did not take the single-pass tuned kernel"
            assert (got == want).all(), (chunks, inplace)
            back = run_inv(want, P, bits, poly, inplace=inplace)
            assert capi.lib().gpuntt_b200_last_launch_count() == 1, "inverse did not take the single-pass tuned kernel"
            assert (back == x).all(), (chunks, inplace)
        capi.lib().gpuntt_b200_force_generic_path(1)
        try:
            assert (run_fwd(x, P, bits, poly) == want).all()
            assert (run_inv(want, P, bits, poly) == x).all()
        finally:
            capi.lib().gpuntt_b200_force_generic_path(0)
    # a batch that ends inside a chunk: from 2^22 elements up it is split -- whole chunks on the tuned kernel (one launch), the ragged
    # tail (fewer polynomials than a chunk holds) on the generic kernel; smaller ones are the generic kernel alone
    for batch, split in ((((3 * chunk) >> logn) + 1, False), (((1 << 22) >> logn) + 1, True), (((1 << 22) >> logn) + (chunk >> logn) - 1, True)):
        if (batch << logn) % chunk:
            x = O.example_input(P.modulus, batch << logn, seed=99 + batch % 7)
            want = O.merge_ntt(x, P)
            for inplace in (True, False):
                assert (run_fwd(x, P, bits, poly, inplace=inplace) == want).all(), (batch, inplace)
                if split:
                    capi.lib().gpuntt_b200_force_generic_path(1)
                    run_fwd(x[: 1 << logn], P, bits, poly)
                    generic_launches = capi.lib().gpuntt_b200_last_launch_count()
                    capi.lib().gpuntt_b200_force_generic_path(0)
                    run_fwd(x, P, bits, poly, inplace=inplace)
                    assert capi.lib().gpuntt_b200_last_launch_count() == 1 + generic_launches, batch
                assert (run_inv(want, P, bits, poly, inplace=inplace) == x).all(), (batch, inplace)


@pytest.mark.parametrize("bits", [64, 32])
@pytest.mark.parametrize("poly", [O.X_N_minus, O.X_N_plus])
def test_one_tile_rings(bits, poly):
    """64-bit N = 2^12 and 32-bit N = 2^13 are exactly one tile: knob ONE_TILE = 2 runs every call of these sizes with the whole
    transform inside the tile (three register rounds, one launch, no hand-off between CTAs), 1 (default) the inverse ones,
    0 none.  All settings against the oracle, every word; ragged and larger-than-the-grid batches."""
    logn = 12 if bits == 64 else 13
    P = O.merge_params(logn, poly, bits)
    try:
        for batch in (1, 3, 150, 297, 700):
            x = O.example_input(P.modulus, batch << logn, seed=logn + batch)
            want = O.merge_ntt(x, P)
            for knob in (2, 1, 0):
                capi.tune(6, knob)
                for inplace in (True, False):
                    got = run_fwd(x, P, bits, poly, inplace=inplace)
                    assert (got == want).all(), (batch, knob, inplace)
                    if knob == 2:
                        assert capi.lib().gpuntt_b200_last_launch_count() == 1
                    back = run_inv(want, P, bits, poly, inplace=inplace)
                    assert (back == x).all(), (batch, knob, inplace)
                    if knob == 2:
                        assert capi.lib().gpuntt_b200_last_launch_count() == 1
    finally:
        capi.tune(6, 1)


@pytest.mark.parametrize("bits,logn,poly", [(64, 17, O.X_N_minus), (64, 19, O.X_N_plus), (64, 21, O.X_N_minus),
                                            (32, 19, O.X_N_minus), (32, 20, O.X_N_plus), (32, 22, O.X_N_minus)])
def test_single_polynomial_tiles_of_large_rings(bits, logn, poly):
    """ONE polynomial of a three-pass ring (64-bit above 2^16, 32-bit above 2^18: the shape of a ZK prover's transform, and what the
    reference's nvbench harness sweeps, bench_merge_ntt.cu:71-75) runs its contiguous pass on 2048- / 4096-element tiles of that
    polynomial; knob SINGLE_POLY_TILES = 0 gives the usual two-polynomial tiles.  Both settings, both directions, in and out of
    place, every word against the oracle."""
    P = O.merge_params(logn, poly, bits)
    x = O.example_input(P.modulus, 1 << logn, seed=logn + 7)
    want = O.merge_ntt(x, P)
    try:
        for knob in (1, 0):
            capi.tune(8, knob)
            for inplace in (True, False):
                assert (run_fwd(x, P, bits, poly, inplace=inplace) == want).all(), (knob, inplace)
                assert capi.lib().gpuntt_b200_last_launch_count() == 3
                assert (run_inv(want, P, bits, poly, inplace=inplace) == x).all(), (knob, inplace)
                assert capi.lib().gpuntt_b200_last_launch_count() == 3
    finally:
        capi.tune(8, 1)


def test_two_host_threads_on_one_stream_do_not_share_scratch_mid_call():
    """scratch (the generic kernel's twiddle companions) is cached per (device, stream); two host threads that issue
    calls on the SAME stream handle must not interleave twiddle_prep_kernel and pass launches.  Every call holds the
    per-stream enqueue lock: both threads' results stay exact (ctypes releases the GIL, so the calls really overlap)."""
    import threading
    bits, logn, batch, rounds = 64, 13, 3, 40
    Ps = [O.merge_params(logn, O.X_N_minus, bits), O.merge_params(logn, O.X_N_plus, bits)]
    xs = [O.example_input(P.modulus, batch << logn, seed=50 + i) for i, P in enumerate(Ps)]
    wants = [O.merge_ntt(x, P) for x, P in zip(xs, Ps)]
    tabs = [to_dev(P.fwd_br, bits) for P in Ps]
    polys = [O.X_N_minus, O.X_N_plus]
    stream = torch.cuda.current_stream().cuda_stream
    bad = []
    capi.lib().gpuntt_b200_force_generic_path(1)
    try:
        def work(i):
            src = to_dev(xs[i], bits)
            out = torch.zeros_like(src)
            for _ in range(rounds):
                capi.merge_ntt(in_ptr=src.data_ptr(), out_ptr=out.data_ptr(), table_ptr=tabs[i].data_ptr(), n_power=logn, batch=batch,
                               element_bits=bits, direction=capi.FORWARD, reduction_poly=polys[i], modulus=Ps[i].modulus, stream=stream)
                torch.cuda.synchronize()
                if not (to_host(out, bits) == wants[i]).all():
                    bad.append(i)
                    return
        threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    finally:
        capi.lib().gpuntt_b200_force_generic_path(0)
    assert not bad, f"thread(s) {sorted(set(bad))} read another call's twiddle companions"
