"""GPU parity tests of the 4-step path through the C ABI (gpuntt_b200_4step_ntt / gpuntt_b200_transpose),
bit-exact against the oracle's restatement of NTT_4STEP_CPU (oracle/ntt_oracle.c, pinned to the reference by
tests/golden/golden.json), mirroring the reference's gpu_4step_ntt_examples / gpu_4step_intt_examples
(example/ntt_4step/test_4step_ntt.cu:147-166, test_4step_intt.cu:82-84,155-166)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from gpu_ntt_b200 import capi  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.gpu_util import to_dev, to_host  # noqa: E402

pytestmark = pytest.mark.gpu


def tables(P, bits, inverse):
    t1, t2, W = (P.t1_inv, P.t2_inv, P.W_inv) if inverse else (P.t1, P.t2, P.W)
    # the examples upload bit-reversed small tables and the natural-layout W (test_4step_ntt.cu:90-110)
    return to_dev(O.bitrev_table(t1), bits), to_dev(O.bitrev_table(t2), bits), to_dev(W, bits)


def transposed(a, rows, cols):
    """per-polynomial rows x cols -> cols x rows"""
    return np.ascontiguousarray(a.reshape(-1, rows, cols).transpose(0, 2, 1)).reshape(a.shape)


@pytest.mark.parametrize("bits", [64, 32])
@pytest.mark.parametrize("logn,batch", [(12, 3), (13, 1), (14, 2), (15, 1), (16, 2), (17, 1), (18, 1), (20, 1)])
def test_4step_fused_and_reference_contracts(bits, logn, batch):
    P = O.fourstep_params(logn, O.X_N_minus, bits)
    assert capi.fourstep_shape(logn) == (P.n1, P.n2)
    n = P.n
    x = O.example_input(P.modulus, batch * n, seed=logn).reshape(batch, n)
    want = O.fourstep_ntt(x, P)
    t1, t2, W = tables(P, bits, False)

    # fused contract: natural in, NTT_4STEP_CPU::ntt order out; out of place and in place
    d = to_dev(x, bits)
    out = torch.zeros_like(d)
    capi.fourstep_ntt(d.view(batch, n), t1, t2, W, P.modulus, logn, out=out.view(batch, n))
    torch.cuda.synchronize()
    assert (to_host(out, bits) == want).all()
    assert (to_host(d, bits) == x).all(), "out-of-place call modified its input"
    capi.fourstep_ntt(d.view(batch, n), t1, t2, W, P.modulus, logn)
    torch.cuda.synchronize()
    assert (to_host(d, bits) == want).all()

    # the reference's call sequence: GPU_Transpose, GPU_4STEP_NTT, GPU_Transpose (test_4step_ntt.cu:147-154)
    a = to_dev(x, bits)
    tmp = torch.zeros_like(a)
    capi.transpose(a.view(batch, n), tmp.view(batch, n), P.n1, P.n2, logn)
    torch.cuda.synchronize()
    assert (to_host(tmp, bits) == transposed(x, P.n1, P.n2)).all()
    capi.fourstep_ntt(tmp.view(batch, n), t1, t2, W, P.modulus, logn, io_contract=capi.FOURSTEP_REFERENCE,
                      out=a.view(batch, n))
    capi.transpose(a.view(batch, n), tmp.view(batch, n), P.n1, P.n2, logn)
    torch.cuda.synchronize()
    assert (to_host(tmp, bits) == want).all()

    # inverse, fused: NTT order in, natural out
    it1, it2, iW = tables(P, bits, True)
    y = to_dev(want, bits)
    capi.fourstep_ntt(y.view(batch, n), it1, it2, iW, P.modulus, logn, direction=capi.INVERSE, mod_inverse=P.n_inv)
    torch.cuda.synchronize()
    assert (to_host(y, bits) == x).all()
    z = O.example_input(P.modulus, batch * n, seed=99 + logn).reshape(batch, n)
    winv = O.fourstep_intt(z, P)
    zd = to_dev(z, bits)
    zo = torch.zeros_like(zd)
    capi.fourstep_ntt(zd.view(batch, n), it1, it2, iW, P.modulus, logn, direction=capi.INVERSE, mod_inverse=P.n_inv,
                      out=zo.view(batch, n))
    torch.cuda.synchronize()
    assert (to_host(zo, bits) == winv).all()

    # inverse, the reference's sequence: host intt_first_transpose, GPU_4STEP_NTT(INVERSE), GPU_Transpose
    # (test_4step_intt.cu:82-84, 155-159)
    pre = to_dev(O.fourstep_intt_first_transpose(z, P), bits)
    r = torch.zeros_like(pre)
    capi.fourstep_ntt(pre.view(batch, n), it1, it2, iW, P.modulus, logn, direction=capi.INVERSE, mod_inverse=P.n_inv,
                      io_contract=capi.FOURSTEP_REFERENCE, out=r.view(batch, n))
    capi.transpose(r.view(batch, n), pre.view(batch, n), P.n1, P.n2, logn)
    torch.cuda.synchronize()
    assert (to_host(pre, bits) == winv).all()


def test_4step_golden_hashes(golden):
    """Outputs hash-identical to what the reference's NTT_4STEP_CPU produced when the fixtures were made."""
    for g in golden["fourstep"]:
        if g["logn"] > 20:
            continue
        bits, logn = g["width"], g["logn"]
        P = O.fourstep_params(logn, O.X_N_minus, bits, inverse_tables=False)
        x = O.example_input(P.modulus, P.n)
        assert str(O.fold_hash(x)) == g["in_hash"]
        t1, t2, W = tables(P, bits, False)
        d = to_dev(x, bits)
        capi.fourstep_ntt(d.view(1, P.n), t1, t2, W, P.modulus, logn)
        torch.cuda.synchronize()
        y = to_host(d, bits)
        assert [str(int(v)) for v in y[:4]] == g["ntt_head"] and str(O.fold_hash(y)) == g["ntt_hash"]


def test_4step_rns_overload_single_modulus_group():
    """GPU_4STEP_NTT RNS overload as the reference's example drives it: mod_count = 1 (test_4step_ntt.cu:147-154)."""
    bits, logn, batch = 64, 14, 3
    P = O.fourstep_params(logn, O.X_N_minus, bits)
    x = O.example_input(P.modulus, batch * P.n, seed=3).reshape(batch, P.n)
    want = O.fourstep_ntt(x, P)
    t1, t2, W = tables(P, bits, False)
    bit, mu = O.modulus(P.modulus, bits)
    mods = to_dev(np.array([P.modulus, bit, mu], dtype=np.uint64), bits)
    d = to_dev(x, bits)
    capi.fourstep_ntt(d.view(batch, P.n), t1, t2, W, 0, logn, mod_count=1, modulus_dev=mods.data_ptr())
    torch.cuda.synchronize()
    assert (to_host(d, bits) == want).all()
    it1, it2, iW = tables(P, bits, True)
    ninv = to_dev(np.array([P.n_inv], dtype=np.uint64), bits)
    capi.fourstep_ntt(d.view(batch, P.n), it1, it2, iW, 0, logn, direction=capi.INVERSE, mod_count=1,
                      modulus_dev=mods.data_ptr(), mod_inverse_dev=ninv.data_ptr())
    torch.cuda.synchronize()
    assert (to_host(d, bits) == x).all()


def _threaded(fn, x, P, threads=None):
    import concurrent.futures as cf
    import os
    threads = threads or min(16, os.cpu_count() or 1)
    rows = x.reshape(-1, P.n)
    out = np.empty_like(rows)

    def work(idx):
        for r in idx:
            out[r] = fn(rows[r], P)
    with cf.ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, [list(range(i, rows.shape[0], threads)) for i in range(threads)]))
    return out.reshape(x.shape)


def test_4step_c4_full_size_every_polynomial():
    """BASELINE config C4 as SURVEY 8(d) prescribes it: NTTParameters4Step<Data64>(24, X_N_minus), batch 16, seed-0 stream;
    ALL 16 polynomials against NTT_4STEP_CPU::ntt (restated; ~4 s per polynomial, sharded over the host threads), the
    SURVEY 8(c) KAT of the first one, and the inverse restores every input word.  The fused-contract forward call must
    not launch a transpose kernel: column pass + pair table + two row passes = 4 launches."""
    bits, logn, batch = 64, 24, 16
    P = O.fourstep_params(logn, O.X_N_minus, bits)
    p = P.modulus
    t1, t2, W = tables(P, bits, False)
    x = O.example_input(p, batch * P.n, seed=0).reshape(batch, P.n)
    want = _threaded(O.fourstep_ntt, x, P)
    d = to_dev(x, bits)
    capi.fourstep_ntt(d.view(batch, P.n), t1, t2, W, p, logn)
    torch.cuda.synchronize()
    assert capi.lib().gpuntt_b200_last_launch_count() == 4
    y = to_host(d, bits).reshape(batch, P.n)
    assert (y == want).all(), f"{int((y != want).sum())} words differ"
    assert O.fold_hash(y[0]) == 10069984314045308296      # SURVEY.md 8(c) KAT captured from the reference
    del t1, t2, W
    it1, it2, iW = tables(P, bits, True)
    capi.fourstep_ntt(d.view(batch, P.n), it1, it2, iW, p, logn, direction=capi.INVERSE, mod_inverse=P.n_inv)
    torch.cuda.synchronize()
    assert (to_host(d, bits).reshape(batch, P.n) == x).all()


@pytest.mark.parametrize("logn,batch", [(19, 2), (21, 2), (22, 1), (23, 2)])
def test_4step_remaining_large_shapes(logn, batch):
    """The 4-step shapes the round-1 suite skipped (n1 x n2 = 32 x 16384, 64 x 32768, 128 x 32768, 128 x 65536): both I/O
    contracts forward, fused inverse; every word against the oracle."""
    bits = 64
    P = O.fourstep_params(logn, O.X_N_minus, bits)
    n = P.n
    x = O.example_input(P.modulus, batch * n, seed=logn).reshape(batch, n)
    want = _threaded(O.fourstep_ntt, x, P)
    t1, t2, W = tables(P, bits, False)
    d = to_dev(x, bits)
    out = torch.zeros_like(d)
    capi.fourstep_ntt(d.view(batch, n), t1, t2, W, P.modulus, logn, out=out.view(batch, n))
    torch.cuda.synchronize()
    assert (to_host(out, bits) == want).all()
    a = torch.zeros_like(d)
    capi.transpose(d.view(batch, n), a.view(batch, n), P.n1, P.n2, logn)
    capi.fourstep_ntt(a.view(batch, n), t1, t2, W, P.modulus, logn, io_contract=capi.FOURSTEP_REFERENCE, out=d.view(batch, n))
    capi.transpose(d.view(batch, n), a.view(batch, n), P.n1, P.n2, logn)
    torch.cuda.synchronize()
    assert (to_host(a, bits) == want).all()
    it1, it2, iW = tables(P, bits, True)
    capi.fourstep_ntt(out.view(batch, n), it1, it2, iW, P.modulus, logn, direction=capi.INVERSE, mod_inverse=P.n_inv)
    torch.cuda.synchronize()
    assert (to_host(out, bits) == x).all()


def test_4step_transposed_and_transpose_kernel_paths_agree():
    """The fused-contract forward call with and without the transposing column pass (gpuntt_b200_tune 4STEP_TRANSPOSED):
    same words, different launch lists."""
    bits, logn, batch = 64, 20, 3
    P = O.fourstep_params(logn, O.X_N_minus, bits, inverse_tables=False)
    x = O.example_input(P.modulus, batch * P.n, seed=5).reshape(batch, P.n)
    want = _threaded(O.fourstep_ntt, x, P)
    t1, t2, W = tables(P, bits, False)
    outs = []
    try:
        for mode in (1, 0):
            capi.tune(3, mode)
            d = to_dev(x, bits)
            capi.lib().gpuntt_b200_set_profiling(1)       # launch kinds: 9 = transpose_kernel
            capi.profile_read()
            capi.fourstep_ntt(d.view(batch, P.n), t1, t2, W, P.modulus, logn)
            torch.cuda.synchronize()
            outs.append((to_host(d, bits), [k for k, _ in capi.profile_read()]))
    finally:
        capi.tune(3, 1)
        capi.lib().gpuntt_b200_set_profiling(0)
    assert (outs[0][0] == want).all() and (outs[1][0] == want).all()
    assert 9 not in outs[0][1] and 9 in outs[1][1]


@pytest.mark.parametrize("logn,batch", [(12, 4), (13, 7), (17, 6), (19, 5), (20, 9), (22, 4)])
def test_4step_column_kernel_with_resident_pairs(logn, batch):
    """Batches of four or more polynomials take the position-major column kernel (merge_wcol.cu: the (W, W') pairs of a tile
    position are loaded once into shared memory and reused by every polynomial); smaller ones and knob 5 = 0 the per-tile
    kernel.  Every n1 (32, 64, 128) x batch parity here, n1 = 256 in the C4 test; both kernels against the oracle."""
    bits = 64
    P = O.fourstep_params(logn, O.X_N_minus, bits, inverse_tables=False)
    x = O.example_input(P.modulus, batch * P.n, seed=logn + batch).reshape(batch, P.n)
    want = _threaded(O.fourstep_ntt, x, P)
    t1, t2, W = tables(P, bits, False)
    try:
        for knob in (1, 0):
            capi.tune(5, knob)
            d = to_dev(x, bits)
            out = torch.zeros_like(d)
            capi.fourstep_ntt(d.view(batch, P.n), t1, t2, W, P.modulus, logn, out=out.view(batch, P.n))
            torch.cuda.synchronize()
            got = to_host(out, bits)
            assert (got == want).all(), f"knob={knob}: {int((got != want).sum())} words differ"
            capi.fourstep_ntt(d.view(batch, P.n), t1, t2, W, P.modulus, logn)
            torch.cuda.synchronize()
            assert (to_host(d, bits) == want).all(), f"knob={knob} in place"
    finally:
        capi.tune(5, 1)


def _kinds_of(call):
    """launch kinds of one call (9 = transpose_kernel)"""
    capi.lib().gpuntt_b200_set_profiling(1)
    try:
        capi.profile_read()
        call()
        torch.cuda.synchronize()
        return [k for k, _ in capi.profile_read()]
    finally:
        capi.lib().gpuntt_b200_set_profiling(0)


@pytest.mark.parametrize("logn,batch", [(12, 4), (13, 5), (16, 4), (17, 6), (18, 4), (20, 4), (21, 4), (22, 5), (23, 4)])
def test_4step_reference_contract_forward_without_transpose_kernel(logn, batch):
    """Reference contract (GPU_Transpose'd input in, n1 x n2 matrix out), batch >= 4: contiguous column pass on the transposed
    input with the pairs resident in shared memory, row passes along that layout, transposing store in the last one -- no
    transpose_kernel inside the call; knob 3 = 0 gives the same words through the transposing path."""
    bits = 64
    P = O.fourstep_params(logn, O.X_N_minus, bits, inverse_tables=False)
    n = P.n
    x = O.example_input(P.modulus, batch * n, seed=7 * logn + batch).reshape(batch, n)
    want = _threaded(O.fourstep_ntt, x, P)
    t1, t2, W = tables(P, bits, False)
    xt = to_dev(transposed(x, P.n1, P.n2), bits)       # what GPU_Transpose(x, row = n1, col = n2) leaves
    try:
        for knob in (1, 0):
            capi.tune(3, knob)
            r = torch.zeros_like(xt)
            kinds = _kinds_of(lambda: capi.fourstep_ntt(xt.view(batch, n), t1, t2, W, P.modulus, logn, io_contract=capi.FOURSTEP_REFERENCE,
                                                        out=r.view(batch, n)))
            got = transposed(to_host(r, bits), P.n1, P.n2)  # the caller's closing GPU_Transpose
            assert (got == want).all(), f"knob={knob}: {int((got != want).sum())} words differ"
            assert (to_host(xt, bits) == transposed(x, P.n1, P.n2)).all(), "the call modified its input"
            lg2 = P.n2.bit_length() - 1
            covered = lg2 in (7, 8) or 4 <= lg2 - (8 if lg2 - 8 >= 4 else 7) <= 8    # fast_fourstep_rows_t_supported
            assert (9 in kinds) == (knob == 0 or not covered), kinds
    finally:
        capi.tune(3, 1)


@pytest.mark.parametrize("logn,batch", [(12, 4), (13, 2), (16, 5), (17, 4), (18, 3), (19, 4), (21, 4), (21, 2), (22, 4), (23, 3)])
def test_4step_inverse_without_transpose_kernel(logn, batch):
    """Inverse, both contracts, against NTT_4STEP_CPU::intt.  Fused contract with n1 >= 64: the first pass reads y as a strided
    pass with a transposing store; reference contract: the product pass stores the n1 x n2 matrix and the last stages run along
    its rows.  Batches of four or more take the position-major product kernel (pairs resident in shared memory), smaller ones
    the per-tile kernel in column-chunk-major order."""
    bits = 64
    P = O.fourstep_params(logn, O.X_N_minus, bits)
    n = P.n
    lg1, lg2 = P.n1.bit_length() - 1, P.n2.bit_length() - 1
    z = O.example_input(P.modulus, batch * n, seed=11 * logn + batch).reshape(batch, n)
    want = _threaded(O.fourstep_intt, z, P)
    it1, it2, iW = tables(P, bits, True)
    da = lg2 if lg2 <= 8 else max(lg2 - 8, 12 - lg1, 4)
    db = lg2 - da
    tuned = (db == 0 and da >= 4) or (db >= 4 and da <= 8)
    try:
        for knob in (1, 0):
            capi.tune(3, knob)
            zd = to_dev(z, bits)
            zo = torch.zeros_like(zd)
            kinds = _kinds_of(lambda: capi.fourstep_ntt(zd.view(batch, n), it1, it2, iW, P.modulus, logn, direction=capi.INVERSE,
                                                        mod_inverse=P.n_inv, out=zo.view(batch, n)))
            got = to_host(zo, bits)
            assert (got == want).all(), f"fused knob={knob}: {int((got != want).sum())} words differ"
            if knob == 1 and tuned and lg1 >= 6:
                assert 9 not in kinds, kinds
            capi.fourstep_ntt(zd.view(batch, n), it1, it2, iW, P.modulus, logn, direction=capi.INVERSE, mod_inverse=P.n_inv)
            torch.cuda.synchronize()
            assert (to_host(zd, bits) == want).all(), f"fused in place knob={knob}"
            pre = to_dev(O.fourstep_intt_first_transpose(z, P), bits)
            r = torch.zeros_like(pre)
            kinds = _kinds_of(lambda: capi.fourstep_ntt(pre.view(batch, n), it1, it2, iW, P.modulus, logn, direction=capi.INVERSE,
                                                        mod_inverse=P.n_inv, io_contract=capi.FOURSTEP_REFERENCE, out=r.view(batch, n)))
            got = transposed(to_host(r, bits), P.n1, P.n2)
            assert (got == want).all(), f"reference knob={knob}: {int((got != want).sum())} words differ"
            if knob == 1 and tuned and da >= 5 and (db == 0 or da + db >= 12):
                assert 9 not in kinds, kinds
    finally:
        capi.tune(3, 1)


def test_4step_errors():
    t = torch.zeros(1 << 12, dtype=torch.int64, device="cuda")
    with pytest.raises(capi.GpuNttError) as ei:
        capi.fourstep_ntt(t.view(1, -1), t, t, t, 17, 11)
    assert ei.value.status == capi.ERR_N_POWER
    with pytest.raises(capi.GpuNttError):
        capi.fourstep_ntt(t.view(1, -1), t, t, t, 576460752303415297, 12, io_contract=capi.FOURSTEP_REFERENCE)  # in == out


def test_4step_rns_overload_one_modulus_takes_the_tuned_kernels_and_can_be_captured():
    """With exactly one device modulus the RNS overload reads that Modulus back once (32 bytes, cached per pointer) and continues
    as the single-modulus form -- same launches, same kernels -- and, once the stream's scratch exists, can be captured into a
    CUDA graph and replayed.  Every result must equal the oracle."""
    bits, logn, batch = 64, 16, 2
    P = O.fourstep_params(logn, O.X_N_minus, bits)
    x = O.example_input(P.modulus, batch * P.n, seed=11).reshape(batch, P.n)
    want = O.fourstep_ntt(x, P)
    t1, t2, W = tables(P, bits, False)
    bit, mu = O.modulus(P.modulus, bits)
    mods = to_dev(np.array([P.modulus, bit, mu], dtype=np.uint64), bits)
    d = to_dev(x, bits).view(batch, P.n)
    out = torch.empty_like(d)
    capi.fourstep_ntt(d, t1, t2, W, P.modulus, logn, out=out)
    torch.cuda.synchronize()
    single_launches = capi.lib().gpuntt_b200_last_launch_count()
    assert (to_host(out, bits).reshape(batch, -1) == want).all()
    out.zero_()
    capi.fourstep_ntt(d, t1, t2, W, 0, logn, out=out, mod_count=1, modulus_dev=mods.data_ptr())
    torch.cuda.synchronize()
    assert capi.lib().gpuntt_b200_last_launch_count() == single_launches
    assert (to_host(out, bits).reshape(batch, -1) == want).all()
    # capture on an explicit stream: scratch cannot be allocated inside a capture, so a call that would have to is declined
    # cleanly (the capture stays valid) ...
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=side):
        with pytest.raises(capi.GpuNttError):
            capi.fourstep_ntt(d, t1, t2, W, 0, logn, out=out, mod_count=1, modulus_dev=mods.data_ptr())
    # ... and once the stream's scratch exists (one call outside a capture; the Modulus behind the pointer is already cached,
    # so nothing synchronises) the same call is captured and replays
    with torch.cuda.stream(side):
        capi.fourstep_ntt(d, t1, t2, W, 0, logn, out=out, mod_count=1, modulus_dev=mods.data_ptr())
    side.synchronize()
    assert (to_host(out, bits).reshape(batch, -1) == want).all()
    g = torch.cuda.CUDAGraph()
    out.zero_()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=side):
        capi.fourstep_ntt(d, t1, t2, W, 0, logn, out=out, mod_count=1, modulus_dev=mods.data_ptr())
    g.replay()
    torch.cuda.synchronize()
    assert (to_host(out, bits).reshape(batch, -1) == want).all()
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert (to_host(out, bits).reshape(batch, -1) == want).all()


def custom_fourstep_params(logn, p):
    """FourStepParams for an arbitrary prime p = 1 (mod 2^logn) (what NTTParameters4Step builds from a factor set)."""
    import copy
    P = copy.copy(O.fourstep_params(logn, O.X_N_minus, 64, inverse_tables=False))
    n = 1 << logn
    assert (p - 1) % n == 0
    omega = 0
    for g in range(2, 1000):
        c = pow(g, (p - 1) // n, p)
        if pow(c, n // 2, p) == p - 1:      # order exactly N
            omega = c
            break
    assert omega
    P.modulus, P.omega, P.root, P.inv_root, P.n_inv = p, omega, omega, pow(omega, p - 2, p), pow(n, p - 2, p)
    P.t1 = np.empty(P.n1 // 2, dtype=np.uint64)
    P.t2 = np.empty(P.n2 // 2, dtype=np.uint64)
    O.lib().ora_4step_small_tables(P.root, p, P.n, P.n1, P.n2, 0, P.t1, P.t2)
    P.W = np.empty(P.n, dtype=np.uint64)
    O.lib().ora_4step_w_table(P.root, p, P.n1, P.n2, 0, P.W)
    P.t1_inv = np.empty(P.n1 // 2, dtype=np.uint64)
    P.t2_inv = np.empty(P.n2 // 2, dtype=np.uint64)
    O.lib().ora_4step_small_tables(P.root, p, P.n, P.n1, P.n2, 1, P.t1_inv, P.t2_inv)
    P.W_inv = np.empty(P.n, dtype=np.uint64)
    O.lib().ora_4step_w_table(P.inv_root, p, P.n1, P.n2, 1, P.W_inv)
    return P


@pytest.mark.parametrize("logn", [16, 17, 18, 19])
def test_4step_moduli_at_the_top_of_the_lazy_range(logn):
    """Primes just below 2^60 - 2^31 (the top of the F60 policy): 20p exceeds 2^64 there, so any stage that lets a value
    pass 16p wraps.  logn 18 is the shape whose row phase opens with a three-stage first round on inputs below 2p
    (Shape<3,2>, in_bound 2) -- the case the round-1 advisor found; extreme inputs (all p-1) maximise the lazy values."""
    from tests.test_moduli_gpu import ntt_prime_below
    p = ntt_prime_below((1 << 60) - (1 << 31), 1 << logn)
    P = custom_fourstep_params(logn, p)
    batch = 4       # (four: the reference-contract call below takes the position-major kernels)
    rng = np.random.default_rng(logn)
    x = rng.integers(0, p, size=(batch, P.n), dtype=np.uint64)
    x[1, :] = p - 1
    x[2, ::2] = p - 1
    x[2, 1::2] = 0
    want = O.fourstep_ntt(x, P)
    t1, t2, W = tables(P, 64, False)
    d = to_dev(x, 64)
    capi.fourstep_ntt(d.view(batch, P.n), t1, t2, W, p, logn)
    torch.cuda.synchronize()
    assert (to_host(d, 64).reshape(batch, -1) == want).all()
    xt = to_dev(transposed(x, P.n1, P.n2), 64)
    r = torch.zeros_like(xt)
    capi.fourstep_ntt(xt.view(batch, P.n), t1, t2, W, p, logn, io_contract=capi.FOURSTEP_REFERENCE, out=r.view(batch, P.n))
    torch.cuda.synchronize()
    assert (transposed(to_host(r, 64), P.n1, P.n2).reshape(batch, -1) == want).all()
    it1, it2, iW = tables(P, 64, True)
    capi.fourstep_ntt(d.view(batch, P.n), it1, it2, iW, p, logn, direction=capi.INVERSE, mod_inverse=P.n_inv)
    torch.cuda.synchronize()
    assert (to_host(d, 64).reshape(batch, -1) == x).all()


@pytest.mark.parametrize("limit", [1 << 30, (1 << 36) + (1 << 30), (1 << 40) - 1, (1 << 60) + (1 << 25), (1 << 60) + (1 << 58) + (1 << 40),
                                   (1 << 62) - 1])
@pytest.mark.parametrize("logn,batch", [(13, 4), (17, 4), (20, 2)])
def test_4step_modulus_ranges(limit, logn, batch):
    """Moduli outside the lazy policies of the tuned 4-step kernels (below 2^40, above 2^60 - 2^31, up to the reference's 2^62 limit,
    modular_arith.cuh:66-67): whichever kernels a call lands on -- generic product passes, exact-policy row transforms -- both
    contracts and both directions equal NTT_4STEP_CPU word for word."""
    from tests.test_moduli_gpu import ntt_prime_below
    p = ntt_prime_below(limit, 1 << logn)
    P = custom_fourstep_params(logn, p)
    rng = np.random.default_rng(logn + limit % 1009)
    x = rng.integers(0, p, size=(batch, P.n), dtype=np.uint64)
    x[1, :] = p - 1
    want = O.fourstep_ntt(x, P)
    t1, t2, W = tables(P, 64, False)
    it1, it2, iW = tables(P, 64, True)
    d = to_dev(x, 64)
    capi.fourstep_ntt(d.view(batch, P.n), t1, t2, W, p, logn)
    torch.cuda.synchronize()
    assert (to_host(d, 64).reshape(batch, -1) == want).all(), f"fused forward p={p}"
    capi.fourstep_ntt(d.view(batch, P.n), it1, it2, iW, p, logn, direction=capi.INVERSE, mod_inverse=P.n_inv)
    torch.cuda.synchronize()
    assert (to_host(d, 64).reshape(batch, -1) == x).all(), f"fused inverse p={p}"
    xt = to_dev(transposed(x, P.n1, P.n2), 64)
    r = torch.zeros_like(xt)
    capi.fourstep_ntt(xt.view(batch, P.n), t1, t2, W, p, logn, io_contract=capi.FOURSTEP_REFERENCE, out=r.view(batch, P.n))
    torch.cuda.synchronize()
    assert (transposed(to_host(r, 64), P.n1, P.n2).reshape(batch, -1) == want).all(), f"reference-contract forward p={p}"
