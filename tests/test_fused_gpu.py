"""GPU parity tests of the single-launch two-pass kernels (gpu_ntt_b200/csrc/merge_fused.cu): the strided and the
contiguous pass of a 2^12..2^16 (64-bit) / 2^13..2^18 (32-bit) transform run in ONE launch, chained through the L2
by per-polynomial counters.  Replaces the reference's two-launch plans (ntt.cuh:628-636, ntt.cu:2104-2141); every
output word is compared with the oracle, and the launch count is asserted through gpuntt_b200_last_launch_count()."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from gpu_ntt_b200 import capi  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.gpu_util import to_dev, to_host  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fused_everywhere():
    """policy 2: the fused kernel wherever the shapes allow (the default policy keeps large 64-bit batches on two
    launches because they are multiplier-bound and measured ~8 % slower fused)"""
    capi.tune(capi.TUNE_FUSED_PASSES, 2)
    yield
    capi.tune(capi.TUNE_FUSED_PASSES, 1)


def _launches():
    return capi.lib().gpuntt_b200_last_launch_count()


def _threaded_oracle(fn, x, P, threads=16):
    """the C oracle releases the GIL inside ctypes: shard the polynomials over host threads"""
    import concurrent.futures as cf
    rows = x.reshape(-1, P.n)
    out = np.empty_like(rows)
    parts = [list(range(i, rows.shape[0], threads)) for i in range(threads)]

    def work(idx):
        for r in idx:
            out[r] = fn(rows[r], P)
    with cf.ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, parts))
    return out.reshape(x.shape)


# batches: 1 (split mode, half-empty contiguous tiles), odd (ragged last tile), a few CTAs per pass, and far more tiles
# than co-resident CTAs (the merged order with its lag is exercised end to end)
CASES = [(64, 12, [1, 3, 64, 9000]), (64, 13, [1, 5, 2500]), (64, 14, [2, 7, 1200]), (64, 15, [1, 9, 700]),
         (64, 16, [1, 2, 5, 33, 300]), (32, 13, [1, 3, 5000]), (32, 14, [1, 6, 4096]), (32, 15, [3, 1111]),
         (32, 16, [1, 640]), (32, 17, [2, 257]), (32, 18, [1, 3, 130])]


@pytest.mark.parametrize("poly", [O.X_N_minus, O.X_N_plus])
@pytest.mark.parametrize("bits,logn,batches", CASES)
def test_fused_two_pass_matches_oracle(bits, logn, batches, poly):
    P = O.merge_params(logn, poly, bits)
    tab, itab = to_dev(P.fwd_br, bits), to_dev(P.inv_br, bits)
    for batch in batches:
        x = O.example_input(P.modulus, batch << logn, seed=logn * 131 + batch)
        want = _threaded_oracle(O.merge_ntt, x, P)
        d = to_dev(x, bits)
        capi.ntt(d.view(batch, -1), tab, P.modulus, logn, poly)
        assert _launches() == 1, "forward transform was not a single launch"
        torch.cuda.synchronize()
        assert (to_host(d, bits) == want).all(), f"forward mismatch batch={batch}"
        # second call on the same stream: the self-cleaning counters must be back at zero
        d2 = to_dev(x, bits)
        out = torch.zeros_like(d2)
        capi.ntt(d2.view(batch, -1), tab, P.modulus, logn, poly, out=out.view(batch, -1))
        assert _launches() == 1
        torch.cuda.synchronize()
        assert (to_host(out, bits) == want).all(), f"out-of-place forward mismatch batch={batch}"
        assert (to_host(d2, bits) == x).all(), "out-of-place call modified its input"
        capi.intt(d.view(batch, -1), itab, P.modulus, P.n_inv, logn, poly)
        assert _launches() == 1, "inverse transform was not a single launch"
        torch.cuda.synchronize()
        assert (to_host(d, bits) == x).all(), f"inverse mismatch batch={batch}"
        back = torch.zeros_like(out)
        capi.intt(out.view(batch, -1), itab, P.modulus, P.n_inv, logn, poly, out=back.view(batch, -1))
        torch.cuda.synchronize()
        assert (to_host(back, bits) == x).all(), f"out-of-place inverse mismatch batch={batch}"


def test_fused_knob_falls_back_to_two_launches():
    P = O.merge_params(16, O.X_N_minus, 64)
    x = O.example_input(P.modulus, 4 << 16, seed=7)
    want = O.merge_ntt(x, P)
    tab = to_dev(P.fwd_br, 64)
    try:
        capi.tune(capi.TUNE_FUSED_PASSES, 0)
        d = to_dev(x, 64)
        capi.ntt(d.view(4, -1), tab, P.modulus, 16, O.X_N_minus)
        assert _launches() == 2
        torch.cuda.synchronize()
        assert (to_host(d, 64) == want).all()
    finally:
        capi.tune(capi.TUNE_FUSED_PASSES, 2)
    for lag in (0, 1, 9):
        try:
            capi.tune(capi.TUNE_FUSED_LAG, lag)
            d = to_dev(np.tile(x, 100), 64)
            capi.ntt(d.view(400, -1), tab, P.modulus, 16, O.X_N_minus)
            assert _launches() == 1
            torch.cuda.synchronize()
            assert (to_host(d, 64).reshape(100, -1) == want.reshape(1, -1)).all(), f"lag={lag}"
        finally:
            capi.tune(capi.TUNE_FUSED_LAG, 4)


def test_fused_concurrent_streams():
    """Two streams transform different batches at the same time: each stream owns its counters."""
    P = O.merge_params(15, O.X_N_plus, 64)
    tab = to_dev(P.fwd_br, 64)
    xs = [O.example_input(P.modulus, 300 << 15, seed=s) for s in (1, 2)]
    wants = [_threaded_oracle(O.merge_ntt, x, P) for x in xs]
    ds = [to_dev(x, 64) for x in xs]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for d, s in zip(ds, streams):
        capi.ntt(d.view(300, -1), tab, P.modulus, 15, O.X_N_plus, stream=s)
    torch.cuda.synchronize()
    for d, w in zip(ds, wants):
        assert (to_host(d, 64) == w).all()


def test_fused_streams_stress():
    """Three streams issue forward / inverse single-launch calls back to back without any synchronisation while a fourth keeps
    the SMs busy with unrelated kernels: full-grid persistent kernels of different calls queue behind each other (a later grid
    starts as the earlier one drains), every stream's counters stay its own, and the round trips end on the input."""
    cases = [(64, 13, 2500), (32, 14, 4096), (64, 15, 300)]
    work = []
    for (bits, logn, batch), seed in zip(cases, (11, 12, 13)):
        P = O.merge_params(logn, O.X_N_plus, bits)
        x = O.example_input(P.modulus, batch << logn, seed=seed)
        work.append((bits, logn, batch, P, x, to_dev(x, bits), to_dev(P.fwd_br, bits), to_dev(P.inv_br, bits), torch.cuda.Stream()))
    noise_stream = torch.cuda.Stream()
    noise = torch.ones(1 << 26, device="cuda")
    torch.cuda.synchronize()
    for it in range(12):
        with torch.cuda.stream(noise_stream):
            noise.mul_(1.0001).add_(0.5)
        for bits, logn, batch, P, x, d, tab, itab, s in work:
            capi.ntt(d.view(batch, -1), tab, P.modulus, logn, O.X_N_plus, stream=s)
            assert _launches() == 1
            capi.intt(d.view(batch, -1), itab, P.modulus, P.n_inv, logn, O.X_N_plus, stream=s)
    for bits, logn, batch, P, x, d, tab, itab, s in work:
        capi.ntt(d.view(batch, -1), tab, P.modulus, logn, O.X_N_plus, stream=s)
    torch.cuda.synchronize()
    for bits, logn, batch, P, x, d, tab, itab, s in work:
        assert (to_host(d, bits) == _threaded_oracle(O.merge_ntt, x, P)).all(), (bits, logn)


def test_default_policy_launch_counts():
    """Default policy: 32-bit two-pass transforms are one launch at every batch size, 64-bit ones while the call is
    launch-bound (at most one strided tile per SM); results identical either way."""
    capi.tune(capi.TUNE_FUSED_PASSES, 1)
    for bits, logn, batch, expect in ((32, 14, 600, 1), (64, 16, 2, 1), (64, 16, 300, 2), (64, 13, 8, 1), (64, 12, 2000, 1), (32, 17, 4, 1)):
        P = O.merge_params(logn, O.X_N_minus, bits)
        x = O.example_input(P.modulus, batch << logn, seed=batch)
        d = to_dev(x, bits)
        capi.ntt(d.view(batch, -1), to_dev(P.fwd_br, bits), P.modulus, logn, O.X_N_minus)
        assert _launches() == expect, (bits, logn, batch)
        torch.cuda.synchronize()
        assert (to_host(d, bits) == _threaded_oracle(O.merge_ntt, x, P)).all()


@pytest.mark.parametrize("poly", [O.X_N_minus, O.X_N_plus])
@pytest.mark.parametrize("logn,batches", [(12, [1, 3, 64, 65]), (13, [1, 5, 32, 33]), (14, [2, 7, 16, 17])])
def test_small_tile_variant_for_launch_bound_calls(logn, batches, poly):
    """64-bit calls of at most 2^18 elements (knob SMALL_TILE_ELEMS) run the single-launch kernel on 1024-element tiles; one
    polynomial more and the call is back on 4096-element tiles.  Same words either way, forward and inverse, one launch."""
    bits = 64
    P = O.merge_params(logn, poly, bits)
    tab, itab = to_dev(P.fwd_br, bits), to_dev(P.inv_br, bits)
    try:
        capi.tune(6, 0)  # (2^12: keep the calls on the two-pass kernel under test, not on the one-tile path)
        for batch in batches:
            x = O.example_input(P.modulus, batch << logn, seed=logn * 17 + batch)
            want = _threaded_oracle(O.merge_ntt, x, P)
            for knob in (1 << 18, 0):
                capi.tune(7, knob)
                d = to_dev(x, bits)
                capi.ntt(d.view(batch, -1), tab, P.modulus, logn, poly)
                assert _launches() == 1
                torch.cuda.synchronize()
                assert (to_host(d, bits) == want).all(), (batch, knob)
                out = torch.zeros_like(d)
                capi.intt(d.view(batch, -1), itab, P.modulus, P.n_inv, logn, poly, out=out.view(batch, -1))
                assert _launches() == 1
                torch.cuda.synchronize()
                assert (to_host(out, bits) == x).all(), (batch, knob)
    finally:
        capi.tune(6, 1)
        capi.tune(7, 1 << 18)
