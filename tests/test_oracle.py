"""CPU-only: pins the C restatement (oracle/ntt_oracle.c) against
  (1) tests/golden/golden.json -- generated from the reference's own CPU code, and
  (2) oracle/_ref itself, live, wherever that library has been built."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O


def _fold(v):
    return str(O.fold_hash(v))


def test_survey_kats():
    """KATs captured from the reference while surveying."""
    P = O.merge_params(12, O.X_N_minus, 64)
    assert (P.modulus, P.omega, P.psi, P.n_inv) == (576460756061519873, 337284104821690767,
                                                    454262583201274046, 576320018572247041)
    x = O.example_input(P.modulus, 4096)
    assert [int(v) for v in x[:3]] == [316369445348223535, 412278604473420992, 347469429231211259]
    y = O.merge_ntt(x, P)
    assert [int(v) for v in y[:3]] == [251762351468011716, 169607255692418897, 365207891856275132]
    assert O.fold_hash(y) == 15830362021633792089
    P = O.merge_params(16, O.X_N_minus, 64)
    assert (P.omega, P.psi, P.n_inv) == (214672953469051179, 520744425723368651, 576451959968440321)
    y = O.merge_ntt(O.example_input(P.modulus, 65536), P)
    assert O.fold_hash(y) == 2501385232060115022
    assert O.modulus(576460756061519873, 64) == (60, 4611685988362617019)


def test_merge_against_golden(golden):
    for g in golden["merge"]:
        P = O.merge_params(g["logn"], g["poly"], g["width"])
        assert str(P.modulus) == g["modulus"] and str(P.omega) == g["omega"] and str(P.psi) == g["psi"]
        assert str(P.n_inv) == g["n_inv"] and str(P.root) == g["root"] and str(P.inv_root) == g["inv_root"]
        assert P.root_size == g["root_size"]
        bit, mu = O.modulus(P.modulus, g["width"])
        assert bit == g["bit"] and str(mu) == g["mu"]
        assert _fold(P.fwd_br) == g["fwd_br_hash"] and _fold(P.inv_br) == g["inv_br_hash"]
        x = O.example_input(P.modulus, g["batch"] * P.n)
        assert _fold(x) == g["in_hash"] and [str(int(v)) for v in x[:4]] == g["in_head"]
        y = O.merge_ntt(x, P)
        z = O.merge_intt(x, P)
        assert _fold(y) == g["ntt_hash"] and [str(int(v)) for v in y[:4]] == g["ntt_head"]
        assert _fold(z) == g["intt_hash"] and [str(int(v)) for v in z[:4]] == g["intt_head"]
        if "ntt_full" in g:
            assert [str(int(v)) for v in y] == g["ntt_full"]
            assert [str(int(v)) for v in z] == g["intt_full"]
        assert (O.merge_intt(y, P) == x).all()


def test_fourstep_against_golden(golden):
    for g in golden["fourstep"]:
        if g["logn"] > 17:
            continue  # the 2^20 record is covered by the slow test below
        P = O.fourstep_params(g["logn"], O.X_N_minus, g["width"])
        assert str(P.modulus) == g["modulus"] and (P.n1, P.n2) == (g["n1"], g["n2"])
        assert str(P.n_inv) == g["n_inv"]
        assert _fold(P.t1) == g["n1_hash"] and _fold(P.t2) == g["n2_hash"] and _fold(P.W) == g["W_hash"]
        assert _fold(P.t1_inv) == g["n1_inv_hash"] and _fold(P.t2_inv) == g["n2_inv_hash"]
        assert _fold(P.W_inv) == g["W_inv_hash"]
        x = O.example_input(P.modulus, P.n)
        assert _fold(x) == g["in_hash"]
        y = O.fourstep_ntt(x, P)
        assert _fold(y) == g["ntt_hash"] and [str(int(v)) for v in y[:4]] == g["ntt_head"]
        z = O.fourstep_intt(x, P)
        assert _fold(z) == g["intt_hash"] and [str(int(v)) for v in z[:4]] == g["intt_head"]
        assert _fold(O.fourstep_intt_first_transpose(x, P)) == g["first_transpose_hash"]


def test_fourstep_2p20_against_golden(golden):
    g = [r for r in golden["fourstep"] if r["logn"] == 20 and r["width"] == 64][0]
    P = O.fourstep_params(20, O.X_N_minus, 64, inverse_tables=False)
    assert _fold(P.W) == g["W_hash"]
    y = O.fourstep_ntt(O.example_input(P.modulus, P.n), P)
    assert _fold(y) == g["ntt_hash"]
    assert O.fold_hash(y) == 6406522035338895268


def test_barrett_against_golden(golden):
    for g in golden["barrett"]:
        p, a, b = int(g["p"]), int(g["a"]), int(g["b"])
        bit, mu = O.modulus(p, g["width"])
        # literal restatement == the reference, including its off-by-one `bit` at 2^61-1
        assert O.lib().ora_barrett_mult(a, b, p, bit, mu, g["width"]) == int(g["r"])
        if bit == p.bit_length():  # wherever the reference's Barrett is valid it is the canonical product
            assert O.lib().ora_mulmod(a, b, p) == int(g["r"]) == (a * b) % p


def test_convolution_property():
    """cpu_merge_ntt_examples (example/ntt_merge/test_cpu_merge_ntt.cu:69-90):
    intt(ntt(a) o ntt(b)) == schoolbook(a*b mod X^N -+ 1)."""
    for width in (64, 32):
        for poly in (O.X_N_minus, O.X_N_plus):
            P = O.merge_params(8, poly, width)
            x = O.example_input(P.modulus, 2 * P.n, seed=3)
            a, b = x[:P.n], x[P.n:]
            fa, fb = O.merge_ntt(a, P), O.merge_ntt(b, P)
            prod = np.array([(int(u) * int(v)) % P.modulus for u, v in zip(fa, fb)], dtype=np.uint64)
            assert (O.merge_intt(prod, P) == O.schoolbook(a, b, P.modulus, poly)).all()


def test_signed_helpers():
    p = 469762049
    x = np.array([-5, 0, 7, -(p - 1)], dtype=np.int64)
    assert O.reduce_signed(x, p).tolist() == [p - 5, 0, 7, 1]
    assert O.centered(np.array([0, p >> 1, (p >> 1) + 1, p - 1], dtype=np.uint64), p).tolist() == \
        [0, p >> 1, (p >> 1) + 1 - p, -1]


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (only possible where /root/reference exists)")
def test_restatement_equals_reference_live():
    R = O.ref()
    for width in (64, 32):
        for poly in (0, 1):
            for logn in (3, 9, 13):
                P = O.merge_params(logn, poly, width)
                x = O.example_input(P.modulus, 2 * P.n, seed=11)
                xr = np.zeros_like(x)
                R.ref_example_input(11, P.modulus, x.size, xr)
                assert (x == xr).all()
                for inverse, fn in ((0, O.merge_ntt), (1, O.merge_intt)):
                    out = np.zeros_like(x)
                    R.ref_merge_transform(logn, poly, width, inverse, x, out, 2)
                    assert (out == fn(x, P)).all()
    h = R.ref_4step_new(14, 1, 64)
    P = O.fourstep_params(14, 1, 64)
    x = O.example_input(P.modulus, P.n, seed=5)
    for op, fn in ((0, O.fourstep_ntt), (1, O.fourstep_intt), (2, O.fourstep_intt_first_transpose)):
        out = np.zeros_like(x)
        R.ref_4step_run(h, 64, op, x, out)
        assert (out == fn(x, P)).all()
    R.ref_4step_free(h, 64)
