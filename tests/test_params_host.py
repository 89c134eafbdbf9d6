"""CPU-only: gpu_ntt_b200.params.NTTParameters (the Python mirror bench.py and the tools build their tables with) against the
oracle's NTTParameters<T> restatement, which tests/golden pins to the reference (nttparameters.cu:22-189)."""
import numpy as np
import pytest

from gpu_ntt_b200.params import NTTParameters, X_N_minus, X_N_plus, bitreverse
from oracle import oracle as O


@pytest.mark.parametrize("bits", [64, 32])
@pytest.mark.parametrize("poly", [X_N_minus, X_N_plus])
@pytest.mark.parametrize("logn", [1, 2, 5, 10, 14, 16])
def test_python_parameters_equal_the_oracle(bits, poly, logn):
    assert (X_N_plus, X_N_minus) == (O.X_N_plus, O.X_N_minus)
    P, Q = NTTParameters(logn, poly, bits), O.merge_params(logn, poly, bits)
    assert (P.modulus, P.omega, P.psi, P.n, P.n_inv, P.root_of_unity, P.inverse_root_of_unity, P.root_of_unity_size) == \
           (Q.modulus, Q.omega, Q.psi, Q.n, Q.n_inv, Q.root, Q.inv_root, Q.root_size)
    assert (P.forward_root_of_unity_table == Q.fwd).all() and (P.inverse_root_of_unity_table == Q.inv).all()
    assert (P.gpu_root_of_unity_table_generator(P.forward_root_of_unity_table).astype(np.uint64) == Q.fwd_br).all()
    assert (P.gpu_root_of_unity_table_generator(P.inverse_root_of_unity_table).astype(np.uint64) == Q.inv_br).all()


def test_bitreverse():
    assert [bitreverse(i, 3) for i in range(8)] == [0, 4, 2, 6, 1, 5, 3, 7]
    assert bitreverse(1234, 12) == int(format(1234, "012b")[::-1], 2)
