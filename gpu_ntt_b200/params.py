"""Host-side mirror of the reference's NTTParameters<T> (src/lib/common/nttparameters.cu:22-189 in
the reference tree): default prime pools, omega/psi, n^-1 and the bit-reversed power tables the
caller uploads.  Pure Python integers (off the hot path, O(N) modular multiplies); used by bench.py
and by tests as the producer of synthetic inputs -- NOT the oracle (oracle/ is the checker)."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

X_N_plus, X_N_minus = 0, 1


def bitreverse(index: int, n_power: int) -> int:
    r = 0
    for _ in range(n_power):
        r = (r << 1) | (index & 1)
        index >>= 1
    return r


def _bitrev_perm(lg: int) -> np.ndarray:
    idx = np.arange(1 << lg, dtype=np.int64)
    out = np.zeros_like(idx)
    for b in range(lg):
        out |= ((idx >> b) & 1) << (lg - 1 - b)
    return out


@dataclass
class NTTParameters:
    """NTTParameters<T>(LOGN, poly) / NTTParameters<T>(LOGN, NTTFactors, poly)."""
    logn: int
    poly_reduction: int = X_N_minus
    element_bits: int = 64
    modulus: int = 0
    omega: int = 0
    psi: int = 0
    n: int = 0
    n_inv: int = 0
    root_of_unity: int = 0
    inverse_root_of_unity: int = 0
    root_of_unity_size: int = 0
    forward_root_of_unity_table: np.ndarray = field(default=None, repr=False)
    inverse_root_of_unity_table: np.ndarray = field(default=None, repr=False)

    def __post_init__(self):
        logn = self.logn
        if self.modulus == 0:  # default pools (nttparameters.cu:84-142)
            if self.element_bits == 32:
                self.modulus = 469762049
                self.omega = pow(900, 1 << (25 - logn), self.modulus)
                self.psi = pow(30, 1 << (25 - logn), self.modulus)
            else:
                self.modulus = 576460756061519873
                self.omega = pow(229929041166717729, 1 << (28 - logn), self.modulus)
                self.psi = pow(4517306222, 1 << (28 - logn), self.modulus)
        p = self.modulus
        self.n = 1 << logn
        self.n_inv = pow(self.n, p - 2, p)
        minus = self.poly_reduction == X_N_minus
        self.root_of_unity = self.omega if minus else self.psi
        self.inverse_root_of_unity = pow(self.root_of_unity, p - 2, p)
        self.root_of_unity_size = self.n >> 1 if minus else self.n
        self.forward_root_of_unity_table = self._powers(self.root_of_unity)
        self.inverse_root_of_unity_table = self._powers(self.inverse_root_of_unity)

    def _powers(self, root: int) -> np.ndarray:
        out = np.empty(self.root_of_unity_size, dtype=np.uint64)
        acc, p = 1, self.modulus
        for i in range(self.root_of_unity_size):
            out[i] = acc
            acc = acc * root % p
        return out

    def gpu_root_of_unity_table_generator(self, table: np.ndarray) -> np.ndarray:
        """table[bitreverse(i, log2 size)] (nttparameters.cu:175-189); dtype follows element_bits."""
        lg = int(self.root_of_unity_size).bit_length() - 1
        out = np.ascontiguousarray(table[_bitrev_perm(lg)])
        return out.astype(np.uint32) if self.element_bits == 32 else out
