"""ctypes/numpy front-end of the CPU oracle (oracle/ntt_oracle.c) and, when it has been
built, of the reference's own CPU code (oracle/_ref/libgpuntt_ref_cpu.so).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(gpu_ntt_b200) never imports this module.

All arrays are numpy uint64 regardless of the element width of the transform
(`width` = 32 or 64 only selects the reference's default parameter pools); callers
cast to uint32 for Data32 device buffers.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libntt_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libgpuntt_ref_cpu.so")

X_N_plus, X_N_minus = 0, 1  # nttparameters.cuh:32-36

_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_lib = None
_ref = None


def build(ref: bool = True) -> None:
    """Compile the C restatement (always) and oracle/_ref (only where /root/reference exists)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = C.CDLL(ORACLE_SO)
        L.ora_modulus.argtypes = [C.c_uint64, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.ora_barrett_mult.restype = C.c_uint64
        L.ora_barrett_mult.argtypes = [C.c_uint64] * 5 + [C.c_int]
        L.ora_mulmod.restype = C.c_uint64
        L.ora_mulmod.argtypes = [C.c_uint64] * 3
        L.ora_expmod.restype = C.c_uint64
        L.ora_expmod.argtypes = [C.c_uint64] * 3
        L.ora_modinv.restype = C.c_uint64
        L.ora_modinv.argtypes = [C.c_uint64] * 2
        L.ora_bitreverse.restype = C.c_int
        L.ora_bitreverse.argtypes = [C.c_int, C.c_int]
        L.ora_merge_params.argtypes = [C.c_int, C.c_int, C.c_int, _u64p]
        L.ora_power_table.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, _u64p]
        L.ora_bitrev_table.argtypes = [_u64p, C.c_uint64, _u64p]
        L.ora_merge_ntt.argtypes = [_u64p, C.c_int, C.c_uint64, _u64p, C.c_int]
        L.ora_merge_intt.argtypes = [_u64p, C.c_int, C.c_uint64, _u64p, C.c_int]
        L.ora_reduce_signed.restype = C.c_uint64
        L.ora_reduce_signed.argtypes = [C.c_int64, C.c_uint64]
        L.ora_centered.restype = C.c_int64
        L.ora_centered.argtypes = [C.c_uint64, C.c_uint64]
        L.ora_schoolbook.argtypes = [_u64p, _u64p, C.c_int, C.c_uint64, C.c_int, _u64p]
        L.ora_4step_params.restype = C.c_int
        L.ora_4step_params.argtypes = [C.c_int, C.c_int, C.c_int, _u64p]
        L.ora_4step_small_tables.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int,
                                             C.c_int, _u64p, _u64p]
        L.ora_4step_w_table.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, _u64p]
        L.ora_4step_ntt.argtypes = [_u64p, _u64p, C.c_int, C.c_int, C.c_uint64, _u64p, _u64p, _u64p]
        L.ora_4step_intt.argtypes = [_u64p, _u64p, C.c_int, C.c_int, C.c_uint64, _u64p, _u64p, _u64p,
                                     C.c_uint64]
        L.ora_4step_intt_first_transpose.argtypes = [_u64p, _u64p, C.c_int, C.c_int]
        L.ora_example_input.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, _u64p]
        L.ora_fold_hash.restype = C.c_uint64
        L.ora_fold_hash.argtypes = [_u64p, C.c_size_t]
        _lib = L
    return _lib


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref():
    """The reference's own CPU implementation (None when oracle/_ref was never built)."""
    global _ref
    if _ref is None and have_ref():
        R = C.CDLL(REF_SO)
        vp = C.c_void_p
        R.ref_merge_params.argtypes = [C.c_int, C.c_int, C.c_int, _u64p, vp, vp, vp, vp]
        R.ref_merge_transform.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _u64p, _u64p, C.c_int]
        R.ref_schoolbook.argtypes = [_u64p, _u64p, C.c_int, C.c_uint64, C.c_int, _u64p]
        R.ref_barrett_mult.restype = C.c_uint64
        R.ref_barrett_mult.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]
        R.ref_example_input.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, _u64p]
        R.ref_4step_new.restype = vp
        R.ref_4step_new.argtypes = [C.c_int, C.c_int, C.c_int]
        R.ref_4step_free.argtypes = [vp, C.c_int]
        R.ref_4step_scalars.argtypes = [vp, C.c_int, _u64p]
        R.ref_4step_table.restype = C.c_uint64
        R.ref_4step_table.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int]
        R.ref_4step_run.argtypes = [vp, C.c_int, C.c_int, _u64p, _u64p]
        R.ref_time_merge_ntt.restype = C.c_double
        R.ref_time_merge_ntt.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32]
        R.ref_time_merge_ntt_io.restype = C.c_double
        R.ref_time_merge_ntt_io.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u64p, _u64p]
        R.ref_hardware_threads.restype = C.c_int
        _ref = R
    return _ref


# --------------------------------------------------------------------------- helpers
def modulus(value: int, width: int = 64):
    b, m = C.c_uint64(), C.c_uint64()
    lib().ora_modulus(value, width, C.byref(b), C.byref(m))
    return int(b.value), int(m.value)


def bitrev_table(t: np.ndarray) -> np.ndarray:
    out = np.empty_like(t)
    lib().ora_bitrev_table(np.ascontiguousarray(t), t.size, out)
    return out


@dataclass
class MergeParams:
    """NTTParameters<T>(LOGN, poly) restated: nttparameters.cu:22-49."""
    logn: int
    poly: int
    width: int
    modulus: int = 0
    omega: int = 0
    psi: int = 0
    n_inv: int = 0
    root: int = 0
    inv_root: int = 0
    root_size: int = 0
    n: int = 0
    fwd: np.ndarray = field(default=None, repr=False)      # natural order
    inv: np.ndarray = field(default=None, repr=False)
    fwd_br: np.ndarray = field(default=None, repr=False)   # what the caller uploads
    inv_br: np.ndarray = field(default=None, repr=False)


def merge_params(logn: int, poly: int = X_N_minus, width: int = 64) -> MergeParams:
    s = np.zeros(8, dtype=np.uint64)
    lib().ora_merge_params(logn, poly, width, s)
    P = MergeParams(logn, poly, width, *[int(x) for x in s])
    P.fwd = np.empty(P.root_size, dtype=np.uint64)
    P.inv = np.empty(P.root_size, dtype=np.uint64)
    lib().ora_power_table(P.root, P.modulus, P.root_size, P.fwd)
    lib().ora_power_table(P.inv_root, P.modulus, P.root_size, P.inv)
    P.fwd_br = bitrev_table(P.fwd)
    P.inv_br = bitrev_table(P.inv)
    return P


def merge_ntt(x: np.ndarray, P: MergeParams) -> np.ndarray:
    """NTTCPU::ntt applied to each row of x (shape [B, n] or [n])."""
    a = np.array(x, dtype=np.uint64, copy=True, order="C")
    rows = a.reshape(-1, P.n)
    for r in rows:
        lib().ora_merge_ntt(r, P.logn, P.modulus, P.fwd, P.poly)
    return a


def merge_intt(x: np.ndarray, P: MergeParams) -> np.ndarray:
    a = np.array(x, dtype=np.uint64, copy=True, order="C")
    rows = a.reshape(-1, P.n)
    for r in rows:
        lib().ora_merge_intt(r, P.logn, P.modulus, P.inv, P.poly)
    return a


def reduce_signed(x: np.ndarray, p: int) -> np.ndarray:
    """modular_arith.cuh:372-385 on an int64 array."""
    x = np.asarray(x, dtype=np.int64)
    return np.where(x < 0, (x + np.int64(p)), x).astype(np.uint64)


def centered(x: np.ndarray, p: int) -> np.ndarray:
    """modular_arith.cuh:389-405."""
    x = np.asarray(x, dtype=np.uint64)
    return np.where(x > np.uint64(p >> 1), x.astype(np.int64) - np.int64(p), x.astype(np.int64))


def schoolbook(a, b, p, poly) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.empty_like(a)
    lib().ora_schoolbook(a, b, a.size, p, poly, out)
    return out


def example_input(p: int, count: int, seed: int = 0) -> np.ndarray:
    """std::mt19937(seed) + std::uniform_int_distribution<uint64_t>(0, p-1), drawn in
    sequence (example/ntt_merge/test_merge_ntt.cu:70-85); generated by <random> itself in
    oracle/example_input.cpp so the stream is the one the reference examples see."""
    out = np.empty(count, dtype=np.uint64)
    lib().ora_example_input(seed, p, count, out)
    return out


# --------------------------------------------------------------------------- 4-step
@dataclass
class FourStepParams:
    """NTTParameters4Step<T>(LOGN, poly) restated: nttparameters.cu:191-225."""
    logn: int
    poly: int
    width: int
    modulus: int = 0
    omega: int = 0
    psi: int = 0
    n_inv: int = 0
    root: int = 0
    inv_root: int = 0
    root_size: int = 0
    n: int = 0
    n1: int = 0
    n2: int = 0
    t1: np.ndarray = field(default=None, repr=False)   # natural order, n1/2
    t2: np.ndarray = field(default=None, repr=False)   # natural order, n2/2
    W: np.ndarray = field(default=None, repr=False)    # n entries
    t1_inv: np.ndarray = field(default=None, repr=False)
    t2_inv: np.ndarray = field(default=None, repr=False)
    W_inv: np.ndarray = field(default=None, repr=False)


def fourstep_params(logn: int, poly: int = X_N_minus, width: int = 64, inverse_tables: bool = True
                    ) -> FourStepParams:
    s = np.zeros(10, dtype=np.uint64)
    if lib().ora_4step_params(logn, poly, width, s) != 0:
        raise ValueError("4-step supports 12 <= logn <= 24")
    P = FourStepParams(logn, poly, width, *[int(x) for x in s])
    P.t1 = np.empty(P.n1 // 2, dtype=np.uint64)
    P.t2 = np.empty(P.n2 // 2, dtype=np.uint64)
    lib().ora_4step_small_tables(P.root, P.modulus, P.n, P.n1, P.n2, 0, P.t1, P.t2)
    P.W = np.empty(P.n, dtype=np.uint64)
    lib().ora_4step_w_table(P.root, P.modulus, P.n1, P.n2, 0, P.W)
    if inverse_tables:
        P.t1_inv = np.empty(P.n1 // 2, dtype=np.uint64)
        P.t2_inv = np.empty(P.n2 // 2, dtype=np.uint64)
        lib().ora_4step_small_tables(P.root, P.modulus, P.n, P.n1, P.n2, 1, P.t1_inv, P.t2_inv)
        P.W_inv = np.empty(P.n, dtype=np.uint64)
        lib().ora_4step_w_table(P.inv_root, P.modulus, P.n1, P.n2, 1, P.W_inv)
    return P


def fourstep_ntt(x: np.ndarray, P: FourStepParams) -> np.ndarray:
    a = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, P.n)
    out = np.empty_like(a)
    for i in range(a.shape[0]):
        lib().ora_4step_ntt(a[i], out[i], P.n1, P.n2, P.modulus, P.t1, P.t2, P.W)
    return out.reshape(np.shape(x))


def fourstep_intt(x: np.ndarray, P: FourStepParams) -> np.ndarray:
    a = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, P.n)
    out = np.empty_like(a)
    for i in range(a.shape[0]):
        lib().ora_4step_intt(a[i], out[i], P.n1, P.n2, P.modulus, P.t1_inv, P.t2_inv, P.W_inv, P.n_inv)
    return out.reshape(np.shape(x))


def fourstep_intt_first_transpose(x: np.ndarray, P: FourStepParams) -> np.ndarray:
    a = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, P.n)
    out = np.empty_like(a)
    for i in range(a.shape[0]):
        lib().ora_4step_intt_first_transpose(a[i], out[i], P.n1, P.n2)
    return out.reshape(np.shape(x))


def fold_hash(v: np.ndarray) -> int:
    v = np.ascontiguousarray(v, dtype=np.uint64).ravel()
    return int(lib().ora_fold_hash(v, v.size))
