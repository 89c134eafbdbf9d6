"""ctypes binding of include/gpuntt_b200.h (the C ABI).  Device buffers are passed as raw
pointers (torch tensors' data_ptr()); nothing here computes anything on the CPU."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPUNTT_B200_LIB") or os.path.join(HERE, "lib", "libgpuntt_b200.so")  # override: A/B builds

# enum values == the reference's (nttparameters.cuh:19-36)
FORWARD, INVERSE = 0, 1
PerPolynomial, PerCoefficient = 0, 1
X_N_plus, X_N_minus = 0, 1

OK, ERR_N_POWER, ERR_LAYOUT, ERR_CUDA, ERR_ARGUMENT, ERR_UNSUPPORTED = range(6)


class GpuNttError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"gpuntt_b200 status {status}: {message}")
        self.status = status
        self.message = message


class MergeDesc(C.Structure):
    """struct gpuntt_b200_merge_desc"""
    _fields_ = [
        ("element_bits", C.c_int), ("is_signed", C.c_int), ("direction", C.c_int), ("n_power", C.c_int),
        ("ntt_layout", C.c_int), ("reduction_poly", C.c_int), ("batch_size", C.c_int), ("mod_count", C.c_int),
        ("in_", C.c_void_p), ("out", C.c_void_p), ("root_of_unity_table", C.c_void_p),
        ("modulus_value", C.c_uint64), ("mod_inverse_value", C.c_uint64),
        ("modulus_dev", C.c_void_p), ("mod_inverse_dev", C.c_void_p), ("stream", C.c_void_p),
        ("modulus_order_dev", C.c_void_p), ("poly_order_dev", C.c_void_p),
    ]


FOURSTEP_REFERENCE, FOURSTEP_FUSED = 0, 1


class FourStepDesc(C.Structure):
    """struct gpuntt_b200_4step_desc"""
    _fields_ = [
        ("element_bits", C.c_int), ("direction", C.c_int), ("n_power", C.c_int), ("batch_size", C.c_int),
        ("mod_count", C.c_int), ("io_contract", C.c_int),
        ("in_", C.c_void_p), ("out", C.c_void_p), ("n1_table", C.c_void_p), ("n2_table", C.c_void_p),
        ("w_table", C.c_void_p), ("modulus_value", C.c_uint64), ("mod_inverse_value", C.c_uint64),
        ("modulus_dev", C.c_void_p), ("mod_inverse_dev", C.c_void_p), ("stream", C.c_void_p),
    ]


def build_library(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into lib/libgpuntt_b200.so (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["bash", os.path.join(HERE, "build.sh")], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libgpuntt_b200.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout + out.stderr)
    return LIB_PATH


_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(gpu_ntt_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, u64, u32, i = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
        L.gpuntt_b200_merge_ntt.restype = i
        L.gpuntt_b200_merge_ntt.argtypes = [C.POINTER(MergeDesc)]
        L.gpuntt_b200_ntt_u64.restype = i
        L.gpuntt_b200_ntt_u64.argtypes = [vp, vp, vp, u64, i, i, i, vp]
        L.gpuntt_b200_intt_u64.restype = i
        L.gpuntt_b200_intt_u64.argtypes = [vp, vp, vp, u64, u64, i, i, i, vp]
        L.gpuntt_b200_ntt_u32.restype = i
        L.gpuntt_b200_ntt_u32.argtypes = [vp, vp, vp, u32, i, i, i, vp]
        L.gpuntt_b200_intt_u32.restype = i
        L.gpuntt_b200_intt_u32.argtypes = [vp, vp, vp, u32, u32, i, i, i, vp]
        L.gpuntt_b200_4step_ntt.restype = i
        L.gpuntt_b200_4step_ntt.argtypes = [C.POINTER(FourStepDesc)]
        L.gpuntt_b200_4step_shape.restype = i
        L.gpuntt_b200_4step_shape.argtypes = [i, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.gpuntt_b200_transpose.restype = i
        L.gpuntt_b200_transpose.argtypes = [i, vp, vp, i, i, i, i, vp]
        L.gpuntt_b200_merge_ntt_host.restype = i
        L.gpuntt_b200_merge_ntt_host.argtypes = [C.POINTER(MergeDesc), vp, C.c_size_t]
        L.gpuntt_b200_last_launch_count.restype = i
        L.gpuntt_b200_total_launch_count.restype = C.c_ulonglong
        L.gpuntt_b200_last_error.restype = C.c_char_p
        L.gpuntt_b200_release_workspaces.restype = None
        L.gpuntt_b200_describe_plan.restype = i
        L.gpuntt_b200_describe_plan.argtypes = [i, i, C.c_char_p, C.c_size_t]
        L.gpuntt_b200_version.restype = i
        L.gpuntt_b200_force_generic_path.restype = None
        L.gpuntt_b200_force_generic_path.argtypes = [i]
        L.gpuntt_b200_example_input.restype = None
        L.gpuntt_b200_example_input.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p]
        for f in (L.gpuntt_b200_scatter_batch, L.gpuntt_b200_gather_batch):
            f.restype = i
            f.argtypes = [C.c_void_p, i, C.POINTER(C.c_void_p), C.POINTER(i), i, C.c_size_t, C.c_longlong, i, C.POINTER(C.c_void_p)]
        L.gpuntt_b200_tune.restype = None
        L.gpuntt_b200_tune.argtypes = [i, i]
        L.gpuntt_b200_set_profiling.restype = None
        L.gpuntt_b200_set_profiling.argtypes = [i]
        L.gpuntt_b200_profile_read.restype = i
        L.gpuntt_b200_profile_read.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_int), i]
        _lib = L
    return _lib


TUNE_FUSED_PASSES, TUNE_FUSED_LAG = 1, 2
TUNE_SINGLE_POLY_TILES = 8


def tune(knob: int, value: int) -> None:
    """gpuntt_b200_tune: A/B knobs (results never depend on them)."""
    lib().gpuntt_b200_tune(knob, value)


def example_input(modulus: int, count: int, seed: int = 0):
    """The reference example drivers' input stream (std::mt19937(seed) + uniform_int_distribution<uint64_t>(0, p-1))
    as a numpy uint64 array."""
    import numpy as np
    out = np.empty(count, dtype=np.uint64)
    lib().gpuntt_b200_example_input(seed, modulus, count, out.ctypes.data)
    return out


def check(status: int) -> None:
    if status != OK:
        raise GpuNttError(status, lib().gpuntt_b200_last_error().decode())


def describe_plan(n_power: int, element_bits: int) -> str:
    buf = C.create_string_buffer(1024)
    lib().gpuntt_b200_describe_plan(n_power, element_bits, buf, 1024)
    return buf.value.decode()


def merge_ntt(*, in_ptr: int, out_ptr: int, table_ptr: int, n_power: int, batch: int, element_bits: int = 64,
              direction: int = FORWARD, reduction_poly: int = X_N_minus, layout: int = PerPolynomial,
              modulus: int = 0, mod_inverse: int = 0, is_signed: bool = False, mod_count: int = 0,
              modulus_dev: int = 0, mod_inverse_dev: int = 0, stream: int = 0, modulus_order_dev: int = 0,
              poly_order_dev: int = 0) -> None:
    """gpuntt_b200_merge_ntt with keyword arguments; pointers are integers (device addresses)."""
    d = MergeDesc(element_bits, int(is_signed), direction, n_power, layout, reduction_poly, batch, mod_count,
                  in_ptr, out_ptr, table_ptr, modulus, mod_inverse, modulus_dev or None, mod_inverse_dev or None,
                  stream or None, modulus_order_dev or None, poly_order_dev or None)
    check(lib().gpuntt_b200_merge_ntt(C.byref(d)))


def _stream_ptr(stream) -> int:
    if stream is None:
        import torch
        return torch.cuda.current_stream().cuda_stream
    return getattr(stream, "cuda_stream", stream)


def ntt(x, table, modulus: int, n_power: int, reduction_poly: int = X_N_minus, out=None, stream=None):
    """GPU_NTT / GPU_NTT_Inplace on torch CUDA tensors (uint64/int64 or uint32/int32 storage).
    x: [batch, N]; table: the bit-reversed root table on the same device. In place when out is None."""
    out = x if out is None else out
    bits = x.element_size() * 8
    merge_ntt(in_ptr=x.data_ptr(), out_ptr=out.data_ptr(), table_ptr=table.data_ptr(), n_power=n_power,
              batch=x.numel() >> n_power, element_bits=bits, direction=FORWARD, reduction_poly=reduction_poly,
              modulus=modulus, stream=_stream_ptr(stream))
    return out


def intt(x, inv_table, modulus: int, n_inv: int, n_power: int, reduction_poly: int = X_N_minus, out=None,
         stream=None):
    """GPU_INTT / GPU_INTT_Inplace on torch CUDA tensors."""
    out = x if out is None else out
    bits = x.element_size() * 8
    merge_ntt(in_ptr=x.data_ptr(), out_ptr=out.data_ptr(), table_ptr=inv_table.data_ptr(), n_power=n_power,
              batch=x.numel() >> n_power, element_bits=bits, direction=INVERSE, reduction_poly=reduction_poly,
              modulus=modulus, mod_inverse=n_inv, stream=_stream_ptr(stream))
    return out


def fourstep_shape(n_power: int):
    """(n1, n2) of the reference's matrix_dimention()."""
    a, b = C.c_int(0), C.c_int(0)
    check(lib().gpuntt_b200_4step_shape(n_power, C.byref(a), C.byref(b)))
    return a.value, b.value


def fourstep_ntt(x, n1_table, n2_table, w_table, modulus: int, n_power: int, *, direction: int = FORWARD,
                 mod_inverse: int = 0, io_contract: int = FOURSTEP_FUSED, out=None, stream=None, mod_count: int = 0,
                 modulus_dev: int = 0, mod_inverse_dev: int = 0):
    """GPU_4STEP_NTT (io_contract=FOURSTEP_REFERENCE) / the fused natural-order form on torch CUDA tensors."""
    out = x if out is None else out
    d = FourStepDesc(x.element_size() * 8, direction, n_power, x.numel() >> n_power, mod_count, io_contract,
                     x.data_ptr(), out.data_ptr(), n1_table.data_ptr(), n2_table.data_ptr(), w_table.data_ptr(),
                     modulus, mod_inverse, modulus_dev or None, mod_inverse_dev or None, _stream_ptr(stream))
    check(lib().gpuntt_b200_4step_ntt(C.byref(d)))
    return out


def transpose(x, out, row: int, col: int, n_power: int, stream=None):
    """GPU_Transpose: out[b][c * row + r] = x[b][r * col + c]."""
    check(lib().gpuntt_b200_transpose(x.element_size() * 8, x.data_ptr(), out.data_ptr(), row, col, n_power,
                                      x.numel() >> n_power, _stream_ptr(stream)))
    return out


def profile_read(max_records: int = 65536):
    """[(kind, ms), ...] for the launches recorded since the last call (see gpuntt_b200_set_profiling)."""
    ms = (C.c_float * max_records)()
    kind = (C.c_int * max_records)()
    n = lib().gpuntt_b200_profile_read(ms, kind, max_records)
    return [(kind[i], ms[i]) for i in range(n)]
