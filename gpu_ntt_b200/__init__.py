"""gpu_ntt_b200 -- B200-native (sm_100a) batched NTT/INTT engine behind the GPU-NTT API.

The product is the C-ABI shared library `lib/libgpuntt_b200.so` (sources in `csrc/`, ABI in
`include/gpuntt_b200.h`) plus the C++17 header mirror of the reference API in `include/gpuntt/`.
This Python package is only the ctypes plumbing the tests and bench.py use to call the C ABI
with torch-owned device memory.  There is no CPU fallback: importing `capi` without the built
library raises, and every call needs a CUDA device.
"""
from . import capi  # noqa: F401
from .capi import (GpuNttError, MergeDesc, lib, merge_ntt, ntt, intt, build_library,  # noqa: F401
                   FORWARD, INVERSE, X_N_plus, X_N_minus, PerPolynomial, PerCoefficient)
