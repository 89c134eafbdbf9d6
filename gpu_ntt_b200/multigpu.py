"""Batch-slice partition of a batched transform across GPUs: polynomials are independent, so rank g
of G owns a contiguous slice of the batch and there is NO collective on the data path.  The only communication is
the measurement plumbing (barrier + max-over-ranks of the device time), which works on any torch.distributed
backend -- nccl on GPUs, gloo in the CPU tests."""
from __future__ import annotations


def batch_slice(rank: int, world: int, batch: int, mod_count: int = 0) -> tuple[int, int]:
    """[begin, end) of the polynomials rank `rank` owns.  With RNS (mod_count > 0) polynomial b uses modulus
    b % mod_count (ntt.cu:613 of the reference), so slice boundaries fall on multiples of mod_count and every rank
    can call the engine with the same modulus array."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world")
    unit = mod_count if mod_count > 0 else 1
    groups = batch // unit
    if groups * unit != batch:
        raise ValueError("batch must be a multiple of mod_count")
    lo = (groups * rank) // world
    hi = (groups * (rank + 1)) // world
    return lo * unit, hi * unit


def aggregate_rate(units_local: int, ms_local: float, dist=None, device=None) -> tuple[float, float, int]:
    """Whole-job throughput: (units/s, max-over-ranks ms, total units).  `dist` = torch.distributed when initialised."""
    import torch
    if dist is None or not dist.is_initialized():
        return units_local / (ms_local * 1e-3), ms_local, units_local
    t = torch.tensor([ms_local], dtype=torch.float64, device=device)
    u = torch.tensor([units_local], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    ms, units = float(t.item()), int(u.item())
    return units / (ms * 1e-3), ms, units


def _slices_call(fn, whole, whole_device, parts, devices, batch, mod_count, streams):
    import ctypes as C
    from . import capi
    n = len(parts)
    ptrs = (C.c_void_p * n)(*[p.data_ptr() for p in parts])
    devs = (C.c_int * n)(*devices)
    sts = None if streams is None else (C.c_void_p * n)(*[getattr(s, "cuda_stream", s) for s in streams])
    poly_bytes = whole.numel() * whole.element_size() // batch
    capi.check(fn(whole.data_ptr(), whole_device, ptrs, devs, n, poly_bytes, batch, mod_count, sts))


def scatter_batch(src, parts, devices, mod_count: int = 0, streams=None) -> None:
    """gpuntt_b200_scatter_batch on torch tensors: slice g of src ([batch, N] on one GPU) -> parts[g] on devices[g]
    (batch_slice(g, len(parts), batch, mod_count) rows), peer copies over NVLink where available."""
    from . import capi
    _slices_call(capi.lib().gpuntt_b200_scatter_batch, src, src.device.index, parts, devices, src.shape[0], mod_count, streams)


def gather_batch(dst, parts, devices, mod_count: int = 0, streams=None) -> None:
    """gpuntt_b200_gather_batch: the reverse of scatter_batch."""
    from . import capi
    _slices_call(capi.lib().gpuntt_b200_gather_batch, dst, dst.device.index, parts, devices, dst.shape[0], mod_count, streams)
