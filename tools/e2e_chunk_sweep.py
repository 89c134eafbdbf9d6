#!/usr/bin/env python
"""End-to-end (pinned host in -> H2D -> NTT -> D2H -> pinned host out) throughput of gpuntt_b200_merge_ntt_host as a function
of its pipeline chunk size, on N ranks at once (torchrun), next to raw copies of the same bytes with the same concurrency:
tells whether the multi-GPU e2e number is bounded by the shared host link or by the pipeline shape.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/e2e_chunk_sweep.py
"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from gpu_ntt_b200 import capi  # noqa: E402
from gpu_ntt_b200.params import NTTParameters, X_N_minus  # noqa: E402

LOGN, BATCH = 16, 1024


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = capi.lib()
    P = NTTParameters(LOGN, X_N_minus, 64)
    h_tab = P.gpu_root_of_unity_table_generator(P.forward_root_of_unity_table)
    h_in = torch.randint(0, P.modulus, (BATCH, 1 << LOGN), dtype=torch.int64).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    d_buf = torch.empty_like(h_in, device="cuda")
    d_buf2 = torch.empty_like(h_in, device="cuda")
    stream = torch.cuda.current_stream()
    desc = capi.MergeDesc(64, 0, capi.FORWARD, LOGN, capi.PerPolynomial, X_N_minus, BATCH, 0, h_in.data_ptr(), h_out.data_ptr(), None,
                          P.modulus, 0, None, None, stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps=5):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        barrier()
        t = (time.perf_counter() - t0) / reps
        if world > 1:
            tt = torch.tensor([t], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        return t

    nbytes = h_in.numel() * 8
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def raw_copies():
        with torch.cuda.stream(s1):
            d_buf.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_buf2, non_blocking=True)
        torch.cuda.synchronize()
    t = timed(raw_copies)
    if rank == 0:
        print(json.dumps({"ranks": world, "what": "raw copies, 512 MiB each way per GPU at the same time", "ms": round(t * 1e3, 2),
                          "aggregate_GBps_each_way": round(world * nbytes / t / 1e9, 1), "ntt_per_s_if_link_bound": round(world * BATCH / t)}), flush=True)
    for mb in (8, 32, 128, 512):
        os.environ["GPUNTT_B200_HOST_CHUNK_MB"] = str(mb)
        t = timed(lambda: capi.check(lib.gpuntt_b200_merge_ntt_host(C.byref(desc), h_tab.ctypes.data, h_tab.size)))
        if rank == 0:
            print(json.dumps({"ranks": world, "what": "gpuntt_b200_merge_ntt_host", "chunk_MiB": mb, "ms": round(t * 1e3, 2),
                              "ntt_per_s": round(world * BATCH / t)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
