#!/usr/bin/env python
"""Host<->device link probe with 1, 2, 4, ... GPUs moving data AT THE SAME TIME (one process per GPU, pinned host memory,
512 MiB each way per GPU, both directions concurrently -- the traffic pattern of bench.py's e2e leg).  Prints per-GPU and
aggregate GB/s per concurrency level; the aggregate ceiling is what bounds the multi-GPU e2e number.

    python tools/pcie_probe_multi.py            # levels 1, 2, 4, 8 up to the visible GPU count
"""
import json
import os
import sys
import time

import torch
import torch.multiprocessing as mp

N = 512 << 20


def worker(rank, world, start_evt, q, write_combined):
    torch.cuda.set_device(rank)
    h_in = torch.empty(N, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(N, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(N, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(N, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h, reps=4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps
    run(True, True, 2)
    q.put(("ready", rank))
    start_evt.wait()
    res = {}
    for name, a, b in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
        res[name] = N / run(a, b) / 1e9
        time.sleep(0.05)
    q.put(("res", rank, res))


def level(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    evt = ctx.Event()
    procs = [ctx.Process(target=worker, args=(r, world, evt, q, False)) for r in range(world)]
    for p in procs:
        p.start()
    for _ in range(world):
        q.get(timeout=300)
    evt.set()
    out = {}
    for _ in range(world):
        _, r, res = q.get(timeout=300)
        out[r] = res
    for p in procs:
        p.join(timeout=60)
    agg = {k: sum(v[k] for v in out.values()) for k in ("h2d", "d2h", "both")}
    return {"gpus": world, "per_gpu_GBps": {k: [round(out[r][k], 1) for r in sorted(out)] for k in ("h2d", "d2h", "both")},
            "aggregate_GBps": {"h2d_alone": round(agg["h2d"], 1), "d2h_alone": round(agg["d2h"], 1),
                               "each_direction_when_both": round(agg["both"], 1)}}


def main():
    n = torch.cuda.device_count()
    try:
        ncpu = len(os.sched_getaffinity(0))
    except Exception:
        ncpu = os.cpu_count()
    print(json.dumps({"visible_gpus": n, "host_cpus": ncpu, "bytes_each_way_per_gpu": N}))
    for world in (1, 2, 4, 8):
        if world > n:
            break
        print(json.dumps(level(world)), flush=True)


if __name__ == "__main__":
    main()
