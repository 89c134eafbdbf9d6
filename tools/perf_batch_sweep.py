#!/usr/bin/env python
"""Throughput against batch size at N = 2^16 (Data64 forward): where the persistent kernels stop being latency-bound."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gpu_ntt_b200 import capi
p, logn = 576460756061519873, int(sys.argv[1]) if len(sys.argv) > 1 else 16
tab = torch.randint(1, p, (1 << (logn - 1),), dtype=torch.int64, device="cuda"); tab[0] = 1
capi.lib().gpuntt_b200_set_profiling(0)
for batch in (1, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048):
    x = torch.randint(0, p, (batch, 1 << logn), dtype=torch.int64, device="cuda")
    g = torch.cuda.CUDAGraph()
    capi.ntt(x, tab, p, logn, 1); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(10):
            capi.ntt(x, tab, p, logn, 1)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    capi.lib().gpuntt_b200_set_profiling(1); capi.profile_read()
    capi.ntt(x, tab, p, logn, 1)
    recs = capi.profile_read(); capi.lib().gpuntt_b200_set_profiling(0)
    print(json.dumps({"logn": logn, "batch": batch, "us_per_call": round(us, 2), "MNTT_per_s": round(batch / us, 4),
                      "pass_us": [round(m * 1e3, 1) for _, m in recs]}), flush=True)
