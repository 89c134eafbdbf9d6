#!/usr/bin/env python
"""Latency of single transforms (batch 1, the shape the reference's nvbench harness sweeps: bench_merge_ntt.cu:71-75),
tuned vs generic kernels.  Timing only (random tables)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from gpu_ntt_b200 import capi
p = 576460756061519873
for bits in (64, 32):
    pp = p if bits == 64 else 469762049
    for logn in (12, 14, 16, 18, 20, 22, 24):
        for batch in (1, 8):
            dt = torch.int64 if bits == 64 else torch.int32
            x = torch.randint(0, pp, (batch, 1 << logn), dtype=dt, device="cuda")
            tab = torch.randint(1, pp, (1 << (logn - 1),), dtype=dt, device="cuda")
            tab[0] = 1
            out = {}
            for label, force in (("tuned", 0), ("generic", 1)):
                capi.lib().gpuntt_b200_force_generic_path(force)
                for _ in range(5):
                    capi.ntt(x, tab, pp, logn, 1)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(50):
                    capi.ntt(x, tab, pp, logn, 1)
                e1.record(); torch.cuda.synchronize()
                out[label] = round(e0.elapsed_time(e1) / 50 * 1e3, 2)
            capi.lib().gpuntt_b200_force_generic_path(0)
            # the same call replayed from a CUDA graph (20 transforms per graph): launch overhead off the host's critical path
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(20):
                    capi.ntt(x, tab, pp, logn, 1)
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                g.replay()
            e1.record(); torch.cuda.synchronize()
            out["graph"] = round(e0.elapsed_time(e1) / 100 * 1e3, 2)
            print(json.dumps({"bits": bits, "logn": logn, "batch": batch, "us_tuned": out["tuned"], "us_tuned_cuda_graph": out["graph"],
                              "us_generic": out["generic"]}), flush=True)
