#!/usr/bin/env python
"""One-tile rings (64-bit 2^12, 32-bit 2^13): whole-polynomial-in-a-tile kernel against the two-pass plan over batch sizes
(knob GPUNTT_B200_TUNE_ONE_TILE); one JSON line per (width, batch, op)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

from gpu_ntt_b200 import capi  # noqa: E402
from gpu_ntt_b200.params import NTTParameters, X_N_minus  # noqa: E402
from perf_configs import dev, time_ms  # noqa: E402

for bits, logn in ((64, 12), (32, 13)):
    P = NTTParameters(logn, X_N_minus, bits)
    tab = dev(P.gpu_root_of_unity_table_generator(P.forward_root_of_unity_table), bits)
    itab = dev(P.gpu_root_of_unity_table_generator(P.inverse_root_of_unity_table), bits)
    for batch in (1, 8, 32, 128, 296, 592, 1184, 4096, 16384):
        x = torch.randint(0, P.modulus, (batch, 1 << logn), dtype=torch.int64 if bits == 64 else torch.int32, device="cuda")
        for op in ("fwd", "inv"):
            fn = (lambda: capi.ntt(x, tab, P.modulus, logn, X_N_minus)) if op == "fwd" else (lambda: capi.intt(x, itab, P.modulus, P.n_inv, logn, X_N_minus))
            res = {}
            for name, knob in (("two_pass", 0), ("one_tile", 2)):
                capi.tune(6, knob)
                res[name] = round(time_ms(fn, 30) * 1e3, 2)
            capi.tune(6, 1)
            print(json.dumps({"bits": bits, "logn": logn, "batch": batch, "op": op, "us_two_pass_plan": res["two_pass"], "us_one_tile": res["one_tile"],
                              "ratio": round(res["two_pass"] / res["one_tile"], 3)}), flush=True)
