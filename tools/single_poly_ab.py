#!/usr/bin/env python
"""ONE polynomial of a large ring (64-bit 2^17..2^28, 32-bit 2^19..2^26, the shape of a ZK prover's transform): contiguous pass on 2048-element tiles
of that polynomial against the usual two-polynomial 4096-element tiles (knob GPUNTT_B200_TUNE_SINGLE_POLY_TILES); timing only
(random tables), one JSON line per (ring size, op)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

from gpu_ntt_b200 import capi  # noqa: E402
from perf_configs import time_ms  # noqa: E402

cases = [(64, l) for l in (17, 18, 20, 22, 24, 26, 28)] + [(32, l) for l in (19, 20, 22, 24, 26)]
if len(sys.argv) > 1:
    cases = [c for c in cases if c[0] == int(sys.argv[1])]
for bits, logn in cases:
    p = 576460756061519873 if bits == 64 else 469762049
    dt = torch.int64 if bits == 64 else torch.int32
    x = torch.randint(0, p, (1, 1 << logn), dtype=dt, device="cuda")
    tab = torch.randint(1, p, (1 << (logn - 1),), dtype=dt, device="cuda")
    tab[0] = 1
    for op in ("fwd", "inv"):
        fn = (lambda: capi.ntt(x, tab, p, logn, 1)) if op == "fwd" else (lambda: capi.intt(x, tab, p, 12345, logn, 1))
        res = {}
        for name, knob in (("two_poly_tiles", 0), ("single_poly_tiles", 1)):
            capi.tune(8, knob)
            res[name] = round(time_ms(fn, 20) * 1e3, 2)
        capi.tune(8, 1)
        print(json.dumps({"bits": bits, "logn": logn, "batch": 1, "op": op, "us_two_poly_tiles": res["two_poly_tiles"],
                          "us_single_poly_tiles": res["single_poly_tiles"], "ratio": round(res["two_poly_tiles"] / res["single_poly_tiles"], 3)}), flush=True)
    del x, tab
