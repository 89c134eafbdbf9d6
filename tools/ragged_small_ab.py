#!/usr/bin/env python
"""Small rings with a batch that ends inside a 2048- / 4096-element chunk: the split (whole chunks on the one-launch tuned kernel +
ragged tail on the generic kernel) against the whole batch on the generic kernel (what such a call took before).  Timing only."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

from gpu_ntt_b200 import capi  # noqa: E402
from perf_configs import time_ms  # noqa: E402

for bits, logn, batch in ((64, 8, 262145), (64, 10, 65537), (64, 10, 1001), (64, 7, 33), (32, 10, 131073), (32, 12, 32769), (32, 9, 99)):
    p = 576460756061519873 if bits == 64 else 469762049
    dt = torch.int64 if bits == 64 else torch.int32
    x = torch.randint(0, p, (batch, 1 << logn), dtype=dt, device="cuda")
    tab = torch.randint(1, p, (1 << (logn - 1),), dtype=dt, device="cuda")
    tab[0] = 1
    res = {}
    for name, force in (("split", 0), ("generic", 1)):
        capi.lib().gpuntt_b200_force_generic_path(force)
        res[name] = round(time_ms(lambda: capi.ntt(x, tab, p, logn, 1), 20) * 1e3, 2)
        res[name + "_launches"] = capi.lib().gpuntt_b200_last_launch_count()
    capi.lib().gpuntt_b200_force_generic_path(0)
    print(json.dumps({"bits": bits, "logn": logn, "batch": batch, "us_split": res["split"], "launches_split": res["split_launches"],
                      "us_generic_whole_batch": res["generic"], "ratio": round(res["generic"] / res["split"], 3)}), flush=True)
