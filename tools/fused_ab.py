#!/usr/bin/env python
"""A/B of the single-launch two-pass kernels (merge_fused.cu) against one launch per pass, same box, same data:
C2, C3 and a ring-size sweep; lag sweep for C2/C3.  One JSON object per line.

    python tools/fused_ab.py [--quick]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch  # noqa: E402

from gpu_ntt_b200 import capi  # noqa: E402
from gpu_ntt_b200.params import NTTParameters, X_N_minus  # noqa: E402
from perf_configs import dev, peak, time_ms  # noqa: E402


def case(logn, batch, bits, iters, inverse=False, fused=2, lag=4):
    P = NTTParameters(logn, X_N_minus, bits)
    p = P.modulus
    tab = dev(P.gpu_root_of_unity_table_generator(P.inverse_root_of_unity_table if inverse else P.forward_root_of_unity_table), bits)
    dt = torch.int64 if bits == 64 else torch.int32
    x = torch.randint(0, p, (batch, 1 << logn), dtype=dt, device="cuda")
    capi.tune(capi.TUNE_FUSED_PASSES, fused)
    capi.tune(capi.TUNE_FUSED_LAG, lag)

    def fn():
        if inverse:
            capi.intt(x, tab, p, P.n_inv, logn, X_N_minus)
        else:
            capi.ntt(x, tab, p, logn, X_N_minus)
    ms = time_ms(fn, iters)
    launches = capi.lib().gpuntt_b200_last_launch_count()
    capi.tune(capi.TUNE_FUSED_PASSES, 1)
    capi.tune(capi.TUNE_FUSED_LAG, 4)
    gbs = 2 * (1 << logn) * (bits // 8) * batch / (ms * 1e-3) / 1e9
    print(json.dumps({"logn": logn, "batch": batch, "bits": bits, "op": "inv" if inverse else "fwd", "fused": fused, "lag": lag,
                      "launches": launches, "ms": round(ms, 4), "us": round(ms * 1e3, 2), "ntt_per_s": round(batch / (ms * 1e-3), 1),
                      "alg_GBps": round(gbs, 1), "frac_hbm": round(gbs / peak(), 4)}), flush=True)
    del x
    torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    it = 5 if args.quick else 20
    capi.lib()
    for fused in (2, 0):
        case(16, 1024, 64, it, fused=fused)
        case(16, 1024, 64, it, inverse=True, fused=fused)
        case(14, 4096, 32, it, fused=fused)
        case(14, 4096, 32, it, inverse=True, fused=fused)
    for lag in (2, 3, 5, 6):
        case(16, 1024, 64, it, lag=lag)
        case(14, 4096, 32, it, lag=lag)
    if args.quick:
        return
    for fused in (2, 0):
        for logn in (12, 13, 14, 15):
            case(logn, (1 << 26) >> logn, 64, it, fused=fused)
        for logn in (13, 15, 16, 17, 18):
            case(logn, (1 << 27) >> logn, 32, it, fused=fused)
        # small batches: the launch-bound regime
        for batch in (1, 4, 8, 32, 128):
            case(16, batch, 64, 50, fused=fused)
            case(13, batch, 64, 50, fused=fused)
            case(14, batch, 32, 50, fused=fused)


if __name__ == "__main__":
    main()
