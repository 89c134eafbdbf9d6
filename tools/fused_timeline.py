#!/usr/bin/env python
"""Timeline of a launch-bound call inside the single-launch two-pass kernel (lab build with -DGPUNTT_TIMELINE, see merge_fused.cu):
SM cycle counter at the hand-off points of every CTA's first tile, in microseconds from the CTA's own entry, plus the entry
skew between CTAs from the global timer.  usage: GPUNTT_B200_LIB=gpu_ntt_b200/lib/libgpuntt_b200_timeline.so fused_timeline.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

from gpu_ntt_b200 import capi  # noqa: E402
from gpu_ntt_b200.params import NTTParameters, X_N_minus, X_N_plus  # noqa: E402
from perf_configs import dev  # noqa: E402

lib = capi.lib()
lib.gpuntt_b200_timeline_read.argtypes = [C.POINTER(C.c_longlong), C.c_int]
NAMES = ["entry", "init done", "loader: at tile 0", "loader: dependency met, TMA issued", "consumers: twiddles built", "consumers: tile landed",
         "consumers: rounds done", "storer: saw done", "storer: smem read by TMA", "storer: store complete + signalled", "storer: all stores complete",
         "CTA: all roles done", "exit"]
mhz = 1965.0
capi.tune(6, 0)  # (the one-tile path would bypass the kernel under study)
capi.tune(7, int(os.environ.get('SMALL_TILE_ELEMS', 1 << 18)))
for bits, logn, batch, poly in ((64, 12, 8, X_N_plus), (64, 13, 8, X_N_plus), (64, 14, 8, X_N_plus), (32, 14, 8, X_N_minus), (64, 16, 8, X_N_plus)):
    P = NTTParameters(logn, poly, bits)
    tab = dev(P.gpu_root_of_unity_table_generator(P.forward_root_of_unity_table), bits)
    x = torch.randint(0, P.modulus, (batch, 1 << logn), dtype=torch.int64 if bits == 64 else torch.int32, device="cuda")
    for _ in range(3):
        capi.ntt(x, tab, P.modulus, logn, poly)
    torch.cuda.synchronize()
    buf = (C.c_longlong * (512 * 16))()
    lib.gpuntt_b200_timeline_read(buf, 1)
    capi.ntt(x, tab, P.modulus, logn, poly)
    assert lib.gpuntt_b200_timeline_read(buf, 1) > 0
    rows = [[buf[c * 16 + k] for k in range(16)] for c in range(512)]
    live = [c for c in range(512) if rows[c][0] != 0]
    g0 = min(rows[c][15] for c in live)
    print(f"== {bits}-bit N=2^{logn} batch {batch}: {len(live)} CTAs (us from each CTA's entry; 'skew' = its entry after the first CTA's, global timer)")
    # first-pass CTAs signal (slot 9 set), second-pass CTAs do not
    for role, pick in (("first-pass CTAs", lambda c: rows[c][9] != 0), ("second-pass CTAs", lambda c: rows[c][9] == 0)):
        cs = [c for c in live if pick(c)]
        if not cs:
            continue
        print(f"  {role}: {len(cs)}")
        skew = [(rows[c][15] - g0) / 1e3 for c in cs]
        print(f"    {'entry skew':44s} min {min(skew):6.2f}  median {sorted(skew)[len(skew) // 2]:6.2f}  max {max(skew):6.2f}")
        for k in range(1, 13):
            v = [(rows[c][k] - rows[c][0]) / mhz for c in cs if rows[c][k] != 0]
            if v:
                print(f"    {NAMES[k]:44s} min {min(v):6.2f}  median {sorted(v)[len(v) // 2]:6.2f}  max {max(v):6.2f}")
