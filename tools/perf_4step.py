#!/usr/bin/env python
"""C4 and two smaller 4-step shapes, every (direction x I/O contract), with the transposeless forms (default) and with knob
4STEP_TRANSPOSED = 0 (the round-1 sequences with transpose kernels); per-launch times from the engine's profiling API."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from gpu_ntt_b200 import capi
import perf_configs as pc
it = 8
for knob in (1, 0):
    capi.tune(3, knob)
    tag = "" if knob else " [knob 3 = 0: transpose kernels]"
    pc.fourstep_case("C4 4-step fused" + tag, 24, 16, it, capi.FOURSTEP_FUSED)
    pc.fourstep_case("C4 4-step reference contract" + tag, 24, 16, it, capi.FOURSTEP_REFERENCE)
    pc.fourstep_case("C4 4-step inverse fused" + tag, 24, 16, it, capi.FOURSTEP_FUSED, inverse=True)
    pc.fourstep_case("C4 4-step inverse reference contract" + tag, 24, 16, it, capi.FOURSTEP_REFERENCE, inverse=True)
    pc.fourstep_case("4-step fused logN=20" + tag, 20, 64, it, capi.FOURSTEP_FUSED)
    pc.fourstep_case("4-step reference logN=20" + tag, 20, 64, it, capi.FOURSTEP_REFERENCE)
    pc.fourstep_case("4-step inverse fused logN=22" + tag, 22, 16, it, capi.FOURSTEP_FUSED, inverse=True)
    pc.fourstep_case("4-step inverse reference logN=20" + tag, 20, 64, it, capi.FOURSTEP_REFERENCE, inverse=True)
