#!/usr/bin/env python
"""A/B timing of library builds on the headline workload (C2): for each .so given on the command line, a fresh
process times 30 forward calls (CUDA events) and prints per-pass device times from the engine's profiling API, and
checks 4 polynomials against the oracle.   usage: ab_c2.py lib1.so lib2.so ..."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch
from gpu_ntt_b200 import capi
from gpu_ntt_b200.params import NTTParameters, X_N_minus
from oracle import oracle as O
lib = capi.lib()
P = NTTParameters(16, X_N_minus, 64); p = P.modulus
tab = torch.from_numpy(P.gpu_root_of_unity_table_generator(P.forward_root_of_unity_table).view(np.int64)).cuda()
x = torch.randint(0, p, (1024, 65536), dtype=torch.int64, device="cuda")
x0 = x[:4].clone()
capi.ntt(x, tab, p, 16, X_N_minus); torch.cuda.synchronize()
PO = O.merge_params(16, O.X_N_minus, 64)
ok = bool((x[:4].cpu().numpy().view(np.uint64) == O.merge_ntt(x0.cpu().numpy().view(np.uint64), PO)).all())
for _ in range(5): capi.ntt(x, tab, p, 16, X_N_minus)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30): capi.ntt(x, tab, p, 16, X_N_minus)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 30
lib.gpuntt_b200_set_profiling(1); capi.profile_read()
for _ in range(10): capi.ntt(x, tab, p, 16, X_N_minus)
recs = capi.profile_read()
p1 = [m for k, m in recs if k == 1]; p2 = [m for k, m in recs if k == 2]
print(json.dumps({"lib": os.path.basename(os.environ.get("GPUNTT_B200_LIB", "default")), "ok": ok, "ms": round(ms, 4), "MNTT_s": round(1024 / ms / 1e3, 4),
                  "pass1_ms": round(sum(p1) / max(1, len(p1)), 4), "pass2_ms": round(sum(p2) / max(1, len(p2)), 4)}))
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, GPUNTT_B200_LIB=os.path.abspath(lib))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or ("FAILED " + lib + "\n" + r.stderr[-800:]), flush=True)
