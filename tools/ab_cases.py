#!/usr/bin/env python
"""A/B timing of library builds over a set of cases: for each .so on the command line a fresh process runs the cases
(tools/perf_configs.py helpers: CUDA events around whole calls + the engine's per-launch times) and prints one JSON line per
case, tagged with the library.  The libraries are visited round-robin `--rounds` times so that clock / thermal drift of the
box shows up as spread inside one library instead of as a difference between libraries.
usage: ab_cases.py [--rounds R] [--cases c2inv,c3inv,c4inv,c4invref,c2fwd,c3fwd,c4fwd,c4ref,small] lib1.so lib2.so ..."""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, io, contextlib
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tools"))
from gpu_ntt_b200 import capi
from gpu_ntt_b200.params import X_N_minus, X_N_plus
import perf_configs as pc
cases = sys.argv[1].split(",")
tag = os.path.basename(os.environ.get("GPUNTT_B200_LIB", "default"))
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    for c in cases:
        if c == "c2fwd": pc.merge_case("C2 fwd", 16, 1024, 64, X_N_minus, 20)
        elif c == "c2inv": pc.merge_case("C2 inv", 16, 1024, 64, X_N_minus, 20, inverse=True)
        elif c == "c2neginv": pc.merge_case("C2 negacyclic inv", 16, 1024, 64, X_N_plus, 20, inverse=True)
        elif c == "c3fwd": pc.merge_case("C3 fwd", 14, 4096, 32, X_N_minus, 20)
        elif c == "c3inv": pc.merge_case("C3 inv", 14, 4096, 32, X_N_minus, 20, inverse=True)
        elif c == "c4fwd": pc.fourstep_case("C4 fused fwd", 24, 16, 6, capi.FOURSTEP_FUSED)
        elif c == "c4ref": pc.fourstep_case("C4 reference fwd", 24, 16, 6, capi.FOURSTEP_REFERENCE)
        elif c == "c4inv": pc.fourstep_case("C4 fused inv", 24, 16, 6, capi.FOURSTEP_FUSED, inverse=True)
        elif c == "c4invref": pc.fourstep_case("C4 reference inv", 24, 16, 6, capi.FOURSTEP_REFERENCE, inverse=True)
        elif c == "small": 
            pc.merge_case("u64 2^10 inv", 10, 65536, 64, X_N_minus, 20, inverse=True)
            pc.merge_case("u32 2^12 inv", 12, 32768, 32, X_N_minus, 20, inverse=True)
        elif c == "big": pc.merge_case("u64 2^20 inv", 20, 64, 64, X_N_minus, 20, inverse=True)
for line in buf.getvalue().splitlines():
    try:
        d = json.loads(line)
    except ValueError:
        continue
    print(json.dumps({"lib": tag, "case": d["case"], "ms": d["ms"], "launches_kind_ms": d.get("launches_kind_ms")}), flush=True)
''' % (ROOT, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--rounds", type=int, default=2)
ap.add_argument("--cases", default="c2inv,c3inv,c4inv,c4invref")
ap.add_argument("libs", nargs="+")
args = ap.parse_args()
for r in range(args.rounds):
    for lib in args.libs:
        env = dict(os.environ, GPUNTT_B200_LIB=os.path.abspath(lib))
        p = subprocess.run([sys.executable, "-c", CHILD, args.cases], env=env, capture_output=True, text=True)
        print(p.stdout.strip() or ("FAILED " + lib + "\n" + p.stderr[-800:]), flush=True)
