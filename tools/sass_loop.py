#!/usr/bin/env python
"""Opcode histogram of the hottest loop (largest backward-branch body) of each kernel matching a substring.
usage: sass_loop.py file kernel-substring [divisor]"""
import re, subprocess, sys, collections
f = sys.argv[1]; sub = sys.argv[2]; div = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["cuobjdump", "-sass", f], capture_output=True, text=True).stdout
kern = None; ins = {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: kern = m.group(1); ins[kern] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", line)
    if m and kern: ins[kern].append((int(m.group(1), 16), m.group(2), m.group(3)))
FMA = ("IMAD", "FFMA", "FMUL", "HFMA2", "FADD")
for k, lst in ins.items():
    if sub not in k: continue
    best = None
    for addr, op, rest in lst:
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", rest)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < addr and (best is None or addr - tgt > best[1] - best[0]): best = (tgt, addr)
    if not best: continue
    body = [(a, o) for a, o, r in lst if best[0] <= a <= best[1]]
    h = collections.Counter(o for a, o in body)
    fma = sum(c * (2.6 if ".WIDE" in o else 2.14 if ".HI" in o else 1) for o, c in h.items() if o.split(".")[0] in FMA)
    dp = sum(c for o, c in h.items() if o.split(".")[0] in ("DFMA", "DADD", "DMUL"))
    print(f"== {k[:70]}: loop {len(body)} instrs ({len(body)/div:.1f}/unit)  fma-pipe IMAD-eq {fma/div:.1f}  fp64 {dp/div:.1f}")
    print("   " + "  ".join(f"{o}:{c/div:.2f}" for o, c in h.most_common(30)))
