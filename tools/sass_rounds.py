#!/usr/bin/env python
"""Per-round instruction mix of a fast_pass_kernel instantiation: the consumer loop is cut at every
LDS-after-STS boundary (one segment per register round) and each segment is costed with the pipe model
tools/arith_bench.cu measured on B200 (IMAD.WIDE / IMAD.HI 4 cycles, other fma-pipe ops 2, alu ops 2 per warp
instruction per SM sub-partition).  usage: sass_rounds.py lib.so kernel-substring"""
import collections, re, subprocess, sys
f, sub = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", f], capture_output=True, text=True).stdout
ALU = ("IADD3", "IADD", "LOP3", "SHF", "ISETP", "SEL", "MOV", "PRMT", "VIMNMX", "LEA", "PLOP3", "VIADD", "LOP", "FSEL", "P2R", "R2P")
cur = None; ins = {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: cur = m.group(1); ins[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur: ins[cur].append(m.group(2))
for k, ops in ins.items():
    if sub not in k: continue
    print("==", k[:110], len(ops), "instrs")
    seg = []; cur = []; state = "ld"
    for op in ops:
        if op.startswith("LDS") and state == "st": seg.append(cur); cur = []; state = "ld"
        if op.startswith("STS"): state = "st"
        cur.append(op)
    seg.append(cur)
    for s in seg:
        h = collections.Counter(s)
        nm = h["IMAD"] / 4.0
        if nm < 8: continue
        fm = sum(c * (4 if (".WIDE" in o or ".HI" in o) else 2) for o, c in h.items() if o.split(".")[0] in ("IMAD", "HFMA2"))
        al = sum(c * 2 for o, c in h.items() if o.split(".")[0] in ALU)
        print(f"  {len(s)} instrs, {nm:.0f} muls: per mul issue {len(s)/nm:.1f} fmaheavy {fm/nm:.1f} alu {al/nm:.1f} :: " + " ".join(f"{o}:{c/nm:.2f}" for o, c in h.most_common(14)))
