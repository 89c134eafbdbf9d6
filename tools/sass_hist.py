#!/usr/bin/env python
"""Histogram of SASS opcodes per kernel in a cubin/.so (cuobjdump -sass), grouped by issue pipe.
usage: sass_hist.py file [kernel-substring] [divisor]"""
import re, subprocess, sys, collections
FMA = ("IMAD", "FFMA", "FMUL", "HFMA2", "FADD", "DFMA")
ALU = ("IADD3", "IADD", "LOP3", "SHF", "ISETP", "SEL", "MOV", "PRMT", "VIMNMX", "IABS", "LEA", "PLOP3", "FSEL", "SGXT", "IMNMX", "VIADD", "LOP", "BMSK", "FLO", "POPC")
def pipe(op):
    b = op.split(".")[0]
    if b in FMA: return "fma"
    if b in ALU: return "alu"
    if b in ("LDG","STG","LDS","STS","LD","ST","LDSM","ATOMS","ATOMG","RED","LDL","STL","LDC","LDCU","SHFL","UBLKCP","LDGSTS","UTMALDG","UTMASTG"): return "lsu"
    if b.startswith("U"): return "uniform"
    return "other"
f = sys.argv[1]; sub = sys.argv[2] if len(sys.argv) > 2 else ""; div = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["cuobjdump", "-sass", f], capture_output=True, text=True).stdout
cur = None; hist = {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: cur = m.group(1); hist[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur: hist[cur][m.group(1)] += 1
for k, h in hist.items():
    if sub not in k: continue
    tot = sum(h.values()); pp = collections.Counter()
    for op, c in h.items(): pp[pipe(op)] += c
    print(f"== {k}: {tot} instrs  " + "  ".join(f"{p}={c/div:.1f}" for p, c in pp.most_common()))
    print("   " + "  ".join(f"{op}:{c/div:.1f}" for op, c in h.most_common(40)))

# estimated pipe cycles per SMSP warp-instruction on B200 (tools/microbench.cu): IMAD.WIDE/IMAD.HI 4, other fma-pipe 2, alu 2
def cycles(h):
    f = a = 0
    for op, c in h.items():
        pp = pipe(op)
        if pp == "fma":
            f += c * (4 if (".WIDE" in op or ".HI" in op) else 2)
        elif pp == "alu":
            a += c * 2
    return f, a
for k, h in hist.items():
    if sub not in k: continue
    f, a = cycles(h)
    print(f"## {k[:60]}: fma-pipe cycles {f/div:.1f}  alu-pipe cycles {a/div:.1f}  issue {sum(h.values())/div:.1f}")
