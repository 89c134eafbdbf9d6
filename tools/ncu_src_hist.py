#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` export: executed warp-instructions and stall samples per opcode,
for the first kernel in the file.  usage: ncu_src_hist.py src.csv [launch_index]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
# split per kernel instance
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
s = starts[which]; e = starts[which + 1] if which + 1 < len(starts) else len(rows)
hdr = rows[s + 1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops = collections.Counter(); smp = collections.Counter(); tot = 0; tots = 0
for r in rows[s + 2:e]:
    if len(r) <= iex: continue
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
    if not m: continue
    op = m.group(1)
    n = int(r[iex] or 0); k = int(r[ismp] or 0)
    ops[op] += n; smp[op] += k; tot += n; tots += k
print(f"kernel: {rows[s][1][:90]}  total warp-instr {tot}  samples {tots}")
for op, n in ops.most_common(45):
    print(f"  {op:28s} {n:12d} {100*n/tot:6.2f}%   samples {100*smp[op]/max(1,tots):6.2f}%")
