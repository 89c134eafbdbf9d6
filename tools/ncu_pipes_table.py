#!/usr/bin/env python
"""Compact table from an `ncu --metrics ... --csv` log: one row per kernel launch, selected pipe/stall metrics."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = None; data = collections.OrderedDict()
for r in rows:
    if len(r) > 10 and r[0] == "ID": hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        data.setdefault((d["ID"], d["Kernel Name"][:70]), {})[d["Metric Name"]] = d["Metric Value"]
short = {"gpu__time_duration.sum": "ns", "sm__cycles_elapsed.max": "cyc", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active": "fmah%",
         "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu%", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue%",
         "smsp__inst_executed.sum": "inst", "sm__inst_executed_pipe_fmaheavy.sum": "i_fmah", "sm__inst_executed_pipe_alu.sum": "i_alu",
         "smsp__warps_active.avg.per_cycle_active": "warps"}
for k, v in data.items():
    print(k[0], k[1])
    print("    " + "  ".join(f"{short[m]}={v[m]}" for m in short if m in v))
    st = {m.split("stalled_")[1].split("_per_warp")[0]: v[m] for m in v if "stalled" in m}
    print("    stalls: " + "  ".join(f"{a}={b}" for a, b in sorted(st.items(), key=lambda t: -float(t[1].replace(',', '') or 0))))
