#!/usr/bin/env python
"""families.csv (ncu --csv metrics log of tools/kernel_families.py, long format) -> one block per library kernel launch.
usage: families_table.py gpurun_out/families.csv gpurun_out/families.log > profiles/r2_v4_kernel_families_ncu.txt"""
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
launches = {}
order = []
for r in rd:
    if len(r) < len(hdr):
        continue
    k = r[ix["ID"]]
    if k not in launches:
        launches[k] = {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]], "m": {}}
        order.append(k)
    launches[k]["m"][r[ix["Metric Name"]]] = (r[ix["Metric Value"]], r[ix["Metric Unit"]])

labels = []
if len(sys.argv) > 2:
    labels = [l.strip() for l in open(sys.argv[2]) if re.match(r"^\d+ ", l)]


def short(name):
    name = re.sub(r"gpuntt_b200::", "", name)
    name = re.sub(r"unsigned long", "u64", name)
    name = re.sub(r"unsigned int", "u32", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:150]


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return float("nan")


print("# one row block per kernel launch of tools/kernel_families.py under ncu (--clock-control none; serialised, cold caches:")
print("# durations are for shares, not for bench values).  torch's own kernels (random fills) are dropped.")
print("# cases in launch order:")
for l in labels:
    print("#   " + l)
print()
for k in order:
    L = launches[k]
    if "at::" in L["name"] or "elementwise" in L["name"]:
        continue
    m = L["m"]
    g = lambda key: num(m[key][0]) if key in m else float("nan")
    unit = lambda key: m[key][1] if key in m else ""
    dur = g("gpu__time_duration.sum")
    dur_us = dur / 1000.0 if unit("gpu__time_duration.sum") in ("ns", "nsecond") else dur
    rd_b, wr_b = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    rd_b *= scale.get(unit("dram__bytes_read.sum"), 1.0)
    wr_b *= scale.get(unit("dram__bytes_write.sum"), 1.0)
    conf, wav = g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"), g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
    print(f"{short(L['name'])}")
    print(f"    grid {L['grid']} block {L['block']} regs {g('launch__registers_per_thread'):.0f}  time {dur_us:9.1f} us"
          f"  dram rd {rd_b / 1e6:9.1f} MB wr {wr_b / 1e6:9.1f} MB ({g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} % of peak)")
    print(f"    inst {g('smsp__inst_executed.sum') / 1e6:8.1f} M  issue {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} %"
          f"  fmaheavy {g('sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f} %"
          f"  alu {g('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active'):.1f} %"
          f"  warps active {g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} %"
          f"  smem wavefronts {wav / 1e6:.2f} M, bank-conflict wavefronts {conf / 1e6:.2f} M ({100.0 * conf / wav if wav else 0:.1f} %)")
