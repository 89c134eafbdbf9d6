#!/usr/bin/env python
"""api_bench JSON lines of the two libraries (tools/api_bench.cu run as `b200` and as `reference`) -> one comparison table.
usage: same_box_table.py gpurun_out/api_b200.jsonl gpurun_out/api_ref.jsonl > profiles/..._api_bench_same_box.txt"""
import json
import sys

rows_b = [json.loads(l) for l in open(sys.argv[1])]
rows_r = [json.loads(l) for l in open(sys.argv[2])]
key = lambda r: (r["case"], r["bits"], r["logn"], r["batch"], r.get("mod_count", 0), r["op"])  # noqa: E731
R = {key(r): r for r in rows_r}
print("# same caller source (tools/api_bench.cu) linked against the reference's kernels built for sm_100 and against this repository, same B200, CUDA events")
print("# parity of first/last polynomial vs NTTCPU checked in the run; latency-* rows: one call between two events (host enqueue + launch + kernels),")
print("# host = host time per call in a stream of 300 calls, stream = device-side time per call in that stream")
below = nlat = 0
for b in rows_b:
    r = R.get(key(b))
    if not r:
        continue
    ratio = r["ms"] / b["ms"]
    if b["case"].startswith("latency"):
        nlat += 1
        below += ratio < 1.0
        print(f"{b['case']:15s} bits={b['bits']} logN={b['logn']:2d} batch={b['batch']:5d} {b['op']:4s} ref {r['ms'] * 1e3:7.1f} us (host {r['host_us_per_call']:4.1f}, stream {r['stream_us_per_call']:6.1f})"
              f" | b200 {b['ms'] * 1e3:7.1f} us (host {b['host_us_per_call']:4.1f}, stream {b['stream_us_per_call']:6.1f}) | ratio {ratio:5.2f} | parity {r.get('parity_vs_NTTCPU')} {b.get('parity_vs_NTTCPU')}")
    else:
        print(f"{b['case']:24s} bits={b['bits']} logN={b['logn']:2d} batch={b['batch']:7d} {b['op']:8s} ref {r['ms']:8.4f} ms | b200 {b['ms']:8.4f} ms | ratio {ratio:5.2f}"
              f" | parity ref {r.get('parity_vs_NTTCPU')} b200 {b.get('parity_vs_NTTCPU')}")
print(f"# launch-bound rows below 1.0: {below} of {nlat}")
