#!/bin/bash
# round 2, GPU pass 15: launch-bound table (C++ caller) with the one-tile path on / off, kernel durations under ncu for both libraries
mkdir -p gpurun_out
tools/bin/api_bench_reference reference latency 2>&1 | grep "^{" > gpurun_out/api_latency_ref.jsonl
GPUNTT_B200_ONE_TILE_BATCH=296 tools/bin/api_bench_b200 b200 latency 2>&1 | grep "^{" > gpurun_out/api_latency_b200.jsonl
GPUNTT_B200_ONE_TILE_BATCH=0 tools/bin/api_bench_b200 b200 latency 2>&1 | grep "^{" > gpurun_out/api_latency_b200_two_pass.jsonl
python - <<'PY'
import json
R=[json.loads(l) for l in open('gpurun_out/api_latency_ref.jsonl')]
for tag,f in (("one-tile<=296",'gpurun_out/api_latency_b200.jsonl'),("two-pass",'gpurun_out/api_latency_b200_two_pass.jsonl')):
    B=[json.loads(l) for l in open(f)]
    below=0
    print("#", tag)
    for b,r in zip(B,R):
        ratio=r['ms']/b['ms']; below+= ratio<1.0
        if b['logn']<=13: print(f"{b['case']:15s} logN={b['logn']} batch={b['batch']:4d} {b['op']}  ref {r['ms']*1e3:6.1f} us (host {r['host_us_per_call']:5.1f}, stream {r['stream_us_per_call']:5.1f}) | b200 {b['ms']*1e3:6.1f} us (host {b['host_us_per_call']:5.1f}, stream {b['stream_us_per_call']:5.1f}) | ref/b200 {ratio:4.2f}")
    print("# rows below 1.0:", below, "of", len(B))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/lat_ncu_b200.csv tools/bin/api_bench_b200 b200 latency > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/lat_ncu_ref.csv tools/bin/api_bench_reference reference latency > /dev/null 2>&1
python - <<'PY'
import csv,re
for f in ('gpurun_out/lat_ncu_b200.csv','gpurun_out/lat_ncu_ref.csv'):
    rows=[r for r in csv.reader(l for l in open(f) if l.startswith('"'))]
    h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value"); ig=h.index("Grid Size")
    print("#",f)
    seen=0
    for r in rows[1:]:
        name=re.sub(r"\(.*","",r[ik])[:70]
        print(f"  {name:70s} grid {r[ig]:>14s} {float(r[iv].replace(',',''))/1e3:8.2f} us")
        seen+=1
        if seen>=60: break
PY
