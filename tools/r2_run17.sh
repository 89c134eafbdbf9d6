#!/bin/bash
# round 2, GPU pass 17: small-tile variant for launch-bound calls -- parity, latency table for knob settings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_fused_gpu.py tests/test_merge_gpu.py -q -x 2>&1 | tail -3
tools/bin/api_bench_reference reference latency 2>&1 | grep "^{" > gpurun_out/api_latency_ref.jsonl
tools/bin/api_bench_b200 b200 latency 2>&1 | grep "^{" > gpurun_out/api_latency_b200.jsonl
GPUNTT_TUNE="7=0" tools/bin/api_bench_b200 b200 latency 2>&1 | grep "^{" > gpurun_out/api_latency_b200_k12.jsonl
GPUNTT_TUNE="6=0" tools/bin/api_bench_b200 b200 latency 2>&1 | grep "^{" > gpurun_out/api_latency_b200_no1tile.jsonl
GPUNTT_TUNE="7=4194304,6=0" tools/bin/api_bench_b200 b200 latency 2>&1 | grep "^{" > gpurun_out/api_latency_b200_k10wide.jsonl
python - <<'PY'
import json
R=[json.loads(l) for l in open('gpurun_out/api_latency_ref.jsonl')]
for tag,f in (("default (small tiles <= 2^19 elements, one-tile <= 296)",'gpurun_out/api_latency_b200.jsonl'),("knob 7 = 0 (4096-element tiles)",'gpurun_out/api_latency_b200_k12.jsonl'),("knob 6 = 0 (no one-tile path)",'gpurun_out/api_latency_b200_no1tile.jsonl'),("small tiles <= 2^22 elements, no one-tile",'gpurun_out/api_latency_b200_k10wide.jsonl')):
    B=[json.loads(l) for l in open(f)]
    below=0
    print("#", tag)
    for b,r in zip(B,R):
        ratio=r['ms']/b['ms']; below+= ratio<1.0
        if b['logn']<=14: print(f"{b['case']:15s} logN={b['logn']} batch={b['batch']:4d} {b['op']}  ref {r['ms']*1e3:6.1f} us (stream {r['stream_us_per_call']:5.1f}) | b200 {b['ms']*1e3:6.1f} us (host {b['host_us_per_call']:4.1f}, stream {b['stream_us_per_call']:5.1f}) | ref/b200 {ratio:4.2f} {b['parity_vs_NTTCPU']}")
    print("# rows below 1.0:", below, "of", len(B))
PY
