#!/bin/bash
# round 2, GPU pass 13: A/B after keeping the last-round code out of non-final kernels; launch-bound table with host cost per call
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_4step_gpu.py tests/test_fused_gpu.py -q -x 2>&1 | tail -3
timeout 900 python tools/ab_cases.py --rounds 1 --cases c2inv,c3inv,c4inv,c4invref,small,big gpu_ntt_b200/lib/libgpuntt_b200_base.so gpu_ntt_b200/lib/libgpuntt_b200.so > gpurun_out/ab_inverse.jsonl 2> gpurun_out/ab_err.txt; tail -3 gpurun_out/ab_err.txt; cat gpurun_out/ab_inverse.jsonl
tools/bin/api_bench_b200 b200 latency 2>&1 | grep "^{" > gpurun_out/api_latency_b200.jsonl
tools/bin/api_bench_reference reference latency 2>&1 | grep "^{" > gpurun_out/api_latency_ref.jsonl
python - <<'PY'
import json
B=[json.loads(l) for l in open('gpurun_out/api_latency_b200.jsonl')]
R=[json.loads(l) for l in open('gpurun_out/api_latency_ref.jsonl')]
below=0
for b,r in zip(B,R):
    assert (b['case'],b['logn'],b['batch'],b['op'])==(r['case'],r['logn'],r['batch'],r['op'])
    ratio=r['ms']/b['ms']; below+= ratio<1.0
    print(f"{b['case']:15s} logN={b['logn']} batch={b['batch']:4d} {b['op']}  ref {r['ms']*1e3:6.1f} us (host {r['host_us_per_call']:5.1f}, stream {r['stream_us_per_call']:5.1f}) | b200 {b['ms']*1e3:6.1f} us (host {b['host_us_per_call']:5.1f}, stream {b['stream_us_per_call']:5.1f}) | ref/b200 {ratio:4.2f} | parity {b['parity_vs_NTTCPU']}")
print("# rows below 1.0:", below, "of", len(B))
PY
