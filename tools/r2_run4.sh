#!/bin/bash
# round 2, GPU pass 4: whole -m gpu suite, bench (both arms), full fused A/B, per-config perf, ncu of fused C2 and the C4 kernels
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -8 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench_err.txt; cat gpurun_out/bench.json | cut -c1-1500; tail -3 gpurun_out/bench_err.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_err.txt; cut -c1-400 gpurun_out/bench_reference.json
timeout 600 python tools/r2_fused_ab.py > gpurun_out/fused_ab.jsonl 2> gpurun_out/fused_ab_err.txt; tail -3 gpurun_out/fused_ab_err.txt
timeout 900 python tools/perf_configs.py > gpurun_out/perf_configs.jsonl 2> gpurun_out/perf_err.txt; tail -3 gpurun_out/perf_err.txt
cat > /tmp/one.py <<'PY'
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
from gpu_ntt_b200 import capi
from gpu_ntt_b200.params import NTTParameters, X_N_minus
from perf_configs import dev
logn, batch, bits, fused = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
P = NTTParameters(logn, X_N_minus, bits)
tab = dev(P.gpu_root_of_unity_table_generator(P.forward_root_of_unity_table), bits)
x = torch.randint(0, P.modulus, (batch, 1 << logn), dtype=torch.int64 if bits == 64 else torch.int32, device='cuda')
capi.tune(capi.TUNE_FUSED_PASSES, fused)
for _ in range(4):
    capi.ntt(x, tab, P.modulus, logn, X_N_minus)
torch.cuda.synchronize()
PY
cat > /tmp/one4.py <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
from gpu_ntt_b200 import capi
from perf_configs import dev
logn, batch = 24, 16
n1, n2 = capi.fourstep_shape(logn)
p = 576460753175838721
rng = np.random.default_rng(1)
t1 = dev(rng.integers(1, p, n1 // 2, dtype=np.uint64), 64); t2 = dev(rng.integers(1, p, n2 // 2, dtype=np.uint64), 64)
t1[0] = 1; t2[0] = 1
w = torch.randint(0, p, (1 << logn,), dtype=torch.int64, device='cuda')
x = torch.randint(0, p, (batch, 1 << logn), dtype=torch.int64, device='cuda')
for _ in range(2):
    capi.fourstep_ntt(x, t1, t2, w, p, logn)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused2 -s 2 -c 1 -o gpurun_out/r2_fused3_c2 -f python /tmp/one.py 16 1024 64 2 > gpurun_out/ncu1.log 2>&1; tail -1 gpurun_out/ncu1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fast_pass|w_pairs" -s 4 -c 4 -o gpurun_out/r2_c4_kernels -f python /tmp/one4.py > gpurun_out/ncu4.log 2>&1; tail -1 gpurun_out/ncu4.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu_bench.log 2>&1
