#!/bin/bash
# round 2, GPU pass 3: fused kernel v3 (loader + storer threads) and the transposing 4-step forward path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_gpu.py -x -q > gpurun_out/pytest_fused.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fused.txt; tail -5 gpurun_out/pytest_fused.txt
timeout 900 python -m pytest tests/test_4step_gpu.py -x -q > gpurun_out/pytest_4step.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_4step.txt; tail -15 gpurun_out/pytest_4step.txt
timeout 600 python tools/r2_fused_ab.py --quick > gpurun_out/fused_ab.jsonl 2> gpurun_out/fused_ab_err.txt; tail -3 gpurun_out/fused_ab_err.txt; cat gpurun_out/fused_ab.jsonl | cut -c1-200
timeout 600 python tools/perf_configs.py --quick > gpurun_out/perf_quick.jsonl 2> gpurun_out/perf_err.txt; tail -3 gpurun_out/perf_err.txt; cut -c1-400 gpurun_out/perf_quick.jsonl
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
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused2 -s 2 -c 1 -o gpurun_out/r2_fused3_c3 -f python /tmp/one.py 14 4096 32 1 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
