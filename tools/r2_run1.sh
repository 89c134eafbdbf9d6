#!/bin/bash
# round 2, GPU pass 1: the fused two-pass kernels -- parity, A/B timing, ncu captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
timeout 900 python -m pytest tests/test_fused_gpu.py -x -q > gpurun_out/pytest_fused.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fused.txt; tail -5 gpurun_out/pytest_fused.txt
timeout 600 python tools/r2_fused_ab.py > gpurun_out/fused_ab.jsonl 2> gpurun_out/fused_ab_err.txt; tail -3 gpurun_out/fused_ab_err.txt; head -20 gpurun_out/fused_ab.jsonl
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_fused_gpu.py > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -5 gpurun_out/pytest_gpu.txt
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
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused2 -s 2 -c 1 -o gpurun_out/r2_fused_c2 -f python /tmp/one.py 16 1024 64 1 > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused2 -s 2 -c 1 -o gpurun_out/r2_fused_c3 -f python /tmp/one.py 14 4096 32 1 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fast_pass -s 4 -c 2 -o gpurun_out/r2_unfused_c3 -f python /tmp/one.py 14 4096 32 0 > gpurun_out/ncu3.log 2>&1; tail -2 gpurun_out/ncu3.log
