#!/bin/bash
# The last state-of-the-tree pass of round 2 (run under gpurun; everything lands in gpurun_out/): smoke, the whole GPU suite, bench.py
# (both arms), the ncu launch list of the bench command, one full capture of the C2 step (traffic), the randomised parity test with
# whatever time is left.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -1 gpurun_out/smoke.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -4 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench_err.txt; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench_err.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_err.txt; cut -c1-200 gpurun_out/bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/launches.csv | cut -c1-200
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
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fast_pass -s 4 -c 2 -o gpurun_out/final2_c2 -f python /tmp/one.py 16 1024 64 1 > gpurun_out/ncu1.log 2>&1; tail -1 gpurun_out/ncu1.log
python tools/ncu_summary.py gpurun_out/final2_c2.ncu-rep > gpurun_out/final2_c2_summary.txt 2>&1; head -4 gpurun_out/final2_c2_summary.txt | cut -c1-200
timeout 200 python tools/fuzz_parity.py ${FUZZ_SECONDS:-60} 11 0.15 > gpurun_out/fuzz_seed11.jsonl 2>&1; tail -c 400 gpurun_out/fuzz_seed11.jsonl
