#!/bin/bash
# The state-of-the-tree pass (run under gpurun; everything lands in gpurun_out/, the summaries are then copied to profiles/):
# smoke, the whole GPU suite, bench.py (both arms), per-configuration timings, the 4-step contracts, the same-box API bench incl. the
# launch-bound table, ncu launch list of the bench command, one ncu metrics row per kernel family, two full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -1 gpurun_out/smoke.txt
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -6 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench_err.txt; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench_err.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_err.txt; cut -c1-200 gpurun_out/bench_reference.json
timeout 900 python tools/perf_configs.py > gpurun_out/perf_configs.jsonl 2> gpurun_out/perf_err.txt; tail -3 gpurun_out/perf_err.txt
timeout 600 python tools/perf_4step.py > gpurun_out/perf_4step.jsonl 2> gpurun_out/perf_4step_err.txt; tail -3 gpurun_out/perf_4step_err.txt
rm -f gpurun_out/api_b200.jsonl gpurun_out/api_ref.jsonl
for c in c2 c2inv c3 c4 sweep small latency; do tools/bin/api_bench_b200 b200 $c 2>&1 | grep "^{" >> gpurun_out/api_b200.jsonl; tools/bin/api_bench_reference reference $c 2>&1 | grep "^{" >> gpurun_out/api_ref.jsonl; done
python tools/same_box_table.py gpurun_out/api_b200.jsonl gpurun_out/api_ref.jsonl > gpurun_out/api_bench_same_box.txt; tail -1 gpurun_out/api_bench_same_box.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu_bench.log 2>&1
timeout 1500 ncu --clock-control none --csv --log-file gpurun_out/families.csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active python tools/kernel_families.py > gpurun_out/families.log 2>&1; tail -3 gpurun_out/families.log
python tools/families_table.py gpurun_out/families.csv gpurun_out/families.log > gpurun_out/kernel_families_ncu.txt
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
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused2 -s 2 -c 1 -o gpurun_out/final_fused_c3 -f python /tmp/one.py 14 4096 32 1 > gpurun_out/ncu2.log 2>&1; tail -1 gpurun_out/ncu2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fast_pass -s 4 -c 2 -o gpurun_out/final_c2 -f python /tmp/one.py 16 1024 64 1 > gpurun_out/ncu1.log 2>&1; tail -1 gpurun_out/ncu1.log
