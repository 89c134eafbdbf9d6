#!/bin/bash
# round-1 (session 4) GPU pass: parity, bench, per-config perf, lab microbench, ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
for v in 0 8 11 14 15; do ./tools/bin/bfly_lab_v$v; done > gpurun_out/bfly_lab.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench_err.txt
timeout 600 python tools/perf_configs.py > gpurun_out/perf_configs.jsonl 2> gpurun_out/perf_err.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/pytest_gpu.txt; cat gpurun_out/bfly_lab.txt; cat gpurun_out/bench.json; cat gpurun_out/perf_configs.jsonl; tail -5 gpurun_out/perf_err.txt
