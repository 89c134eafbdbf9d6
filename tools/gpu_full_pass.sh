#!/bin/bash
# state-of-the-tree pass: parity, smoke, bench, per-config perf, ncu launch list of the bench command, ncu --set full of the two hot kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -4 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench_err.txt; cat gpurun_out/bench.json; tail -3 gpurun_out/bench_err.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_err.txt; cat gpurun_out/bench_reference.json
timeout 900 python tools/perf_configs.py > gpurun_out/perf_configs.jsonl 2> gpurun_out/perf_err.txt; tail -3 gpurun_out/perf_err.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fast_pass -s 6 -c 2 -o gpurun_out/prof_fast_pass_final -f python bench.py --steps 2 --warmup 3 --quick > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
