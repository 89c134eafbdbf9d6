#!/bin/bash
# round-end pass on one B200: smoke, bench (both arms), per-config perf, ncu launch list of the bench command,
# ncu --set full of the two C2 kernels and of the single-pass small-ring kernel (N = 2^11, the 4th case of `api_bench small`)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -1 gpurun_out/smoke.txt
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench_err.txt; cat gpurun_out/bench.json; tail -3 gpurun_out/bench_err.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_err.txt; cat gpurun_out/bench_reference.json
timeout 900 python tools/perf_configs.py > gpurun_out/perf_configs.jsonl 2> gpurun_out/perf_err.txt; tail -3 gpurun_out/perf_err.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fast_pass -s 6 -c 2 -o gpurun_out/prof_fast_pass_v9 -f python bench.py --steps 2 --warmup 3 --quick > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fast_pass -s 47 -c 1 -o gpurun_out/prof_small_ring_v9 -f tools/bin/api_bench_b200 b200 small > gpurun_out/ncu_small.log 2>&1; tail -2 gpurun_out/ncu_small.log
