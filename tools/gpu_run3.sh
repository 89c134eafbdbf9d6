#!/bin/bash
# Session 3, run A: arithmetic microbench + state of the tree (tests, bench, ncu of the fast path).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== arith_bench"; timeout 300 ./tools/arith_bench 2>&1 | tee gpurun_out/arith_bench.txt
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
echo "== bench" ; timeout 600 python bench.py --steps 50 --warmup 5 2>gpurun_out/bench_err.txt | tee gpurun_out/bench.json
tail -3 gpurun_out/bench_err.txt
echo "== ncu launches" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu_bench.log 2>&1 ; tail -1 gpurun_out/ncu_bench.log | cut -c1-300
echo "== ncu full" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:fast_pass -s 6 -c 2 -o gpurun_out/prof_fast_pass -f python bench.py --steps 2 --warmup 3 --quick > gpurun_out/ncu_full.log 2>&1 ; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
