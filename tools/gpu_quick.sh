#!/bin/bash
# quick loop: parity tests + bench (+ optional ncu of the fast path with NCU=1)
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
echo "== bench" ; timeout 600 python bench.py --steps 50 --warmup 5 2>gpurun_out/bench_err.txt | tee gpurun_out/bench.json
tail -3 gpurun_out/bench_err.txt
if [ -n "$NCU" ]; then
echo "== ncu launches" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu_bench.log 2>&1 ; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
echo "== ncu full" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:fast_pass -s 6 -c 2 -o gpurun_out/prof_fast_pass -f python bench.py --steps 2 --warmup 3 --quick > gpurun_out/ncu_full.log 2>&1 ; tail -2 gpurun_out/ncu_full.log
fi
