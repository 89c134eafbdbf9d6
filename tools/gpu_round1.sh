#!/bin/bash
# First GPU session: smoke, parity tests, microbench, bench, ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
echo "== microbench" ; timeout 120 ./tools/microbench 2>&1 | tee gpurun_out/microbench.txt
echo "== bench" ; timeout 600 python bench.py --steps 50 --warmup 5 2>gpurun_out/bench_err.txt | tee gpurun_out/bench.json
tail -3 gpurun_out/bench_err.txt
echo "== ncu launches" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu_bench.log 2>&1 ; tail -2 gpurun_out/ncu_bench.log
echo "== ncu full" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:merge_pass -s 4 -c 2 -o gpurun_out/prof_merge_pass -f python bench.py --steps 2 --warmup 3 --quick > gpurun_out/ncu_full.log 2>&1 ; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
