#!/bin/bash
mkdir -p gpurun_out
echo "== pytest fast" ; timeout 900 python -m pytest tests/test_merge_gpu.py -m gpu -x -q -k "fast_path or c2 or golden" 2>&1 | tail -15
echo "== sanitizer (memcheck) on a small fast-path case"; timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_merge_gpu.py -m gpu -x -q -k "fast_path_and_generic and minus and 5" 2>&1 | tail -8
echo "== pytest all" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== bench" ; timeout 600 python bench.py --steps 50 --warmup 5 2>gpurun_out/bench_err.txt | tee gpurun_out/bench2.json
tail -3 gpurun_out/bench_err.txt
echo "== ncu launches" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches2.csv python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu_bench2.log 2>&1 ; tail -1 gpurun_out/ncu_bench2.log | cut -c1-300
echo "== ncu full" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:fast_pass -s 6 -c 2 -o gpurun_out/prof_fast_pass -f python bench.py --steps 2 --warmup 3 --quick > gpurun_out/ncu_full2.log 2>&1 ; tail -2 gpurun_out/ncu_full2.log
