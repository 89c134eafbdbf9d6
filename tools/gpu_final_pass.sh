#!/bin/bash
# A shorter state-of-the-tree pass than tools/gpu_full_pass.sh (no per-family ncu sweep, no full captures): smoke, the whole GPU suite,
# bench.py (both arms), per-configuration timings, the 4-step contracts, the same-box API bench, the A/B tools of the latest additions,
# the sanitizer over the newest shapes, ncu launch list of the bench command.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -1 gpurun_out/smoke.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -4 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench_err.txt; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench_err.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_err.txt; cut -c1-200 gpurun_out/bench_reference.json
timeout 200 python tools/ragged_small_ab.py > gpurun_out/ragged_small_ab.jsonl 2>&1; cat gpurun_out/ragged_small_ab.jsonl
timeout 600 python tools/perf_configs.py > gpurun_out/perf_configs.jsonl 2> gpurun_out/perf_err.txt; tail -3 gpurun_out/perf_err.txt
timeout 300 python tools/perf_4step.py > gpurun_out/perf_4step.jsonl 2> gpurun_out/perf_4step_err.txt; tail -3 gpurun_out/perf_4step_err.txt
timeout 200 python tools/perf_batch1.py > gpurun_out/perf_batch1.jsonl 2>&1
rm -f gpurun_out/api_b200.jsonl gpurun_out/api_ref.jsonl
for c in c2 c2inv c3 c4 sweep small latency; do timeout 200 tools/bin/api_bench_b200 b200 $c 2>&1 | grep "^{" >> gpurun_out/api_b200.jsonl; timeout 200 tools/bin/api_bench_reference reference $c 2>&1 | grep "^{" >> gpurun_out/api_ref.jsonl; done
python tools/same_box_table.py gpurun_out/api_b200.jsonl gpurun_out/api_ref.jsonl > gpurun_out/api_bench_same_box.txt; tail -1 gpurun_out/api_bench_same_box.txt
bash tools/gpu_sanitize_single_poly.sh
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/launches.csv | cut -c1-200
