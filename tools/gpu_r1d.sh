#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench_err.txt
M=sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_alu.sum,smsp__inst_executed.sum,gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct,smsp__warp_issue_stalled_wait_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct,smsp__warp_issue_stalled_not_selected_per_warp_active.pct,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_barrier_per_warp_active.pct,smsp__warps_active.avg.per_cycle_active
timeout 600 ncu --metrics $M --clock-control none -k regex:fast_pass -c 2 --csv --log-file gpurun_out/fast_pass_pipes.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/ncu_pipes.log 2>&1
tail -15 gpurun_out/pytest_gpu.txt; cat gpurun_out/bench.json; tail -3 gpurun_out/bench_err.txt
