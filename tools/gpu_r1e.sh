#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -25 gpurun_out/pytest_gpu.txt
timeout 600 python tools/perf_configs.py > gpurun_out/perf_configs.jsonl 2> gpurun_out/perf_err.txt
cat gpurun_out/perf_configs.jsonl; tail -5 gpurun_out/perf_err.txt
