#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -8 gpurun_out/pytest_gpu.txt
tools/bin/api_bench_b200 b200 latency > gpurun_out/api_latency_b200.txt 2>&1; tail -5 gpurun_out/api_latency_b200.txt
tools/bin/api_bench_reference reference latency > gpurun_out/api_latency_ref.txt 2>&1; tail -3 gpurun_out/api_latency_ref.txt
