#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_4step_gpu.py -q -x > gpurun_out/pytest_4step.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_4step.txt; tail -12 gpurun_out/pytest_4step.txt
timeout 600 python tools/perf_configs.py --quick > gpurun_out/perf_quick.jsonl 2> gpurun_out/perf_err.txt; tail -3 gpurun_out/perf_err.txt; grep "4-step" gpurun_out/perf_quick.jsonl | cut -c1-420
timeout 2400 python -m pytest tests -m gpu -q -x --deselect tests/test_4step_gpu.py > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -8 gpurun_out/pytest_gpu.txt
