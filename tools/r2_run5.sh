#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -8 gpurun_out/pytest_gpu.txt
gpu_ntt_b200/lib/gpu_merge_examples 13 3 > gpurun_out/examples.txt 2>&1; gpu_ntt_b200/lib/gpu_4step_examples 20 2 >> gpurun_out/examples.txt 2>&1; cat gpurun_out/examples.txt
