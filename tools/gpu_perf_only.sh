#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/perf_configs.py > gpurun_out/perf_configs.jsonl 2> gpurun_out/perf_err.txt
tail -3 gpurun_out/perf_err.txt
