#!/bin/bash
# round 2, GPU pass 16: timeline of launch-bound calls inside the single-launch kernel; quick parity; latency table after the twiddle-build change
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_gpu.py tests/test_merge_gpu.py -q -x -k "fused or small or one_tile or config_c3 or forward_and_inverse" 2>&1 | tail -3
GPUNTT_B200_LIB=$PWD/gpu_ntt_b200/lib/libgpuntt_b200_timeline.so timeout 300 python tools/fused_timeline.py > gpurun_out/fused_timeline.txt 2>&1; cat gpurun_out/fused_timeline.txt
