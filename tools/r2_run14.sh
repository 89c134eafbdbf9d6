#!/bin/bash
# round 2, GPU pass 14: one-tile rings (parity + batch sweep against the two-pass plan), inverse A/B after the code-size fixes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_merge_gpu.py tests/test_fused_gpu.py tests/test_moduli_gpu.py -q -x 2>&1 | tail -3
timeout 600 python tools/one_tile_ab.py > gpurun_out/one_tile_ab.jsonl 2> gpurun_out/one_tile_err.txt; tail -3 gpurun_out/one_tile_err.txt; cat gpurun_out/one_tile_ab.jsonl
timeout 900 python tools/ab_cases.py --rounds 1 --cases c2inv,c4inv,big gpu_ntt_b200/lib/libgpuntt_b200_base.so gpu_ntt_b200/lib/libgpuntt_b200.so > gpurun_out/ab_inverse.jsonl 2> gpurun_out/ab_err.txt; tail -3 gpurun_out/ab_err.txt; cat gpurun_out/ab_inverse.jsonl
