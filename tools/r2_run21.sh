#!/bin/bash
# round 2, GPU pass 21: A/B of the three-operand additions (base2 = the commit before), restored / new tests
mkdir -p gpurun_out
timeout 900 python tools/ab_cases.py --rounds 2 --cases c2fwd,c2inv,c3fwd,c3inv,c4fwd,small,big gpu_ntt_b200/lib/libgpuntt_b200_base2.so gpu_ntt_b200/lib/libgpuntt_b200.so > gpurun_out/ab_adds.jsonl 2> gpurun_out/ab_err.txt; tail -3 gpurun_out/ab_err.txt; cat gpurun_out/ab_adds.jsonl
timeout 1500 python -m pytest tests/test_merge_gpu.py tests/test_4step_gpu.py tests/test_moduli_gpu.py tests/test_fused_gpu.py -q -x 2>&1 | tail -4
