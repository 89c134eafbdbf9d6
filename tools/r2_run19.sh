#!/bin/bash
# round 2, GPU pass 19: PerCoefficient on the tuned kernels -- parity + timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_merge_gpu.py -q -x -k "per_coefficient or signed" 2>&1 | tail -3
python - <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import perf_configs as pc
pc.percoef_case("PerCoefficient 2^9 x 131072 (X^N+1)", 9, 131072, 20)
pc.percoef_case("PerCoefficient 2^8 x 262144 (X^N+1)", 8, 262144, 20)
pc.percoef_case("PerCoefficient 2^6 x 1048576 (X^N+1)", 6, 1048576, 20)
pc.percoef_case("PerCoefficient 2^9 x 1024 (the reference example's shape)", 9, 1024, 20)
PY
