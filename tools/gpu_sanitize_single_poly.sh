#!/bin/bash
# compute-sanitizer memcheck + racecheck over the single-polynomial contiguous-pass shapes (64-bit 2048-element tiles, 32-bit
# 4096-element tiles; forward lazy / exact policy and inverse), every word against the oracle under the tool.
mkdir -p gpurun_out
cat > /tmp/san_sp.py <<'PY'
import sys
sys.path.insert(0, ".")
from gpu_ntt_b200 import capi
from oracle import oracle as O
from tests.test_merge_gpu import run_fwd, run_inv
for bits, logn, poly in ((64, 17, O.X_N_minus), (64, 18, O.X_N_plus), (32, 19, O.X_N_minus), (32, 19, O.X_N_plus)):
    P = O.merge_params(logn, poly, bits)
    x = O.example_input(P.modulus, 1 << logn, seed=logn)
    want = O.merge_ntt(x, P)
    for inplace in (True, False):
        assert (run_fwd(x, P, bits, poly, inplace=inplace) == want).all()
        assert capi.lib().gpuntt_b200_last_launch_count() == 3
        assert (run_inv(want, P, bits, poly, inplace=inplace) == x).all()
    print("ok single-polynomial tiles", bits, logn, poly, flush=True)
PY
for tool in memcheck racecheck; do
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python /tmp/san_sp.py > gpurun_out/sanitizer_single_poly_$tool.txt 2>&1
    echo "$tool rc=$?" >> gpurun_out/sanitizer_single_poly_$tool.txt
    tail -4 gpurun_out/sanitizer_single_poly_$tool.txt
done
