#!/bin/bash
# compute-sanitizer passes over small cases of the tuned kernels (memcheck + racecheck); sanitizer slows kernels ~50x
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from gpu_ntt_b200 import capi
from oracle import oracle as O
for bits, logn, batch in ((64, 16, 3), (64, 13, 2), (32, 14, 3), (64, 17, 1)):
    P = O.merge_params(logn, O.X_N_minus, bits)
    x = O.example_input(P.modulus, batch << logn, seed=1)
    if bits == 64:
        d = torch.from_numpy(x.view(np.int64)).cuda(); tab = torch.from_numpy(P.fwd_br.view(np.int64)).cuda(); itab = torch.from_numpy(P.inv_br.view(np.int64)).cuda()
        back = lambda t: t.cpu().numpy().view(np.uint64)
    else:
        d = torch.from_numpy(x.astype(np.uint32).view(np.int32)).cuda(); tab = torch.from_numpy(P.fwd_br.astype(np.uint32).view(np.int32)).cuda(); itab = torch.from_numpy(P.inv_br.astype(np.uint32).view(np.int32)).cuda()
        back = lambda t: t.cpu().numpy().view(np.uint32).astype(np.uint64)
    capi.ntt(d.view(batch, -1), tab, P.modulus, logn, O.X_N_minus); torch.cuda.synchronize()
    assert (back(d) == O.merge_ntt(x, P)).all()
    capi.intt(d.view(batch, -1), itab, P.modulus, P.n_inv, logn, O.X_N_minus); torch.cuda.synchronize()
    assert (back(d) == x).all()
    print("ok", bits, logn, batch, flush=True)
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --kernel-regex kns=fast_pass python /tmp/san_case.py > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok |Error|hazard" gpurun_out/sanitizer_$tool.txt | head -20
done
# the RNS, ordered and 4-step entry points on the tuned kernels (small parametrisations of the parity tests)
timeout 1200 compute-sanitizer --tool memcheck --kernel-regex kns=fast_pass python -m pytest tests/test_merge_gpu.py tests/test_4step_gpu.py -m gpu -x -q \
    -k "test_rns_form_tuned_kernels and (64-12 or 64-13 or 32-14) or test_rns_ordered_entry_points and 12-8-4 or test_4step_fused_and_reference_contracts and (12-3 or 13-1 or 16-2) and 64" \
    > gpurun_out/sanitizer_memcheck_entry_points.txt 2>&1
echo "== memcheck over RNS / ordered / 4-step tests rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Error" gpurun_out/sanitizer_memcheck_entry_points.txt | head
