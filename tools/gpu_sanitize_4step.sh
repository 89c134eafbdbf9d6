#!/bin/bash
# compute-sanitizer memcheck + racecheck over the transposeless 4-step forms (reference-contract forward, both inverse contracts;
# position-major product kernel with one and with several twiddle ranges, transposing stores) -> profiles/r2_compute_sanitizer_4step.txt
mkdir -p gpurun_out
cat > /tmp/san_4step.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from gpu_ntt_b200 import capi
from oracle import oracle as O
from tests.gpu_util import to_dev, to_host
from tests.test_4step_gpu import tables, transposed
for logn, batch in ((12, 4), (16, 4), (17, 5)):
    P = O.fourstep_params(logn, O.X_N_minus, 64)
    n = P.n
    x = O.example_input(P.modulus, batch * n, seed=logn).reshape(batch, n)
    want = np.stack([O.fourstep_ntt(r, P) for r in x])
    t1, t2, W = tables(P, 64, False)
    xt = to_dev(transposed(x, P.n1, P.n2), 64); r = torch.zeros_like(xt)
    capi.fourstep_ntt(xt.view(batch, n), t1, t2, W, P.modulus, logn, io_contract=capi.FOURSTEP_REFERENCE, out=r.view(batch, n)); torch.cuda.synchronize()
    assert (transposed(to_host(r, 64), P.n1, P.n2).reshape(-1) == want.reshape(-1)).all()
    it1, it2, iW = tables(P, 64, True)
    y = to_dev(want, 64); o = torch.zeros_like(y)
    capi.fourstep_ntt(y.view(batch, n), it1, it2, iW, P.modulus, logn, direction=capi.INVERSE, mod_inverse=P.n_inv, out=o.view(batch, n)); torch.cuda.synchronize()
    assert (to_host(o, 64).reshape(-1) == x.reshape(-1)).all()
    pre = to_dev(np.stack([O.fourstep_intt_first_transpose(r_, P) for r_ in want]), 64); r2 = torch.zeros_like(pre)
    capi.fourstep_ntt(pre.view(batch, n), it1, it2, iW, P.modulus, logn, direction=capi.INVERSE, mod_inverse=P.n_inv, io_contract=capi.FOURSTEP_REFERENCE, out=r2.view(batch, n)); torch.cuda.synchronize()
    assert (transposed(to_host(r2, 64), P.n1, P.n2).reshape(-1) == x.reshape(-1)).all()
    print("ok", logn, batch, flush=True)
PY
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool python /tmp/san_4step.py > gpurun_out/sanitizer_4step_$tool.txt 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|Traceback|assert" gpurun_out/sanitizer_4step_$tool.txt | head -10; grep -c "^ok " gpurun_out/sanitizer_4step_$tool.txt
done
