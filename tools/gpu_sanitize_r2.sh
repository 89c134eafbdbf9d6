#!/bin/bash
# compute-sanitizer memcheck + racecheck over the round-2 additions: small-tile single-launch kernel (single modulus and RNS), one-tile
# rings, RNS on small rings, PerCoefficient on the tuned kernels, inverse last round (n^-1 folded, twiddle-1 butterflies)
mkdir -p gpurun_out
cat > /tmp/san_r2.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from gpu_ntt_b200 import capi
from oracle import oracle as O
from tests.gpu_util import to_dev, to_host
from tests.test_merge_gpu import run_fwd, run_inv, _rns_roundtrip, rns_primes
# small-tile single-launch kernel and the one-tile rings (knob 6: 2 = every call, 0 = never)
for knob6 in (0, 2):
    capi.tune(6, knob6)
    for bits, logn, batch in ((64, 12, 5), (64, 13, 8), (64, 14, 3), (32, 13, 6)):
        for poly in (O.X_N_minus, O.X_N_plus):
            P = O.merge_params(logn, poly, bits)
            x = O.example_input(P.modulus, batch << logn, seed=logn)
            want = O.merge_ntt(x, P)
            assert (run_fwd(x, P, bits, poly) == want).all()
            assert (run_inv(want, P, bits, poly) == x).all()
            print("ok merge", knob6, bits, logn, batch, poly, flush=True)
capi.tune(6, 1)
# RNS: small tiles (2^12, 2^13) and small rings
for bits, logn, batch, tops in ((64, 12, 8, (59, 61)), (64, 13, 6, (59, 59, 58)), (64, 9, 10, (59, 61)), (64, 11, 6, (59, 59, 59)), (32, 10, 14, (29, 25))):
    primes = [rns_primes(bits, logn, 1 + i, t)[i] for i, t in enumerate(tops)]
    _rns_roundtrip(bits, logn, batch, len(tops), primes)
    print("ok rns", bits, logn, batch, flush=True)
# PerCoefficient on the tuned kernels
for logh, w in ((6, 64), (8, 32), (9, 256)):
    P = O.merge_params(logh, O.X_N_plus, 64)
    h = 1 << logh
    x = O.example_input(P.modulus, h * w, seed=logh).reshape(h, w)
    want = O.merge_ntt(np.ascontiguousarray(x.T), P).reshape(w, h).T
    d = to_dev(x, 64)
    s = torch.cuda.current_stream().cuda_stream
    capi.merge_ntt(in_ptr=d.data_ptr(), out_ptr=d.data_ptr(), table_ptr=to_dev(P.fwd_br, 64).data_ptr(), n_power=logh, batch=w, element_bits=64,
                   direction=capi.FORWARD, reduction_poly=O.X_N_plus, layout=capi.PerCoefficient, modulus=P.modulus, stream=s)
    torch.cuda.synchronize()
    assert (to_host(d, 64).reshape(h, w) == want).all()
    capi.merge_ntt(in_ptr=d.data_ptr(), out_ptr=d.data_ptr(), table_ptr=to_dev(P.inv_br, 64).data_ptr(), n_power=logh, batch=w, element_bits=64,
                   direction=capi.INVERSE, reduction_poly=O.X_N_plus, layout=capi.PerCoefficient, modulus=P.modulus, mod_inverse=P.n_inv, stream=s)
    torch.cuda.synchronize()
    assert (to_host(d, 64).reshape(h, w) == x).all()
    print("ok percoef", logh, w, flush=True)
# inverse last round on two- and three-pass plans, small rings
for bits, logn, batch in ((64, 16, 2), (64, 17, 1), (64, 10, 4), (32, 12, 2), (32, 14, 2)):
    P = O.merge_params(logn, O.X_N_minus, bits)
    x = O.example_input(P.modulus, batch << logn, seed=3)
    assert (run_inv(O.merge_ntt(x, P), P, bits, O.X_N_minus) == x).all()
    print("ok inv", bits, logn, batch, flush=True)
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python /tmp/san_r2.py > gpurun_out/sanitizer_r2_$tool.txt 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|Traceback|assert" gpurun_out/sanitizer_r2_$tool.txt | head -10; grep -c "^ok " gpurun_out/sanitizer_r2_$tool.txt
done
