#!/bin/bash
# compute-sanitizer memcheck + racecheck over the single-pass small-ring kernels (two- and three-round shapes, both widths, both rings)
mkdir -p gpurun_out
cat > /tmp/san_small.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from gpu_ntt_b200 import capi
from oracle import oracle as O
for bits, logn, batch in ((64, 11, 3), (64, 10, 6), (64, 9, 12), (64, 8, 24), (64, 7, 48), (32, 12, 3), (32, 11, 6), (32, 10, 12), (32, 9, 24), (32, 8, 48)):
    for poly in (O.X_N_minus, O.X_N_plus):
        P = O.merge_params(logn, poly, bits)
        x = O.example_input(P.modulus, batch << logn, seed=1)
        if bits == 64:
            d = torch.from_numpy(x.view(np.int64)).cuda(); tab = torch.from_numpy(P.fwd_br.view(np.int64)).cuda(); itab = torch.from_numpy(P.inv_br.view(np.int64)).cuda()
            back = lambda t: t.cpu().numpy().view(np.uint64)
        else:
            d = torch.from_numpy(x.astype(np.uint32).view(np.int32)).cuda(); tab = torch.from_numpy(P.fwd_br.astype(np.uint32).view(np.int32)).cuda(); itab = torch.from_numpy(P.inv_br.astype(np.uint32).view(np.int32)).cuda()
            back = lambda t: t.cpu().numpy().view(np.uint32).astype(np.uint64)
        capi.ntt(d.view(batch, -1), tab, P.modulus, logn, poly); torch.cuda.synchronize()
        assert capi.lib().gpuntt_b200_last_launch_count() == 1
        assert (back(d) == O.merge_ntt(x, P)).all()
        capi.intt(d.view(batch, -1), itab, P.modulus, P.n_inv, logn, poly); torch.cuda.synchronize()
        assert capi.lib().gpuntt_b200_last_launch_count() == 1
        assert (back(d) == x).all()
        print("ok", bits, logn, batch, poly, flush=True)
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --kernel-regex kns=fast_pass python /tmp/san_small.py > gpurun_out/sanitizer_small_$tool.txt 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitizer_small_$tool.txt | head -20; grep -c "^ok " gpurun_out/sanitizer_small_$tool.txt
done
