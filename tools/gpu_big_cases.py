#!/usr/bin/env python
"""One-off parity checks at the edges of the tuned plans (largest 32-bit rings, big batches); slower than the test suite."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from gpu_ntt_b200 import capi
from oracle import oracle as O
def dev(a, bits):
    return torch.from_numpy(a.view(np.int64)).cuda() if bits == 64 else torch.from_numpy(a.astype(np.uint32).view(np.int32)).cuda()
def host(t, bits):
    a = t.cpu().numpy()
    return a.view(np.uint64) if bits == 64 else a.view(np.uint32).astype(np.uint64)
for bits, logn, batch, poly in ((32, 25, 1, O.X_N_minus), (32, 26, 1, O.X_N_minus), (64, 24, 3, O.X_N_minus), (32, 18, 5, O.X_N_plus), (32, 13, 7, O.X_N_plus)):
    t0 = time.time()
    P = O.merge_params(logn, poly, bits)
    x = O.example_input(P.modulus, batch << logn, seed=logn)
    want = O.merge_ntt(x, P)
    d = dev(x, bits).view(batch, -1)
    capi.ntt(d, dev(P.fwd_br, bits), P.modulus, logn, poly); torch.cuda.synchronize()
    ok1 = bool((host(d, bits).ravel() == want.ravel()).all())
    capi.intt(d, dev(P.inv_br, bits), P.modulus, P.n_inv, logn, poly); torch.cuda.synchronize()
    ok2 = bool((host(d, bits).ravel() == x.ravel()).all())
    print("case", bits, logn, batch, poly, "fwd", ok1, "inv", ok2, "launches", capi.lib().gpuntt_b200_last_launch_count(), round(time.time() - t0, 1), "s", flush=True)
# big batch: C5's whole 8192-polynomial stream on one GPU, sample check + round trip
P = O.merge_params(16, O.X_N_minus, 64)
g = torch.Generator(device="cuda").manual_seed(5)
x = torch.randint(0, P.modulus, (8192, 1 << 16), dtype=torch.int64, device="cuda", generator=g)
x0 = x[[0, 1, 4095, 8191]].clone()
tab, itab = dev(P.fwd_br, 64), dev(P.inv_br, 64)
capi.ntt(x, tab, P.modulus, 16, O.X_N_minus); torch.cuda.synchronize()
ok = all(bool((host(x[b], 64) == O.merge_ntt(host(x0[i], 64), P)).all()) for i, b in enumerate((0, 1, 4095, 8191)))
capi.intt(x, itab, P.modulus, P.n_inv, 16, O.X_N_minus); torch.cuda.synchronize()
ok2 = bool((x[[0, 1, 4095, 8191]] == x0).all())
print("batch 8192 N=2^16: sample parity", ok, "round trip", ok2)
