#!/bin/bash
# round 2, GPU pass 20: RNS small rings on the tuned kernels, restored 4-step tests, enqueue lock -- parity
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_merge_gpu.py tests/test_4step_gpu.py -q -x -k "rns or 4step_errors or captured or lazy_range or per_coefficient" 2>&1 | tail -5
python - <<'PY'
import sys, json
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import numpy as np, torch
from gpu_ntt_b200 import capi
from perf_configs import dev, time_ms
from tests.test_merge_gpu import rns_primes
# RNS small rings: tuned vs generic
for logn, batch, mc in ((10, 65536, 4), (11, 32768, 4), (8, 262144, 4)):
    primes = [p for p, _ in rns_primes(64, logn, mc, 59)]
    n = 1 << logn
    rng = np.random.default_rng(2)
    tab = dev(np.concatenate([rng.integers(1, p, n, dtype=np.uint64) for p in primes]), 64)
    mods = dev(np.array([[p, p.bit_length(), 0] for p in primes], dtype=np.uint64).ravel(), 64)
    x = torch.randint(0, min(primes), (batch, n), dtype=torch.int64, device="cuda")
    def fn():
        capi.merge_ntt(in_ptr=x.data_ptr(), out_ptr=x.data_ptr(), table_ptr=tab.data_ptr(), n_power=logn, batch=batch, element_bits=64,
                       direction=capi.FORWARD, reduction_poly=0, mod_count=mc, modulus_dev=mods.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    res = {}
    for tag, g in (("tuned", 0), ("generic", 1)):
        capi.lib().gpuntt_b200_force_generic_path(g)
        res[tag] = round(time_ms(fn, 20), 4)
    capi.lib().gpuntt_b200_force_generic_path(0)
    print(json.dumps({"case": f"RNS x{mc} small ring 2^{logn} x {batch}", "ms_tuned": res["tuned"], "ms_generic": res["generic"]}), flush=True)
PY
