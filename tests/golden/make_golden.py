"""Regenerates tests/golden/golden.json from the REFERENCE's own CPU code
(oracle/_ref/libgpuntt_ref_cpu.so, compiled from /root/reference by oracle/Makefile).
Run in the build container only:  python tests/golden/make_golden.py

Every record pins: the NTTParameters / NTTParameters4Step scalars, a hash of the
bit-reversed tables the caller uploads, and NTTCPU / NTT_4STEP_CPU outputs on the
example drivers' seeded input (std::mt19937(0), uniform_int_distribution<u64>(0,p-1)).
hash = fold(h = h*1000003 + v mod 2^64).
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def fold(v):
    return str(O.fold_hash(np.ascontiguousarray(v, dtype=np.uint64)))


def main():
    R = O.ref()
    assert R is not None, "build oracle/_ref first (make -C oracle ref)"
    recs = {"merge": [], "fourstep": [], "barrett": []}
    for width in (64, 32):
        for poly in (1, 0):
            for logn in (1, 2, 3, 4, 5, 8, 10, 11, 12, 14, 16, 17):
                n = 1 << logn
                scal = np.zeros(10, dtype=np.uint64)
                size = n // 2 if poly == 1 else n
                fwd_br = np.zeros(size, dtype=np.uint64)
                inv_br = np.zeros(size, dtype=np.uint64)
                R.ref_merge_params(logn, poly, width, scal, None, None,
                                   fwd_br.ctypes.data_as(C.c_void_p), inv_br.ctypes.data_as(C.c_void_p))
                p = int(scal[0])
                batch = 2 if logn <= 12 else 1
                x = np.zeros(batch * n, dtype=np.uint64)
                R.ref_example_input(0, p, x.size, x)
                y = np.zeros_like(x)
                R.ref_merge_transform(logn, poly, width, 0, x, y, batch)
                z = np.zeros_like(x)
                R.ref_merge_transform(logn, poly, width, 1, x, z, batch)
                rec = dict(width=width, poly=poly, logn=logn, batch=batch,
                           modulus=str(p), bit=int(scal[1]), mu=str(int(scal[2])), omega=str(int(scal[3])),
                           psi=str(int(scal[4])), n_inv=str(int(scal[5])), root=str(int(scal[6])),
                           inv_root=str(int(scal[7])), root_size=int(scal[8]),
                           fwd_br_hash=fold(fwd_br), inv_br_hash=fold(inv_br),
                           in_head=[str(int(v)) for v in x[:4]], in_hash=fold(x),
                           ntt_head=[str(int(v)) for v in y[:4]], ntt_hash=fold(y),
                           intt_head=[str(int(v)) for v in z[:4]], intt_hash=fold(z))
                if logn <= 4:
                    rec["in_full"] = [str(int(v)) for v in x]
                    rec["ntt_full"] = [str(int(v)) for v in y]
                    rec["intt_full"] = [str(int(v)) for v in z]
                recs["merge"].append(rec)
    for width in (64, 32):
        for logn in (12, 13, 15, 16, 17, 20):
            h = R.ref_4step_new(logn, 1, width)
            scal = np.zeros(12, dtype=np.uint64)
            R.ref_4step_scalars(h, width, scal)
            n = 1 << logn
            p = int(scal[0])
            tabs = {}
            for which, name in enumerate(("n1", "n2", "W", "n1_inv", "n2_inv", "W_inv")):
                sz = R.ref_4step_table(h, width, which, None, 0)
                t = np.zeros(sz, dtype=np.uint64)
                R.ref_4step_table(h, width, which, t.ctypes.data_as(C.c_void_p), 0)
                tabs[name + "_hash"] = fold(t)
            x = np.zeros(n, dtype=np.uint64)
            R.ref_example_input(0, p, n, x)
            y = np.zeros_like(x)
            R.ref_4step_run(h, width, 0, x, y)
            z = np.zeros_like(x)
            R.ref_4step_run(h, width, 1, x, z)
            ft = np.zeros_like(x)
            R.ref_4step_run(h, width, 2, x, ft)
            R.ref_4step_free(h, width)
            recs["fourstep"].append(dict(width=width, logn=logn, modulus=str(p), bit=int(scal[1]),
                                         mu=str(int(scal[2])), root=str(int(scal[6])), n_inv=str(int(scal[5])),
                                         n1=int(scal[10]), n2=int(scal[11]), **tabs, in_hash=fold(x),
                                         ntt_head=[str(int(v)) for v in y[:4]], ntt_hash=fold(y),
                                         intt_head=[str(int(v)) for v in z[:4]], intt_hash=fold(z),
                                         first_transpose_hash=fold(ft)))
    rng = np.random.RandomState(7)
    for width, p in ((64, 576460756061519873), (64, (1 << 61) - 1), (64, 4611686018427387847 - 0),
                     (32, 469762049), (32, 1073741789)):
        for _ in range(8):
            a = int(rng.randint(0, 2**31)) * int(rng.randint(0, 2**31)) % p
            b = int(rng.randint(0, 2**31)) * int(rng.randint(0, 2**31)) % p
            recs["barrett"].append(dict(width=width, p=str(p), a=str(a), b=str(b),
                                        r=str(int(R.ref_barrett_mult(a, b, p, width)))))
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json"), "w") as f:
        json.dump(recs, f, indent=0)
    print("wrote", len(recs["merge"]), "merge,", len(recs["fourstep"]), "4-step,", len(recs["barrett"]), "barrett records")


if __name__ == "__main__":
    main()
