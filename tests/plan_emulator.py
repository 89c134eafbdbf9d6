"""Pure-Python emulation of the index algebra of gpu_ntt_b200/csrc/merge_ntt.cu (tile decode,
register rounds, twiddle indexing) driven by the REAL launch plan the library reports
(gpuntt_b200_describe_plan).  Arithmetic is plain canonical `%` -- the point is to check, without
a GPU, that plan + indexing compute the reference transform."""
import re

import numpy as np


def parse_plan(text):
    passes = []
    for m in re.finditer(r"pass\d+\{tile=2\^(\d+) (\w+) lo=(\d+) d=(\d+) c=(\d+) rounds=([\d,]+)\}", text):
        passes.append(dict(k=int(m.group(1)), lo=int(m.group(3)), d=int(m.group(4)), c=int(m.group(5)),
                           rounds=[int(x) for x in m.group(6).split(",")]))
    return passes


def swz(l, bits):
    return l ^ (((l >> 4) & 7) << 1) if bits == 64 else l ^ (((l >> 5) & 7) << 2)


def emulate(x, n, p, table_br, plus, passes, inverse, n_inv=None, mod_count=0, moduli=None, n_invs=None):
    """x: flat python-int list of batch*N canonical residues.  table_br: caller's (bit-reversed) table; for
    RNS a dict slice->list.  Returns the transformed list."""
    N = 1 << n
    total = len(x)
    data = list(x)
    order = list(reversed(passes)) if inverse else passes
    for ps in order:
        k, lo, d, c = ps["k"], ps["lo"], ps["d"], ps["c"]
        tile_elems = 1 << k
        ntiles = total >> k if lo > 0 else (total + tile_elems - 1) >> k
        # local low bit of every round (rounds listed high -> low)
        lbs, acc = [], c + d
        for r in ps["rounds"]:
            acc -= r
            lbs.append(acc)
        rr = list(zip(ps["rounds"], lbs))
        if inverse:
            rr.reverse()
        for tile in range(ntiles):
            if lo > 0:
                hi = lo + d
                cc_bits, pre_bits = lo - c, n - hi
                cc = tile & ((1 << cc_bits) - 1)
                P = (tile >> cc_bits) & ((1 << pre_bits) - 1)
                poly0 = tile >> (cc_bits + pre_bits)
                gbase = (poly0 << n) + (P << hi) + (cc << c)
                jrow_tile = P << d
            else:
                gbase = tile << k
                jrow_tile = gbase & (N - 1)
                poly0 = gbase >> n
            cmask = (1 << c) - 1
            poly_shift = n - lo
            g_of = [gbase + ((l >> c) << lo) + (l & cmask) for l in range(tile_elems)]
            sm = [data[g] if g < total else 0 for g in g_of]
            for (R, lb) in rr:
                items = tile_elems >> R
                stage_hi = n - 1 - lo - (lb - c)
                rb0 = lb - c
                for item in range(items):
                    l_base = ((item >> lb) << (lb + R)) | (item & ((1 << lb) - 1))
                    row = l_base >> c
                    jrow = (jrow_tile | row) & ((1 << poly_shift) - 1)
                    if mod_count:
                        mi = (poly0 + (row >> poly_shift)) % mod_count
                        pp, tw = moduli[mi], table_br[mi]
                    else:
                        pp, tw = p, table_br
                    e = [sm[l_base | (a << lb)] for a in range(1 << R)]
                    abs_ = range(R) if inverse else range(R - 1, -1, -1)
                    for ab in abs_:
                        s = stage_hi - ab
                        tb = (plus << s) + (jrow >> (rb0 + ab + 1))
                        for xx in range((1 << R) >> (ab + 1)):
                            w = tw[tb + xx]
                            for y in range(1 << ab):
                                a0 = (xx << (ab + 1)) | y
                                a1 = a0 | (1 << ab)
                                if not inverse:
                                    t = e[a1] * w % pp
                                    e[a0], e[a1] = (e[a0] + t) % pp, (e[a0] - t) % pp
                                else:
                                    e[a0], e[a1] = (e[a0] + e[a1]) % pp, (e[a0] - e[a1]) * w % pp
                    for a in range(1 << R):
                        sm[l_base | (a << lb)] = e[a]
            for l, g in enumerate(g_of):
                if g < total:
                    data[g] = sm[l]
    if inverse:
        for g in range(total):
            if mod_count:
                mi = (g >> n) % mod_count
                data[g] = data[g] * n_invs[mi] % moduli[mi]
            else:
                data[g] = data[g] * n_inv % p
    return data


def bank_conflicts(passes, bits):
    """Worst-case shared-memory wavefronts per warp access of every round of a plan (1 = conflict
    free) for the round loads/stores: a warp = 32 consecutive items (blockDim 256)."""
    worst = {}
    esz = bits // 8
    for pi, ps in enumerate(passes):
        k, c, d = ps["k"], ps["c"], ps["d"]
        acc = c + d
        for R in ps["rounds"]:
            acc -= R
            lb = acc
            w = 0
            for warp in range(min(8, (1 << k >> R) // 32 or 1)):
                vn = 16 // esz
                vec = lb == 0 and (1 << R) >= vn   # 16-byte accesses when the elements are adjacent
                asz = 16 if vec else esz
                for a in range(0, 1 << R, vn if vec else 1):
                    addrs = []
                    for lane in range(32):
                        item = warp * 32 + lane
                        if item >= (1 << k >> R):
                            continue
                        l_base = ((item >> lb) << (lb + R)) | (item & ((1 << lb) - 1))
                        addrs.append(swz(l_base | (a << lb), bits) * esz)
                    # hardware processes 128 B of requests per wavefront when no two lanes of the
                    # group hit different addresses in the same bank
                    group = 128 // asz
                    for g0 in range(0, len(addrs), group):
                        banks = {}
                        for ad in addrs[g0:g0 + group]:
                            for word in range(ad // 4, (ad + asz) // 4):
                                banks.setdefault(word % 32, set()).add(word)
                        w = max(w, max(len(v) for v in banks.values()))
            worst[(pi, lb, R)] = w
    return worst
