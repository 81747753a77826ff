#!/usr/bin/env python
"""Research script (test infrastructure): the tie audit of the device-arithmetic model against the compiled reference.

For every frame: reference palette (oracle/_ref), exact-integer model palette + audit flags (oracle).  Reports
 * frames where the model differs from the reference (all must be flagged),
 * the flag rate by decision type D1..D5.
Usage: python tests/research/sweep_tie_audit.py W H K first_seed n_frames [kind]
"""
import ctypes as C
import os
import sys
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import Oracle, Reference, _ptr, _u32p  # noqa: E402


def one(args):
    w, h, k, seed, kind = args
    o, r = Oracle(), Reference()
    px = o.generate(kind, w, h, seed)
    ref_pal, _ = r.quant_varpart_fast(px, k)
    fn = o.lib.oracle_quant_varpart_fast_exact_audit
    fn.restype = C.c_int
    fn.argtypes = [C.c_uint32, _u32p, C.c_uint32, C.c_uint32, _u32p, _u32p, C.c_int, C.c_int, C.c_int, C.c_int, _u32p]
    ct = np.zeros(k, np.uint32)
    nk = C.c_uint32(k)
    flags = np.zeros(6, np.uint32)
    fn(px.size, _ptr(px), 1, px.size, C.byref(nk), _ptr(ct), 8, 1, 10, 0, _ptr(flags))
    pal = ct[:nk.value]
    same = pal.size == ref_pal.size and np.array_equal(pal, ref_pal)
    return seed, same, flags.tolist()


def main():
    w, h, k, s0, n = (int(x) for x in sys.argv[1:6])
    kind = int(sys.argv[6]) if len(sys.argv) > 6 else 1
    with Pool(8) as pool:
        res = pool.map(one, [(w, h, k, s0 + i, kind) for i in range(n)], chunksize=1)
    diff = [r for r in res if not r[1]]
    flagged = [r for r in res if r[2][0]]
    print(f"frames {n}: differ from reference {len(diff)}, flagged {len(flagged)}")
    tot = np.sum([r[2] for r in res], axis=0)
    print("flag counts [total, D1 axis, D2 cut, D3 hyperplane, D4 tse, D5 round]:", tot.tolist())
    for r in diff:
        print("  DIFF seed", r[0], "flags", r[2], "" if r[2][0] else "<-- NOT FLAGGED")
    for r in flagged:
        if r[1]:
            print("  flagged but equal: seed", r[0], r[2])


if __name__ == "__main__":
    main()
