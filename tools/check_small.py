#!/usr/bin/env python
"""Every kernel family once on small inputs, checked against the oracle (a quick end-to-end sanity run on a GPU box).
usage: python tools/check_small.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import Oracle, muted  # noqa: E402

pkg = importlib.import_module("clusteringsegmentation-1_b200")
o = Oracle()
dq = pkg.DivQuant()
rng = np.random.default_rng(5)
bad = 0


def check(name, ok):
    global bad
    print(("PASS " if ok else "FAIL ") + name, flush=True)
    bad += 0 if ok else 1


for n, k, uq in ((16, 4, 0), (4096, 16, 0), (6000, 64, 0), (3000, 8, 1), (9000, 300, 0)):
    px = rng.integers(0, 1 << 24, n, dtype=np.uint32)
    with muted():
        r_out, r_pal = o.quant_recurse(px, k, uq)
    with muted((2,)):
        out, pal = dq.quant_recurse(px, k, uq)
    check(f"quant_recurse n={n} k={k} uniq={uq}", np.array_equal(pal, r_pal) and np.array_equal(out, r_out))
g1s = o.generate(1, 300, 200, 3)   # below the ordered path's limit of 65536 colours: the reference's result bit for bit
with muted():
    r_out, r_pal = o.quant_recurse(g1s, 64, 0)
with muted((2,)):
    out, pal = dq.quant_recurse(g1s, 64, 0)
check(f"quant_recurse G1 300x200 k=64 (U={np.unique(g1s & 0xFFFFFF).size}, ordered path)", np.array_equal(pal, r_pal) and np.array_equal(out, r_out))
g1 = o.generate(1, 400, 300, 3)    # 68 296 colours: above the limit, exact-integer kernels (one palette entry differs from the
                                   # reference here: an exact tie)
ctx = dq.lib.dq_default_context()
dq.lib.dq_context_set_exact_small(ctx, 0)   # the exact-integer kernels on the same input
with muted():
    model, _ = o.quant_varpart_fast(g1, 64, exact_counts=True)
with muted((2,)):
    pal, _ = dq.quant_varpart_fast(g1, 64)
dq.lib.dq_context_set_exact_small(ctx, 1)
check("integer kernels G1 400x300 k=64", np.array_equal(pal, model))
pal = rng.integers(0, 1 << 24, 125, dtype=np.uint32)
q = dq.map_colors_mps(g1, pal)
check("map_colors_mps", np.array_equal(q, o.map_colors_mps(g1, pal)))
check("block vote", np.array_equal(dq.block_vote(q, 400, 300, 4), o.block_vote(q, 400, 300, 4)))
im = rng.integers(0, 64, (37, 51, 3)).astype(np.uint8) * 4
check("srm edges", np.array_equal(dq.srm_sorted_edges(im), o.srm_sorted_edges(im)))
keys, counts = dq.pixel_histogram(g1)
ek, ec = np.unique(g1 & 0xFFFFFF, return_counts=True)
check("pixel histogram", np.array_equal(keys, ek) and np.array_equal(counts, ec))
print("failures:", bad)
sys.exit(1 if bad else 0)
