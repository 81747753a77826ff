#!/usr/bin/env python
"""Fuzz (not a test): random small inputs through the CUDA path against the compiled reference (oracle/_ref, or the
oracle restatement when it is absent).  Usage: python tools/fuzz_small.py [trials] [seed]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import Oracle, Reference, muted  # noqa: E402

pkg = importlib.import_module("clusteringsegmentation-1_b200")
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
NMAX = int(sys.argv[3]) if len(sys.argv) > 3 else 4000
dq = pkg.DivQuant()
o = Oracle()
try:
    ref, kind = Reference(), "reference"
except FileNotFoundError:
    ref, kind = o, "oracle"
bad = small = 0
for t in range(trials):
    mode = t % 4
    n = int(rng.integers(1, NMAX))
    if mode == 0:
        px = rng.integers(0, 1 << 24, n, dtype=np.uint32)
    elif mode == 1:
        pal = rng.integers(0, 1 << 24, int(rng.integers(1, 60)), dtype=np.uint32)
        px = pal[rng.integers(0, pal.size, n)]
    elif mode == 2:
        g = rng.integers(0, 256, int(rng.integers(1, 30)))
        pal = (g * 0x010101).astype(np.uint32)
        px = np.tile(pal, n // pal.size + 1)[:n]
    else:
        c = rng.integers(0, 4, (n, 3)) * 85
        px = ((c[:, 0] << 16) | (c[:, 1] << 8) | c[:, 2]).astype(np.uint32)
    k = int(rng.choice([1, 2, 3, 5, 16, 64, 256, 300, 700]))
    uniq = int(rng.integers(0, 2)) if t % 7 == 0 else 0
    u = np.unique(px & 0xFFFFFF).size
    with muted():
        r_out, r_pal = ref.quant_recurse(px, k, uniq)
    with muted((2,)):
        out, pal = dq.quant_recurse(px, k, uniq)
    ok = np.array_equal(pal, r_pal) and np.array_equal(out, r_out)
    small += u <= 65536
    if not ok:
        bad += 1
        print(f"trial {t}: MISMATCH mode={mode} n={n} U={u} k={k} uniq={uniq} palettes_equal={np.array_equal(pal, r_pal)}", flush=True)
print(f"{trials} trials against the {kind}: {bad} mismatches ({small} inputs with U <= 65536)")
import time
g1 = o.generate(1, 1920, 1080, 7).reshape(1080, 1920)
cases = [("random", rng.integers(0, 1 << 24, n, dtype=np.uint32), k) for n, k in ((65536, 256), (16384, 256), (4096, 256), (4000, 16), (1000, 256), (100, 16))]
cases += [(f"g1 crop {s}x{s}", np.ascontiguousarray(g1[100:100 + s, 200:200 + s]).ravel(), 256) for s in (120, 200, 300, 420)]
for name, px, k in cases:
    for _ in range(2):
        t0 = time.perf_counter()
        with muted((2,)):
            dq.quant_recurse(px, k, 0)
        dt = time.perf_counter() - t0
    st = dq.last_stats()
    print(f"timing {name} n={px.size} U={st['num_points']} k={k}: {dt * 1e3:.3f} ms per call (host clock), launches {st['kernel_launches']}")
