#!/usr/bin/env python
"""Ad-hoc GPU bring-up check (not a test): compares the CUDA path with the oracle on a few inputs and
prints diagnostics.  Usage on a GPU box: python tools/gpu_check.py [quick|full]"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import Oracle, muted  # noqa: E402

pkg = importlib.import_module("clusteringsegmentation-1_b200")
mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
o = Oracle()
dq = pkg.DivQuant()
print(dq.lib.dq_version().decode(), flush=True)
bad = 0


def check(name, ok, extra=""):
    global bad
    print(("PASS " if ok else "FAIL ") + name + (" " + extra if extra else ""), flush=True)
    bad += 0 if ok else 1


# 1. histogram
rng = np.random.default_rng(3)
px = rng.integers(0, 1 << 24, 5000, dtype=np.uint32)
px[::3] = px[0]
col, cnt = dq.histogram(px | 0xFF000000)
uc, ucnt = np.unique(px, return_counts=True)
order = np.argsort(col)
check("histogram", np.array_equal(col[order], uc) and np.array_equal(cnt[order], ucnt), f"U={col.size}")

# 2. KATs (uniform path is integer-exact in the reference: must match bit for bit)
kats = json.load(open(os.path.join(ROOT, "tests/golden/divquant_kat.json")))["cases"]
for kat in kats:
    p = np.array(kat["pixels"], np.uint32)
    for uq in (1, 0):
        with muted((2,)):
            out, pal = dq.quant_recurse(p, kat["k"], uq)
        ok = [int(x) for x in pal] == kat["palette"] and [int(x) for x in out] == kat["out_pixels"]
        check(f"KAT {kat['name']} uniq={uq}", ok, str([hex(x) for x in pal]) if not ok else "")

# 3. map_colors on random palettes
for t in range(6):
    n = int(rng.integers(1, 100000))
    p = rng.integers(0, 1 << 24, n, dtype=np.uint32)
    pal = rng.integers(0, 1 << 24, int(rng.integers(1, 300)), dtype=np.uint32)
    check(f"map random n={n} k={pal.size}", np.array_equal(dq.map_colors_mps(p, pal), o.map_colors_mps(p, pal)))

# 4. synthetic / fixture images
cases = [("g1_640", o.generate(1, 640, 360), 64), ("g2_320", o.generate(2, 320, 180), 256)]
if mode == "full":
    cases += [("batman", np.load(os.path.join(ROOT, "tests/golden/batman_px.npz"))["px"].ravel(), 256),
              ("cookie", np.load(os.path.join(ROOT, "tests/golden/cookie_px.npz"))["px"].ravel(), 256),
              ("g1_1080", o.generate(1, 1920, 1080), 256), ("g1_4k", o.generate(1, 3840, 2160), 256)]
for name, p, k in cases:
    t0 = time.time()
    with muted((2,)):
        out, pal = dq.quant_recurse(p, k, 0)
    t1 = time.time()
    st = dq.last_stats()
    with muted():
        oout, opal = o.quant_recurse(p, k, 0)
    t2 = time.time()
    check(f"quant_recurse {name} K={k}", np.array_equal(pal, opal) and np.array_equal(out, oout),
          f"gpu {1e3*(t1-t0):.1f} ms cpu {1e3*(t2-t1):.1f} ms stats {st}")
    if not np.array_equal(pal, opal):
        print("   palette diff entries:", int((pal[:min(len(pal), len(opal))] != opal[:min(len(pal), len(opal))]).sum()), len(pal), len(opal))
print("FAILED" if bad else "ALL OK", bad)
sys.exit(1 if bad else 0)
