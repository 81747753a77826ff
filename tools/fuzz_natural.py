#!/usr/bin/env python
"""Fuzz (not a test): random crops / subsamplings of the reference's own test images and of the synthetic generators
through the CUDA path against the compiled reference.  Usage: python tools/fuzz_natural.py [trials] [seed]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import Oracle, Reference, muted  # noqa: E402

pkg = importlib.import_module("clusteringsegmentation-1_b200")
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
dq = pkg.DivQuant()
o = Oracle()
try:
    ref, kind = Reference(), "reference"
except FileNotFoundError:
    ref, kind = o, "oracle"
images = []
for name in ("batman", "cookie"):
    z = np.load(os.path.join(ROOT, "tests", "golden", f"{name}_px.npz"))
    h, w = (int(v) for v in z["shape"])
    images.append((name, z["px"].reshape(h, w)))
images.append(("g1", o.generate(1, 1920, 1080, 99).reshape(1080, 1920)))
bad = large = worst = 0
for t in range(trials):
    name, img = images[t % len(images)]
    h, w = img.shape
    ch, cw = int(rng.integers(40, min(h, 700))), int(rng.integers(40, min(w, 700)))
    y, x = int(rng.integers(0, h - ch + 1)), int(rng.integers(0, w - cw + 1))
    step = int(rng.choice([1, 1, 2, 3]))
    px = np.ascontiguousarray(img[y:y + ch:step, x:x + cw:step]).ravel()
    k = int(rng.choice([4, 16, 64, 125, 256]))
    u = np.unique(px & 0xFFFFFF).size
    with muted():
        r_out, r_pal = ref.quant_recurse(px, k, 0)
    with muted((2,)):
        out, pal = dq.quant_recurse(px, k, 0)
    ok = np.array_equal(pal, r_pal) and np.array_equal(out, r_out)
    large += u > 65536
    if not ok:
        bad += 1
        if pal.size == r_pal.size:
            d = np.abs(((pal[:, None] >> np.array([16, 8, 0])) & 0xFF).astype(int) - ((r_pal[:, None] >> np.array([16, 8, 0])) & 0xFF).astype(int))
            lab = o.colortable_indexes(out, pal) if hasattr(o, "colortable_indexes") else None
            r_lab = o.colortable_indexes(r_out, r_pal) if lab is not None else None
            nlab = int((lab != r_lab).sum()) if lab is not None else -1
            worst = max(worst, int(d.max()))
            print(f"trial {t}: MISMATCH {name} crop {ch}x{cw}/{step} n={px.size} U={u} k={k}: {int((pal != r_pal).sum())} palette entries differ, "
                  f"max channel difference {int(d.max())} LSB, {nlab} of {px.size} labels differ", flush=True)
        else:
            worst = 999
            print(f"trial {t}: MISMATCH {name} n={px.size} U={u} k={k}: palette sizes {pal.size} vs {r_pal.size}", flush=True)
print(f"{trials} natural crops against the {kind}: {bad} mismatches ({large} inputs with U > 65536), worst palette difference {worst} LSB")
