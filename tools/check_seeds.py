#!/usr/bin/env python
"""Runs quant_recurse on given G1 seeds and compares palette / output hashes with tests/golden/frames.npz.
usage: python tools/check_seeds.py c4|bench seed [seed ...]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import Oracle  # input generator + hashing only

tag = sys.argv[1]
w, h, k = (1920, 1080, 64) if tag == "c4" else (3840, 2160, 256)
fr = np.load(os.path.join(ROOT, "tests", "golden", "frames.npz"))
pkg = importlib.import_module("clusteringsegmentation-1_b200")
dq = pkg.DivQuant(timings=False)
o = Oracle()
for seed in (int(x) for x in sys.argv[2:]):
    i = int(np.where(fr[f"{tag}_seeds"] == seed)[0][0])
    px = o.generate(1, w, h, seed)
    out, pal = dq.quant_recurse(px, k, 0)
    st = dq.last_stats()
    print(seed, "pal ok" if o.hash_words(pal) == int(fr[f"{tag}_pal_hash"][i]) else "PAL DIFF",
          "out ok" if o.hash_words(out) == int(fr[f"{tag}_out_hash"][i]) else "OUT DIFF",
          {k2: st[k2] for k2 in ("tie_flags", "ordered_rerun", "tie_resolved", "cut_overrides")}, "model mask", int(fr[f"{tag}_tie_mask"][i]), flush=True)
