#!/usr/bin/env python
"""Fuzz (not a test): mid-size inputs (4 096 < U <= ~300 000 distinct colours, the range the exact-integer split kernels and
their tie audit serve) through the CUDA path against the compiled reference.  Crops of the synthetic generators at random
seeds / sizes / subsamplings, plus noise-perturbed crops that make near-ties more likely.
Usage: python tools/fuzz_mid.py [trials] [seed]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import Oracle, Reference, muted  # noqa: E402

pkg = importlib.import_module("clusteringsegmentation-1_b200")
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
MIN_U = int(os.environ.get("FUZZ_MIN_U", "0"))     # skip inputs with fewer distinct colours (e.g. 262144: beyond the ordered path)
KMAX = int(os.environ.get("FUZZ_KMAX", "100000"))  # clamp the requested colours (e.g. 512: the audited kernel only)
dq = pkg.DivQuant()
o = Oracle()
ref = Reference()
bad = flagged = rerun = resolved = 0
hist_u = []
masks = {}
for t in range(trials):
    gen = 1 + (t % 2)
    img = o.generate(gen, 1920, 1080, int(rng.integers(1, 1 << 20))).reshape(1080, 1920)
    ch, cw = int(rng.integers(80, 1080)), int(rng.integers(80, 1500))
    y, x = int(rng.integers(0, 1080 - ch + 1)), int(rng.integers(0, 1920 - cw + 1))
    step = int(rng.choice([1, 1, 1, 2]))
    px = np.ascontiguousarray(img[y:y + ch:step, x:x + cw:step]).ravel()
    if t % 5 == 4:  # quantise the low bits: many equal colours, symmetric clusters, more exact ties
        px = px & np.uint32(0xFFF8F8F8 if t % 10 == 4 else 0xFFFCFCFC)
    k = int(rng.choice([16, 64, 256, 256, 300, 512, 700]))
    k = min(k, KMAX)
    u = np.unique(px & 0xFFFFFF).size
    if u < MIN_U:
        continue
    with muted():
        r_out, r_pal = ref.quant_recurse(px, k, 0)
    with muted((2,)):
        out, pal = dq.quant_recurse(px, k, 0)
    st = dq.last_stats()
    flagged += st["tie_flags"] != 0
    if st["tie_flags"]:
        key = (st["tie_flags"], "perturbed" if t % 5 == 4 else "plain", "U>262144" if u > 262144 else "U<=262144",
               "rerun" if st["ordered_rerun"] else "forced cut" if st["cut_overrides"] else "resolved" if st["tie_resolved"] else "reported")
        masks[key] = masks.get(key, 0) + 1
    rerun += st["ordered_rerun"]
    resolved += st["tie_resolved"] > 0
    hist_u.append(u)
    if not (np.array_equal(pal, r_pal) and np.array_equal(out, r_out)):
        bad += 1
        print(f"trial {t}: MISMATCH gen={gen} crop {ch}x{cw}/{step} n={px.size} U={u} k={k} tie_flags={st['tie_flags']} "
              f"palette entries differing {int((pal != r_pal).sum()) if pal.size == r_pal.size else -1}", flush=True)
hu = np.array(hist_u)
print(f"{len(hist_u)} mid-size inputs against the reference: {bad} mismatches; U min/median/max {hu.min()}/{int(np.median(hu))}/{hu.max()}, "
      f"{int((hu > 4096).sum())} above the ordered path's default limit; tie audit flagged {flagged} "
      f"({resolved} resolved in place, {rerun} re-run in the reference's order)")
for key in sorted(masks):
    print("  flags", key, masks[key])
