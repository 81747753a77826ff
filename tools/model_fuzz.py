#!/usr/bin/env python
"""CPU only: soundness fuzz of the tie audit + resolver METHOD (not of the device code): the oracle's model of the device's
default large-input path (oracle.quant_varpart_resolved) against the compiled reference on random crops / subsamplings of the
synthetic generators and of the reference's test images, random K.  Whenever the model says nothing is left flagged, its
palette must be the reference's -- a difference there would mean a bound of csrc/dq_tie.cuh is too tight.
usage: python tools/model_fuzz.py [trials] [seed] [processes]"""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import Oracle, Reference, muted  # noqa: E402

_IMAGES = None


def images(o):
    global _IMAGES
    if _IMAGES is None:
        _IMAGES = []
        for name in ("batman", "cookie"):
            z = np.load(os.path.join(ROOT, "tests", "golden", f"{name}_px.npz"))
            h, w = (int(v) for v in z["shape"])
            _IMAGES.append(z["px"].reshape(h, w))
    return _IMAGES


def one(args):
    t, seed = args
    rng = np.random.default_rng(seed * 1000003 + t)
    o, ref = Oracle(), Reference()
    src = t % 4
    if src < 2:
        img = o.generate(1 + src, 1920, 1080, int(rng.integers(1, 1 << 20))).reshape(1080, 1920)
    else:
        img = images(o)[src - 2]
    h, w = img.shape
    ch, cw = int(rng.integers(40, min(h, 900))), int(rng.integers(40, min(w, 900)))
    y, x = int(rng.integers(0, h - ch + 1)), int(rng.integers(0, w - cw + 1))
    step = int(rng.choice([1, 1, 2, 3]))
    px = np.ascontiguousarray(img[y:y + ch:step, x:x + cw:step]).ravel()
    if t % 7 == 6:
        px = px & np.uint32(0xFFFCFCFC)
    k = int(rng.choice([4, 16, 64, 125, 256, 300, 512]))
    u = int(np.unique(px & 0xFFFFFF).size)
    if u <= 4096:  # the device sums these in the reference's own order (ordered path): nothing to model
        return None
    with muted():
        rp, _ = ref.quant_varpart_fast(px, k)
    pal, info = o.quant_varpart_resolved(px, k)
    return bool(np.array_equal(pal, rp)), info, u, k


if __name__ == "__main__":
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    procs = int(sys.argv[3]) if len(sys.argv) > 3 else (os.cpu_count() or 1)
    with mp.Pool(procs) as pool:
        res = pool.map(one, [(t, seed) for t in range(trials)], chunksize=8)
    skipped = sum(1 for r in res if r is None)
    res = [r for r in res if r is not None]
    clean = [r for r in res if r[1]["left"] == 0]
    unsound = [r for r in clean if not r[0]]
    left = [r for r in res if r[1]["left"] != 0]
    print(f"{len(res)} inputs above the ordered path's default limit of 4096 colours (unique colours {min(r[2] for r in res)}.."
          f"{max(r[2] for r in res)}; {skipped} smaller ones skipped), seed {seed}:")
    print(f"  nothing left flagged: {len(clean)}; of those NOT equal to the reference: {len(unsound)}   <- must be 0")
    print(f"  roundings resolved {sum(r[1]['roundings'] for r in res)}, cuts confirmed {sum(r[1]['cuts_confirmed'] for r in res)}, "
          f"cuts forced {sum(r[1]['cuts_forced'] for r in res)}")
    print(f"  a decision left flagged (axis / hyperplane / TSE: the device re-runs these in the reference's order): {len(left)}; "
          f"of those not equal to the reference as they stand: {sum(1 for r in left if not r[0])}")
    print(f"    kinds among those: axis (two channel variances) {sum(1 for r in left if r[1]['left_axis'])}, hyperplane "
          f"{sum(1 for r in left if r[1]['left_hyperplane'])}, TSE arg-max {sum(1 for r in left if r[1]['left_tse'])}; "
          f"with more than 262144 colours (beyond the re-run): {sum(1 for r in left if r[2] > 262144)}")
    for r in unsound[:10]:
        print("  UNSOUND:", r)
    sys.exit(1 if unsound else 0)
