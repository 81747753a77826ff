#!/usr/bin/env python
"""Golden fingerprints of the compiled reference (oracle/_ref) for every frame of BASELINE config 4 and bench.py.

    python tests/golden/make_golden_frames.py          (build container; ~1 min on 8 cores)

Writes tests/golden/frames.npz:
  c4_*     the 1024 frames of config 4: 1920x1080 G1, K=64, frame f has seed 12345+f  
  bench_*  the frames bench.py times: 3840x2160 G1, K=256, seeds 12345..12345+255 (8 ranks x 32 ring slots)
  g2_*     the stress inputs of SURVEY.md 8c/8d: G2 uniform-random 1920x1080; g2_4k_*: 3840x2160 (--g2-4k, ~6 min)
For each frame: palette hash, out hash (oracle.hash_words, FNV-1a over u32 words), palette size, and the TieBit mask
the CPU model of the device's tie audit raises (oracle_quant_varpart_fast_exact_audit; 0 = the exact-integer path is
guaranteed to equal the reference).
"""
import ctypes as C
import os
import sys
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import Oracle, Reference, _ptr, _u32p  # noqa: E402

TIE_BITS = (1, 2, 4, 8, 16)  # D1..D5 -> TieBit (csrc/dq_split.cuh)


def model_flags(o, px, k):
    fn = o.lib.oracle_quant_varpart_fast_exact_audit
    fn.restype = C.c_int
    fn.argtypes = [C.c_uint32, _u32p, C.c_uint32, C.c_uint32, _u32p, _u32p, C.c_int, C.c_int, C.c_int, C.c_int, _u32p]
    ct = np.zeros(k, np.uint32)
    nk = C.c_uint32(k)
    flags = np.zeros(6, np.uint32)
    fn(px.size, _ptr(px), 1, px.size, C.byref(nk), _ptr(ct), 8, 1, 10, 0, _ptr(flags))
    mask = 0
    for d in range(1, 6):
        if flags[d]:
            mask |= TIE_BITS[d - 1]
    return mask, ct[:nk.value].copy()


def one(args):
    kind, w, h, k, seed = args
    o, r = Oracle(), Reference()
    px = o.generate(kind, w, h, seed)
    out, pal = r.quant_recurse(px, k, 0)
    mask, model_pal = model_flags(o, px, k)
    return (o.hash_words(pal), o.hash_words(out), pal.size, mask, int(np.array_equal(model_pal, r.quant_varpart_fast(px, k)[0])))


def run(pool, kind, w, h, k, seeds):
    res = pool.map(one, [(kind, w, h, k, s) for s in seeds], chunksize=1)
    return {"pal_hash": np.array([r[0] for r in res], np.uint64), "out_hash": np.array([r[1] for r in res], np.uint64),
            "pal_size": np.array([r[2] for r in res], np.uint32), "tie_mask": np.array([r[3] for r in res], np.uint32),
            "model_equal": np.array([r[4] for r in res], np.uint8), "seeds": np.array(list(seeds), np.uint64)}


def main():
    """--only TAG[,TAG...] (c4, bench, g2, g2_4k) recomputes just those sets and keeps the rest of an existing frames.npz."""
    path = os.path.join(HERE, "frames.npz")
    only = None
    if "--only" in sys.argv:
        only = set(sys.argv[sys.argv.index("--only") + 1].split(","))
    out = {}
    if only is not None and os.path.exists(path):
        out = dict(np.load(path))
    sets = (("c4", 1, 1920, 1080, 64, range(12345, 12345 + 1024)),
            ("bench", 1, 3840, 2160, 256, range(12345, 12345 + 256)),
            ("g2", 2, 1920, 1080, 256, range(12345, 12346)),
            ("g2_4k", 2, 3840, 2160, 256, range(12345, 12346)))  # ~6 min of CPU: only with --g2-4k or --only g2_4k
    with Pool(8) as pool:
        for tag, kind, w, h, k, seeds in sets:
            if only is not None and tag not in only:
                continue
            if only is None and tag == "g2_4k" and "--g2-4k" not in sys.argv:
                continue
            for key, val in run(pool, kind, w, h, k, seeds).items():
                out[f"{tag}_{key}"] = val
            print(tag, "frames", len(out[f"{tag}_seeds"]), "flagged", int((out[f"{tag}_tie_mask"] != 0).sum()),
                  "model != reference", int((out[f"{tag}_model_equal"] == 0).sum()), flush=True)
    # an unflagged frame must equal the reference: the audit's contract
    for tag in ("c4", "bench", "g2", "g2_4k"):
        if f"{tag}_seeds" not in out:
            continue
        bad = (out[f"{tag}_model_equal"] == 0) & (out[f"{tag}_tie_mask"] == 0)
        assert not bad.any(), (tag, out[f"{tag}_seeds"][bad])
    np.savez_compressed(path, **out)
    print("written", path)


if __name__ == "__main__":
    main()
