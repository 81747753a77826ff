#!/usr/bin/env python
"""Experiment: the ordered path's default limit.  For the 36 seeded natural crops of tests/golden (4 000 .. 60 000 colours)
runs quant_recurse with the ordered path up to `limit` colours and the audited exact-integer kernels above it; reports
parity against the compiled reference's palettes, how the tie audit disposed of the frames and the time.
usage: python tools/limit_check.py limit [limit ...]"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden import crop_pixels  # noqa: E402
from oracle import Oracle, muted  # noqa: E402

pkg = importlib.import_module("clusteringsegmentation-1_b200")
dq = pkg.DivQuant(timings=False)
o = Oracle()
golden = np.load(os.path.join(ROOT, "tests", "golden", "reference_outputs.npz"))
shaped = {}
for name in ("batman", "cookie"):
    z = np.load(os.path.join(ROOT, "tests", "golden", f"{name}_px.npz"))
    shaped[name] = z["px"].reshape(int(z["shape"][0]), int(z["shape"][1]))
shaped["g1"] = o.generate(1, 1920, 1080, 99).reshape(1080, 1920)
specs = golden["crop_specs"]
ctx = dq.lib.dq_default_context()
for limit in (int(x) for x in sys.argv[1:]):
    dq.lib.dq_context_set_exact_max_points(ctx, limit)
    ok = flagged = resolved = rerun = ordered = 0
    t_total = 0.0
    for i, spec in enumerate(specs):
        px = crop_pixels(shaped, spec)
        with muted((2,)):
            dq.quant_recurse(px, int(spec[6]), 0)
            t0 = time.perf_counter()
            out, pal = dq.quant_recurse(px, int(spec[6]), 0)
            t_total += time.perf_counter() - t0
        st = dq.last_stats()
        ok += int(np.array_equal(pal, golden[f"crop{i}_palette"]) and o.hash_words(out) == int(golden[f"crop{i}_out_hash"][0]))
        flagged += int(st["tie_flags"] != 0)
        resolved += int(st["tie_resolved"] > 0)
        rerun += st["ordered_rerun"]
        ordered += int(int(golden[f"crop{i}_unique"][0]) <= limit)
    print(f"limit {limit:6d}: {ok}/{len(specs)} bit-exact, {ordered} on the ordered path by size, flagged {flagged} (resolver {resolved}, ordered re-run {rerun}), "
          f"{1e3 * t_total / len(specs):.2f} ms per crop", flush=True)
dq.lib.dq_context_set_exact_max_points(ctx, 65536)
