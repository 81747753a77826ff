import importlib, sys, time, numpy as np
sys.path.insert(0, '.')
from oracle import Oracle, Reference, muted
pkg = importlib.import_module("clusteringsegmentation-1_b200")
dq = pkg.DivQuant(); o = Oracle(); ref = Reference()
rng = np.random.default_rng(5)
for (w, h, k) in ((4, 4, 4), (4, 4, 125), (10, 10, 256), (40, 25, 125), (40, 25, 1500), (16, 16, 300), (400, 250, 125)):
    px = o.generate(1, w, h, 777)
    with muted():
        r_out, r_pal = ref.quant_recurse(px, k, 0)
        for _ in range(3): out, pal = dq.quant_recurse(px, k, 0)
        ts = []
        for _ in range(15):
            t0 = time.perf_counter(); dq.quant_recurse(px, k, 0); ts.append(time.perf_counter() - t0)
    print(w * h, k, "ok" if np.array_equal(pal, r_pal) and np.array_equal(out, r_out) else "MISMATCH", f"{1e3*np.median(ts):.3f} ms", flush=True)
