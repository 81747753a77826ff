import importlib, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("clusteringsegmentation-1_b200")
dq = pkg.DivQuant()
name, ks = sys.argv[1], [int(x) for x in sys.argv[2:]]
px = np.load(os.path.join(ROOT, "tests/golden", f"{name}_px.npz"))["px"].ravel()
g = np.load(os.path.join(ROOT, "tests/golden/reference_outputs.npz"))
for k in ks:
    out, pal = dq.quant_recurse(px, k, 0)
    print(name, k, "palette ok" if np.array_equal(pal, g[f"{name}_k{k}_palette"]) else "PALETTE DIFF", dq.last_stats(), flush=True)
