import importlib, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("clusteringsegmentation-1_b200")
from oracle import Oracle, muted
o = Oracle(); dq = pkg.DivQuant()
g = np.load(os.path.join(ROOT, "tests/golden/reference_outputs.npz"))
ids = [int(x) for x in sys.argv[1:]] or [12]
for i in ids:
    px, k = g[f"small{i}_in"], int(g[f"small{i}_k"][0])
    for uq in (1, 0):
        out, pal = dq.quant_recurse(px, k, uq)
        st = dq.last_stats()
        ref = g[f"small{i}_u{uq}_palette"]
        print(i, "n", px.size, "U", np.unique(px & 0xFFFFFF).size, "K", k, "uq", uq, "ok" if np.array_equal(pal, ref) else "DIFF",
              [hex(x) for x in pal[:6]], [hex(x) for x in ref[:6]], {k2: st[k2] for k2 in ("split_rounds", "splits_computed", "actual_colors", "empty_clusters")}, flush=True)
