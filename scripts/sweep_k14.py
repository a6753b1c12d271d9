import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dequan_b200 import api
from dequan_b200.model import nqueens
for n in (13, 14, 15):
    m = api.Model(nqueens(n))
    for k in (5, 6, 7, 8):
        best = 1e9
        for rep in range(4):
            r = m.solve_tree("count", engine="lane", split_depth=k)
            best = min(best, r.kernel_ms)
        print(f"N={n} K={k} records={r.n_prefixes} ms={best:.3f} Gnodes/s={r.nodes/best/1e6:.1f} sols={r.solutions} nodes={r.nodes}", flush=True)
