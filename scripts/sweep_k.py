import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dequan_b200 import api
from dequan_b200.model import nqueens
for n in (13, 14, 15, 16):
    m = api.Model(nqueens(n))
    for k in (5, 6, 7, 8, 9):
        if (n >= 17 and k < 4) or k > n - 3: continue
        try:
            r = m.solve_tree("count", engine="lane", split_depth=k, time_kernels=True)
            r = m.solve_tree("count", engine="lane", split_depth=k, time_kernels=True)
            print(f"N={n} K={k} records={r.n_prefixes} ms={r.kernel_ms:.3f} search_ms={r.search_kernel_ms:.3f} Gnodes/s={r.nodes/r.kernel_ms/1e6:.1f} sols={r.solutions} nodes={r.nodes}", flush=True)
        except Exception as e:
            print(n, k, "EXC", e, flush=True)
