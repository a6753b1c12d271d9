"""N-Queens count-all kernel timing.  usage: queens_bench.py n [n ...]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dequan_b200 import api
from dequan_b200.model import nqueens
KNOWN = {12: (14200, 641974), 13: (73712, 3456855), 14: (365596, 19787662), 15: (2279184, 121498513), 16: (14772512, 795563572), 17: (95815104, 5474619051), 18: (666090624, 39749028012)}
for n in map(int, sys.argv[1:] or ["14", "15", "16", "17"]):
    m = api.Model(nqueens(n))
    best = None
    for rep in range(4):
        r = m.solve_tree("count")
        assert (r.solutions, r.nodes) == KNOWN[n], (n, r)
        best = r.kernel_ms if best is None else min(best, r.kernel_ms)
    print(f"N={n} K={r.split_depth} records={r.n_prefixes} kernel_ms={best:.3f} Gnodes/s={r.nodes / best / 1e6:.1f}", flush=True)
