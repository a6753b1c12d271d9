"""Split-depth sweep of the N-Queens bucket search.  usage: sweep_k.py [n,n,...] [k,k,...]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dequan_b200 import api
from dequan_b200.model import nqueens
ns = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [13, 14, 15, 16, 17]
ks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [5, 6, 7, 8, 9, 10]
for n in ns:
    m = api.Model(nqueens(n))
    for k in ks:
        if k > n - 4 or (n >= 17 and k < 7) or (n >= 18 and k > 9):
            continue
        try:
            best = None
            for _ in range(3):
                r = m.solve_tree("count", engine="lane", split_depth=k, time_kernels=True)
                if best is None or r.kernel_ms < best.kernel_ms:
                    best = r
            r = best
            print(f"N={n} K={k} records={r.n_prefixes} ms={r.kernel_ms:.3f} search_ms={r.search_kernel_ms:.3f} Gnodes/s={r.nodes/r.kernel_ms/1e6:.1f} sols={r.solutions} nodes={r.nodes}", flush=True)
        except Exception as e:
            print(n, k, "EXC", e, flush=True)
