"""One N-Queens all-solutions solve, repeated (ncu target).  usage: queens_once.py [n] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dequan_b200 import api
from dequan_b200.model import nqueens
n = int(sys.argv[1]) if len(sys.argv) > 1 else 17
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
m = api.Model(nqueens(n))
for _ in range(reps):
    r = m.solve_tree("count", time_kernels=True)
    print(n, r.solutions, r.nodes, f"kernel_ms={r.kernel_ms:.3f} search_ms={r.search_kernel_ms:.3f}", flush=True)
