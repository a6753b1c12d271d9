"""Profiling driver: N-Queens count-all on one engine (for ncu)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dequan_b200 import api
from dequan_b200.model import nqueens
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
eng = sys.argv[2] if len(sys.argv) > 2 else "lane"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
m = api.Model(nqueens(n))
for _ in range(reps):
    r = m.solve_tree("count", engine=eng)
print(n, eng, r.solutions, r.nodes, r.kernel_ms, r.nodes / r.kernel_ms / 1e6, "Gnodes/s")
