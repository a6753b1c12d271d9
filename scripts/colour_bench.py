"""Colouring batch timing.  usage: colour_bench.py [count] [c] [k] [budget]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from dequan_b200 import api, generators as G
count = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
c = float(sys.argv[2]) if len(sys.argv) > 2 else 4.2
k = int(sys.argv[3]) if len(sys.argv) > 3 else 3
budget = int(sys.argv[4]) if len(sys.argv) > 4 else 1_000_000
t = time.time()
off, edges = G.colouring_batch(count, 200, c)
print(f"generated {count} instances, {off[-1]} edges in {time.time() - t:.1f}s", flush=True)
for rep in range(3):
    r = api.solve_batch_graphs(200, k, off, edges, node_budget=budget)
    print(f"count={count} c={c} k={k} budget={budget} ms={r.kernel_ms:.2f} sat={r.n_sat} unsat={r.n_unsat} budget_hit={r.n_budget} nodes={r.total_nodes} "
          f"inst/s={count / r.kernel_ms * 1e3:.0f} Gnodes/s={r.total_nodes / r.kernel_ms / 1e6:.2f} max_nodes={int(r.nodes.max())} median={int(np.median(r.nodes))}", flush=True)
