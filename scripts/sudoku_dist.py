import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from dequan_b200 import api, generators as G
from dequan_b200.model import sudoku_template
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
cells = G.sudoku_batch(n, 30)
t = api.Model(sudoku_template())
r = t.solve_batch_cells(cells)
nodes = np.sort(r.nodes)[::-1]
print("total", nodes.sum(), "ms", r.kernel_ms)
print("top", nodes[:12].tolist())
for q in (0.5, 0.9, 0.99, 0.999, 0.9999): print(q, int(np.quantile(r.nodes, q)))
for thr in (1000, 4096, 16384, 65536, 262144):
    m = r.nodes > thr
    print(f"> {thr}: {m.sum()} instances, {r.nodes[m].sum()/nodes.sum()*100:.1f}% of nodes")
for b in (2048, 8192, 32768):
    rb = t.solve_batch_cells(cells, node_budget=b)
    print(f"budget {b}: ms={rb.kernel_ms:.2f} unfinished={rb.n_budget} nodes={rb.total_nodes}")
