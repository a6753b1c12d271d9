"""Generic warp engine against the class-specific engines on the BASELINE shapes (numbers quoted in DESIGN.md)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from dequan_b200 import api, generators as G
from dequan_b200.model import nqueens, sudoku_template
for n in (12, 14, 15):
    m = api.Model(nqueens(n))
    for eng in ("warp", "lane"):
        best = min(m.solve_tree("count", engine=eng).kernel_ms for _ in range(3))
        r = m.solve_tree("count", engine=eng)
        print(f"queens N={n} engine={eng} kernel_ms={best:.3f} Gnodes/s={r.nodes / best / 1e6:.2f}", flush=True)
cells = G.sudoku_batch(200000, 30)
t = api.Model(sudoku_template())
for eng in ("warp", "lane"):
    best = min(t.solve_batch_cells(cells, engine=eng).kernel_ms for _ in range(2))
    r = t.solve_batch_cells(cells, engine=eng)
    print(f"sudoku 200k g30 engine={eng} kernel_ms={best:.2f} Mpuzzles/s={200000 / best / 1e3:.2f} Gnodes/s={r.total_nodes / best / 1e6:.2f}", flush=True)
off, edges = G.colouring_batch(1024, 200, 4.2)
for eng in ("warp", "auto"):
    best = min(api.solve_batch_graphs(200, 3, off, edges, node_budget=100000, engine=eng).kernel_ms for _ in range(2))
    r = api.solve_batch_graphs(200, 3, off, edges, node_budget=100000, engine=eng)
    print(f"colouring 1024 x G(200,4.2) k=3 engine={'register' if eng == 'auto' else eng} kernel_ms={best:.2f} Gnodes/s={r.total_nodes / best / 1e6:.2f}", flush=True)
