"""Sudoku batch timing on one engine.  usage: sudoku_bench.py [n] [engine] [task_nodes] [givens]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from dequan_b200 import api, generators as G
from dequan_b200.model import sudoku_template
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
eng = sys.argv[2] if len(sys.argv) > 2 else "lane"
tn = int(sys.argv[3]) if len(sys.argv) > 3 else 0
giv = int(sys.argv[4]) if len(sys.argv) > 4 else 30
cells = G.sudoku_batch(n, giv)
t = api.Model(sudoku_template())
for rep in range(3):
    r = t.solve_batch_cells(cells, engine=eng, task_nodes=tn)
    print(f"n={n} engine={eng} task_nodes={tn} givens={giv} ms={r.kernel_ms:.2f} launches={r.launches} sat={r.n_sat} nodes={r.total_nodes} "
          f"Mpuzzles/s={n / r.kernel_ms / 1e3:.2f} Gnodes/s={r.total_nodes / r.kernel_ms / 1e6:.2f}", flush=True)
