"""Where a drop-in-sized FIRST-mode solve spends its time: the reference's SudokuTest (rows + columns only) and friends."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dequan_b200 import api
from dequan_b200.model import sudoku, REFERENCE_SUDOKU, nqueens
for name, csp, mode in (("SudokuTest rows+cols alldiff", sudoku(REFERENCE_SUDOKU, boxes=False, alldiff=True), "first"),
                        ("Sudoku boxes binary", sudoku(REFERENCE_SUDOKU, boxes=True, alldiff=False), "first"),
                        ("nqueens12 first", nqueens(12), "first"), ("nqueens20 first", nqueens(20), "first"), ("nqueens24 first", nqueens(24), "first")):
    m = api.Model(csp)
    for eng in ("auto", "warp"):
        for _ in range(2):
            r = m.solve_tree(mode, engine=eng)
        t0 = time.perf_counter()
        for _ in range(3):
            r = m.solve_tree(mode, engine=eng)
        wall = (time.perf_counter() - t0) / 3 * 1e3
        print(f"{name:30s} engine={eng:5s} used={r.engine:5s} nodes={r.nodes:9d} split={r.split_depth:3d} prefixes={r.n_prefixes:7d} launches={r.launches:4d} "
              f"device_span_ms={r.kernel_ms:7.3f} search_ms={r.search_kernel_ms:7.3f} wall_ms={wall:7.3f}", flush=True)
