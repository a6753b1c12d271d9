"""Quick on-GPU diagnostics (not a test): prints engine vs oracle for a handful of cases."""
import sys, os, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from dequan_b200 import api, generators as G
from dequan_b200.model import nqueens, sudoku, sudoku_template, colouring, REFERENCE_SUDOKU
from randmodels import model_suite

def case(name, fn):
    try:
        t = time.time(); r = fn(); print(f"[{name}] {time.time()-t:.3f}s {r}", flush=True)
    except Exception as e:
        print(f"[{name}] EXC {e}", flush=True); traceback.print_exc()

print(api.device_info())
case("int_peak", lambda: api.measure_int_peak())
engines = sys.argv[1].split(",") if len(sys.argv) > 1 else ["warp"]
for eng in engines:
    for n in (2, 3, 4, 6, 8, 10, 12):
        csp = nqueens(n)
        def f(mode):
            g = api.Model(csp).solve_tree(mode, engine=eng); w = O.solve(csp, mode)
            ok = (g.status, g.solutions, g.nodes, g.first) == (w.status, w.solutions, w.nodes, w.first)
            return ("OK" if ok else "MISMATCH", g, (w.solutions, w.nodes, w.first))
        case(f"{eng} q{n} count", lambda: f("count"))
        if eng != "lane": case(f"{eng} q{n} first", lambda: f("first"))
    for n in (14, 15, 16, 17):
        mm = api.Model(nqueens(n))
        case(f"{eng} q{n} count", lambda: mm.solve_tree("count", engine=eng))
        case(f"{eng} q{n} count again", lambda: mm.solve_tree("count", engine=eng))
if "--quick" in sys.argv: sys.exit(0)
def suite():
    bad = 0
    for i, csp in enumerate(model_suite(120)):
        for mode in ("first", "count"):
            g = api.Model(csp).solve_tree(mode); w = O.solve(csp, mode)
            if (g.status, g.solutions, g.nodes, g.first) != (w.status, w.solutions, w.nodes, w.first):
                bad += 1
                if bad < 6: print("  mismatch", i, mode, g, w)
    return f"mismatches={bad}"
case("random suite", suite)
def sud():
    cells = G.sudoku_batch(2000, 30)
    r = api.Model(sudoku_template()).solve_batch_cells(cells)
    bad = 0
    for i in range(0, 2000, 10):
        o = O.solve(sudoku(cells[i]), "first")
        if (int(r.nodes[i]), r.solution[i].tolist()) != (o.nodes, o.first): bad += 1
    return f"bad={bad} sat={r.n_sat} nodes={r.total_nodes} ms={r.kernel_ms:.3f}"
case("sudoku 2000", sud)
def sudbig():
    cells = G.sudoku_batch(1_000_000, 30)
    t = api.Model(sudoku_template())
    r = t.solve_batch_cells(cells)
    r = t.solve_batch_cells(cells)
    return f"sat={r.n_sat} nodes={r.total_nodes} ms={r.kernel_ms:.3f} puzzles/s={1e6/(r.kernel_ms*1e-3):.3e} nodes/s={r.total_nodes/(r.kernel_ms*1e-3):.3e}"
case("sudoku 1M", sudbig)
def col():
    off, e = G.colouring_batch(64, 200, 3.6, seed=11)
    r = api.solve_batch_graphs(200, 3, off, e, node_budget=50000)
    bad = 0
    for i in range(0, 64, 4):
        o = O.solve(colouring(200, 3, e[off[i]:off[i+1]]), "first", 50000)
        if (api.OUTCOME[r.status[i]], int(r.nodes[i])) != (o.status, o.nodes): bad += 1
    return f"bad={bad} sat={r.n_sat} unsat={r.n_unsat} budget={r.n_budget} nodes={r.total_nodes} ms={r.kernel_ms:.3f}"
case("colouring", col)
