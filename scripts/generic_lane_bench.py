"""Generic-path engines on small trees.  usage: generic_lane_bench.py [n]   (run with DQ_NO_CLASS=1 for the plain N-Queens rows)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dequan_b200 import api
from dequan_b200.model import nqueens, Op, OpConstraint, colouring
from dequan_b200 import generators as G

def bench(name, csp, engines=("auto", "reg", "warp")):
    m = api.Model(csp)
    for eng in engines:
        try:
            for _ in range(2):
                r = m.solve_tree("count", engine=eng)
            t0 = time.perf_counter()
            for _ in range(3):
                r = m.solve_tree("count", engine=eng)
            wall = (time.perf_counter() - t0) / 3 * 1e3
            print(f"{name:34s} engine={eng:5s} used={r.engine:5s} split={r.split_depth:2d} prefixes={r.n_prefixes:8d} nodes={r.nodes:11d} sols={r.solutions:9d} "
                  f"kernel_ms={r.kernel_ms:8.3f} search_ms={r.search_kernel_ms:8.3f} wall_ms={wall:8.3f} Gnodes/s(kernel)={r.nodes / r.kernel_ms / 1e6:7.2f}", flush=True)
        except api.DequanError as e:
            print(f"{name:34s} engine={eng:5s} -> {e}", flush=True)

ns = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [12, 13, 14]
for n in ns:
    bench(f"nqueens{n}" + (" (DQ_NO_CLASS)" if os.environ.get("DQ_NO_CLASS") else " (class engine on auto)"), nqueens(n))
    c = nqueens(n); c.AddConstraint(OpConstraint(0, n - 1, Op.Inf, 0)); c.FinalizeModel()
    bench(f"nqueens{n} + q0 < q{n-1}", c)
e = G.colouring_instance(30, 3.2, 5, 1)
bench("colouring 30 vertices k=3 count", colouring(30, 3, e))
e = G.colouring_instance(32, 4.5, 5, 2)
bench("colouring 32 vertices k=4 count", colouring(32, 4, e))
