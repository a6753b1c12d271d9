"""Register small-tree engine (dq_small_tree.cuh) against the generic warp engine on models of at most 32 variables."""
import json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from dequan_b200 import api
from dequan_b200.model import nqueens, CSP, OpConstraint, Op
from randmodels import random_model
import random

def colouring(nv, k, c, seed):
    rng = random.Random(seed)
    csp = CSP()
    for _ in range(nv):
        csp.AddIntVar(0, k)
    seen = set()
    while len(seen) < int(c * nv / 2):
        a, b = rng.sample(range(nv), 2)
        if (min(a, b), max(a, b)) in seen: continue
        seen.add((min(a, b), max(a, b)))
        csp.AddConstraint(OpConstraint(a, b, Op.NotEqual, 0))
    csp.FinalizeModel()
    return csp

cases = [("queens12_first", nqueens(12), "first"), ("queens13_count_generic", nqueens(13), "count"),
         ("colour30_k3_c3.0_count", colouring(30, 3, 3.0, 1), "count"), ("colour32_k4_c6_count", colouring(32, 4, 6.0, 2), "count"),
         ("colour24_k4_c5_first", colouring(24, 4, 5.0, 3), "first")]
for name, csp, mode in cases:
    m = api.Model(csp)
    row = {"case": name}
    for eng in ("warp", "reg"):
        best = None
        for _ in range(4):
            r = m.solve_tree(mode, engine=eng)
            best = r if best is None or r.kernel_ms < best.kernel_ms else best
        row[eng] = {"ms": round(best.kernel_ms, 3), "nodes": best.nodes, "Mnodes_s": round(best.nodes / best.kernel_ms / 1e3, 1), "sol": best.solutions}
    assert row["warp"]["nodes"] == row["reg"]["nodes"] and row["warp"]["sol"] == row["reg"]["sol"]
    print(json.dumps(row))

# wall-clock latency of one call, the way the drop-in header makes it (engine left to the library)
for name, csp, mode in [("queens8_first", nqueens(8), "first"), ("queens8_count_generic", nqueens(8), "count"), ("queens12_first", nqueens(12), "first"),
                        ("colour24_k4_c5_first", colouring(24, 4, 5.0, 3), "first"), ("colour30_k3_c3.0_count", colouring(30, 3, 3.0, 1), "count")]:
    m = api.Model(csp)
    eng = "warp" if "generic" in name else "auto"
    best = 1e9
    for _ in range(20):
        t0 = time.perf_counter(); r = m.solve_tree(mode, engine=eng); dt = time.perf_counter() - t0
        best = min(best, dt)
    print(json.dumps({"case": name, "wall_us": round(best * 1e6, 1), "kernel_ms": round(r.kernel_ms, 3), "nodes": r.nodes, "engine": r.engine, "launches": r.launches}))
