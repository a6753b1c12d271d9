"""Colouring batch sweep over engines / lanes per instance.  usage: colour_sweep.py [c] [k] [budget] [counts,..] [groups,..]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from dequan_b200 import api, generators as G  # noqa: E402

c = float(sys.argv[1]) if len(sys.argv) > 1 else 4.2
k = int(sys.argv[2]) if len(sys.argv) > 2 else 3
budget = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000
counts = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1024, 8192]
groups = sys.argv[5].split(",") if len(sys.argv) > 5 else ["lane", "reg"]
base = min(max(counts), 8192)
t = time.time()
off0, edges0 = G.colouring_batch(base, 200, c)
print(f"generated {base} instances, {off0[-1]} edges in {time.time() - t:.1f}s", flush=True)
for count in counts:
    if count <= base:
        off, edges = off0[:count + 1], edges0[:off0[count]]
    else:   # larger batches repeat the generated instances
        reps = count // base
        edges = np.ascontiguousarray(np.tile(edges0, (reps, 1)))
        off = np.concatenate([[0], (np.arange(reps)[:, None] * off0[-1] + off0[1:][None, :]).ravel()]).astype(np.int64)
    for g in groups:
        engine = g
        best = None
        for rep in range(3):
            r = api.solve_batch_graphs(200, k, off, edges, node_budget=budget, engine=engine)
            if best is None or r.kernel_ms < best.kernel_ms:
                best = r
        r = best
        print(f"count={count} c={c} k={k} budget={budget} group={g} ms={r.kernel_ms:.3f} search_ms={r.search_kernel_ms:.3f} sat={r.n_sat} unsat={r.n_unsat} "
              f"budget_hit={r.n_budget} nodes={r.total_nodes} inst/s={count / r.kernel_ms * 1e3:.0f} Gnodes/s={r.total_nodes / r.kernel_ms / 1e6:.2f}", flush=True)
