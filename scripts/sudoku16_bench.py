"""16x16 Sudoku (256 variables on 1..16, AllDifferent over 16 rows, 16 columns, 16 boxes: no structural class, the generic path)
against the CPU oracle: nodes, first solution, wall times.  usage: sudoku16_bench.py [givens,..] [seed]"""
import os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from dequan_b200 import api
from dequan_b200.model import CSP, AllDifferentConstraint


def grid16(seed):
    rng = random.Random(seed)
    base = [[(4 * (r % 4) + r // 4 + c) % 16 for c in range(16)] for r in range(16)]
    perm = list(range(16)); rng.shuffle(perm)
    rows = [4 * b + r for b in rng.sample(range(4), 4) for r in rng.sample(range(4), 4)]
    cols = [4 * b + c for b in rng.sample(range(4), 4) for c in rng.sample(range(4), 4)]
    return [[perm[base[r][c]] + 1 for c in cols] for r in rows]


def model(cells):
    csp = CSP()
    for g in cells:
        csp.AddFixedVar(g) if g else csp.AddIntVar(1, 17)
    for i in range(16):
        csp.AddConstraint(AllDifferentConstraint([16 * i + c for c in range(16)]))
        csp.AddConstraint(AllDifferentConstraint([16 * r + i for r in range(16)]))
        br, bc = 4 * (i // 4), 4 * (i % 4)
        csp.AddConstraint(AllDifferentConstraint([16 * (br + r) + bc + c for r in range(4) for c in range(4)]))
    csp.FinalizeModel()
    return csp


if __name__ == "__main__":
    import oracle_lib as O
    givens = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [150, 130, 115]
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    for g in givens:
        full = [v for row in grid16(seed) for v in row]
        rng = random.Random(seed * 1000 + g)
        keep = set(rng.sample(range(256), g))
        cells = [v if i in keep else 0 for i, v in enumerate(full)]
        csp = model(cells)
        t0 = time.perf_counter(); want = O.solve(csp, "first", 200_000_000); t_cpu = time.perf_counter() - t0
        m = api.Model(csp)
        m.solve_tree("first")
        t0 = time.perf_counter(); r = m.solve_tree("first"); t_gpu = time.perf_counter() - t0
        ok = (r.status, r.nodes, r.first) == (want.status, want.nodes, want.first)
        print(f"sudoku16 givens={g} status={want.status} nodes={want.nodes} parity={'ok' if ok else 'MISMATCH'} engine={r.engine} "
              f"gpu_ms={1e3 * t_gpu:.3f} ({r.nodes / t_gpu / 1e6:.1f} M nodes/s) oracle_cpu_ms={1e3 * t_cpu:.3f} ({want.nodes / max(t_cpu, 1e-9) / 1e6:.1f} M nodes/s, one thread)",
              flush=True)
