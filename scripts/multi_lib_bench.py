"""Single-process multi-GPU solve (dq_solve_tree_multi): N-Queens all-solutions on 1..G devices of this box.
usage: multi_lib_bench.py [boards,..] [steps]      prints one line per (board, device count)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dequan_b200 import api  # noqa: E402
from dequan_b200.model import nqueens  # noqa: E402

boards = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [17, 18]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
gold = {}
for name in ("reference.json", "reference_large.json"):
    for n, g in json.load(open(os.path.join(ROOT, "tests", "golden", name)))["nqueens"].items():
        gold[int(n)] = (g["count"]["solutions"], g["count"]["nodes"])
G = torch.cuda.device_count()
counts = [c for c in (1, 2, 4, 8) if c <= G]
for n in boards:
    m = api.Model(nqueens(n))
    base = None
    for c in counts:
        devs = tuple(range(c))
        for _ in range(3):
            r = m.solve_tree_multi("count", devs)
        assert (r.solutions, r.nodes) == gold[n], (n, c, r)
        t0 = time.perf_counter()
        kern = 0.0
        for _ in range(steps):
            r = m.solve_tree_multi("count", devs)
            kern += r.kernel_ms
        wall = (time.perf_counter() - t0) / steps * 1e3
        rt = m.solve_tree_multi("count", devs, time_kernels=True)
        base = base or wall
        print(f"nqueens{n} devices={c} wall_ms={wall:.3f} slowest_device_queue_ms={kern / steps:.3f} bucket_kernel_ms={rt.search_kernel_ms:.3f} "
              f"levels_and_copies_ms={rt.kernel_ms - rt.search_kernel_ms:.3f} host_tail_ms={wall - kern / steps:.3f} "
              f"Gnodes/s={r.nodes / wall / 1e6:.1f} speedup={base / wall:.2f} efficiency={base / wall / c:.3f}", flush=True)
