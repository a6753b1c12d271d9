"""Per-partition kernel time of the 17-Queens tree for 1/2/4/8 partitions run one after another on one GPU."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dequan_b200 import api
from dequan_b200.model import nqueens
n = int(sys.argv[1]) if len(sys.argv) > 1 else 17
m = api.Model(nqueens(n))
for world in (1, 2, 4, 8):
    tot_s = tot_n = 0; times = []
    for r in range(world):
        m.solve_tree("count", part_rank=r, part_count=world)
        x = m.solve_tree("count", part_rank=r, part_count=world)
        tot_s += x.solutions; tot_n += x.nodes; times.append(x.kernel_ms)
    print(f"parts={world} sols={tot_s} nodes={tot_n} max_ms={max(times):.3f} min_ms={min(times):.3f} mean_ms={sum(times)/world:.3f}", flush=True)
