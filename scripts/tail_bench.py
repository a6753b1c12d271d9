"""Host-side tail of a one-process-per-GPU solve: solve_tree call overhead and multi.reduce_tree, timed apart (world size 1 is
enough for the launch / copy / Python costs; the NCCL exchange itself adds its latency at N > 1).  usage: tail_bench.py [n]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from dequan_b200 import api, multi
from dequan_b200.model import nqueens
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
m = api.Model(nqueens(n))
for _ in range(20):
    loc = m.solve_tree("count"); g = multi.reduce_tree(loc, m.nodes_upto, "count", n, device=dev)
R = 200
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(R):
    loc = m.solve_tree("count")
torch.cuda.synchronize(); t1 = time.perf_counter()
for _ in range(R):
    g = multi.reduce_tree(loc, m.nodes_upto, "count", n, device=dev)
torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"nqueens{n}: solve_tree call {1e6 * (t1 - t0) / R:.1f} us (kernel queue {1e3 * loc.kernel_ms:.1f} us), reduce_tree {1e6 * (t2 - t1) / R:.1f} us")
dist.destroy_process_group()
