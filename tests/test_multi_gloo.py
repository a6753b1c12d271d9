"""CPU: world_size-2/3 gloo runs of the multi-GPU reduction tail (dequan_b200/multi.py), with the
oracle's prefix-partition mode standing in for each rank's GPU solve."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from dequan_b200 import multi
from dequan_b200.model import nqueens
from randmodels import random_model


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cases, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = []
    for (kind, arg, mode, depth) in cases:
        csp = nqueens(arg) if kind == "q" else random_model(arg, 7, 9, 5, "ne")
        local = O.solve(csp, mode, split_depth=depth, part_rank=rank, part_count=world)
        upto = lambda key: O.solve(csp, mode, split_depth=depth, part_rank=rank, part_count=world, upto_key=key).nodes
        g = multi.reduce_tree(local, upto, mode, len(csp.domains))
        res.append((g.status, g.solutions, g.nodes, g.first))
    if rank == 0:
        out.put(res)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_reduce_tree_matches_unsplit(world):
    cases = [("q", 8, "count", 2), ("q", 8, "first", 2), ("q", 6, "first", 3), ("q", 9, "count", 3), ("q", 3, "count", 1),
             ("r", 1005, "count", 2), ("r", 1010, "first", 2), ("r", 1015, "first", 1)]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cases, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for (kind, arg, mode, depth), g in zip(cases, got):
        csp = nqueens(arg) if kind == "q" else random_model(arg, 7, 9, 5, "ne")
        want = O.solve(csp, mode)
        assert g == (want.status, want.solutions, want.nodes, want.first), (kind, arg, mode, depth, g, want)


def _batch_worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dequan_b200 import generators as G
    from dequan_b200.model import sudoku
    lo, hi = multi.shard_range(n, rank, world)
    cells = G.sudoku_batch(hi - lo, givens=36, seed=5, start=lo)          # each rank generates and solves only its shard
    outs = [O.solve(sudoku(c), "first") for c in cells]
    g = multi.reduce_batch(sum(o.status == "sat" for o in outs), sum(o.status == "unsat" for o in outs), 0, sum(o.nodes for o in outs))
    if rank == 0:
        out.put((g.n_sat, g.n_unsat, g.n_budget, g.nodes))
    dist.destroy_process_group()


def test_sharded_batch_totals():
    """Batches shard by contiguous instance ranges; shards regenerate the same instances the whole batch holds."""
    from dequan_b200 import generators as G
    from dequan_b200.model import sudoku
    n, world = 24, 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_batch_worker, args=(r, world, port, n, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = [O.solve(sudoku(c), "first") for c in G.sudoku_batch(n, givens=36, seed=5)]
    assert got == (sum(o.status == "sat" for o in whole), 0, 0, sum(o.nodes for o in whole))
    assert [multi.shard_range(10, r, 3) for r in range(3)] == [(0, 3), (3, 6), (6, 10)]
