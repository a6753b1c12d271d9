"""Multi-GPU reduction tail for a prefix-split single-tree solve.

One process per GPU (torch.distributed, NCCL over NVLink on the GPU box, gloo in the CPU tests).
Prefix i of the DFS-ordered frontier belongs to rank i % world; the search itself needs no
exchange.  What the north star names — the solution-count / node-count sum and the
lexicographic-min first-solution reduction — is one small record per rank:
    {DFS key of the rank's first solution, solutions, nodes, solution vector}
COUNT mode: ONE all_gather of the records and one device->host read; every rank then takes
the sum of the counts and the solution of the lowest key (= a MIN reduction whose payload
rides along, so no second round is needed to find and broadcast from the owner).
FIRST mode: the node count of the reference's sequential search depends on the GLOBAL
minimum key (`nodes_upto`), so the keys are MIN-reduced first and the records gathered
second; the sum then equals the reference's stats.assigned_vars.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional

import torch
import torch.distributed as dist

I64_MAX = 2**63 - 1
I64_MIN = -(2**63)
U64_MAX = 2**64 - 1


@dataclass
class GlobalTreeResult:
    status: str
    solutions: int
    nodes: int
    first: Optional[List[int]]
    first_key: int


def _key_to_i64(k: int) -> int:
    return I64_MAX if k >= I64_MAX else int(k)


_BUFS = {}


def _gather_rows(rec: List[int], device, group):
    """One record per rank -> [world][len(rec)] int64 rows on the host (numpy).  On a GPU the record goes through cached
    pinned / device buffers: one async H2D, one all_gather_into_tensor, one async D2H, one stream sync — the tail is a
    few tens of microseconds next to a solve that takes one to two milliseconds per rank at 8 GPUs."""
    import numpy as np
    world = dist.get_world_size(group)
    n = len(rec)
    dev = torch.device(device)
    if dev.type != "cuda":
        t = torch.tensor(rec, dtype=torch.int64)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t, group=group)
        return torch.stack(out).numpy()
    key = (n, world, str(dev), id(group))
    b = _BUFS.get(key)
    if b is None:
        b = (torch.empty(n, dtype=torch.int64).pin_memory(), torch.empty(n, dtype=torch.int64, device=dev),
             torch.empty(world * n, dtype=torch.int64, device=dev), torch.empty(world * n, dtype=torch.int64).pin_memory())
        _BUFS[key] = b
    pin_in, dev_in, dev_out, pin_out = b
    pin_in.numpy()[:] = rec
    dev_in.copy_(pin_in, non_blocking=True)
    dist.all_gather_into_tensor(dev_out, dev_in, group=group)
    pin_out.copy_(dev_out, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()           # the one device->host read
    return pin_out.numpy().reshape(world, n).copy()


def reduce_tree(local, nodes_upto: Callable[[int], int], mode: str, n_vars: int, device="cpu", group=None) -> GlobalTreeResult:
    """`local` needs .solutions .nodes .first .first_key; nodes_upto(key) -> this rank's share."""
    my_key = _key_to_i64(local.first_key if local.first is not None else U64_MAX)
    if mode == "count":
        nodes = int(local.nodes)
    else:
        key = torch.tensor([my_key], dtype=torch.int64, device=device)
        dist.all_reduce(key, op=dist.ReduceOp.MIN, group=group)
        gk = int(key.item())
        nodes = int(nodes_upto(gk if gk != I64_MAX else U64_MAX))
    rec = [my_key, int(local.solutions) if mode == "count" else 0, nodes]
    rec += [int(v) for v in local.first] if local.first is not None else [I64_MIN] * n_vars
    rows = _gather_rows(rec, device, group)
    owner = int(rows[:, 0].argmin())                    # lowest DFS key; keys of different ranks never tie
    gkey = int(rows[owner, 0])
    have = gkey != I64_MAX
    solutions = int(rows[:, 1].sum()) if mode == "count" else (1 if have else 0)
    return GlobalTreeResult("sat" if solutions else "unsat", solutions, int(rows[:, 2].sum()),
                            [int(v) for v in rows[owner, 3:3 + n_vars]] if have else None, gkey if have else U64_MAX)


def shard_range(n: int, rank: int, world: int):
    """Contiguous instance range of `rank` when a batch of n instances is cut into `world` shards."""
    return rank * n // world, (rank + 1) * n // world


@dataclass
class GlobalBatchTotals:
    n_sat: int
    n_unsat: int
    n_budget: int
    nodes: int


def reduce_batch(n_sat: int, n_unsat: int, n_budget: int, nodes: int, device="cpu", group=None) -> GlobalBatchTotals:
    """The one collective of a sharded batch: every rank solved its own contiguous shard (outputs stay local);
    the totals are summed."""
    acc = torch.tensor([int(n_sat), int(n_unsat), int(n_budget), int(nodes)], dtype=torch.int64, device=device)
    dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return GlobalBatchTotals(*[int(x) for x in acc.tolist()])
