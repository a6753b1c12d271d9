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


def _gather_rows(rec: torch.Tensor, group) -> torch.Tensor:
    world = dist.get_world_size(group)
    out = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(out, rec, group=group)
    return torch.stack(out).cpu()                      # the one device->host read


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
    rows = _gather_rows(torch.tensor(rec, dtype=torch.int64, device=device), group)
    owner = int(torch.argmin(rows[:, 0]).item())        # lowest DFS key; keys of different ranks never tie
    gkey = int(rows[owner, 0].item())
    have = gkey != I64_MAX
    solutions = int(rows[:, 1].sum().item()) if mode == "count" else (1 if have else 0)
    return GlobalTreeResult("sat" if solutions else "unsat", solutions, int(rows[:, 2].sum().item()),
                            rows[owner, 3:3 + n_vars].tolist() if have else None, gkey if have else U64_MAX)


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
