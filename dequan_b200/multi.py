"""Multi-GPU reduction tail for a prefix-split single-tree solve.

One process per GPU (torch.distributed, NCCL over NVLink on the GPU box, gloo in the CPU tests).
Prefix i of the DFS-ordered frontier belongs to rank i % world; the search itself needs no
exchange.  The only collectives are the ones the north star names:
  * all_reduce(SUM)  of {solutions, nodes}
  * all_reduce(MIN)  of the DFS index of the prefix holding each rank's first solution
  * all_reduce(MAX)  of the owner's solution vector (everyone else contributes INT64_MIN),
    which is a broadcast from the owner without a second round to discover who the owner is.
In FIRST mode the node count is re-asked per rank for the GLOBAL minimum key
(`nodes_upto`), so the sum equals the reference's sequential stats.assigned_vars.
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


def reduce_tree(local, nodes_upto: Callable[[int], int], mode: str, n_vars: int, device="cpu", group=None) -> GlobalTreeResult:
    """`local` needs .solutions .nodes .first .first_key; nodes_upto(key) -> this rank's share."""
    key = torch.tensor([_key_to_i64(local.first_key if local.first is not None else U64_MAX)], dtype=torch.int64, device=device)
    dist.all_reduce(key, op=dist.ReduceOp.MIN, group=group)
    gkey = int(key.item())
    have = gkey != I64_MAX
    if mode == "count":
        acc = torch.tensor([int(local.solutions), int(local.nodes)], dtype=torch.int64, device=device)
    else:
        acc = torch.tensor([0, int(nodes_upto(gkey if have else U64_MAX))], dtype=torch.int64, device=device)
    dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    sol = torch.full((max(n_vars, 1),), I64_MIN, dtype=torch.int64, device=device)
    if have and local.first is not None and _key_to_i64(local.first_key) == gkey:
        sol[:n_vars] = torch.tensor(local.first, dtype=torch.int64, device=device)
    dist.all_reduce(sol, op=dist.ReduceOp.MAX, group=group)
    solutions = int(acc[0].item()) if mode == "count" else (1 if have else 0)
    return GlobalTreeResult("sat" if solutions else "unsat", solutions, int(acc[1].item()),
                            sol[:n_vars].tolist() if have else None, gkey if have else U64_MAX)


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
