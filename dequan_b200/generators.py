"""Seeded synthetic instances of BASELINE.json's configs.

All randomness is splitmix64 keyed by (seed, instance index, draw index), so the same
instance can be regenerated anywhere (numpy here; trivially restated in C/CUDA).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """One splitmix64 output per input counter value (vectorised, wraps mod 2^64)."""
    with np.errstate(over="ignore"):
        z = (x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _keys(seed: int, idx: np.ndarray, stream: int, width: int) -> np.ndarray:
    """[len(idx), width] independent u64 keys for instance indices `idx`."""
    with np.errstate(over="ignore"):
        base = splitmix64(np.uint64(seed) ^ splitmix64(idx.astype(np.uint64) * np.uint64(0x100000001B3) + np.uint64(stream)))
        ctr = base[:, None] + np.arange(1, width + 1, dtype=np.uint64)[None, :] * np.uint64(0xD1342543DE82EF95)
        return splitmix64(ctr)


def _perm(seed: int, idx: np.ndarray, stream: int, width: int) -> np.ndarray:
    return np.argsort(_keys(seed, idx, stream, width), axis=1, kind="stable")


def sudoku_batch(n: int, givens: int = 30, seed: int = 20261018, start: int = 0, chunk: int = 65536) -> np.ndarray:
    """n puzzles as uint8[n, 81], 0 = blank.  Pattern grid g[r][c] = ((3r + r//3 + c) % 9) + 1,
    shuffled by digit relabelling, rows within bands, bands, columns within stacks, stacks;
    then `givens` cells chosen uniformly are kept."""
    out = np.zeros((n, 81), dtype=np.uint8)
    r = np.arange(9)
    base = ((3 * r[:, None] + r[:, None] // 3 + r[None, :]) % 9).astype(np.int64)  # digits 0..8
    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        idx = np.arange(start + c0, start + c0 + m, dtype=np.uint64)
        digit = _perm(seed, idx, 1, 9)                                  # [m,9]
        bands = _perm(seed, idx, 2, 3)                                  # [m,3]
        rows_in = np.stack([_perm(seed, idx, 3 + b, 3) for b in range(3)], axis=1)   # [m,3,3]
        stacks = _perm(seed, idx, 6, 3)
        cols_in = np.stack([_perm(seed, idx, 7 + b, 3) for b in range(3)], axis=1)
        # row_map[m, 9]: destination row i takes source row bands[i//3]*3 + rows_in[i//3][i%3]
        row_map = (bands[:, :, None] * 3 + rows_in).reshape(m, 9)
        col_map = (stacks[:, :, None] * 3 + cols_in).reshape(m, 9)
        g = base[row_map[:, :, None], col_map[:, None, :]]               # [m,9,9]
        g = np.take_along_axis(digit, g.reshape(m, 81), axis=1) + 1      # relabel -> 1..9
        keep = _perm(seed, idx, 10, 81)[:, :givens]                      # cells kept as givens
        mask = np.zeros((m, 81), dtype=bool)
        np.put_along_axis(mask, keep, True, axis=1)
        out[c0:c0 + m] = np.where(mask, g, 0).astype(np.uint8)
    return out


def sudoku_lines(cells: np.ndarray) -> List[str]:
    return ["".join(chr(48 + int(v)) for v in row) for row in cells]


def max_cardinality_order(n: int, edges: np.ndarray) -> np.ndarray:
    """Maximum-cardinality search numbering: repeatedly pick the unnumbered vertex with most
    numbered neighbours (ties -> smallest id).  Returns new_id[old_id]."""
    adj: List[List[int]] = [[] for _ in range(n)]
    for u, v in edges:
        adj[int(u)].append(int(v))
        adj[int(v)].append(int(u))
    weight = np.zeros(n, dtype=np.int64)
    done = np.zeros(n, dtype=bool)
    new_id = np.zeros(n, dtype=np.int64)
    for k in range(n):
        w = np.where(done, -1, weight)
        u = int(np.argmax(w))
        done[u] = True
        new_id[u] = k
        for v in adj[u]:
            weight[v] += 1
    return new_id


def colouring_instance(n: int, c: float, seed: int, index: int, renumber: bool = True) -> np.ndarray:
    """G(n, p = c/(n-1)) edge list uint8[m,2] (u < v), vertices renumbered by maximum-cardinality order
    (with id order dequan cannot solve these; the numbering is part of the generator)."""
    iu, ju = np.triu_indices(n, 1)
    keys = _keys(seed, np.array([index], dtype=np.uint64), 21, iu.size)[0]
    p = c / (n - 1)
    thresh = np.uint64(int(p * 2.0**64)) if p < 1.0 else _M64
    sel = keys < thresh
    edges = np.stack([iu[sel], ju[sel]], axis=1)
    if renumber and edges.size:
        nid = max_cardinality_order(n, edges)
        edges = nid[edges]
        edges = np.stack([edges.min(axis=1), edges.max(axis=1)], axis=1)
        edges = edges[np.lexsort((edges[:, 1], edges[:, 0]))]
    return edges.astype(np.uint8 if n <= 256 else np.int32)


def colouring_batch(count: int, n: int, c: float, seed: int = 20261018, start: int = 0, chunk: int = 1024) -> Tuple[np.ndarray, np.ndarray]:
    """CSR batch: edge_off int64[count+1], edges uint8[total,2].  Instance i is `colouring_instance(n, c, seed, start + i)`;
    the maximum-cardinality numbering runs on `chunk` instances at a time (dense adjacency, vectorised over the chunk)."""
    assert n <= 256
    iu, ju = np.triu_indices(n, 1)
    p = c / (n - 1) if n > 1 else 0.0
    thresh = np.uint64(int(p * 2.0**64)) if p < 1.0 else _M64
    lists: List[np.ndarray] = []
    for c0 in range(0, count, chunk):
        m = min(chunk, count - c0)
        if iu.size == 0:
            lists.extend(np.zeros((0, 2), dtype=np.uint8) for _ in range(m))
            continue
        idx = np.arange(start + c0, start + c0 + m, dtype=np.uint64)
        sel = _keys(seed, idx, 21, iu.size) < thresh                      # [m, n(n-1)/2]
        adj = np.zeros((m, n, n), dtype=bool)
        adj[:, iu, ju] = sel
        adj |= adj.transpose(0, 2, 1)
        # maximum-cardinality search on all m graphs at once: pick the unnumbered vertex with most numbered neighbours
        # (ties -> smallest id, np.argmax returns the first maximum)
        weight = np.zeros((m, n), dtype=np.int64)
        done = np.zeros((m, n), dtype=bool)
        new_id = np.zeros((m, n), dtype=np.int64)
        rows = np.arange(m)
        for step in range(n):
            u = np.argmax(np.where(done, -1, weight), axis=1)
            done[rows, u] = True
            new_id[rows, u] = step
            weight += adj[rows, u]
        for i in range(m):
            e = np.stack([iu[sel[i]], ju[sel[i]]], axis=1)
            if e.size:
                e = new_id[i][e]
                e = np.stack([e.min(axis=1), e.max(axis=1)], axis=1)
                e = e[np.lexsort((e[:, 1], e[:, 0]))]
            lists.append(e.astype(np.uint8))
    off = np.zeros(count + 1, dtype=np.int64)
    for i, e in enumerate(lists):
        off[i + 1] = off[i] + len(e)
    edges = np.concatenate(lists, axis=0) if lists else np.zeros((0, 2), dtype=np.uint8)
    return off, np.ascontiguousarray(edges.astype(np.uint8))


def graph_lines(edge_off: np.ndarray, edges: np.ndarray, n: int) -> List[str]:
    out = []
    for i in range(len(edge_off) - 1):
        e = edges[edge_off[i]:edge_off[i + 1]]
        out.append(f"{n} {len(e)} " + " ".join(f"{int(u)} {int(v)}" for u, v in e))
    return out
