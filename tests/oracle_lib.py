"""ctypes binding of the CPU oracle (oracle/dq_oracle.c).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from dequan_b200.model import CSP, dq_model_desc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libdq_oracle.so")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "dequan_ref")
U64_MAX = 2**64 - 1


class dqo_result(C.Structure):
    _fields_ = [("outcome", C.c_int32), ("solutions", C.c_uint64), ("nodes", C.c_uint64),
                ("validated_constraints", C.c_uint64), ("applied_arcs", C.c_uint64),
                ("first_key", C.c_uint64), ("n_prefixes", C.c_uint64)]


class dqo_opts(C.Structure):
    _fields_ = [("mode", C.c_int32), ("split_depth", C.c_int32), ("part_rank", C.c_int32),
                ("part_count", C.c_int32), ("node_budget", C.c_uint64), ("upto_key", C.c_uint64)]


_lib = None


def build_oracle() -> str:
    src = os.path.join(ROOT, "oracle", "dq_oracle.c")
    if (not os.path.exists(ORACLE_SO)) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return ORACLE_SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle())
        _lib.dqo_solve.argtypes = [C.POINTER(dq_model_desc), C.POINTER(dqo_opts), C.POINTER(dqo_result),
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        _lib.dqo_solve.restype = C.c_int
    return _lib


@dataclass
class OracleOut:
    status: str
    solutions: int
    nodes: int
    first: Optional[List[int]]
    order: List[int]
    validated_constraints: int
    applied_arcs: int
    first_key: int
    n_prefixes: int


STATUS = {0: "unsat", 1: "sat", 2: "budget"}


def solve(csp: CSP, mode: str = "first", budget: int = 0, split_depth: int = 0, part_rank: int = 0,
          part_count: int = 1, upto_key: int = U64_MAX) -> OracleOut:
    desc, keep = csp.desc()
    nv = len(csp.domains)
    first = np.zeros(max(nv, 1), dtype=np.int32)
    order = np.zeros(max(nv, 1), dtype=np.int32)
    o = dqo_opts(1 if mode == "count" else 0, split_depth, part_rank, part_count, budget, upto_key)
    r = dqo_result()
    rc = lib().dqo_solve(C.byref(desc), C.byref(o), C.byref(r),
                         first.ctypes.data_as(C.POINTER(C.c_int32)), order.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0
    del keep
    have = r.solutions > 0
    return OracleOut(STATUS[r.outcome], r.solutions, r.nodes, first[:nv].tolist() if have else None,
                     order[:nv].tolist(), r.validated_constraints, r.applied_arcs, r.first_key, r.n_prefixes)


def enumerate_solutions(csp: CSP, cap: int = 100000):
    """All solutions in the reference's visiting order -> (list of value lists by var id, total count)."""
    desc, keep = csp.desc()
    nv = len(csp.domains)
    out = np.zeros((max(cap, 1), max(nv, 1)), dtype=np.int32)
    r = dqo_result()
    L = lib()
    L.dqo_enumerate.argtypes = [C.c_void_p, C.POINTER(dqo_result), C.POINTER(C.c_int32), C.c_uint64]
    rc = L.dqo_enumerate(C.byref(desc), C.byref(r), out.ctypes.data_as(C.POINTER(C.c_int32)), cap)
    assert rc == 0
    del keep
    return out[:min(r.solutions, cap), :nv].tolist(), r.solutions
