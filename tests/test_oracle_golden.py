"""CPU: the C restatement (oracle/dq_oracle.c) against fixtures captured from the unmodified reference."""
import hashlib

import numpy as np
import pytest

import oracle_lib as O
from dequan_b200 import generators as G
from dequan_b200.model import (CSP, REFERENCE_SUDOKU, Op, OpConstraint, colouring, nqueens, sudoku)
from randmodels import model_suite


def test_reference_scenarios(golden):
    # /root/reference/test/main-test.cpp:187-233 OpInequalityTest
    csp = CSP()
    v = [csp.AddIntVar(0, 10), csp.AddIntVar(0, 10), csp.AddFixedVar(6), csp.AddFixedVar(5)]
    csp.AddConstraint(OpConstraint(v[0], v[2], Op.Inf, 0))
    csp.AddConstraint(OpConstraint(v[0], v[3], Op.SupEqual, 0))
    csp.AddConstraint(OpConstraint(v[1], v[2], Op.InfEqual, 0))
    csp.AddConstraint(OpConstraint(v[1], v[3], Op.Sup, 0))
    cases = {
        "OpInequalityTest": csp,
        "NQueensTest8": nqueens(8),
        "SudokuTest_rows_cols_alldiff": sudoku(REFERENCE_SUDOKU, boxes=False, alldiff=True),
        "Sudoku_boxes_alldiff": sudoku(REFERENCE_SUDOKU, boxes=True, alldiff=True),
        "Sudoku_boxes_binary": sudoku(REFERENCE_SUDOKU, boxes=True, alldiff=False),
    }
    for name, model in cases.items():
        g = golden["reference_tests"][name]
        o = O.solve(model, "first")
        assert (o.status, o.nodes, o.first, o.order) == (g["status"], g["nodes"], g["first"], g["order"]), name
        assert (o.applied_arcs, o.validated_constraints) == (g["applied_arcs"], g["validated_constraints"]), name
    assert golden["reference_tests"]["OpInequalityTest"]["first"][:2] == [5, 6]
    assert golden["reference_tests"]["NQueensTest8"]["nodes"] == 88


@pytest.mark.parametrize("n", range(1, 12))
def test_nqueens_counts(golden, n):
    g = golden["nqueens"][str(n)]
    c = O.solve(nqueens(n), "count")
    assert (c.solutions, c.nodes, c.first) == (g["count"]["solutions"], g["count"]["nodes"], g["count"]["first"])
    f = O.solve(nqueens(n), "first")
    assert (f.status, f.nodes, f.first) == (g["first"]["status"], g["first"]["nodes"], g["first"]["first"])


def test_known_counts_oeis(golden):
    # OEIS A000170
    a000170 = [1, 0, 0, 2, 10, 4, 40, 92, 352, 724, 2680, 14200, 73712, 365596]
    for n, want in enumerate(a000170, start=1):
        assert golden["nqueens"][str(n)]["count"]["solutions"] == want


@pytest.mark.parametrize("mode,budget,key", [("first", 0, "first"), ("count", 0, "count"), ("first", 7, "first_budget7"),
                                             ("count", 25, "count_budget25")])
def test_random_suite(golden, mode, budget, key):
    gs = golden["random_suite"]
    suite = model_suite(gs["n"], gs["seed0"])
    h = hashlib.sha256()
    for c in suite:
        h.update(c.to_text().encode())
    assert h.hexdigest() == gs["sha256"], "random model generator drifted from the golden fixtures"
    for i, (csp, g) in enumerate(zip(suite, gs[key])):
        o = O.solve(csp, mode, budget)
        assert (o.status, o.solutions, o.nodes, o.first, o.order) == (g["status"], g["solutions"], g["nodes"], g["first"], g["order"]), i
        assert (o.applied_arcs, o.validated_constraints) == (g["applied_arcs"], g["validated_constraints"]), i


@pytest.mark.parametrize("giv", [30, 40, 24])
def test_sudoku_generator_and_oracle(golden, giv):
    g = golden[f"sudoku_g{giv}"]
    cells = G.sudoku_batch(g["n"], givens=giv, seed=g["seed"])
    assert hashlib.sha256(cells.tobytes()).hexdigest() == g["sha256"]
    n = g["n"] if giv != 24 else 12
    for i in range(n):
        o = O.solve(sudoku(cells[i]), "first")
        assert o.status == g["status"][i] and o.nodes == g["nodes"][i]
        assert "".join(map(str, o.first)) == g["solution"][i]


def test_colouring_oracle(golden):
    for case in golden["colouring"]:
        off, edges = G.colouring_batch(case["count"], case["n_vertices"], case["c"])
        assert hashlib.sha256(edges.tobytes()).hexdigest() == case["sha256"]
        for i in range(case["count"]):
            o = O.solve(colouring(case["n_vertices"], case["k"], edges[off[i]:off[i + 1]]), "first", case["budget"])
            assert (o.status, o.nodes, o.first) == (case["status"][i], case["nodes"][i], case["first"][i])


def test_partition_emulation_sums():
    """The oracle's prefix-partition mode (the multi-GPU split) re-adds to the unsplit run."""
    csp = nqueens(8)
    whole = O.solve(csp, "count")
    for depth in (1, 2, 3):
        for world in (2, 3):
            parts = [O.solve(csp, "count", split_depth=depth, part_rank=r, part_count=world) for r in range(world)]
            assert sum(p.solutions for p in parts) == whole.solutions
            assert sum(p.nodes for p in parts) == whole.nodes
    first = O.solve(csp, "first")
    for depth in (1, 2, 3):
        for world in (2, 3):
            loc = [O.solve(csp, "first", split_depth=depth, part_rank=r, part_count=world) for r in range(world)]
            key = min(p.first_key for p in loc)
            again = [O.solve(csp, "first", split_depth=depth, part_rank=r, part_count=world, upto_key=key) for r in range(world)]
            assert sum(p.nodes for p in again) == first.nodes
            owner = [p for p in again if p.first_key == key][0]
            assert owner.first == first.first


def test_oracle_enumeration_matches_reference():
    """Every solution, in visiting order, as recorded by the unmodified reference (tests/golden/make_enumerate_golden.py)."""
    import json
    import os
    from enum_models import enum_models
    with open(os.path.join(os.path.dirname(__file__), "golden", "enumerate_reference.json")) as f:
        gold = json.load(f)
    for name, csp in enum_models():
        g = gold["models"][name]
        sols, total = O.enumerate_solutions(csp, gold["cap"])
        assert total == g["solutions"], name
        assert sols == g["all"], name
        assert O.solve(csp, "count").nodes == g["nodes"], name


def test_oracle_duplicate_values_vs_reference():
    """Values domains that list a value more than once (SURVEY.md par. 9 Q2): the C restatement against the unmodified
    reference on 150 random models, both modes."""
    import hashlib
    import json
    import os
    from randmodels import dup_suite
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = json.load(open(os.path.join(root, "tests", "golden", "reference_dups.json")))["duplicate_values"]
    suite = dup_suite(g["n"], g["seed0"])
    assert hashlib.sha256("".join(c.to_text() for c in suite).encode()).hexdigest() == g["sha256"]
    for mode in ("first", "count"):
        for i, (csp, want) in enumerate(zip(suite, g[mode])):
            got = O.solve(csp, mode)
            assert (got.status, got.solutions, got.nodes, got.first, got.order) == \
                   (want["status"], want["solutions"], want["nodes"], want["first"], want["order"]), (mode, i)
