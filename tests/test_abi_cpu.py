"""CPU: the C-ABI library loads, exports every declared symbol, and the host model compiler
(dq_compile = CSP::FinalizeModel + Assignment::Reset) agrees with the reference on ordering."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_lib as O
from dequan_b200 import api
from dequan_b200.model import (CSP, AllDifferentConstraint, Domain, DomainType, Op, OpConstraint, REFERENCE_SUDOKU,
                               colouring, nqueens, sudoku, sudoku_template)
from randmodels import model_suite

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_every_declared_symbol(product_lib):
    hdr = open(os.path.join(ROOT, "include", "dequan_b200.h")).read()
    declared = set(re.findall(r"^(?:int|void|const char \*)\s*\*?(dq_[a-z_0-9]+)\s*\(", hdr, re.M))
    assert declared, "no declarations parsed"
    assert declared == set(api.EXPORTS)
    for name in declared:
        assert hasattr(product_lib, name), name
    assert b"sm_100a" in product_lib.dq_version()


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", api.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out


def test_compile_order_matches_reference(golden, product_lib):
    gs = golden["random_suite"]
    suite = model_suite(gs["n"], gs["seed0"])
    for csp, g in zip(suite, gs["first"]):
        m = api.Model(csp)
        assert m.order() == g["order"]
        m.close()
    m = api.Model(sudoku(REFERENCE_SUDOKU))
    assert m.order() == golden["reference_tests"]["Sudoku_boxes_binary"]["order"]


def test_model_classes(product_lib):
    assert api.Model(nqueens(8)).info()["model_class"] == "queens"
    assert api.Model(nqueens(17)).info() == {"n_vars": 17, "max_dom": 17, "n_arcs": 17 * 16, "model_class": "queens"}
    s = api.Model(sudoku_template()).info()
    assert s["model_class"] == "sudoku9" and s["n_arcs"] == 1620 and s["n_vars"] == 81
    assert api.Model(sudoku(REFERENCE_SUDOKU, alldiff=True)).info()["model_class"] == "generic"  # givens have other value lists
    assert api.Model(colouring(5, 3, [(0, 1), (1, 2)])).info()["model_class"] == "ne_same"
    csp = CSP()
    a, b = csp.AddIntVar(0, 4), csp.AddIntVar(0, 4)
    csp.AddConstraint(OpConstraint(a, b, Op.Inf, 0))
    assert api.Model(csp).info()["model_class"] == "generic"


def test_compile_rejects_out_of_scope(product_lib):
    csp = CSP()
    csp.AddIntVar(0, 70)  # > 64 values
    with pytest.raises(api.DequanError) as e:
        api.Model(csp)
    assert e.value.code == -2
    csp = CSP()
    csp.AddIntVar(Domain(DomainType.Values, [1, 1, 2]))  # duplicate values compile: every copy is a position
    assert api.Model(csp).info()["max_dom"] == 3
    csp = CSP()
    a = csp.AddIntVar(0, 3)
    csp.AddConstraint(OpConstraint(a, 5, Op.Equal, 0))  # bad var id
    with pytest.raises(api.DequanError) as e:
        api.Model(csp)
    assert e.value.code == -1
    csp = CSP()
    a = csp.AddIntVar(0, 3)
    csp.AddConstraint(AllDifferentConstraint([a, a]))
    with pytest.raises(api.DequanError):
        api.Model(csp)


def test_no_cpu_fallback(product_lib):
    """Without a device every solve must fail loudly (DQ_ERR_CUDA), never answer from the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = api.Model(nqueens(6))
    with pytest.raises(api.DequanError) as e:
        m.solve_tree("count")
    assert e.value.code == -3
    with pytest.raises(api.DequanError):
        api.Model(sudoku_template()).solve_batch_cells(np.zeros((2, 81), dtype=np.uint8))


def test_product_does_not_touch_oracle():
    """The product sources never load, link or execute anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dequan_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".sh")):
                txt = open(os.path.join(dirpath, f)).read()
                for needle in ("libdq_oracle", "oracle_lib", "dqo_solve", "dequan_ref", "dq_oracle.c", "import oracle"):
                    assert needle not in txt, (f, needle)


def test_explicit_assign_order_is_validated_and_kept():
    from dequan_b200 import api
    from dequan_b200.model import nqueens
    csp = nqueens(5)
    csp.assign_order = [4, 3, 2, 1, 0]
    assert api.Model(csp).order() == [4, 3, 2, 1, 0]
    assert O.solve(csp, "count").order == [4, 3, 2, 1, 0]
    assert O.solve(csp, "count").solutions == 10
    csp.assign_order = [0, 1, 1, 2, 3]
    with pytest.raises(api.DequanError):
        api.Model(csp)


def test_on_disk_formats():
    from dequan_b200 import api, generators as G
    cells = G.sudoku_batch(5, givens=30)
    text = "# five puzzles\n" + "\n".join(G.sudoku_lines(cells)).replace("0", ".", 7) + "\n\n"
    assert (api.parse_sudoku_lines(text) == cells).all()
    assert api.parse_sudoku_lines("").shape == (0, 81)
    with pytest.raises(api.DequanError):
        api.parse_sudoku_lines("123\n")
    with pytest.raises(api.DequanError):
        api.parse_sudoku_lines("x" * 81 + "\n")
    nv, edges = api.parse_dimacs_col("c triangle plus a tail\np edge 4 4\ne 1 2\ne 2 3\ne 1 3\ne 3 4\n")
    assert nv == 4 and edges.tolist() == [[0, 1], [1, 2], [0, 2], [2, 3]]
    with pytest.raises(api.DequanError):
        api.parse_dimacs_col("p edge 3 1\ne 1 4\n")
    with pytest.raises(api.DequanError):
        api.parse_dimacs_col("e 1 2\n")
