import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_large():
    """Minutes-of-CPU runs of the unmodified reference (tests/golden/make_golden_large.py): 15/16/17-Queens,
    config C4 as stated (G(200, c/199), k=3/4), the first 10 000 puzzles of the 1 M Sudoku batch."""
    with open(os.path.join(ROOT, "tests", "golden", "reference_large.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def product_lib():
    """The CUDA library must exist (built by __graft_entry__.build()); building needs only nvcc."""
    from dequan_b200 import api
    if not os.path.exists(api.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return api.lib()
