"""The drop-in C++ header (include/dequan.h) against the unmodified reference header.

tests/cpp/dropin_scenarios.cpp is the reference's own test program (test/main-test.cpp:27-233) plus
further models, written only against the API both headers share.  tests/golden/dropin_*.jsonl is
its output when compiled with the reference header (tests/golden/make_dropin_golden.sh); here it is
compiled with include/dequan.h + libdequan_b200.so and must print the same lines: outcome, solution,
stats.assigned_vars, assign_order and the whole post-solve Assignment (current_domains, saved_domains).
"""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def scen_bin(product_lib, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("dropin") / "dropin_scen")
    libdir = os.path.join(ROOT, "dequan_b200", "lib")
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-o", out,
                           os.path.join(ROOT, "tests", "cpp", "dropin_scenarios.cpp"), "-L" + libdir, "-ldequan_b200",
                           "-Wl,-rpath," + libdir])
    return out


def _lines(text):
    return [json.loads(l) for l in text.splitlines() if l.strip()]


def test_domain_ops_match_reference(scen_bin):
    got = _lines(subprocess.run([scen_bin, "domains"], capture_output=True, text=True, check=True).stdout)
    want = _lines(open(os.path.join(GOLD, "dropin_domains.jsonl")).read())
    assert len(got) == len(want) == 400
    assert got == want


def test_solve_fails_loudly_without_device(scen_bin):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    p = subprocess.run([scen_bin], capture_output=True, text=True)
    assert p.returncode != 0 and "dequan::b200::Error" in p.stderr and p.stdout == ""


@pytest.mark.gpu
def test_scenarios_match_reference(scen_bin):
    got = _lines(subprocess.run([scen_bin], capture_output=True, text=True, check=True, timeout=600).stdout)
    want = _lines(open(os.path.join(GOLD, "dropin_reference.jsonl")).read())
    assert [g["name"] for g in got] == [w["name"] for w in want]
    for g, w in zip(got, want):
        assert g == w, f"scenario {w['name']} differs from the reference"


@pytest.mark.gpu
def test_b200_extensions(product_lib, tmp_path):
    """dequan::b200::CountAll / EnumerateAll (no reference counterpart): counts, node counts and the full solution
    lists in visiting order against the enumeration goldens recorded from the unmodified reference."""
    out = str(tmp_path / "ext")
    libdir = os.path.join(ROOT, "dequan_b200", "lib")
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-o", out,
                           os.path.join(ROOT, "tests", "cpp", "extensions.cpp"), "-L" + libdir, "-ldequan_b200", "-Wl,-rpath," + libdir])
    got = {g["name"]: g for g in _lines(subprocess.run([out], capture_output=True, text=True, check=True, timeout=300).stdout)}
    gold = json.load(open(os.path.join(GOLD, "enumerate_reference.json")))["models"]
    for name in ("nqueens6", "nqueens8", "ordered_values"):
        g, w = got[name], gold[name]
        assert (g["count"], g["nodes"], g["enum_nodes"]) == (w["solutions"], w["nodes"], w["nodes"]), name
        assert g["all"] == w["all"] and g["first_in_a"] and g["small_refused"], name
    assert got["nqueens3"]["count"] == 0 and got["nqueens3"]["all"] == [] and not got["nqueens3"]["first_in_a"]
