"""Regenerates tests/golden/reference.json by running the UNMODIFIED reference
(oracle/_ref/dequan_ref, built from /root/reference/dequan.h by oracle/Makefile).
Run in the build container only (the GPU box has no /root/reference):
    make -C oracle ref && python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from dequan_b200 import generators as G  # noqa: E402
from dequan_b200.model import colouring, nqueens, sudoku  # noqa: E402
from randmodels import model_suite  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "dequan_ref")
KEEP = ("status", "solutions", "nodes", "first", "order", "applied_arcs", "validated_constraints")


def ref_solve(csps, mode, budget=0):
    txt = "".join(c.to_text() for c in csps)
    out = subprocess.run([REF, "solve", mode, str(budget)], input=txt, capture_output=True, text=True, check=True).stdout
    res = [json.loads(line) for line in out.splitlines()]
    assert len(res) == len(csps)
    return [{k: r[k] for k in KEEP} for r in res]


def digest(csps):
    h = hashlib.sha256()
    for c in csps:
        h.update(c.to_text().encode())
    return h.hexdigest()


def main():
    gold = {"generated_by": "tests/golden/make_golden.py", "reference": "nsweb/dequan dequan.h (unmodified)"}
    out = subprocess.run([REF, "tests"], capture_output=True, text=True, check=True).stdout
    gold["reference_tests"] = {}
    for line in out.splitlines():
        o = json.loads(line)
        gold["reference_tests"][o["test"]] = {k: o["result"][k] for k in KEEP}

    gold["nqueens"] = {}
    for n in range(1, 15):
        m = nqueens(n)
        c = ref_solve([m], "count")[0]
        f = ref_solve([m], "first")[0]
        gold["nqueens"][str(n)] = {"count": {k: c[k] for k in ("solutions", "nodes", "first")},
                                   "first": {k: f[k] for k in ("status", "nodes", "first")}}
        print("nqueens", n, gold["nqueens"][str(n)]["count"]["solutions"], gold["nqueens"][str(n)]["count"]["nodes"], flush=True)

    suite = model_suite(300)
    gold["random_suite"] = {"n": 300, "seed0": 1000, "sha256": digest(suite),
                            "first": ref_solve(suite, "first"), "count": ref_solve(suite, "count"),
                            "first_budget7": ref_solve(suite, "first", 7), "count_budget25": ref_solve(suite, "count", 25)}

    for giv, cnt in ((30, 400), (24, 60), (40, 100)):
        cells = G.sudoku_batch(cnt, givens=giv)
        models = [sudoku(row) for row in cells]
        res = ref_solve(models, "first")
        gold[f"sudoku_g{giv}"] = {"n": cnt, "givens": giv, "seed": 20261018, "sha256": hashlib.sha256(cells.tobytes()).hexdigest(),
                                  "nodes": [r["nodes"] for r in res], "status": [r["status"] for r in res],
                                  "solution": ["".join(map(str, r["first"])) if r["first"] else None for r in res]}
        print("sudoku", giv, sum(r["nodes"] for r in res) / cnt, flush=True)

    gold["colouring"] = []
    for (k, c, cnt, budget) in ((3, 3.0, 12, 20000), (3, 4.2, 12, 20000), (4, 6.0, 8, 20000)):
        off, edges = G.colouring_batch(cnt, 60, c)
        models = [colouring(60, k, edges[off[i]:off[i + 1]]) for i in range(cnt)]
        res = ref_solve(models, "first", budget)
        gold["colouring"].append({"n_vertices": 60, "k": k, "c": c, "count": cnt, "budget": budget,
                                  "sha256": hashlib.sha256(edges.tobytes()).hexdigest(),
                                  "status": [r["status"] for r in res], "nodes": [r["nodes"] for r in res],
                                  "first": [r["first"] for r in res]})
        print("colouring", k, c, [r["status"] for r in res], flush=True)

    with open(os.path.join(ROOT, "tests", "golden", "reference.json"), "w") as f:
        json.dump(gold, f, separators=(",", ":"))
    print("wrote reference.json", os.path.getsize(os.path.join(ROOT, "tests", "golden", "reference.json")))


if __name__ == "__main__":
    main()
