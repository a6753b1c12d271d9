"""Regenerates tests/golden/reference_large.json by running the UNMODIFIED reference
(oracle/_ref/dequan_ref, built from /root/reference/dequan.h by oracle/Makefile) on the
BASELINE.json shapes that make_golden.py leaves out because they take minutes of CPU:

  * N-Queens N = 15, 16, 17 (and, section queens18, 18) all-solutions (solutions, nodes = stats.assigned_vars, DFS-first solution);
    the tree is split on the first variable's value with singleton domains, one reference solve per
    value (SURVEY.md §8c "parallel CPU split"), and the per-value results are kept as well;
  * config C4 as stated: G(200, c/199), k=3 c in {4.0, 4.2, 4.4, 4.69}, k=4 c in {6, 7}, 64 instances
    each, node budget 100 000 (status, nodes, colours per instance);
  * config C3 per-instance: the first 10 000 puzzles of the 1 M batch (30 givens): nodes + solution;
  * 150 random models whose Values domains list values more than once (SURVEY.md par. 9 Q2), first + count.

Run in the build container only (the GPU box has no /root/reference):
    make -C oracle ref && python tests/golden/make_golden_large.py [section ...]
Sections: queens colouring sudoku (default: all), dups, queens18 (18-Queens, about two hours on six cores).  An existing file is updated section by section.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from dequan_b200 import generators as G  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "dequan_ref")
OUT = os.path.join(ROOT, "tests", "golden", "reference_large.json")
THREADS = int(os.environ.get("DQ_GOLDEN_THREADS", "6"))

COLOURING = [(3, 4.0), (3, 4.2), (3, 4.4), (3, 4.69), (4, 6.0), (4, 7.0)]
COLOURING_COUNT = 64
COLOURING_BUDGET = 100000
SUDOKU_COUNT = 10000
SUDOKU_GIVENS = 30


def run(args):
    out = subprocess.run([REF] + [str(a) for a in args], capture_output=True, text=True, check=True).stdout
    return [json.loads(line) for line in out.splitlines()]


def queens(gold, sizes=(15, 16, 17)):
    from concurrent.futures import ThreadPoolExecutor
    gold.setdefault("nqueens", {})
    for n in sizes:
        # one single-threaded reference solve per first-row value v (variable 0 gets the singleton domain {v},
        # stays first in assign_order): nodes = 1 + the subtree below, so the sum over v is the whole tree.
        def task(v):
            r = run(["nqueens", n, "count", 1, v])[0]
            print("nqueens", n, "first value", v, r["solutions"], r["nodes"], f'{r["seconds"]:.1f}s', flush=True)
            return {"solutions": r["solutions"], "nodes": r["nodes"], "first": r["first"]}
        with ThreadPoolExecutor(THREADS) as ex:
            per_value = list(ex.map(task, range(n)))
        first = next(p["first"] for p in per_value if p["first"])
        gold["nqueens"][str(n)] = {"count": {"solutions": sum(p["solutions"] for p in per_value),
                                             "nodes": sum(p["nodes"] for p in per_value), "first": first},
                                   "per_first_value": [{"solutions": p["solutions"], "nodes": p["nodes"]} for p in per_value]}
        print("nqueens", n, gold["nqueens"][str(n)]["count"], flush=True)
        save(gold)


def colouring(gold):
    gold["colouring200"] = []
    for k, c in COLOURING:
        off, edges = G.colouring_batch(COLOURING_COUNT, 200, c)
        with tempfile.NamedTemporaryFile("w", suffix=".graphs", delete=False) as f:
            f.write("\n".join(G.graph_lines(off, edges, 200)) + "\n")
        res = run(["color", f.name, k, COLOURING_BUDGET, THREADS])
        os.unlink(f.name)
        res = [r for r in res if "summary" not in r]
        assert len(res) == COLOURING_COUNT
        gold["colouring200"].append({"n_vertices": 200, "k": k, "c": c, "count": COLOURING_COUNT, "budget": COLOURING_BUDGET,
                                     "seed": 20261018, "sha256": hashlib.sha256(edges.tobytes()).hexdigest(),
                                     "status": [r["status"] for r in res], "nodes": [r["nodes"] for r in res],
                                     "first": [r["first"] for r in res]})
        st = [r["status"] for r in res]
        print("colouring", k, c, {s: st.count(s) for s in set(st)}, sum(r["nodes"] for r in res), flush=True)
    save(gold)


def sudoku(gold):
    cells = G.sudoku_batch(SUDOKU_COUNT, givens=SUDOKU_GIVENS)
    with tempfile.NamedTemporaryFile("w", suffix=".sdk", delete=False) as f:
        f.write("\n".join(G.sudoku_lines(cells)) + "\n")
    res = run(["sudoku", f.name, "boxes", THREADS])
    os.unlink(f.name)
    res = [r for r in res if "summary" not in r]
    assert len(res) == SUDOKU_COUNT
    assert all(r["status"] == "sat" for r in res)
    sol = "".join("".join(map(str, r["first"])) for r in res)
    gold["sudoku10k"] = {"n": SUDOKU_COUNT, "givens": SUDOKU_GIVENS, "seed": 20261018,
                         "sha256": hashlib.sha256(cells.tobytes()).hexdigest(),
                         "nodes": [r["nodes"] for r in res],
                         # 810 000 digits; kept as one string (the per-puzzle solution is sol[81*i : 81*i+81])
                         "solutions": sol, "solutions_sha256": hashlib.sha256(sol.encode()).hexdigest()}
    print("sudoku10k", sum(r["nodes"] for r in res), flush=True)
    save(gold)


def dups(gold):
    """Random models whose Values domains list values more than once (SURVEY.md par. 9 Q2), first and count modes."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from randmodels import dup_suite
    suite = dup_suite(150)
    txt = "".join(c.to_text() for c in suite)
    keep = ("status", "solutions", "nodes", "first", "order")
    out = {}
    for mode in ("first", "count"):
        res = subprocess.run([REF, "solve", mode, "0"], input=txt, capture_output=True, text=True, check=True).stdout
        rows = [json.loads(line) for line in res.splitlines()]
        assert len(rows) == len(suite)
        out[mode] = [{k: r[k] for k in keep} for r in rows]
    dup = {"generated_by": "tests/golden/make_golden_large.py dups", "reference": "nsweb/dequan dequan.h (unmodified)",
           "duplicate_values": {"n": 150, "seed0": 5000, "sha256": hashlib.sha256(txt.encode()).hexdigest(), **out}}
    with open(os.path.join(ROOT, "tests", "golden", "reference_dups.json"), "w") as f:     # (a file of its own)
        json.dump(dup, f, separators=(",", ":"))
    print("dups", sum(r["nodes"] for r in out["count"]), flush=True)


def save(gold):
    with open(OUT, "w") as f:
        json.dump(gold, f, separators=(",", ":"))


def main():
    gold = {"generated_by": "tests/golden/make_golden_large.py", "reference": "nsweb/dequan dequan.h (unmodified)"}
    if os.path.exists(OUT):
        with open(OUT) as f:
            gold.update(json.load(f))
    sections = sys.argv[1:] or ["colouring", "sudoku", "queens"]
    for s in sections:
        {"queens": queens, "queens18": lambda g: queens(g, (18,)), "colouring": colouring, "sudoku": sudoku, "dups": dups}[s](gold)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
