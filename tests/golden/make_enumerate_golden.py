"""Regenerates tests/golden/enumerate_reference.json: every solution, in visiting order, that the UNMODIFIED reference
(oracle/_ref/dequan_ref `enumerate`, a counting Constraint snapshotting inst_vars on each hit) finds for a fixed set of
models.  Run in the build container only (the GPU box has no /root/reference):
    make -C oracle ref && python tests/golden/make_enumerate_golden.py
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from enum_models import enum_models  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "dequan_ref")
CAP = 400


def main():
    names, csps = zip(*enum_models())
    txt = "".join(c.to_text() for c in csps)
    out = subprocess.run([REF, "enumerate", str(CAP)], input=txt, capture_output=True, text=True, check=True).stdout
    res = [json.loads(line) for line in out.splitlines()]
    assert len(res) == len(csps)
    gold = {"generated_by": "tests/golden/make_enumerate_golden.py", "cap": CAP,
            "models": {n: {"solutions": r["solutions"], "nodes": r["nodes"], "all": r["all"]} for n, r in zip(names, res)}}
    with open(os.path.join(ROOT, "tests", "golden", "enumerate_reference.json"), "w") as f:
        json.dump(gold, f, separators=(",", ":"))
    print({n: r["solutions"] for n, r in zip(names, res)})


if __name__ == "__main__":
    main()
