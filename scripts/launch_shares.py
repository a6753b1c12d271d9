"""Per-workload kernel shares from an ncu launch list (gpu__time_duration.sum CSV of `python bench.py`).
usage: launch_shares.py gpurun_out/r1_launches_bench.csv > profiles/r1_launch_shares.txt"""
import collections
import csv
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
seq = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"].split("(")[0]
    if name.startswith("void at::"):
        name = "at::FillFunctor (L2 flush, untimed)"
    seq.append((name, float(row["Metric Value"].replace(",", ""))))
groups = collections.OrderedDict()


def add(g, name, ns):
    t = groups.setdefault(g, collections.OrderedDict()).setdefault(name, [0, 0.0])
    t[0] += 1
    t[1] += ns


buf = []
for name, ns in seq:
    if name in ("dq::k_queens_level", "dq::k_queens_level_wide", "dq::k_queens_levels_head", "dq::k_queens_first_warp"):
        buf.append((name, ns))          # belongs to the bucket launch that follows
    elif "k_queens_bucket" in name:
        g = "17-Queens count-all (main workload)" if ns > 3e6 else "14-Queens count-all (extra.nqueens14_1gpu)"
        for n2, t2 in buf:
            add(g, n2, t2)
        buf = []
        add(g, name, ns)
    elif name.startswith("dq::k_sudoku"):
        add("1 M Sudoku batch (sudoku)", name, ns)
    elif "k_batch_graphs" in name or "k_graphs" in name:
        add("G(200) 3-colouring batches (extra.colouring_*)", name, ns)
    else:
        add("other", name, ns)
out = ["# ncu --metrics gpu__time_duration.sum --clock-control none -c 600 : python bench.py --steps 2 --warmup 3 --no-cpu",
       "# per-launch times are cold-cache and SERIALISED (k_queens_first_warp runs on a side stream next to the level and",
       "# bucket kernels in a real step; under ncu it is timed alone).  What must agree with bench.py is each kernel's SHARE."]
for g, d in groups.items():
    tot = sum(v[1] for k, v in d.items() if k != "dq::k_queens_first_warp")
    out.append(f"== {g}")
    for k, (c, t) in d.items():
        if g == "other":
            share = ""
        elif k == "dq::k_queens_first_warp":
            share = "side stream (overlapped)"
        else:
            share = f"{100 * t / tot:5.1f}% of the serial chain"
        out.append(f"{k:36s} launches={c:4d} total_us={t / 1e3:11.1f} avg_us={t / 1e3 / c:10.2f}  {share}")
print("\n".join(out))
