"""Per-source-line share of executed warp instructions from an .ncu-rep.  usage: ncu_lines.py rep [min_pct]"""
import csv, subprocess, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
tot, lines, fname = 0, [], ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]
    if len(r) > 8 and r[0].strip().isdigit() and r[7].strip().isdigit() and r[8].strip().isdigit():
        n, th = int(r[7]), int(r[8])
        lines.append((fname, int(r[0]), n, th, r[1].strip()[:120]))
        tot += n
print("total warp inst", tot)
for f, ln, n, th, src in lines:
    if n > tot * thr / 100:
        print(f"{f[:22]:22s} {ln:4d} {100 * n / tot:5.1f}%  act={th / max(n, 1):5.1f}  {src}")
