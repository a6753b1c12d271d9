"""Per-SASS-instruction hot spots from an .ncu-rep source page.  usage: ncu_hot.py rep nodes [kernel_substr]"""
import csv, subprocess, sys
rep, nodes = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]
ia, it, isrc, iav, idv, ismp = (h.index(x) for x in ('Instructions Executed', 'Thread Instructions Executed', 'Source', 'Avg. Threads Executed', 'Divergent Branches', '# Samples'))
body = [r for r in rows[2:] if len(r) > ia and r[ia].isdigit()]
tot = sum(int(r[ia]) for r in body)
print(f"warp-inst/node {tot/nodes:.3f}   thread-inst/node {sum(int(r[it]) for r in body)/nodes:.2f}   avg active {sum(int(r[it]) for r in body)/tot:.2f}")
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
for i, r in enumerate(body):
    n = int(r[ia])
    if n > tot * thr:
        print(f"{i:4d} {n/nodes:7.3f} thr={r[iav]:>3s} div={int(r[idv])/nodes:6.3f} smp={r[ismp]:>6s}  {r[isrc].strip()[:100]}")
