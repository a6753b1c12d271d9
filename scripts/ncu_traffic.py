"""Records the DRAM traffic of one kernel launch from an `ncu --set full` report into profiles/r2_traffic.json
(what bench.py reports as roofline.traffic).  usage: ncu_traffic.py REPORT.ncu-rep KEY [KERNEL_REGEX] [--sum]
KEY names the kernel/workload (e.g. k_queens_bucket/nqueens17); --sum adds up all matching launches (a pipeline)."""
import csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, key = sys.argv[1], sys.argv[2]
pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else None
do_sum = "--sum" in sys.argv
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
ki, ri, wi, ti = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rd = wr = ms = 0.0
names = []
for r in rows[2:]:
    if pat and not pat.search(r[ki]):
        continue
    rd_i, wr_i = float(r[ri]) * scale[units[ri]], float(r[wi]) * scale[units[wi]]
    t_i = float(r[ti]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[ti], 1e-6)
    if do_sum:
        if r[ki][:60] in names:
            continue                       # one launch per kernel: a pipeline pass
        rd, wr, ms = rd + rd_i, wr + wr_i, ms + t_i
        names.append(r[ki][:60])
    else:
        rd, wr, ms, names = rd_i, wr_i, t_i, [r[ki][:60]]
commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip()
path = os.path.join(ROOT, "profiles", "r2_traffic.json")
db = json.load(open(path)) if os.path.exists(path) else {}
db[key] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "gpu_time_ms_under_ncu": ms, "kernels": sorted(set(names)),
           "report": os.path.basename(rep), "commit": commit}
json.dump(db, open(path, "w"), indent=1, sort_keys=True)
print(key, db[key])
