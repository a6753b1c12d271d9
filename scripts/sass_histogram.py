"""Static opcode counts per kernel from cuobjdump -sass (run here, no GPU).  usage: sass_histogram.py > profiles/r2_sass_histograms.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "dequan_b200", "lib", "libdequan_b200.so")
want = re.compile(sys.argv[1] if len(sys.argv) > 1 else
                  r"k_graphs_adjacencyILb0|k_sudoku_count|k_sudoku_first|k_sudoku_digest|k_queens_bucket_tILi6ELi5|k_queens_bucket_tILi4ELi5|k_queens_bucketE|k_queens_first_warp|k_queens_level_wide|k_graphs_lane|k_tree_lanes")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
print("# cuobjdump -sass dequan_b200/lib/libdequan_b200.so: static opcode counts per kernel (whole kernel, all paths).")
print("# UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier, REDUX/VOTE/SHFL = warp collectives, LDS/STS = shared memory.")
print("# ALU pipe: LOP3 SHF ISETP VIMNMX IADD3 LEA SEL PRMT ...; FMA pipe: IMAD; XU: POPC FLO BREV.")
name, ops = None, None
keys = ["UBLKCP", "SYNCS", "LOP3", "SHF", "VOTE", "VOTEU", "REDUX", "SHFL", "LDS", "STS", "LDG", "STG", "POPC", "FLO", "IMAD", "VIMNMX3", "VIMNMX", "LEA", "ISETP"]


def flush():
    if name and want.search(name):
        tot = sum(ops.values())
        print(f"== {name}  ({tot} instructions)")
        print("   " + "  ".join(f"{k}:{v}" for k, v in ops.most_common(28)))
        print("   key: " + "  ".join(f"{k}={ops.get(k, 0)}" for k in keys))


for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        name, ops = m.group(1), collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and ops is not None:
        ops[m.group(1)] += 1
flush()
