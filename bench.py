#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: search nodes/sec (N-Queens all-solutions) and Sudoku
puzzles/sec, on 1/2/4/8 B200, next to the reference dequan timed on the host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun ... bench.py --gpus N ...                      (N > 1, one rank per GPU, NCCL)

One "step" = one complete solve of the workload through the C ABI.  The workload is the same at every N, so that the
1/2/4/8-GPU series is a strong-scaling series of ONE job: BASELINE config C5, 17-Queens all-solutions (95 815 104
solutions / 5 474 619 051 nodes; 24 ms on one B200 — the largest single-tree configuration and the one the north star
names), FC-surviving prefixes dealt to the ranks by key, {solutions, nodes} summed with NCCL.  The line also carries
`sudoku` (config C3, 1M puzzles, sharded over the ranks), and at N == 1 `extra.nqueens14_1gpu` (config C2, 14-Queens:
value, e2e and roofline of its own) and `extra.colouring_*` (config C4).
`value`  : inputs/tables already resident in HBM (compiled model re-used), whole-job nodes/s.
`e2e`    : the same through the host-facing C-ABI calls with HOST buffers, every step: dq_compile of the flat
           model descriptor (CSP::FinalizeModel + Assignment::Reset), table upload, solve, result read-back, dq_free.
Every result of every step is checked against the known answers; a mismatch aborts the bench.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

QUEENS = {  # OEIS A000170 + node counts pinned by tests/golden (N<=14) and the engine parity runs
    14: (365596, 19787662), 15: (2279184, 121498513), 16: (14772512, 795563572), 17: (95815104, 5474619051),
}
# forward-checking domain updates per node (A) -> algorithmic lane-ops/node = 5A+4 (queens), 2A+4 (!=,0)
QUEENS_A = {14: 79143794 / 19787662, 15: 498817896 / 121498513, 16: 3342155422 / 795563572, 17: 23537105544 / 5474619051}
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "dequan_ref")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        # the busiest half of the samples = "under load"
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": reasons}


def run_ref(args_list, timeout=900):
    out = subprocess.run([REF_BIN] + [str(a) for a in args_list], capture_output=True, text=True, timeout=timeout, check=True).stdout
    return [json.loads(line) for line in out.splitlines() if line.strip()]


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
def cpu_baseline_queens(n_gpus: int):
    """Reference dequan on the host cores, bounded sample of the same workload."""
    thr = cpu_threads()
    if not os.path.exists(REF_BIN):
        return cpu_baseline_port_queens(n_gpus)
    t = time.time()
    o = run_ref(["nqueens", 17, "count", thr, 8, 3])[0]
    wall = time.time() - t
    return {"value": o["nodes"] / o["seconds"], "unit": "nodes/s", "cores": min(thr, 17), "kind": "reference",
            "sample": f"17-Queens subtree under prefix (8,3), one depth-3 subtree per thread ({o['nodes']} nodes, {wall:.1f}s wall)"}


def cpu_baseline_port_queens(n_gpus: int):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from dequan_b200.model import nqueens
    n = 11
    t = time.time()
    o = oracle_lib.solve(nqueens(n), "count")
    dt = time.time() - t
    return {"value": o.nodes / dt, "unit": "nodes/s", "cores": 1, "kind": "port",
            "sample": f"{n}-Queens all-solutions with the C restatement (oracle/_ref absent)"}


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation, all host threads, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    thr = cpu_threads()
    n = 17
    cmd = ["nqueens", 17, "count", thr, 8, 3]
    times, nodes = [], 0
    kind = "reference" if os.path.exists(REF_BIN) else "port"
    for i in range(args.warmup + args.steps):
        if kind == "reference":
            o = run_ref(cmd)[0]
            dt, nd = o["seconds"], o["nodes"]
        else:
            b = cpu_baseline_port_queens(args.gpus)
            nd = 1
            dt = 1.0 / b["value"]
        if i >= args.warmup:
            times.append(dt)
            nodes += nd
    total = sum(times)
    val = nodes / total
    sample = "17-Queens subtree under prefix (8,3) split over the host threads (bounded sample of the 5.47e9-node tree)"
    line = {"impl": "reference", "metric": "search_nodes_per_sec", "value": val, "unit": "nodes/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(len(times), 1),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(args.gpus, n),
            "cpu_baseline": {"value": val, "unit": "nodes/s", "cores": min(thr, n), "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, n):
    which = {17: "BASELINE config C5 (the same job at every N: strong scaling; C2 = 14-Queens is reported in extra.nqueens14_1gpu)",
             14: "BASELINE config C2"}.get(n, "")
    return {"workload": f"nqueens{n}_count_all", "baseline_config": which,
            "model": "N vars AddIntVar(0,N), 3 OpConstraint NotEqual per pair (main-test.cpp:36-49)",
            "parallelism": "single tree, FC-surviving prefixes in DFS order" + (f", dealt by key to {n_gpus} GPUs, NCCL sum" if n_gpus > 1 else ""),
            "l2": "flushed between steps (256 MiB write, untimed)"}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default="auto")
    ap.add_argument("--sudoku-n", type=int, default=1_000_000)
    ap.add_argument("--givens", type=int, default=30)
    ap.add_argument("--no-sudoku", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from dequan_b200 import api, multi
    from dequan_b200.model import nqueens

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    api.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure_queens(n, steps, warmup):
        """value / e2e legs of the N-Queens all-solutions workload; every step's result is checked."""
        want_sols, want_nodes = QUEENS[n]
        csp = nqueens(n)
        model = api.Model(csp)

        def solve_step(m):
            loc = m.solve_tree("count", part_rank=rank, part_count=world, engine=args.engine)
            if world == 1:
                return loc, loc.solutions, loc.nodes
            g = multi.reduce_tree(loc, m.nodes_upto, "count", n, device=dev)
            return loc, g.solutions, g.nodes

        dom = {"ms": 0.0, "frontier_nodes": 0, "records": 0}

        def timed_steps(step_fn, k, w):
            tot, kern, launches = 0.0, 0.0, 0
            dom["ms"] = 0.0
            for i in range(w + k):
                flush.fill_(i & 0xFF)
                sync_all()
                t0 = time.perf_counter()
                loc, sols, nodes = step_fn()
                sync_all()
                dt = time.perf_counter() - t0
                assert (sols, nodes) == (want_sols, want_nodes), f"parity failure in bench step: {(sols, nodes)}"
                if i >= w:
                    tot += dt
                    kern += loc.kernel_ms
                    launches += loc.launches
            if world > 1:
                t = torch.tensor([tot, kern], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                tot, kern = float(t[0]), float(t[1])
            return tot, kern, launches

        tot, kern_ms, launches = timed_steps(lambda: solve_step(model), steps, warmup)
        # roofline leg: the search kernel's own duration from CUDA events on its stream (DQ_TREE_TIME_KERNELS; the steps
        # above replay the solve as one CUDA graph, whose inner events cannot be read back)
        dom["ms"] = 0.0
        for i in range(steps):
            flush.fill_(i & 0xFF)
            sync_all()
            loc = model.solve_tree("count", part_rank=rank, part_count=world, engine=args.engine, time_kernels=True)
            assert (loc.solutions, loc.nodes) == (want_sols, want_nodes) or world > 1
            dom["ms"] += loc.search_kernel_ms
            dom["frontier_nodes"], dom["records"] = loc.frontier_nodes, loc.n_prefixes
        dom_ms, dom_frontier, dom_records = dom["ms"], dom["frontier_nodes"], dom["records"]
        desc_keep = csp.desc()          # the host-side flat descriptor: the input buffers of the C-ABI call

        def e2e_step():
            m = api.Model(csp, desc_keep)   # dq_compile: FinalizeModel + Reset on the host, tables uploaded by the solve
            r = solve_step(m)
            m.close()
            return r
        e_tot, _, _ = timed_steps(e2e_step, steps, warmup)
        return {"n": n, "nodes": want_nodes, "steps": steps, "tot": tot, "kern_ms": kern_ms, "launches": launches, "e_tot": e_tot,
                "dom_ms": dom_ms, "dom_frontier": dom_frontier, "dom_records": dom_records, "table_bytes": model.table_bytes(),
                "engine": model.solve_tree("count", part_rank=rank, part_count=world, engine=args.engine).engine}

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n = 17
    Q = measure_queens(n, args.steps, args.warmup)
    want_nodes, tot, kern_ms, launches, e_tot = Q["nodes"], Q["tot"], Q["kern_ms"], Q["launches"], Q["e_tot"]
    dom_ms, dom_frontier, dom_records, table_bytes, engine_used = Q["dom_ms"], Q["dom_frontier"], Q["dom_records"], Q["table_bytes"], Q["engine"]
    Q14 = measure_queens(14, max(args.steps, 10), args.warmup) if world == 1 and not args.no_extra else None

    sudoku = None
    if not args.no_sudoku:
        sudoku = sudoku_section(args, torch, api, dev, world, rank, dist, flush)      # still inside the clock-sampling window

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    clocks = sampler.stop()
    int_peak, _ = api.measure_int_peak()
    value = want_nodes * args.steps / tot
    kernel_nodes_per_s = want_nodes * args.steps / (kern_ms * 1e-3)
    ops_per_node = 5 * QUEENS_A[n] + 4
    hbm_peak, peak_src = peaks()
    line = {
        "metric": "search_nodes_per_sec", "value": value, "unit": "nodes/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": workload_config(world, n),
        "clocks": clocks,
        "e2e": {"value": want_nodes * args.steps / e_tot, "unit": "nodes/s", "h2d_bytes_per_step": table_bytes + 64,
                "d2h_bytes_per_step": 64 + 4 * n + 8 * 64, "ms_per_step": 1e3 * e_tot / args.steps},
        "gpu_launches": launches,
        "engine": engine_used,
        "kernel_ms_per_step": kern_ms / args.steps,
        "roofline": roofline_queens(n, world, want_nodes, dom_frontier, dom_records, dom_ms / args.steps, int_peak, hbm_peak, peak_src),
    }
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline_queens(world)
    extra = {}
    if world == 1 and not args.no_extra:
        # BASELINE config C2: 14-Queens all-solutions on one B200 (365 596 solutions / 19 787 662 nodes)
        q = Q14
        extra["nqueens14_1gpu"] = {
            "config": workload_config(1, 14), "value": q["nodes"] * q["steps"] / q["tot"], "unit": "nodes/s",
            "ms_per_step": 1e3 * q["tot"] / q["steps"], "kernel_ms_per_step": q["kern_ms"] / q["steps"], "steps": q["steps"],
            "e2e": {"value": q["nodes"] * q["steps"] / q["e_tot"], "unit": "nodes/s", "ms_per_step": 1e3 * q["e_tot"] / q["steps"],
                    "h2d_bytes_per_step": q["table_bytes"] + 64, "d2h_bytes_per_step": 64 + 4 * 14 + 8 * 64},
            "gpu_launches": q["launches"], "engine": q["engine"],
            "roofline": roofline_queens(14, 1, q["nodes"], q["dom_frontier"], q["dom_records"], q["dom_ms"] / q["steps"], int_peak, hbm_peak, peak_src)}
        # BASELINE config C4: G(200, c/199) 3-colouring near the phase transition, batched, node budget per instance
        from dequan_b200 import generators as G
        # 1 024 instances leave 7 warps per SM (one warp per instance): a latency-bound launch; 8 192 fill the machine
        off_all, edges_all = G.colouring_batch(8192, 200, 4.2)
        for count in (1024, 8192):
            off, edges = off_all[:count + 1], edges_all[:off_all[count]]
            cr = api.solve_batch_graphs(200, 3, off, edges, node_budget=100_000)
            cr = api.solve_batch_graphs(200, 3, off, edges, node_budget=100_000)
            extra[f"colouring_g200_c4.2_k3_{count}"] = {"instances": count, "node_budget": 100_000, "kernel_ms": cr.kernel_ms,
                                                        "instances_per_sec": count / (cr.kernel_ms * 1e-3),
                                                        "nodes_per_sec": cr.total_nodes / (cr.kernel_ms * 1e-3),
                                                        "sat": cr.n_sat, "unsat": cr.n_unsat, "budget": cr.n_budget,
                                                        "engine": "register-resident warp engine (dq_reg_graphs.cuh)"}
    if sudoku is not None:
        line["sudoku"] = sudoku_rooflines(sudoku, hbm_peak, peak_src, int_peak)
    line["extra"] = extra
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of k_queens_bucket (profiles/r1_ncu_queens14_bucket.txt, r1_ncu_queens17_bucket.txt)
NCU_TRAFFIC = {(14, 1): 13093120 + 256, (17, 1): 1982841000 + 5133824}


def roofline_queens(n, world, nodes, frontier_nodes, records, lane_ms, int_peak, hbm_peak, peak_src):
    """Roofline of the dominant kernel, k_queens_bucket (depth-bucketed subtree search; profiles/r1_launch_shares.txt).
    It is integer-issue bound: algorithmic work = (5A+4) lane-ops per node, A = forward-checking domain
    updates per node; the peak is the LOP3 rate measured in this run.  Its HBM side is shown next to it: one 16-byte
    record read per subtree."""
    ops_per_node = 5 * QUEENS_A[n] + 4
    lane_nodes = (nodes - frontier_nodes) if world == 1 else None          # N>1: this rank's share is not split out
    per_gpu_nodes = (nodes / world) if lane_nodes is None else lane_nodes
    achieved = per_gpu_nodes * ops_per_node / (lane_ms * 1e-3) if lane_ms else 0.0
    algo_bytes = records * 16 if world == 1 else None
    return {"bound": "int32-alu", "kernel": "k_queens_bucket", "achieved": achieved / 1e12, "peak": int_peak / 1e12,
            "unit": "Tlane-op/s per GPU", "frac": achieved / int_peak, "kernel_ms": lane_ms,
            "nodes_per_launch": per_gpu_nodes, "ops_per_node": ops_per_node,
            "peak_source": "LOP3 microbenchmark (dq_measure_int_peak), this run",
            "traffic": NCU_TRAFFIC.get((n, world)), "algorithmic_bytes": algo_bytes,
            "hbm": {"achieved": (algo_bytes / (lane_ms * 1e-3) / 1e9) if algo_bytes and lane_ms else None, "peak": hbm_peak,
                    "unit": "GB/s", "peak_source": peak_src},
            "note": "not HBM- or tensor-bound: no dense contraction on this path, one 16 B record per subtree from HBM"}


# ncu, dram__bytes_read.sum + dram__bytes_write.sum summed over the seven kernels of one 1 M-puzzle pipeline pass
# (profiles/r1_ncu_sudoku_traffic.txt): 856 MB read + 366 MB written, of which 288 B per puzzle are the digest (written
# once, read by k_sudoku_first and again per counting task) and the rest task / snapshot records of the counting stage
NCU_SUDOKU_TRAFFIC_1M = 856_260_000 + 366_400_000


def sudoku_rooflines(out, hbm_peak, peak_src, int_peak):
    kpps, nps = out.pop("_kernel_pps"), out["nodes_per_sec"]
    out["roofline"] = {"bound": "hbm", "achieved": kpps * 174 / 1e9, "peak": hbm_peak, "unit": "GB/s",
                       "frac": kpps * 174 / 1e9 / hbm_peak,
                       "traffic": NCU_SUDOKU_TRAFFIC_1M if out.get("shard", out["n"]) == 1_000_000 else None,
                       "algorithmic_bytes": 174 * out.get("shard", out["n"]), "bytes_per_puzzle": 174, "peak_source": peak_src,
                       "note": "instance stream only; the search itself is integer-issue bound (see roofline_int)"}
    out["roofline_int"] = {"ops_per_node": 2 * 10 + 4, "achieved": nps * 24 / 1e12, "peak": int_peak / 1e12, "unit": "Tlane-op/s",
                           "frac": nps * 24 / int_peak}
    return out


def sudoku_section(args, torch, api, dev, world=1, rank=0, dist=None, flush=None):
    """BASELINE config C3: batch of synthetic 9x9 Sudoku (810 binary != arcs), first solution each.  With N GPUs the
    batch is cut into N contiguous shards, one per rank; the only collective is the final sum of {solved, nodes}."""
    from dequan_b200 import generators as G
    from dequan_b200.model import sudoku_template
    n_total = args.sudoku_n
    from dequan_b200 import multi
    lo, hi = multi.shard_range(n_total, rank, world)
    n = hi - lo
    cells = G.sudoku_batch(n, givens=args.givens, start=lo)
    tmpl = api.Model(sudoku_template())

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # HBM-resident leg
    d_cells = torch.from_numpy(cells).to(dev)
    d_sol = torch.empty_like(d_cells)
    d_nodes = torch.empty(n, dtype=torch.int64, device=dev)
    d_status = torch.empty(n, dtype=torch.uint8, device=dev)
    steps, warm = max(3, min(args.steps, 5)), 3
    kern, tot, launches, total_nodes = 0.0, 0.0, 0, 0
    for i in range(warm + steps):
        if flush is not None:
            flush.fill_(i & 0xFF)               # a shard may fit the 126 MB L2: evict it between steps (untimed)
        sync_all()
        t0 = time.perf_counter()
        st = tmpl.solve_batch_cells_ptr(d_cells.data_ptr(), n, 81, d_sol.data_ptr(), d_nodes.data_ptr(), d_status.data_ptr(), device=True)
        sync_all()
        dt = time.perf_counter() - t0
        assert st.n_sat == n
        if i >= warm:
            tot += dt; kern += st.kernel_ms; launches += st.kernel_launches; total_nodes = st.total_nodes
    tot, kern = reduce_max(tot), reduce_max(kern)
    # host-buffer leg (pinned), copies inside the timed region
    h_cells = torch.from_numpy(cells).pin_memory()
    h_sol = torch.empty((n, 81), dtype=torch.uint8).pin_memory()
    h_nodes = torch.empty(n, dtype=torch.int64).pin_memory()
    h_status = torch.empty(n, dtype=torch.uint8).pin_memory()
    e_tot = 0.0
    for i in range(warm + steps):
        if flush is not None:
            flush.fill_(i & 0xFF)
        sync_all()
        t0 = time.perf_counter()
        st = tmpl.solve_batch_cells_ptr(h_cells.data_ptr(), n, 81, h_sol.data_ptr(), h_nodes.data_ptr(), h_status.data_ptr())
        sync_all()
        dt = time.perf_counter() - t0
        if i >= warm:
            e_tot += dt
    e_tot = reduce_max(e_tot)
    if world > 1:      # the one collective of a sharded batch: totals
        g = multi.reduce_batch(st.n_sat, st.n_unsat, st.n_budget, total_nodes, device=dev)
        assert g.n_sat == n_total
        total_nodes = g.nodes
    # checks: device and host legs agree; solutions are valid grids consistent with the givens
    sol = h_sol.numpy()
    assert (d_sol.cpu().numpy() == sol).all() and (d_nodes.cpu().numpy() == h_nodes.numpy()).all()
    g = sol.reshape(n, 9, 9)
    want = np.arange(1, 10)
    assert (np.sort(g, axis=2) == want).all() and (np.sort(g, axis=1) == want[:, None]).all()
    assert (sol[cells != 0] == cells[cells != 0]).all()
    if rank != 0:
        return None
    n_shard, n = n, n_total
    pps = n * steps / tot
    kpps = n * steps / (kern * 1e-3)
    out = {"metric": "sudoku_puzzles_per_sec", "value": pps, "unit": "puzzles/s", "n": n, "givens": args.givens, "n_gpus": world,
           "scaling": "strong", "shard": n_shard,
           "steps": steps, "ms_per_step": 1e3 * tot / steps, "kernel_ms_per_step": kern / steps,
           "nodes_per_puzzle": total_nodes / n, "nodes_per_sec": total_nodes * steps / (kern * 1e-3),
           "config": {"workload": f"sudoku_1M_g{args.givens}" if n == 1_000_000 else f"sudoku_{n}_g{args.givens}",
                      "l2": "flushed between steps (256 MiB write, untimed); inputs+outputs are 171 B per puzzle"},
           "e2e": {"value": n * steps / e_tot, "unit": "puzzles/s", "h2d_bytes_per_step": n * 81, "d2h_bytes_per_step": n * 90,
                   "ms_per_step": 1e3 * e_tot / steps},
           "gpu_launches": launches, "engine": "lane pipeline (digest, first, strong, walk, count, finish)",
           "_kernel_pps": kpps}
    if not args.no_cpu and os.path.exists(REF_BIN):
        sample = min(n_shard, 16000)
        path = "/tmp/dq_bench_sudoku.txt"
        with open(path, "w") as f:
            f.write("\n".join(G.sudoku_lines(cells[:sample])) + "\n")
        thr = cpu_threads()
        o = run_ref(["sudoku", path, "boxes", thr, sample, "quiet"])[-1]
        # parity at scale against the unmodified reference: the node total of the sample must be the reference's
        assert int(h_nodes.numpy()[:sample].sum()) == int(o["nodes"]), ("sudoku node total differs from the reference",
                                                                      int(h_nodes.numpy()[:sample].sum()), o["nodes"])
        out["cpu_baseline"] = {"value": o["puzzles"] / o["wall_seconds"], "unit": "puzzles/s", "cores": thr, "kind": "reference",
                               "sample": f"first {sample} puzzles of the same batch, one puzzle per thread, model build included "
                                         f"(solve-only {o['puzzles'] / o['solve_seconds_sum'] * thr:.0f} puzzles/s)",
                               "nodes_per_sec": o["nodes"] / o["wall_seconds"]}
    return out


if __name__ == "__main__":
    main()
