"""Prints the headline metrics of every kernel in an .ncu-rep (run here on CPU)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
want = ['gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem',
 'sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__thread_inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio',
 'sm__inst_executed.avg.per_cycle_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__warps_eligible.avg.per_cycle_active',
 'sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_lsu.sum','sm__inst_executed_pipe_cbu.sum','sm__inst_executed_pipe_adu.sum','sm__inst_executed_pipe_uniform.sum','sm__inst_executed_pipe_xu.sum',
 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__t_sector_hit_rate.pct',
 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio','smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__average_warps_issue_stalled_drain_per_issue_active.ratio','smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__cycles_active.avg','smsp__cycles_active.avg']
ki = h.index('Kernel Name')
for r in rows[2:]:
    print("====", r[ki][:80])
    for i, n in enumerate(h):
        if n in want:
            print(f"  {n:86s} {r[i]:>18s} {u[i]}")
