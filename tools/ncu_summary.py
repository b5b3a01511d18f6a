"""Condenses an `ncu --page raw --csv` dump to the metrics DESIGN.md / profiles/ quote.
Usage: ncu -i X.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv"""
import csv
import sys

WANT = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.avg.per_cycle_active',
    'smsp__thread_inst_executed_per_inst_executed.ratio',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
    'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum',
    'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
    'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum',
    'smsp__average_warp_latency_per_inst_issued.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_drain_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
only = sys.argv[2] if len(sys.argv) > 2 else None
seen = set()
for r in rows[2:]:
    name = r[idx['Kernel Name']].split('(')[0]
    if only and only not in name:
        continue
    if name in seen:
        continue
    seen.add(name)
    print(f"===== {name}  (launch id {r[idx['ID']]})")
    for w in WANT:
        if w in idx:
            print(f"  {w:82s} {r[idx[w]]:>18s} {units[idx[w]]}")
