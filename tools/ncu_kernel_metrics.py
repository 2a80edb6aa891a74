"""Key counters of the kernels in an ncu report (raw page): time, instructions, issue / stall picture, occupancy limits.
usage: python tools/ncu_kernel_metrics.py report.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_xu.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv", "--print-units", "base"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
h = rows[0]
for r in rows[2:]:
    print("----", r[h.index("Kernel Name")][:90])
    for w in WANT:
        if w in h:
            print(f"  {w:90s} {r[h.index(w)][:24]}")
