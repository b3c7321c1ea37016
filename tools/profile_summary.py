"""Text summary of an ncu report for profiles/: key counters of every captured launch plus the
instruction mix by source line of one kernel.
usage: python tools/profile_summary.py <report.ncu-rep> [<object.o> <kernel-substring> <source.cu>]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
print(f"# {rep}")
for r in rows[2:]:
    d = dict(zip(h, r))
    print(f"\n## launch {d.get('ID')}: {d.get('Kernel Name')}")
    for k in KEYS:
        if k in d and d[k] != "":
            print(f"{k:95s} {d[k]:>18s} {units[h.index(k)]}")
if len(sys.argv) > 4:
    print("\n## instruction mix by source line (tools/ncu_lines.py)")
    sys.stdout.flush()
    import os
    env = dict(os.environ, TOP="30")
    subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "ncu_lines.py"), rep, *sys.argv[2:5]], env=env)
