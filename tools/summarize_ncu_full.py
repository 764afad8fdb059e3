"""Key metrics of an `ncu --set full` report as a small CSV (one column per captured launch).
    python tools/summarize_ncu_full.py report.ncu-rep "header comment" > profiles/<name>_ncu_full.csv"""
import csv, io, subprocess, sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum",
        "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"]

rep, comment = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, launches = rows[0], rows[1], rows[2:]
print(f"# ncu --set full --clock-control none --import-source on; {comment}")
ki = hdr.index("Kernel Name")
print("# " + "; ".join(sorted({r[ki].split('(')[0] for r in launches})))
w = csv.writer(sys.stdout, lineterminator="\n")
w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(launches))])
for m in KEEP:
    if m in hdr:
        i = hdr.index(m)
        w.writerow([m, units[i]] + [r[i] for r in launches])
