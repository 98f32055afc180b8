"""Print the handful of ncu raw-page metrics we quote (usage: ncu_keys.py report.ncu-rep [kernel index])."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.max", "smsp__average_warps_issue_stalled",
        "sm__throughput.avg.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__cycles_active.avg", "sm__cycles_active.avg"]
for vals in rows[2:]:
    print("-" * 60)
    for h, u, v in zip(hdr, units, vals):
        if any(h == k or (k.endswith("stalled") and h.startswith(k)) or (k.endswith("pct") and h.startswith(k)) for k in KEYS):
            if h.startswith("smsp__average_warps_issue_stalled") and not h.endswith("per_issue_active.ratio"):
                continue
            print(f"{h:90s} {v} {u}")
