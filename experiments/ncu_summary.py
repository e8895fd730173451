"""Prints the metrics that matter from an .ncu-rep (raw page): python experiments/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.avg", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "sm__inst_executed_pipe_uniform", "lts__t_sector_hit_rate.pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("-----")
    for h, u, v in zip(hdr, rows[1], r):
        if any(h == k or (k in h and len(k) > 25) for k in keys):
            print(f"{h} [{u}] = {v}")
