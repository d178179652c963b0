"""Text summary of an ncu report (run here, no GPU needed): python scripts/ncu_summary.py rep.ncu-rep > profiles/x.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    print("=" * 100)
    print(r[ki])
    rd = wr = None
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("  %-75s %18s %s" % (w, r[i], units[i]))
            if w == "dram__bytes_read.sum": rd = (float(r[i].replace(",", "")), units[i])
            if w == "dram__bytes_write.sum": wr = (float(r[i].replace(",", "")), units[i])
    if rd and wr:
        print("  %-75s %18.4f %s" % ("traffic = dram read + write", rd[0] + wr[0], rd[1]))
