"""Print the headline metrics of `ncu -i X.ncu-rep --page raw --csv` (one kernel per row)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d.get("Kernel Name"))
    for k in want:
        if k in d:
            print(f"  {k:90s} {d[k]:>16s} {units[hdr.index(k)]}")
    for k in hdr:
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
            v = float(d[k] or 0)
            if v >= 0.15:
                print(f"  stall {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:40s} {v:6.2f}")
