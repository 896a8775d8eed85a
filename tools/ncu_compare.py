"""Key metrics of one or more raw ncu csv exports side by side.  usage: python tools/ncu_compare.py a.csv b.csv ..."""
import csv, sys
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_lg.sum"]
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    print("==", path)
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k} [{units[i]}]:", " | ".join(r[i][:60] for r in rows[2:]))
    st = [(h, [r[hdr.index(h)] for r in rows[2:]]) for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
    st.sort(key=lambda x: -float(x[1][0]) if x[1][0] else 0)
    for h, v in st[:8]:
        print("  stall", h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v)
