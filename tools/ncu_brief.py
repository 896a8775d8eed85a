"""Key metrics of every kernel launch in an ncu report, as a markdown table row set.
usage: python tools/ncu_brief.py report.ncu-rep [more.ncu-rep ...]"""
import csv, subprocess, sys

KEYS = [("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
        ("smsp__inst_executed.sum", "warp insts"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait")]
print("| kernel | " + " | ".join(k[1] for k in KEYS) + " |")
print("|---|" + "---|" * len(KEYS))
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0].split("::")[-1]
        cells = []
        for k, _ in KEYS:
            if k in d:
                u = units[hdr.index(k)]
                v = d[k]
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {u}".strip())
            else:
                cells.append("-")
        print(f"| `{name}` | " + " | ".join(cells) + " |")
