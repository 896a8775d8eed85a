"""Turns a `ncu --set full` capture of the two traversal launches of one primary+AO pass (tools/profile_target.py) into
  profiles/<tag>_traverse_ncu<suffix>.csv     key metrics per launch (primary, AO)
  profiles/<tag>_traverse_counters.json      per-RAY counters per launch kind, read by bench.py's roofline block:
                                            warp instructions, active lanes, L2 bytes, DRAM bytes, L1 data-pipe use
usage: python tools/summarize_ncu_r02.py <report.ncu-rep> <target log with the ray counts> <tag> [suffix]"""
import csv, json, os, subprocess, sys

rep, log, tag = sys.argv[1], sys.argv[2], sys.argv[3]
suffix = sys.argv[4] if len(sys.argv) > 4 else ""
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__cycles_elapsed.avg", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
with open(os.path.join(out, f"{tag}_traverse_ncu{suffix}.csv"), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:traverse_fast3 -s 2 -c 2  python tools/profile_target.py\n")
    f.write("# launch 1 = camera rays (generated in the kernel) + AO spawn, launch 2 = AO rays of the same 1920x1080x4spp pass\n")
    f.write("metric,unit," + ",".join(f"launch{i + 1}" for i in range(len(rows) - 2)) + "\n")
    for k in keep + sorted(stall):
        if k in hdr:
            j = hdr.index(k)
            f.write(f"{k},{units[j]}," + ",".join('"' + r[j] + '"' for r in rows[2:]) + "\n")
last = [l for l in open(log).read().splitlines() if l.split() and l.split()[0] in ("sphere_grid", "terrain", "instanced")][-1].split()
rays = [int(last[1]), int(last[2])]


def val(r, k):
    d = dict(zip(hdr, r))
    v = float(d[k].replace(",", ""))
    u = units[hdr.index(k)].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


res = {"source": f"profiles/{tag}_traverse_ncu{suffix}.csv (ncu --set full of tools/profile_target.py, {last[0]}, 1920x1080x4 spp; "
                 "per-ray = launch total / rays of that launch)", "scene": last[0]}
for kind, r, n in zip(("primary", "ao"), rows[2:4], rays):
    res[kind] = {"rays": n, "warp_inst_per_ray": val(r, "smsp__inst_executed.sum") / n,
                 "lanes": val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                 "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                 "l2_bytes_per_ray": 32.0 * val(r, "lts__t_sectors.sum") / n,  # 32-byte sectors through the L2 tag stage
                 "dram_bytes_per_ray": (val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")) / n,
                 "l1_wavefront_pct": val(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                 "duration_ms_under_ncu": val(r, "gpu__time_duration.sum") *
                 {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[hdr.index("gpu__time_duration.sum")], 1.0)}
json.dump(res, open(os.path.join(out, f"{tag}_traverse_counters.json"), "w"), indent=1)
print(json.dumps(res, indent=1))
