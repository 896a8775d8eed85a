"""Turns gpurun_out/{launches.csv, prof_trav.ncu-rep} into the tracked summaries under profiles/.
usage: python tools/summarize_ncu.py r01 [suffix]"""
import collections, csv, json, os, subprocess, sys

tag = sys.argv[1]
suffix = sys.argv[2] if len(sys.argv) > 2 else ""
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles")
os.makedirs(out, exist_ok=True)

# ---- launch list
rows = list(csv.reader(open(os.path.join(root, "gpurun_out", f"launches{suffix}.csv"))))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
per_launch = []
for r in data:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0]
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    per_launch.append((r[0], name, v))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(out, f"{tag}_launches{suffix}.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python tools/profile_target.py\n")
    f.write("# (build of the 100,002-triangle scene + two primary+AO passes at 1920x1080x4spp; cold-cache, serialised)\n")
    f.write("id,kernel,duration_us\n")
    for i, n, v in per_launch:
        f.write(f"{i},{n},{v:.2f}\n")
with open(os.path.join(out, f"{tag}_launch_shares{suffix}.md"), "w") as f:
    f.write(f"| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f}% |\n")
    f.write(f"| total | | {tot:.1f} | |\n")

# ---- full capture of the traversal kernel
rep = os.path.join(root, "gpurun_out", f"prof_trav{suffix}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
         and "not_issued" not in h]
with open(os.path.join(out, f"{tag}_traverse_fast_ncu{suffix}.csv"), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:traverse_fast -s 2 -c 2  python tools/profile_target.py\n")
    f.write("# launch 1 = primary rays (8,294,400), launch 2 = AO rays (4,551,027) of the same pass\n")
    f.write("metric,unit," + ",".join(f"launch{i + 1}" for i in range(len(rows) - 2)) + "\n")
    for k in keep + sorted(stall):
        if k in hdr:
            j = hdr.index(k)
            f.write(f"{k},{units[j]}," + ",".join('"' + r[j] + '"' for r in rows[2:]) + "\n")
dram = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    def b(k):
        v = float(d[k].replace(",", ""))
        u = units[hdr.index(k)].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    dram.append(b("dram__bytes_read.sum") + b("dram__bytes_write.sum"))
# rays of the captured launches: last line of the target's log ("scene primary_rays ao_rays ...")
rays = None
try:
    last = [l for l in open(os.path.join(root, "gpurun_out", "pt2.log")).read().splitlines() if l.startswith("sphere_grid")][-1].split()
    rays = [int(last[1]), int(last[2])][:len(dram)]
except Exception:
    pass
json.dump({"dram_bytes_per_launch": sum(dram) / len(dram), "per_launch": dram, "rays_per_launch": rays,
           "dram_bytes_per_ray": (sum(dram) / sum(rays)) if rays else None,
           "source": f"profiles/{tag}_traverse_fast_ncu{suffix}.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"},
          open(os.path.join(out, f"{tag}_traverse_traffic.json"), "w"), indent=1)
print("wrote", sorted(os.listdir(out)))
