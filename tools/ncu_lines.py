"""Per-source-line summary of an ncu report's `--page source --print-source cuda,sass --csv` export.
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import collections, csv, subprocess, sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur, hdr = r[1].split("/")[-1], None
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[2] != "-":  # keep the cuda-line rows (Address == "-"), skip the sass rows under them
        continue
    d = dict(zip(hdr[4:], r[4:]))
    try:
        ie, te, s = int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), int(d["Warp Stall Sampling (All Samples)"])
    except (KeyError, ValueError):
        continue
    if ie or s:
        out.append((cur, int(r[0]), ie, te, s, r[1].strip()[:80]))
ti, ts = sum(o[2] for o in out), sum(o[4] for o in out)
print(f"warp instructions {ti}  stall samples {ts}")
byfile = collections.Counter()
for o in out:
    byfile[o[0]] += o[2]
print({k: f"{100 * v / ti:.1f}%" for k, v in byfile.items()})
for o in sorted(sorted(out, key=lambda o: -o[2])[:top_n], key=lambda o: (o[0], o[1])):
    print(f"{o[0]}:{o[1]:4d} inst {100 * o[2] / ti:5.2f}% lanes {o[3] / max(o[2], 1):5.1f} stall {100 * o[4] / ts:5.2f}%  {o[5]}")
