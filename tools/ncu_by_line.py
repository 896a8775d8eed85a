"""Per-source-line cost of one kernel launch: joins the SASS page of an .ncu-rep (instructions executed, thread
instructions executed, stall samples per SASS instruction) with nvdisasm's line info of the same kernel in the .o.

usage: tools/ncu_by_line.py <rep.ncu-rep> <launch index> <object.o> <mangled-name substring> [rays]
Prints, per (file:line) and inlined-at chain head, warp instructions, average active lanes, share of the launch's
issue slots and of its stall samples; with `rays`, warp instructions per ray."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_page(rep, launch):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip",
                          str(launch), "--launch-count", "1"], capture_output=True, text=True, check=True).stdout
    lines = out.splitlines()
    heads = [i for i, ln in enumerate(lines) if ln.startswith('"Kernel Name"')] + [len(lines)]
    sect = lines[heads[0]:heads[1]]  # ncu prints the selected launch's section twice; take the first copy
    rows = list(csv.DictReader(io.StringIO("\n".join(sect[1:]))))
    return sect[0], rows


def line_info(obj, pattern):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cubin)], capture_output=True,
                         text=True)
    if txt.returncode != 0:
        txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True,
                             text=True, check=True)
    out, on, cur = [], False, ("?", 0, "")
    for ln in txt.stdout.splitlines():
        if ln.startswith(".text."):
            on = pattern in ln
            continue
        if not on:
            continue
        m = re.match(r"\s*//## File \"([^\"]+)\", line (\d+)(.*)", ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3).strip())
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return out


def main():
    rep, launch, obj, pat = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
    rays = float(sys.argv[5]) if len(sys.argv) > 5 else None
    name, rows = sass_page(rep, launch)
    info = line_info(obj, pat)
    if len(info) != len(rows):
        print(f"warning: {len(rows)} SASS rows in the report vs {len(info)} in the object", file=sys.stderr)
    agg = collections.OrderedDict()
    tot_w = tot_t = tot_s = 0
    for r, (_, _, loc) in zip(rows, info):
        w, t = int(r["Instructions Executed"]), int(r["Thread Instructions Executed"])
        s = int(r["Warp Stall Sampling (All Samples)"] or 0)
        a = agg.setdefault(loc[:2], [0, 0, 0, 0])
        a[0] += w
        a[1] += t
        a[2] += s
        a[3] += 1
        tot_w += w
        tot_t += t
        tot_s += s
    print(name[:200])
    print(f"total: {tot_w} warp inst, {tot_t / max(tot_w, 1):.2f} lanes, {tot_s} samples" +
          (f", {tot_w / rays:.1f} warp inst / ray, {tot_t / rays:.0f} thread inst / ray" if rays else ""))
    print(f"{'file:line':28s} {'sass':>5s} {'warp inst':>12s} {'lanes':>6s} {'issue %':>8s} {'stall %':>8s}" +
          (f" {'winst/ray':>10s}" if rays else ""))
    for (f, l), (w, t, s, k) in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
        if w == 0:
            continue
        print(f"{f + ':' + str(l):28s} {k:5d} {w:12d} {t / w:6.2f} {100.0 * w / tot_w:8.2f} {100.0 * s / max(tot_s, 1):8.2f}" +
              (f" {w / rays:10.2f}" if rays else ""))


if __name__ == "__main__":
    main()
