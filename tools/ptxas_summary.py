"""Registers / spills of every traversal kernel instantiation, from the ptxas -v log of the last build.
usage: python tools/ptxas_summary.py [substring]"""
import os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
flt = sys.argv[1] if len(sys.argv) > 1 else "fast3"
txt = open(os.path.join(root, "nanort_b200", "csrc", "traverse.o.ptxas.log")).read()
for b in re.split(r"ptxas info\s+: Compiling entry function '", txt)[1:]:
    name = b.split("'")[0]
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().replace("nrt::", "")
    dem = dem[:dem.find("(")]
    if flt not in dem or ", 512," in dem:
        continue
    m = re.search(r"Used (\d+) registers", b)
    st = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b)
    print(m.group(1), "regs, stack/spill-st/spill-ld", st.groups() if st else None, dem[:160])
