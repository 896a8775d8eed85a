"""Build-time probe: python tools/build_probe.py scene [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nanort_b200 import api, scenes as S
scene = sys.argv[1] if len(sys.argv) > 1 else "terrain"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
t0 = time.time(); v, f = S.make_scene(scene); print(f"{scene}: {len(f)} tris generated in {time.time()-t0:.1f}s", flush=True)
for r in range(reps):
    acc = api.BVHAccel()
    t0 = time.time(); acc.Build(len(f), v, f); t1 = time.time()
    st = acc.GetStatistics()
    print(f"  build {r}: wall {1e3*(t1-t0):.1f} ms, device {st['build_secs']*1e3:.2f} ms, nodes {st['num_leaf_nodes']+st['num_branch_nodes']}, depth {st['max_tree_depth']}", flush=True)
    acc.free()
if len(sys.argv) > 3 and sys.argv[3] == "ref":
    for r in range(2):
        acc = api.BVHAccel()
        t0 = time.time(); acc.Build(len(f), v, f, flags=api.BUILD_REFERENCE_TREE); t1 = time.time()
        st = acc.GetStatistics()
        print(f"  conformance build {r}: wall {1e3*(t1-t0):.1f} ms, device {st['build_secs']*1e3:.2f} ms, nodes "
              f"{st['num_leaf_nodes']+st['num_branch_nodes']}, depth {st['max_tree_depth']}", flush=True)
        acc.free()
