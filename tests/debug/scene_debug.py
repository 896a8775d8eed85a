"""Debug helper: fast two-level traversal vs the oracle on the row scene; prints the mismatch classes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nanort_b200 import api, scenes as S
from oracle import orc
import test_gpu_scene as T

kind = sys.argv[1] if len(sys.argv) > 1 else "row"
insts = S.instances_mixed() if kind == "mixed" else S.instances_row()
port = orc.PortScene(insts, cpp11=True)
sc = T._gpu_scene(insts, api.BUILD_FAST, api.BUILD_FAST)
rays = T._rays_for(insts, 200000, seed=6)
if kind == "row":
    rays = T._row_rays(rays)
ph, pm = port.traverse(rays, threads=8)
gh, gm = sc.Traverse(rays)
print("mask mismatches", (pm != gm).sum(), "of", len(rays), "hits", pm.sum())
bad = np.nonzero(pm != gm)[0]
for i in bad[:10]:
    print(i, rays[i], "port", pm[i], ph[i], "gpu", gm[i], gh[i])
    print("   list", port.list_node_intersections(rays[i])[0][:6], len(port.list_node_intersections(rays[i])[0]))
hit = (pm == 1) & (gm == 1)
same = hit & (ph["node_id"] == gh["node_id"]) & (ph["prim_id"] == gh["prim_id"])
print("same pick", same.sum(), "bit-equal", ph[same].tobytes() == gh[same].tobytes())
other = hit & ~same
rel = np.abs(ph["t"][other] - gh["t"][other]) / np.maximum(ph["t"][other], 1e-6)
print("other", other.sum(), "rel max", rel.max() if other.any() else 0, "n>1e-5", (rel > 1e-5).sum())
for i in np.nonzero(other)[0][np.argsort(-rel)][:8]:
    print(i, rays[i], "port", ph[i], "gpu", gh[i], len(port.list_node_intersections(rays[i])[0]))
