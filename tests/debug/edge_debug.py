import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nanort_b200 import api
from oracle import orc
from edge_cases import degenerate_mesh, hostile_rays
port = orc.Port()
v, f = degenerate_mesh()
nodes, idx, _ = port.build(v, f, None, orc.MODE_CPP11)
rays = hostile_rays(v[:34 * 3].min(axis=0) - 1, v[:34 * 3].max(axis=0) + 1)
wh, wm = port.traverse(nodes, idx, v, f, rays, cpp11=True)
acc = api.BVHAccel(); acc.Adopt(nodes, idx, v, f)
gh, gm = acc.Traverse(rays, flags=api.TRAVERSE_CONFORMANCE)
bad = np.nonzero(gm != wm)[0]
print("conformance mask mismatches", len(bad), "classes", sorted(set(((b // 8) % 14) if b % 8 == 0 else -1 for b in bad)))
for b in bad[:12]:
    print(b, (b // 8) % 14 if b % 8 == 0 else -1, rays[b], "want", wm[b], wh[b], "got", gm[b], gh[b])
both = (gm == 1) & (wm == 1)
diff = np.nonzero(both & (gh.view(np.uint8).reshape(-1, 16) != wh.view(np.uint8).reshape(-1, 16)).any(axis=1))[0]
print("record mismatches", len(diff), sorted(set(((b // 8) % 14) if b % 8 == 0 else -1 for b in diff)))
for b in diff[:8]:
    print(b, (b // 8) % 14 if b % 8 == 0 else -1, rays[b], "want", wh[b], "got", gh[b])
fast = api.BVHAccel(); fast.Build(len(f), v, f)
fh, fm = fast.Traverse(rays)
bad = np.nonzero(fm != wm)[0]
print("fast mask mismatches", len(bad), sorted(set(((b // 8) % 14) if b % 8 == 0 else -1 for b in bad)))
for b in bad[:10]:
    print(b, (b // 8) % 14 if b % 8 == 0 else -1, rays[b], "want", wm[b], wh[b], "got", fm[b], fh[b])
print("---- class 6 investigation")
from helpers import check_tree_structure
fh2, fm2 = acc.Traverse(rays)  # fast kernel over the adopted reference tree
print("fast kernel on reference tree: mask mismatches", (fm2 != wm).sum(), sorted(set(((b // 8) % 14) if b % 8 == 0 else -1 for b in np.nonzero(fm2 != wm)[0])))
fn, fi = fast.GetNodes(), fast.GetIndices()
try:
    check_tree_structure(fn, fi, v, f)
    print("fast-built tree structure ok", len(fn))
except AssertionError as e:
    print("STRUCTURE FAIL", e)
oh, om = port.traverse(fn, fi, v, f, rays, cpp11=True)
print("oracle over fast-built tree vs oracle over reference tree: mask mismatches", (om != wm).sum())
ch, cm = fast.Traverse(rays, flags=api.TRAVERSE_CONFORMANCE)
print("conformance kernel over fast-built tree: mask mismatches vs oracle-same-tree", (cm != om).sum(), " fast vs oracle-same-tree", (fm != om).sum())
