"""Dumps fast-path vs oracle mismatches for offline analysis (gpurun_out/mismatch_*.npz)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import orc
from nanort_b200 import api, scenes as S

port = orc.Port()
os.makedirs("gpurun_out", exist_ok=True)
for scene, kw, fix in (("cornell", {}, True), ("cornell", {}, False), ("sphere_grid", dict(nx=3, nz=3), True)):
    v, f = S.make_scene(scene, **kw)
    nodes, idx, _ = port.build(v, f, mode=orc.MODE_CPP11 | (orc.MODE_FIXBINS if fix else 0))
    cam = S.scene_camera(scene, 192, 160)
    rays = np.concatenate([S.primary_rays(cam, 192, 160, spp=1, seed=11),
                           S.incoherent_rays(v.min(axis=0), v.max(axis=0), 60000, seed=5)])
    want_h, want_m = port.traverse(nodes, idx, v, f, rays, threads=8)
    acc = api.BVHAccel(); acc.Adopt(nodes, idx, v, f)
    got_h, got_m = acc.Traverse(rays, flags=api.TRAVERSE_FAST)
    conf_h, conf_m = acc.Traverse(rays, flags=api.TRAVERSE_CONFORMANCE)
    both = got_m.astype(bool) & want_m.astype(bool)
    bad = (got_m != want_m) | (both & (got_h["prim_id"] != want_h["prim_id"])) | (both & (got_h["t"] != want_h["t"]))
    print(scene, fix, "n", len(rays), "bad", int(bad.sum()), "conf==want",
          np.array_equal(conf_m, want_m) and np.array_equal(conf_h[both].view(np.uint32), want_h[both].view(np.uint32)))
    np.savez(f"gpurun_out/mismatch_{scene}_{int(fix)}.npz", rays=rays[bad], got=got_h[bad], want=want_h[bad],
             got_m=got_m[bad], want_m=want_m[bad], idx=np.nonzero(bad)[0])
