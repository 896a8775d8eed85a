"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
production + conformance build, both traversal kernels, the AO and path passes, a two-level scene, the fp64 accel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nanort_b200 import api, scenes as S

v, f = S.make_scene("sphere_grid", nx=2, nz=2)
acc = api.BVHAccel(); acc.Build(len(f), v, f)
ref = api.BVHAccel(); ref.Build(len(f), v, f, flags=api.BUILD_REFERENCE_TREE)
cam = S.scene_camera("sphere_grid", 96, 64)
rays = np.concatenate([S.primary_rays(cam, 96, 64, spp=1, seed=7), S.incoherent_rays(v.min(axis=0), v.max(axis=0), 4096, seed=8)])
for a in (acc, ref):
    for fl in (api.TRAVERSE_FAST, api.TRAVERSE_CONFORMANCE):
        h, m = a.Traverse(rays, flags=fl)
print("traverse ok", int(m.sum()))
p = api.AoParams()
for i in range(12): p.cam[i] = float(cam[i])
p.width, p.height, p.spp, p.sample0, p.seed = 96, 64, 2, 0, 1
p.tile_w, p.tile_h, p.shard, p.n_shards = 64, 8, 0, 1
p.ray_min_t, p.ray_max_t, p.ao_min_t, p.ao_max_t = 1e-3, 1e30, 1e-3, 2.0
accum = torch.zeros(96 * 64, dtype=torch.float32, device="cuda")
r = acc.RenderAO(p, accum.data_ptr())
print("ao ok", r.primary_rays, r.ao_rays)
# round 2, third session: any-hit occlusion rays (AO launch, nrt_traverse), 32-byte ray records
p.flags = api.TRAVERSE_ANY_HIT
accum2 = torch.zeros(96 * 64, dtype=torch.float32, device="cuda")
r2 = acc.RenderAO(p, accum2.data_ptr())
assert torch.equal(accum, accum2) and r2.ao_hits == r.ao_hits
p.flags = 0
ha, ma = acc.Traverse(rays, flags=api.TRAVERSE_ANY_HIT)
assert np.array_equal(ma, m)
r32 = np.ascontiguousarray(rays.view(np.uint8).reshape(-1, 36)[:, :32]).view(np.dtype((np.void, 32))).reshape(-1)
h32, _ = acc.Traverse(r32, flags=api.TRAVERSE_RAY32, mask=False)
assert np.array_equal(h32.view(np.uint32), acc.Traverse(rays)[0].view(np.uint32))
print("any-hit / ray32 ok")
insts = S.instances_mixed(7)
accels, sc = {}, api.Scene()
for iv, jf, x in insts:
    key = (iv.ctypes.data, jf.ctypes.data)
    if key not in accels:
        accels[key] = api.BVHAccel(); accels[key].Build(len(jf), iv, jf)
    sc.AddNode(accels[key], x)
sc.Commit()
srays = S.incoherent_rays(np.float32([-8, -4, -8]), np.float32([8, 4, 8]), 4096, seed=3)
srays["min_t"] = 0.0
for fl in (api.TRAVERSE_FAST, api.TRAVERSE_CONFORMANCE):
    sh, sm = sc.Traverse(srays, flags=fl)
print("scene ok", int(sm.sum()))
ps = api.AoParams()
scam = S.look_at((0.0, 6.0, 14.0), (0.0, 0.0, 0.0), aspect=96 / 64)
for i in range(12): ps.cam[i] = float(scam[i])
ps.width, ps.height, ps.spp, ps.sample0, ps.seed = 96, 64, 2, 0, 1
ps.tile_w, ps.tile_h, ps.shard, ps.n_shards = 64, 8, 0, 1
ps.ray_min_t, ps.ray_max_t, ps.ao_min_t, ps.ao_max_t = 1e-3, 1e30, 1e-3, 2.0
saccum = torch.zeros(96 * 64, dtype=torch.float32, device="cuda")
rs = sc.RenderAO(ps, saccum.data_ptr())
assert float(saccum.double().sum().item()) == float(rs.primary_rays - rs.ao_hits)
print("scene ao ok", rs.primary_rays, rs.ao_rays, rs.ao_hits)
a64 = api.BVHAccelF64(); a64.Build(len(f), v.astype(np.float64), f)
r64 = np.zeros(2048, api.RAY64_DTYPE)
r64["org"], r64["dir"], r64["min_t"], r64["max_t"] = rays["org"][:2048], rays["dir"][:2048], 0.0, 1e30
r64b = np.zeros(4096, api.RAY64_DTYPE)
r64b["org"], r64b["dir"], r64b["min_t"], r64b["max_t"] = rays["org"][-4096:], rays["dir"][-4096:], 0.0, 1e30
for rr in (r64, r64b):
    h64, m64 = a64.Traverse(rr)  # fast kernel (f64_fast.cuh) + lazily derived PairNodeD / TriD layout
    c64h, c64m = a64.Traverse(rr, flags=api.TRAVERSE_CONFORMANCE)
    assert np.array_equal(m64, c64m) and np.array_equal(h64["t"], c64h["t"])
print("f64 ok", int(m64.sum()))
# zero-copy small-call path (<= 64 rays): float fast / conformance, fp64 reference order
for n in (1, 33, 64):
    for fl in (api.TRAVERSE_FAST, api.TRAVERSE_CONFORMANCE):
        hs, ms = acc.Traverse(rays[-n:], flags=fl)
    hs64, ms64 = a64.Traverse(r64b[-n:], flags=api.TRAVERSE_CONFORMANCE)
print("small calls ok", int(ms.sum()), int(ms64.sum()))
big_v, big_f = S.make_scene("terrain", n=96)
big = api.BVHAccel(); big.Build(len(big_f), big_v, big_f)
print("terrain build ok", big.GetStatistics()["num_leaf_nodes"])
# round 2: middle phase + segmented small blocks at several sizes (one subtree, one mid node, level-synchronous + mid + subtrees)
for n_side in (8, 30, 96):
    tv, tf = S.make_scene("terrain", n=n_side)
    for ml in (1, 4):
        t = api.BVHAccel(); t.Build(len(tf), tv, tf, options=api.BVHBuildOptions(min_leaf_primitives=ml))
        st = t.GetStatistics()
        assert st["num_leaf_nodes"] == st["num_branch_nodes"] + 1
print("builder sizes ok")
c64 = api.BVHAccelF64(); c64.Build(len(f), v.astype(np.float64) * (1 + 1e-12), f, flags=api.BUILD_REFERENCE_TREE)
h64b, m64b = c64.Traverse(r64)
print("f64 conformance build ok", int(m64b.sum()), c64.GetStatistics()["max_tree_depth"])
centers = np.random.default_rng(1).uniform(-1, 1, (500, 3)).astype(np.float32)
sp = api.BVHAccel(); sp.BuildSpheres(centers, np.full(500, 0.05, np.float32))
sph, spm = sp.Traverse(srays)
bx = api.BVHAccel(); bx.BuildBoxes(np.concatenate([centers - 0.05, centers + 0.05], axis=1))
lh, lc = bx.ListNodeIntersections(srays[:512], max_intersections=8)
print("prims ok", int(spm.sum()), int(lc.sum()))
