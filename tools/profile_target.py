"""Small fixed workload for ncu captures: config-2 scene, 1920x1080, 4 spp, two passes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nanort_b200 import api, scenes as S

scene = sys.argv[1] if len(sys.argv) > 1 else "sphere_grid"
W, H, spp = 1920, 1080, 4
v, f = S.make_scene(scene)
acc = api.BVHAccel()
acc.Build(len(f), v, f)
cam = S.scene_camera(scene, W, H)
bmin, bmax = acc.BoundingBox()
p = api.AoParams()
for i in range(12): p.cam[i] = float(cam[i])
p.width, p.height, p.spp, p.sample0, p.seed = W, H, spp, 0, 1
p.tile_w, p.tile_h, p.shard, p.n_shards = 64, 8, 0, 1
p.ray_min_t, p.ray_max_t, p.ao_min_t, p.ao_max_t = 1e-3, 1e30, 1e-3, 0.25 * float(np.linalg.norm(bmax - bmin))
accum = torch.zeros(W * H, dtype=torch.float32, device="cuda")
for it in range(2):
    r = acc.RenderAO(p, accum.data_ptr())
print(scene, r.primary_rays, r.ao_rays, r.total_ms, r.traverse_ms)
