"""configs[4] of BASELINE.json from one rank's point of view: 1,002,528-triangle terrain, 4096x4096, 256 spp, primary +
1-bounce AO, tiles of 64x64 pixels dealt round-robin over 8 ranks -- this process traces shard 0 of 8."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nanort_b200 import api, scenes as S

W = H = 4096
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 256
v, f = S.make_scene("terrain")
acc = api.BVHAccel(); acc.Build(len(f), v, f)
cam = S.scene_camera("terrain", W, H)
bmin, bmax = acc.BoundingBox()
p = api.AoParams()
for i in range(12): p.cam[i] = float(cam[i])
p.width, p.height, p.spp, p.sample0, p.seed = W, H, spp, 0, 1
p.tile_w, p.tile_h, p.shard, p.n_shards = 64, 64, 0, 8
p.ray_min_t, p.ray_max_t, p.ao_min_t, p.ao_max_t = 1e-3, 1e30, 1e-3, 0.25 * float(np.linalg.norm(bmax - bmin))
accum = torch.zeros(W * H, dtype=torch.float32, device="cuda")
for it in range(2):
    accum.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = acc.RenderAO(p, accum.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"shard 0/8 of {W}x{H}x{spp}spp on {len(f)} tris: {r.primary_rays} primary + {r.ao_rays} AO rays in {ms:.1f} ms "
          f"-> {(r.primary_rays + r.ao_rays) / ms / 1e3:.1f} Mrays/s ({r.launches} launches); build {acc.GetStatistics()['build_secs']*1e3:.2f} ms", flush=True)
