"""ncu target: one two-level traversal of the config-4 instanced scene (4K primary rays)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nanort_b200 import api, scenes as S

W, H = 3840, 2160
base = S.sphere_grid()
insts = S.instances_grid(10, 10, base=base)
blas = api.BVHAccel(); blas.Build(len(base[1]), base[0], base[1])
sc = api.Scene()
for v, f, x in insts:
    sc.AddNode(blas, x)
sc.Commit()
cam = S.scene_camera("instanced", W, H)
rays = S.primary_rays(cam, W, H, spp=1, seed=1, min_t=0.0)
d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1, 36)).cuda()
n = len(rays)
d_hits = torch.zeros(n, 32, dtype=torch.uint8, device="cuda")
d_mask = torch.zeros(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    sc.TraverseDevice(d_rays.data_ptr(), n, d_hits.data_ptr(), d_mask.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
