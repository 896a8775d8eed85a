"""Experiment: how much does grouping the AO rays by direction octant (and by octant + origin cell) help the AO
traversal launch?  Uses the exported AO ray set of one 1920x1080x8spp pass of the config-2 scene."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nanort_b200 import api, scenes as S

W, H, spp = 1920, 1080, 8
v, f = S.make_scene("sphere_grid")
acc = api.BVHAccel(); acc.Build(len(f), v, f)
cam = S.scene_camera("sphere_grid", W, H)
bmin, bmax = acc.BoundingBox()
p = api.AoParams()
for i in range(12): p.cam[i] = float(cam[i])
p.width, p.height, p.spp, p.sample0, p.seed = W, H, spp, 0, 1
p.tile_w, p.tile_h, p.shard, p.n_shards = 64, 8, 0, 1
p.ray_min_t, p.ray_max_t, p.ao_min_t, p.ao_max_t = 1e-3, 1e30, 1e-3, 0.25 * float(np.linalg.norm(bmax - bmin))
n = W * H * spp
accum = torch.zeros(W * H, dtype=torch.float32, device="cuda")
d_p = torch.empty(n * 36, dtype=torch.uint8, device="cuda")
d_a = torch.empty(n * 36, dtype=torch.uint8, device="cuda")
n_p, n_a = acc.ExportAOWorkload(p, accum.data_ptr(), d_p.data_ptr(), d_a.data_ptr())
ao = d_a[: n_a * 36].view(torch.float32).view(-1, 9)
st = torch.cuda.current_stream().cuda_stream
d_hits = torch.empty(n_a * 16, dtype=torch.uint8, device="cuda")

def timed(rays_t, label):
    fn = lambda: acc.TraverseDevice(rays_t.data_ptr(), n_a, d_hits.data_ptr(), stream=st)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{label:40s} {ms:7.3f} ms  {n_a/ms/1e3:8.1f} Mrays/s")
    return ms

timed(ao, "queue order (as spawned)")
d = ao[:, 3:6]
octant = ((d[:, 0] < 0).int() | ((d[:, 1] < 0).int() << 1) | ((d[:, 2] < 0).int() << 2))
order = torch.sort(octant, stable=True).indices
timed(ao[order].contiguous(), "stable sort by octant (8 groups)")
# per 64K-ray block (a bounded on-chip-sized window) sorted by octant: what a blocked queue could do
blk = torch.arange(n_a, device="cuda") // 65536
order2 = torch.sort(blk * 8 + octant, stable=True).indices
timed(ao[order2].contiguous(), "sort by octant inside 64K-ray blocks")
blk = torch.arange(n_a, device="cuda") // 4096
order3 = torch.sort(blk * 8 + octant, stable=True).indices
timed(ao[order3].contiguous(), "sort by octant inside 4K-ray blocks")
blk = torch.arange(n_a, device="cuda") // 256
order4 = torch.sort(blk * 8 + octant, stable=True).indices
timed(ao[order4].contiguous(), "sort by octant inside 256-ray blocks")
