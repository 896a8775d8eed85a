"""configs[0] of BASELINE.json on the GPU: Cornell box (34 triangles), 512x512 primary rays."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nanort_b200 import api, scenes as S

v, f = S.make_scene("cornell")
acc = api.BVHAccel(); acc.Build(len(f), v, f)
cam = S.scene_camera("cornell", 512, 512)
rays = S.primary_rays(cam, 512, 512, spp=1, seed=1)
n = len(rays)
d_r = torch.from_numpy(rays.view(np.uint8).reshape(-1, 36)).cuda()
d_h = torch.empty(n * 16, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(5): acc.TraverseDevice(d_r.data_ptr(), n, d_h.data_ptr(), stream=st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): acc.TraverseDevice(d_r.data_ptr(), n, d_h.data_ptr(), stream=st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
pr = api.PinnedArray(n, S.RAY_DTYPE); pr.array[:] = rays
ph, pm = api.PinnedArray(n, S.HIT_DTYPE), api.PinnedArray(n, np.uint8)
for _ in range(3): acc.Traverse(pr.array, hits=ph.array, mask=pm.array)
t0 = time.perf_counter()
for _ in range(20): acc.Traverse(pr.array, hits=ph.array, mask=pm.array)
host_ms = (time.perf_counter() - t0) / 20 * 1e3
print(f"cornell {len(f)} tris, {n} primary rays: device-resident {ms*1e3:.1f} us = {n/ms/1e3:.0f} Mrays/s; host rays -> host hits "
      f"{host_ms*1e3:.0f} us = {n/host_ms/1e3:.0f} Mrays/s; build {acc.GetStatistics()['build_secs']*1e3:.3f} ms; hit rate {pm.array.mean():.3f}")
