"""Policy sweep of the unified two-level kernel on the config-4 instanced scene: 4K primary rays (coherent) and
random interior rays (incoherent).  Every variant must reproduce the default's output bit for bit."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nanort_b200 import api, scenes as S

base = S.sphere_grid()
insts = S.instances_grid(10, 10, base=base)
blas = api.BVHAccel(); blas.Build(len(base[1]), base[0], base[1])
sc = api.Scene()
for v, f, x in insts:
    sc.AddNode(blas, x)
sc.Commit()
cam = S.scene_camera("instanced", 3840, 2160)
prim = S.primary_rays(cam, 3840, 2160, spp=1, seed=1, min_t=0.0)
lo, hi = sc.GetBoundingBox()
inc = S.incoherent_rays(lo, hi + np.float32([0, 3, 0]), 4 << 20, seed=4, axis_parallel_fraction=0.0)
inc["min_t"] = 0.0
st = torch.cuda.current_stream().cuda_stream
names = {0: "default <8,16,8>", 2: "<6,16,8>", 3: "<8,16,8>", 4: "<5,16,8>", 5: "<7,8,8>",
         6: "<7,24,8>", 7: "<7,16,4>", 8: "<7,16,12>", 9: "<7,16,16>"}
for label, rays in (("primary 4K", prim), ("incoherent 4Mi", inc)):
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1, 36)).cuda()
    n = len(rays)
    d_hits = torch.zeros(n, 32, dtype=torch.uint8, device="cuda")
    d_mask = torch.zeros(n, dtype=torch.uint8, device="cuda")
    ref = None
    for var in sorted(names):
        fn = lambda: sc.TraverseDevice(d_rays.data_ptr(), n, d_hits.data_ptr(), d_mask.data_ptr(), flags=(var << 8), stream=st)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        out = d_hits.cpu().numpy().tobytes()
        if ref is None: ref = out
        print(f"{label:16s} variant {var} {names[var]:18s} {ms:7.3f} ms {n/ms/1e3:8.1f} Mrays/s  same output {out == ref}  hit {d_mask.float().mean().item():.3f}")
