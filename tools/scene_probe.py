"""Config-4 probe: the instanced scene as a two-level scene (100 instances of one 100K-triangle grid) against the
same triangles flattened into one 10M-triangle accel: commit/build time and 4K primary-ray throughput."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nanort_b200 import api, scenes as S

W, H = 3840, 2160
cx = int(sys.argv[1]) if len(sys.argv) > 1 else 10
base = S.sphere_grid()
insts = S.instances_grid(cx, cx, base=base)
t0 = time.time()
blas = api.BVHAccel(); blas.Build(len(base[1]), base[0], base[1])
sc = api.Scene()
for v, f, x in insts:
    sc.AddNode(blas, x)
t1 = time.time()
assert sc.Commit()
t2 = time.time()
for _ in range(3):
    ta = time.time(); sc.Commit(); tb = time.time()
print(f"two-level: BLAS build wall {1e3*(t1-t0):.1f} ms, commit of {len(insts)} instances wall {1e3*(tb-ta):.2f} ms")
cam = S.scene_camera("instanced", W, H)
rays = S.primary_rays(cam, W, H, spp=1, seed=1, min_t=0.0)
d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1, 36)).cuda()
n = len(rays)
d_hits = torch.zeros(n, 32, dtype=torch.uint8, device="cuda")
d_mask = torch.zeros(n, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = timed(lambda: sc.TraverseDevice(d_rays.data_ptr(), n, d_hits.data_ptr(), d_mask.data_ptr(), stream=st))
print(f"two-level fast (unified walk): {ms:.3f} ms  {n/ms/1e3:.1f} Mrays/s  hit rate {d_mask.float().mean().item():.3f}")
m_u = d_mask.cpu().numpy().copy(); h_u = d_hits.cpu().numpy().copy()
sc.TraverseDevice(d_rays.data_ptr(), n, d_hits.data_ptr(), d_mask.data_ptr(), stream=st)
torch.cuda.synchronize()
ms_c = timed(lambda: sc.TraverseDevice(d_rays.data_ptr(), n, d_hits.data_ptr(), d_mask.data_ptr(), flags=api.TRAVERSE_CONFORMANCE, stream=st), reps=3)
print(f"two-level list kernel: {ms_c:.3f} ms  {n/ms_c/1e3:.1f} Mrays/s")
m2 = d_mask.cpu().numpy().copy(); h2 = d_hits.cpu().numpy().view(api.SCENE_HIT_DTYPE).reshape(-1).copy()
if cx <= 10:
    v, f = S.instanced(cx, cx)
    flat = api.BVHAccel(); t0 = time.time(); flat.Build(len(f), v, f); t1 = time.time()
    print(f"flattened: {len(f)} tris, build wall {1e3*(t1-t0):.1f} ms (device {1e3*flat.GetStatistics()["build_secs"]:.2f} ms)")
    d_h16 = torch.zeros(n, 16, dtype=torch.uint8, device="cuda")
    ms_f = timed(lambda: flat.TraverseDevice(d_rays.data_ptr(), n, d_h16.data_ptr(), d_mask.data_ptr(), stream=st))
    print(f"flattened fast: {ms_f:.3f} ms  {n/ms_f/1e3:.1f} Mrays/s")
    m1 = d_mask.cpu().numpy(); h1 = d_h16.cpu().numpy().view(api.HIT_DTYPE).reshape(-1)
    print("mask agreement", (m1 == m2).mean())
    both = (m1 == 1) & (m2 == 1)
    nf = len(base[1])
    same = (h2["node_id"][both].astype(np.int64) * nf + h2["prim_id"][both]) == h1["prim_id"][both]
    print("same triangle", same.mean(), "max rel dt", (np.abs(h2["t"][both][same] - h1["t"][both][same]) / h1["t"][both][same]).max())
