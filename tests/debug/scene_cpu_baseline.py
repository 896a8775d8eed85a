"""TEST INFRASTRUCTURE: the unmodified reference scene graph (oracle/_ref) on the host cores beside the GPU's two-level
traversal, on a sample of the config-4 rays -- the CPU leg that used to hang off tools/scene_probe.py --cpu."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from nanort_b200 import api, scenes as S
from oracle import orc

W, H = 3840, 2160
base = S.sphere_grid()
insts = S.instances_grid(10, 10, base=base)
blas = api.BVHAccel(); blas.Build(len(base[1]), base[0], base[1])
sc = api.Scene()
for v, f, x in insts:
    sc.AddNode(blas, x)
assert sc.Commit()
cam = S.scene_camera("instanced", W, H)
rays = S.primary_rays(cam, W, H, spp=1, seed=1, min_t=0.0)
n = len(rays)
idx = np.arange(0, n, max(1, n // 400000))
sample = rays[idx]
h2, m2 = sc.Traverse(sample)
t0 = time.time(); ref = orc.ReferenceScene(insts, cpp11=True); t1 = time.time()
print(f"reference nanosg: AddNode+Commit of {len(insts)} nodes ({len(base[1])} tris each) wall {t1-t0:.2f} s")
th = os.cpu_count()
t0 = time.time(); rh, rm = ref.traverse(sample, threads=th); t1 = time.time()
print(f"reference nanosg Scene::Traverse: {len(sample)} rays, {th} threads, {len(sample)/(t1-t0)/1e6:.3f} Mrays/s")
print("mask agreement vs reference", (rm == m2).mean())
both = (rm == 1) & (m2 == 1)
same = (rh["node_id"][both] == h2["node_id"][both]) & (rh["prim_id"][both] == h2["prim_id"][both])
print("same (instance, triangle)", same.mean(), "records bit-equal where same:", rh[both][same].tobytes() == h2[both][same].tobytes())
