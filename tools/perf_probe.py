"""Quick on-GPU probe: build + primary/AO pass timings (not a bench; prints to stdout)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nanort_b200 import api, scenes as S

def run(name, W, H, spp, kw={}):
    v, f = S.make_scene(name, **kw)
    acc = api.BVHAccel()
    t0 = time.time(); acc.Build(len(f), v, f); t1 = time.time()
    st = acc.GetStatistics()
    print(f"{name}: {len(f)} tris, build wall {t1-t0:.3f}s device {st['build_secs']*1e3:.2f} ms, stats {st}", flush=True)
    cam = S.scene_camera(name, W, H)
    bmin, bmax = acc.BoundingBox()
    diag = float(np.linalg.norm(bmax - bmin))
    p = api.AoParams()
    for i in range(12): p.cam[i] = float(cam[i])
    p.width, p.height, p.spp, p.sample0, p.seed = W, H, spp, 0, 1
    p.tile_w, p.tile_h, p.shard, p.n_shards = 64, 8, 0, 1
    p.ray_min_t, p.ray_max_t, p.ao_min_t, p.ao_max_t = 1e-3, 1e30, 1e-3, 0.25 * diag
    p.flags = 0
    accum = torch.zeros(W * H, dtype=torch.float32, device="cuda")
    for it in range(3):
        accum.zero_()
        r = acc.RenderAO(p, accum.data_ptr())
        rays = r.primary_rays + r.ao_rays
        print(f"   pass {it}: primary {r.primary_rays} ao {r.ao_rays} ao_hits {r.ao_hits} total {r.total_ms:.2f} ms "
              f"trav {r.traverse_ms:.2f} ms -> {rays / r.total_ms / 1e3:.1f} Mrays/s (trav-only {rays / r.traverse_ms / 1e3:.1f}) "
              f"launches {r.launches}", flush=True)
    img = (accum / spp).cpu().numpy().reshape(H, W)
    print("   mean visibility", float(img.mean()), "min", float(img.min()), "max", float(img.max()))
    return acc

if __name__ == "__main__":
    run("cornell", 512, 512, 4)
    run("sphere_grid", 1920, 1080, 4)
    run("terrain", 1920, 1080, 4)
