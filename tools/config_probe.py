"""On-GPU probes of BASELINE.json's other configs (parity-test cases, not bench lines):
  config 2: 1M-triangle terrain, 1920x1080, path_tracer loop (diffuse + area light), rays/s counts every Traverse
  config 3: 10M-triangle flattened instanced scene: Build + 3840x2160 primary rays
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nanort_b200 import api, scenes as S

def path_config(spp=8):
    v, f = S.make_scene("terrain")
    v, f, l0, ln = S.with_area_light(v, f, (0.0, 6.0, 0.0), 2.0, 2.0)
    mats = np.concatenate([S.material(diffuse=(0.7, 0.7, 0.7)), S.material(emission=(20, 20, 20))])
    ids = np.zeros(len(f), np.uint32); ids[l0:] = 1
    emissive = np.arange(l0, l0 + ln, dtype=np.uint32)
    acc = api.BVHAccel(); acc.Build(len(f), v, f)
    W, H = 1920, 1080
    cam = S.scene_camera("terrain", W, H)
    d_m = torch.as_tensor(mats.view(np.float32).reshape(-1), device="cuda")
    d_i = torch.as_tensor(ids.astype(np.int32), device="cuda")
    d_e = torch.as_tensor(emissive.astype(np.int32), device="cuda")
    p = api.PathParams()
    for i in range(12): p.cam[i] = float(cam[i])
    p.width, p.height, p.spp, p.sample0, p.seed = W, H, spp, 0, 3
    p.tile_w, p.tile_h, p.shard, p.n_shards = 64, 8, 0, 1
    p.max_bounces, p.ray_min_t, p.ray_max_t = 10, 1e-3, 1e30
    p.n_materials, p.n_emissive = len(mats), len(emissive)
    p.d_materials, p.d_material_ids, p.d_emissive_faces = d_m.data_ptr(), d_i.data_ptr(), d_e.data_ptr()
    p.d_facevarying_normals, p.flags = None, 0
    accum = torch.zeros(W * H * 3, dtype=torch.float32, device="cuda")
    for it in range(3):
        accum.zero_()
        r = acc.RenderPath(p, accum.data_ptr())
        rays = r.radiance_rays + r.shadow_rays
        print(f"config2 path tracer {W}x{H}x{spp}spp, <=10 bounces: camera {r.camera_rays} radiance {r.radiance_rays} "
              f"shadow {r.shadow_rays} total {r.total_ms:.2f} ms (trav {r.traverse_ms:.2f}) -> {rays / r.total_ms / 1e3:.1f} Mrays/s, "
              f"{r.launches} launches", flush=True)
    img = accum.cpu().numpy().reshape(H, W, 3) / spp
    print("   mean radiance", img.mean(axis=(0, 1)))

def build_config():
    v, f = S.make_scene("instanced")
    acc = api.BVHAccel()
    t0 = time.time(); acc.Build(len(f), v, f); t1 = time.time()
    st = acc.GetStatistics()
    print(f"config3 build: {len(f)} tris, device {st['build_secs']*1e3:.1f} ms, wall incl. upload {1e3*(t1-t0):.1f} ms, "
          f"nodes {st['num_leaf_nodes']+st['num_branch_nodes']}, depth {st['max_tree_depth']}", flush=True)
    W, H = 3840, 2160
    cam = S.scene_camera("instanced", W, H)
    bmin, bmax = acc.BoundingBox()
    p = api.AoParams()
    for i in range(12): p.cam[i] = float(cam[i])
    p.width, p.height, p.spp, p.sample0, p.seed = W, H, 1, 0, 1
    p.tile_w, p.tile_h, p.shard, p.n_shards = 64, 8, 0, 1
    p.ray_min_t, p.ray_max_t, p.ao_min_t, p.ao_max_t = 1e-3, 1e30, 1e-3, 0.02 * float(np.linalg.norm(bmax - bmin))
    n = W * H
    accum = torch.zeros(n, dtype=torch.float32, device="cuda")
    d_p = torch.empty(n * 36, dtype=torch.uint8, device="cuda")
    d_a = torch.empty(n * 36, dtype=torch.uint8, device="cuda")
    n_p, n_a = acc.ExportAOWorkload(p, accum.data_ptr(), d_p.data_ptr(), d_a.data_ptr())
    hits = torch.empty(n * 16, dtype=torch.uint8, device="cuda")
    best = 1e9
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); acc.TraverseDevice(d_p.data_ptr(), n_p, hits.data_ptr()); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"config3 4K primary: {n_p} rays in {best:.3f} ms -> {n_p / best / 1e3:.1f} Mrays/s; primary hits {n_a}", flush=True)

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "path"): path_config(int(sys.argv[2]) if len(sys.argv) > 2 else 8)
    if which in ("all", "build"): build_config()
