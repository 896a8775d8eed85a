"""Kernel-variant sweep (experiment selector in flags bits 8..15): times the traversal of the exported
primary and AO ray sets and checks every variant's hits against the default (variant 0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nanort_b200 import api, scenes as S

variants = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 2, 4, 5, 6, 8, 9, 10, 11, 21, 30, 40, 42]
scenes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["sphere_grid", "terrain"]
W, H, spp = 1920, 1080, int(os.environ.get("NRT_SWEEP_SPP", "2"))
for scene in scenes:
    v, f = S.make_scene(scene)
    acc = api.BVHAccel(); acc.Build(len(f), v, f)
    cam = S.scene_camera(scene, W, H)
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
    hits = torch.empty(n * 16, dtype=torch.uint8, device="cuda")
    ref = {}
    print(f"== {scene}: {len(f)} tris, {n_p} primary + {n_a} AO rays", flush=True)
    for var in variants:
        row = []
        for name, d_r, cnt in (("primary", d_p, n_p), ("ao", d_a, n_a)):
            best = 1e9
            for rep in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                acc.TraverseDevice(d_r.data_ptr(), cnt, hits.data_ptr(), flags=(var << 8))
                e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            h = hits[: cnt * 16].clone()
            if name not in ref:
                ref[name] = h
                same = "ref"
            else:
                a32, b32 = h.view(torch.int32).view(-1, 4), ref[name].view(torch.int32).view(-1, 4)
                nd = int((a32 != b32).any(dim=1).sum().item())
                same = "same" if nd == 0 else f"DIFF({nd})"
            row.append(f"{name} {best:7.3f} ms {cnt / best / 1e3:8.1f} Mrays/s {same}")
        print(f"variant {var:3d}: " + " | ".join(row), flush=True)
