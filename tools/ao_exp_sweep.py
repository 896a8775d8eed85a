"""Sweep of the fused AO pass's experiment policies (NRT_AO_EXP="<primary digit><ao digit>", csrc/traverse.cu): the
whole pass of the bench workload (1920x1080x16 spp) per setting, per-launch-kind times from the pass's own CUDA events,
framebuffer compared with the default's bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nanort_b200 import api, scenes as S

scenes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["sphere_grid", "terrain"]
settings = sys.argv[2].split(",") if len(sys.argv) > 2 else ["00", "10", "20", "30", "40", "50", "60", "70",
                                                             "01", "02", "03", "04", "05", "06", "07"]
W, H, spp = 1920, 1080, 16
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
    p.flags = 0
    accum = torch.zeros(W * H, dtype=torch.float32, device="cuda")
    ref = None
    print(f"== {scene}: {len(f)} tris, {W}x{H}x{spp}", flush=True)
    for st in settings:
        os.environ["NRT_AO_EXP"] = st
        best = None
        for rep in range(5):
            accum.zero_()
            r = acc.RenderAO(p, accum.data_ptr())
            if best is None or r.total_ms < best[0]:
                best = (float(r.total_ms), float(r.primary_traverse_ms), float(r.ao_traverse_ms), int(r.primary_rays), int(r.ao_rays))
        img = accum.clone()
        if ref is None:
            ref, same = img, "ref"
        else:
            same = "same" if bool(torch.equal(img, ref)) else f"DIFF({int((img != ref).sum().item())})"
        t, tp, ta, n_p, n_a = best
        print(f"exp {st}: total {t:7.3f} ms {(n_p + n_a) / t / 1e3:8.1f} Mrays/s | primary {tp:6.3f} ms {n_p / tp / 1e3:8.1f} | "
              f"ao {ta:6.3f} ms {n_a / ta / 1e3:8.1f} | frame {same}", flush=True)
os.environ["NRT_AO_EXP"] = ""
