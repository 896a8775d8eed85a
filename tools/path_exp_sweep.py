"""Sweep of the path tracer's experiment policies (NRT_AO_EXP digits 3 and 4: radiance launch, shadow launch;
csrc/traverse.cu) on bench.py's configs[2] scene at 16 spp: whole-pass device time, mean radiance as a sanity check."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nanort_b200 import api, scenes as S

settings = sys.argv[1].split(",") if len(sys.argv) > 1 else ["0000", "0010", "0020", "0050", "0060", "0002", "0062"]
W, H, spp = 1920, 1080, int(os.environ.get("NRT_SWEEP_SPP", "16"))
dev = torch.device("cuda:0")
v, f = S.make_scene("terrain")
v, f, l0, ln = S.with_area_light(v, f, (0.0, 6.0, 0.0), 2.0, 2.0)
mats = np.concatenate([S.material(diffuse=(0.7, 0.7, 0.7)), S.material(emission=(20, 20, 20))])
ids = np.zeros(len(f), np.uint32); ids[l0:] = 1
emissive = np.arange(l0, l0 + ln, dtype=np.uint32)
acc = api.BVHAccel(); acc.Build(len(f), v, f)
cam = S.scene_camera("terrain", W, H)
d_m = torch.as_tensor(mats.view(np.float32).reshape(-1), device=dev)
d_i = torch.as_tensor(ids.astype(np.int32), device=dev)
d_e = torch.as_tensor(emissive.astype(np.int32), device=dev)
p = api.PathParams()
for i in range(12): p.cam[i] = float(cam[i])
p.width, p.height, p.spp, p.sample0, p.seed = W, H, spp, 0, 3
p.tile_w, p.tile_h, p.shard, p.n_shards = 64, 8, 0, 1
p.max_bounces, p.ray_min_t, p.ray_max_t = 10, 1e-3, 1e30
p.n_materials, p.n_emissive = len(mats), len(emissive)
p.d_materials, p.d_material_ids, p.d_emissive_faces = d_m.data_ptr(), d_i.data_ptr(), d_e.data_ptr()
p.d_facevarying_normals, p.flags = None, 0
accum = torch.zeros(W * H * 3, dtype=torch.float32, device=dev)
print(f"== terrain + area light, {W}x{H}x{spp} spp path loop", flush=True)
for st in settings:
    os.environ["NRT_AO_EXP"] = st
    best = None
    for rep in range(3):
        accum.zero_()
        r = acc.RenderPath(p, accum.data_ptr())
        if best is None or r.total_ms < best[0]:
            best = (float(r.total_ms), float(r.traverse_ms), int(r.radiance_rays + r.shadow_rays))
    mean = float((accum / spp).mean().item())
    print(f"exp {st}: total {best[0]:8.3f} ms trav {best[1]:8.3f} ms {best[2] / best[0] / 1e3:8.1f} Mrays/s mean radiance {mean:.6f}", flush=True)
os.environ["NRT_AO_EXP"] = ""
