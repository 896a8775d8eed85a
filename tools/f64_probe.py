"""BVHAccel<double> throughput probe: host rays -> host records through nrt_traverse_f64 (pinned buffers), fast kernel
and reference-order kernel, on the bench scene's primary rays and on incoherent rays."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nanort_b200 import api, scenes as S

for scene in ("sphere_grid", "terrain"):
    v, f = S.make_scene(scene)
    v64 = v.astype(np.float64)
    acc = api.BVHAccelF64()
    t0 = time.time(); acc.Build(len(f), v64, f); t1 = time.time()
    W, H = 1920, 1080
    cam = S.scene_camera(scene, W, H)
    prim = S.primary_rays(cam, W, H, spp=1, seed=3)
    inc = S.incoherent_rays(v.min(axis=0), v.max(axis=0), 1 << 21, seed=5)
    print(f"== {scene}: {len(f)} triangles, build (wall, incl. upload) {1e3 * (t1 - t0):.1f} ms", flush=True)
    for name, r32 in (("primary", prim), ("incoherent", inc)):
        n = len(r32)
        rays = api.PinnedArray(n, api.RAY64_DTYPE)
        rays.array["org"], rays.array["dir"] = r32["org"], r32["dir"]
        rays.array["min_t"], rays.array["max_t"] = r32["min_t"], r32["max_t"]
        hits, mask = api.PinnedArray(n, api.HIT64_DTYPE), api.PinnedArray(n, np.uint8)
        out = {}
        for label, fl in (("fast", api.TRAVERSE_FAST), ("reference-order", api.TRAVERSE_CONFORMANCE)):
            best = 1e9
            for rep in range(4):
                t0 = time.perf_counter()
                acc.Traverse(rays.array, flags=fl, hits=hits.array, mask=mask.array)
                best = min(best, time.perf_counter() - t0)
            out[label] = (best, int(mask.array.sum()), hits.array["t"].copy())
            print(f"  {name:10s} {label:16s} {n} rays {1e3 * best:8.2f} ms {n / best / 1e6:8.1f} Mrays/s hits {out[label][1]}", flush=True)
        # kernels alone: device-resident rays, CUDA events
        import torch
        d_r = torch.as_tensor(rays.array.view(np.uint8).reshape(-1), device="cuda")
        d_h = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
        for label, fl in (("fast", api.TRAVERSE_FAST), ("reference-order", api.TRAVERSE_CONFORMANCE)):
            best = 1e9
            for rep in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                acc.TraverseDevice(d_r.data_ptr(), n, d_h.data_ptr(), flags=fl)
                e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(f"  {name:10s} {label:16s} device-resident {best:8.3f} ms {n / best / 1e3:8.1f} Mrays/s", flush=True)
        assert out["fast"][1] == out["reference-order"][1]
        assert np.array_equal(out["fast"][2], out["reference-order"][2])  # t is bit-equal even where an exact tie picks another prim
