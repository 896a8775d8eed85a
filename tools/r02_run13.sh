#!/bin/bash
# round-2 GPU call 13: fp64 conformance build (nrt_build_f64_ex), terrain path-parity test, fp64 throughput
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_f64.py tests/test_gpu_path.py tests/test_gpu_dropin.py -q > gpurun_out/r02_t13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t13.log
timeout 600 python - > gpurun_out/r02_f64_rate13.log 2>&1 <<'PY'
import time, numpy as np, sys
sys.path.insert(0, '.')
from nanort_b200 import api, scenes as S
v, f = S.make_scene("sphere_grid")
v64 = v.astype(np.float64)
cam = S.scene_camera("sphere_grid", 1920, 1080)
r32 = S.primary_rays(cam, 1920, 1080, spp=1, seed=1)
r = np.zeros(len(r32), api.RAY64_DTYPE)
r["org"], r["dir"], r["min_t"], r["max_t"] = r32["org"], r32["dir"], 1e-3, 1e30
for flags, name in ((api.BUILD_FAST, "production topology + exact double boxes"), (api.BUILD_REFERENCE_TREE, "reference-identical double tree")):
    acc = api.BVHAccelF64()
    t0 = time.time(); acc.Build(len(f), v64, f, flags=flags); t1 = time.time()
    st = acc.GetStatistics()
    acc.Traverse(r[:100000])
    t2 = time.time(); h, m = acc.Traverse(r); t3 = time.time()
    print(f"fp64 {name}: build wall {1e3*(t1-t0):.1f} ms (device {st['build_secs']*1e3:.2f} ms, depth {st['max_tree_depth']}), "
          f"Traverse {len(r)} primary rays host->host {1e3*(t3-t2):.1f} ms = {len(r)/(t3-t2)/1e6:.1f} Mrays/s, hits {int(m.sum())}")
PY
tail -5 gpurun_out/r02_t13.log; cat gpurun_out/r02_f64_rate13.log
