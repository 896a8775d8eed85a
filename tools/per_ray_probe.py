"""Latency of small nrt_traverse calls (the facade's per-ray Traverse) through the ctypes mirror: wall time per call for
n = 1, 8, 64 rays, one thread and 8 threads.  (ctypes adds ~1-2 us per call to what a C++ caller pays.)"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nanort_b200 import api, scenes as S

v, f = S.make_scene("sphere_grid")
acc = api.BVHAccel(); acc.Build(len(f), v, f)
cam = S.scene_camera("sphere_grid", 256, 144)
rays = S.primary_rays(cam, 256, 144, spp=1, seed=3)
hits, mask = np.zeros(64, S.HIT_DTYPE), np.zeros(64, np.uint8)
for flags, name in ((api.TRAVERSE_FAST, "fast"), (api.TRAVERSE_CONFORMANCE, "conformance")):
    for n in (1, 8, 64):
        for _ in range(200):
            acc.Traverse(rays[:n], flags=flags, hits=hits[:n], mask=mask[:n])
        reps = 3000
        t0 = time.perf_counter()
        for i in range(reps):
            acc.Traverse(rays[i:i + n], flags=flags, hits=hits[:n], mask=mask[:n])
        dt = (time.perf_counter() - t0) / reps
        print(f"{name:12s} n={n:3d}: {dt * 1e6:7.2f} us per call = {n / dt / 1e6:7.3f} Mrays/s (1 thread)", flush=True)

def worker(k, reps, n):
    h, m = np.zeros(n, S.HIT_DTYPE), np.zeros(n, np.uint8)
    for i in range(reps):
        acc.Traverse(rays[(k * 1000 + i) % 30000:][:n], hits=h, mask=m)

for nthr in (8, 16):
    for n in (1, 64):
        reps = 2000
        ts = [threading.Thread(target=worker, args=(k, reps, n)) for k in range(nthr)]
        t0 = time.perf_counter()
        for t in ts: t.start()
        for t in ts: t.join()
        dt = time.perf_counter() - t0
        print(f"fast n={n:3d} x {nthr} threads: {nthr * reps * n / dt / 1e6:7.3f} Mrays/s ({dt / reps * 1e6:7.2f} us per call per thread; the GIL is released inside the call)", flush=True)
