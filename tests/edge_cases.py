"""Hostile inputs shared by the CPU (oracle vs reference) and GPU (CUDA vs oracle) edge-case tests."""
import numpy as np

from nanort_b200 import scenes as S


def hostile_rays(bmin, bmax, n=6000, seed=21):
    """Random interior rays with every 8th ray replaced by a special: zero / NaN / inf / denormal / huge
    components, inverted and degenerate [min_t, max_t] ranges."""
    rays = S.incoherent_rays(bmin, bmax, n, seed=seed)
    nan, inf = np.float32(np.nan), np.float32(np.inf)
    k = 0
    for i in range(0, n, 8):
        r = rays[i]
        c = k % 14
        if c == 0:
            r["dir"] = (0, 0, 0)
        elif c == 1:
            r["dir"][k % 3] = nan
        elif c == 2:
            r["org"][k % 3] = nan
        elif c == 3:
            r["org"][k % 3] = inf
        elif c == 4:
            r["dir"][k % 3] = -inf
        elif c == 5:
            r["min_t"], r["max_t"] = 5.0, 1.0
        elif c == 6:
            r["min_t"], r["max_t"] = -10.0, inf
        elif c == 7:
            r["max_t"] = 0.0
        elif c == 8:
            r["dir"][k % 3] = np.float32(1e-40)  # denormal
        elif c == 9:
            r["org"] = (1e30, -1e30, 1e30)
        elif c == 10:
            r["dir"] = np.float32(1e20) * r["dir"]  # far from unit length
        elif c == 11:
            r["min_t"] = nan
        elif c == 12:
            r["max_t"] = nan
        elif c == 13:
            r["dir"] = (-0.0, -0.0, -1.0)
        k += 1
    return rays


def degenerate_mesh():
    """Cornell box plus zero-area triangles (repeated vertex, collinear vertices), a sliver, a huge and a tiny
    triangle."""
    v, f = S.cornell()
    base = len(v)
    extra_v = np.array([[1, 1, 1], [1, 1, 1], [2, 3, 1],            # repeated vertex
                        [0, 5, 0], [1, 5, 0], [2, 5, 0],            # collinear
                        [-3, 2, -3], [3, 2.000001, 3], [0, 2, 0],   # sliver
                        [-1e6, 4, -1e6], [1e6, 4, -1e6], [0, 4, 1e6],  # huge
                        [0.5, 0.5, 0.5], [0.5 + 1e-6, 0.5, 0.5], [0.5, 0.5 + 1e-6, 0.5]], np.float32)  # tiny
    extra_f = (np.arange(15, dtype=np.uint32) + base).reshape(5, 3)
    return np.concatenate([v, extra_v]).astype(np.float32), np.concatenate([f, extra_f]).astype(np.uint32)
