"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built from
/root/reference/nanort.h by oracle/Makefile).  Run in the authoring container only:

    python tests/golden/make_golden.py

The fixtures pin the oracle port (tests/test_golden.py, CPU) and the CUDA path (tests/test_gpu_golden.py).
Reference build: g++ -O2 -std=c++11 -ffp-contract=off [-DNANORT_USE_CPP11_FEATURE].
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402
from nanort_b200 import scenes as S  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def ray(org, dir, min_t=0.0, max_t=3.4028234663852886e38):
    r = np.zeros(1, S.RAY_DTYPE)
    r["org"], r["dir"], r["min_t"], r["max_t"] = org, dir, min_t, max_t
    return r


def kat_cases():
    """Hand cases for every branch of TriangleIntersector::Intersect (nanort.h:1054-1150) and the
    Traverse epilogue (nanort.h:2552-2553)."""
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0],  # tri 0 (ccw seen from +z)
                  [0, 0, -1], [1, 0, -1], [0, 1, -1]], np.float32)  # tri 1 behind it
    f = np.array([[0, 1, 2], [3, 4, 5]], np.uint32)
    cases = []

    def add(name, r, **topt):
        cases.append((name, r, orc.trace_options(**topt)))

    add("front_hit", ray((0.2, 0.2, 1), (0, 0, -1)))
    add("back_hit_no_cull", ray((0.2, 0.2, -2), (0, 0, 1)))
    add("back_hit_cull", ray((0.2, 0.2, 1), (0, 0, -1), ), cull_back_face=1)
    add("back_side_cull", ray((0.2, 0.2, -2), (0, 0, 1)), cull_back_face=1)
    add("skip_first", ray((0.2, 0.2, 1), (0, 0, -1)), skip_prim_id=0)
    add("range_second_only", ray((0.2, 0.2, 1), (0, 0, -1)), prim_ids_range=(1, 2))
    add("range_empty", ray((0.2, 0.2, 1), (0, 0, -1)), prim_ids_range=(2, 2))
    add("t_equals_max_t_is_miss", ray((0.2, 0.2, 1), (0, 0, -1), 0.0, 1.0))
    add("t_just_inside_max_t", ray((0.2, 0.2, 1), (0, 0, -1), 0.0, np.nextafter(np.float32(1.0), np.float32(2.0))))
    add("t_equals_min_t_is_hit", ray((0.2, 0.2, 1), (0, 0, -1), 1.0, 10.0))
    add("t_below_min_t_hits_second", ray((0.2, 0.2, 1), (0, 0, -1), 1.5, 10.0))
    add("edge_exact_fp64_fallback", ray((0.5, 0.0, 1), (0, 0, -1)))
    add("vertex_exact", ray((0.0, 0.0, 1), (0, 0, -1)))
    add("parallel_miss", ray((0.2, 0.2, 1), (1, 0, 0)))
    add("neg_zero_dir", ray((0.2, 0.2, 1), (-0.0, -0.0, -1)))
    add("tiny_dir_component", ray((0.2, 0.2, 1), (1e-9, 0, -1)))
    add("outside", ray((2, 2, 1), (0, 0, -1)))
    add("oblique", ray((0.9, 0.05, 1), (-0.6, 0.1, -1)))
    return v, f, cases


def main():
    refs = {True: orc.Reference(True), False: orc.Reference(False)}
    assert refs[True].sizes() == [40, 36, 16, 28, 16]

    # ---- 1. hand cases, both build modes of the reference
    v, f, cases = kat_cases()
    rays = np.concatenate([c[1] for c in cases])
    topts = np.concatenate([c[2] for c in cases])
    out = {"verts": v, "faces": f, "rays": rays, "topts": topts, "names": np.array([c[0] for c in cases])}
    for cpp11, ref in refs.items():
        acc = ref.build(v, f)
        hits = np.zeros(len(cases), S.HIT_DTYPE)
        mask = np.zeros(len(cases), np.uint8)
        for i in range(len(cases)):
            h, m = acc.traverse(rays[i:i + 1], topts=topts[i:i + 1])
            hits[i], mask[i] = h[0], m[0]
        out[f"hits_cpp{'11' if cpp11 else '03'}"] = hits
        out[f"mask_cpp{'11' if cpp11 else '03'}"] = mask
    np.savez_compressed(os.path.join(OUT, "kat_intersect.npz"), **out)

    # ---- 2. the reference's only regression program (fp64) + its fp32 restatement
    tv = np.array([[1, 2, -3], [-1, 2, -3], [1, 2, 3]], np.float64)
    tf = np.array([[0, 1, 2]], np.uint32)
    org = np.array([-0.36, 7.93890843, 1.2160368])
    res = {}
    for k, dx in (("plain", 0.0), ("bug", -5.30287619e-17)):
        d = np.array([dx, -8.66025404e-01, -0.5])
        d = d / np.sqrt((d * d).sum())
        r, tuv, prim = refs[True].traverse_one_f64(tv, tf, org, d, 0.0, 1.0e30)
        res[f"f64_{k}"] = np.array([r, *tuv, prim])
        r32 = ray(org.astype(np.float32), d.astype(np.float32), 0.0, 1.0e30)
        acc = refs[True].build(tv.astype(np.float32), tf)
        h, m = acc.traverse(r32)
        res[f"f32_{k}_ray"] = r32
        res[f"f32_{k}_hit"] = h
        res[f"f32_{k}_mask"] = m
    np.savez_compressed(os.path.join(OUT, "regression30.npz"), verts=tv, faces=tf, **res)

    # ---- 3. scenes: tree + hits of the reference itself
    for name, kw, w, h in (("cornell", {}, 64, 64), ("sphere_grid", dict(nx=2, nz=2), 64, 48)):
        v, f = S.make_scene(name, **kw)
        cam = S.scene_camera(name, w, h)
        prim = S.primary_rays(cam, w, h, spp=1, seed=21)
        inc = S.incoherent_rays(v.min(axis=0), v.max(axis=0), 4096, seed=22)
        out = {"verts": v, "faces": f}
        for cpp11, ref in refs.items():
            tag = "11" if cpp11 else "03"
            acc = ref.build(v, f)
            ph, pm = acc.traverse(prim)
            ao, src = S.ao_rays(v, f, prim, ph, pm, seed=23, max_t=0.25 * float(np.linalg.norm(v.max(0) - v.min(0))))
            rays = np.concatenate([prim, ao, inc])
            hits, mask = acc.traverse(rays)
            hits[mask == 0] = np.zeros(1, S.HIT_DTYPE)  # untouched-on-miss records carry no information
            nodes = acc.nodes()
            nodes["axis"][nodes["flag"] != 0] = 0  # the reference leaves leaf.axis uninitialised
            out.update({f"rays_cpp{tag}": rays, f"hits_cpp{tag}": hits, f"mask_cpp{tag}": mask,
                        f"nodes_cpp{tag}": nodes, f"indices_cpp{tag}": acc.indices(),
                        f"stats_cpp{tag}": np.array(list(acc.stats().values()), np.uint32)})
        np.savez_compressed(os.path.join(OUT, f"scene_{name}.npz"), **out)
        print(name, {k: (val.shape, val.dtype) for k, val in out.items() if k.endswith("11")})


if __name__ == "__main__":
    main()
