"""The CUDA path against the committed golden vectors of the unmodified reference (tests/golden/), through the
C-ABI: hand cases for every Intersect branch, the regression-30 program, whole scenes; trace options."""
import os

import numpy as np
import pytest

from helpers import assert_parity, compare_hits

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("tag,cpp11", [("11", True), ("03", False)])
@pytest.mark.parametrize("fast", [False, True])
def test_kat_intersect_branches_on_gpu(tag, cpp11, fast):
    from nanort_b200 import api

    d = np.load(os.path.join(G, "kat_intersect.npz"))
    v, f, rays, topts = d["verts"], d["faces"], d["rays"], d["topts"]
    acc = api.BVHAccel()
    acc.Build(len(f), v, f)
    flags = (api.TRAVERSE_FAST if fast else api.TRAVERSE_CONFORMANCE) | (0 if cpp11 else api.TRAVERSE_CPP03_INVERSE)
    for i, name in enumerate(d["names"]):
        h, m = acc.Traverse(rays[i:i + 1], options=topts[i:i + 1], flags=flags)
        assert m[0] == d[f"mask_cpp{tag}"][i], name
        if m[0]:
            assert h[0].tobytes() == d[f"hits_cpp{tag}"][i].tobytes(), name
        else:
            assert h[0]["prim_id"] == 0xFFFFFFFF and h[0]["t"] == rays[i]["max_t"], name


def test_regression30_on_gpu():
    from nanort_b200 import api

    d = np.load(os.path.join(G, "regression30.npz"))
    v, f = d["verts"].astype(np.float32), d["faces"]
    acc = api.BVHAccel()
    acc.Build(len(f), v, f)
    for k in ("plain", "bug"):
        for flags in (api.TRAVERSE_FAST, api.TRAVERSE_CONFORMANCE):
            h, m = acc.Traverse(d[f"f32_{k}_ray"], flags=flags)
            assert m[0] == 1 and h[0].tobytes() == d[f"f32_{k}_hit"][0].tobytes()
            assert abs(float(h[0]["u"]) - 0.68) < 1e-6 and abs(float(h[0]["v"]) - 0.131201) < 1e-6


@pytest.mark.parametrize("scene", ["cornell", "sphere_grid"])
@pytest.mark.parametrize("tag,cpp11", [("11", True), ("03", False)])
def test_golden_scene_hits(port, scene, tag, cpp11):
    """(1) the reference's own tree from the fixture, walked in conformance mode: bit-identical records;
    (2) the GPU-built tree, fast kernel: same hits up to classified ties."""
    from nanort_b200 import api

    d = np.load(os.path.join(G, f"scene_{scene}.npz"))
    v, f, rays = d["verts"], d["faces"], d[f"rays_cpp{tag}"]
    gh, gm = d[f"hits_cpp{tag}"], d[f"mask_cpp{tag}"]
    mode = 0 if cpp11 else api.TRAVERSE_CPP03_INVERSE
    adopted = api.BVHAccel()
    adopted.Adopt(d[f"nodes_cpp{tag}"], d[f"indices_cpp{tag}"], v, f)
    h, m = adopted.Traverse(rays, flags=api.TRAVERSE_CONFORMANCE | mode)
    hit = gm.astype(bool)
    assert np.array_equal(m, gm) and np.array_equal(h[hit].view(np.uint32), gh[hit].view(np.uint32))
    built = api.BVHAccel()
    built.Build(len(f), v, f)
    h, m = built.Traverse(rays, flags=api.TRAVERSE_FAST | mode)
    assert_parity(compare_hits(port, v, f, rays, h, m, gh, gm, cpp11=cpp11))


@pytest.mark.parametrize("tkw", [dict(cull_back_face=1), dict(skip_prim_id=17), dict(prim_ids_range=(1000, 3000)),
                                 dict(cull_back_face=1, prim_ids_range=(0, 2500), skip_prim_id=2001)])
def test_trace_options_match_oracle(port, tkw):
    from oracle import orc
    from nanort_b200 import api, scenes as S

    v, f = S.make_scene("sphere_grid", nx=2, nz=2)
    rays = S.incoherent_rays(v.min(axis=0), v.max(axis=0), 60000, seed=8)
    rn, ri, _ = port.build(v, f, mode=orc.MODE_CPP11)
    t = orc.trace_options(**tkw)
    want_h, want_m = port.traverse(rn, ri, v, f, rays, topts=t, threads=8)
    acc = api.BVHAccel()
    acc.Build(len(f), v, f)
    for flags in (api.TRAVERSE_FAST, api.TRAVERSE_CONFORMANCE):
        h, m = acc.Traverse(rays, options=api.BVHTraceOptions(**tkw), flags=flags)
        assert_parity(compare_hits(port, v, f, rays, h, m, want_h, want_m, topts=t))
    adopted = api.BVHAccel()
    adopted.Adopt(rn, ri, v, f)
    h, m = adopted.Traverse(rays, options=api.BVHTraceOptions(**tkw), flags=api.TRAVERSE_CONFORMANCE)
    hit = want_m.astype(bool)
    assert np.array_equal(m, want_m) and np.array_equal(h[hit].view(np.uint32), want_h[hit].view(np.uint32))


def test_host_call_chunks_pageable_and_pinned_and_empty(port):
    """nrt_traverse with more rays than one staging chunk (3 pipeline slots), from pageable and pinned memory,
    with and without the mask, and with zero rays."""
    from oracle import orc
    from nanort_b200 import api, scenes as S

    v, f = S.make_scene("sphere_grid", nx=2, nz=2)
    cam = S.scene_camera("sphere_grid", 1600, 1000)
    rays = S.primary_rays(cam, 1600, 1000, spp=2, seed=3)  # 3.2 M rays > 3 chunks of 1 Mi
    acc = api.BVHAccel()
    acc.Build(len(f), v, f)
    h1, m1 = acc.Traverse(rays)
    pin_r, pin_h, pin_m = api.PinnedArray(len(rays), S.RAY_DTYPE), api.PinnedArray(len(rays), S.HIT_DTYPE), \
        api.PinnedArray(len(rays), np.uint8)
    pin_r.array[:] = rays
    acc.Traverse(pin_r.array, hits=pin_h.array, mask=pin_m.array)
    assert np.array_equal(m1, pin_m.array) and h1.tobytes() == pin_h.array.tobytes()
    # against the oracle on a sample
    rn, ri, _ = port.build(v, f, mode=orc.MODE_CPP11)
    idx = np.arange(0, len(rays), 37)
    want_h, want_m = port.traverse(rn, ri, v, f, rays[idx], threads=8)
    assert_parity(compare_hits(port, v, f, rays[idx], h1[idx], m1[idx], want_h, want_m))
    # mask derivable from the records
    assert np.array_equal(m1.astype(bool), h1["prim_id"] != 0xFFFFFFFF)
    h0, m0 = acc.Traverse(rays[:0])
    assert len(h0) == 0 and len(m0) == 0
