"""GPU parity of the traversal kernels against the oracle (through the C-ABI)."""
import numpy as np
import pytest

from helpers import assert_parity, compare_hits

pytestmark = pytest.mark.gpu


def _rays(S, name, verts, faces, w=192, h=160, n_inc=60000):
    cam = S.scene_camera(name, w, h)
    prim = S.primary_rays(cam, w, h, spp=1, seed=11)
    bmin, bmax = verts.min(axis=0), verts.max(axis=0)
    inc = S.incoherent_rays(bmin, bmax, n_inc, seed=5)
    return np.concatenate([prim, inc])


@pytest.mark.parametrize("scene,kw", [("cornell", {}), ("sphere_grid", dict(nx=3, nz=3)), ("terrain", dict(n=64))])
@pytest.mark.parametrize("cpp11", [True, False])
def test_adopted_reference_tree_conformance_is_bit_exact(port, scene, kw, cpp11):
    """GPU conformance walk over the CPU-built tree == oracle, bit for bit, ties included."""
    from oracle import orc
    from nanort_b200 import api, scenes as S

    v, f = S.make_scene(scene, **kw)
    nodes, idx, _ = port.build(v, f, mode=orc.MODE_CPP11 if cpp11 else 0)
    rays = _rays(S, scene, v, f)
    want_h, want_m = port.traverse(nodes, idx, v, f, rays, cpp11=cpp11, threads=8)
    acc = api.BVHAccel()
    acc.Adopt(nodes, idx, v, f)
    flags = api.TRAVERSE_CONFORMANCE | (0 if cpp11 else api.TRAVERSE_CPP03_INVERSE)
    got_h, got_m = acc.Traverse(rays, flags=flags)
    assert np.array_equal(got_m, want_m)
    hit = want_m.astype(bool)
    assert np.array_equal(got_h[hit].view(np.uint32), want_h[hit].view(np.uint32))
    # miss records are {0, 0, max_t, 0xFFFFFFFF}
    assert np.all(got_h[~hit]["prim_id"] == 0xFFFFFFFF)
    assert np.array_equal(got_h[~hit]["t"], rays[~hit]["max_t"])


@pytest.mark.parametrize("scene,kw", [("cornell", {}), ("sphere_grid", dict(nx=3, nz=3)), ("terrain", dict(n=64))])
@pytest.mark.parametrize("fix", [False, True])
def test_adopted_tree_fast_path_matches_oracle(port, scene, kw, fix):
    """Fast kernel (child-pair layout, distance order) over a CPU-built tree: same hits, bit-equal t/u/v,
    prim_id equal except classified exact-t ties."""
    from oracle import orc
    from nanort_b200 import api, scenes as S

    v, f = S.make_scene(scene, **kw)
    nodes, idx, _ = port.build(v, f, mode=orc.MODE_CPP11 | (orc.MODE_FIXBINS if fix else 0))
    rays = _rays(S, scene, v, f)
    want_h, want_m = port.traverse(nodes, idx, v, f, rays, threads=8)
    acc = api.BVHAccel()
    acc.Adopt(nodes, idx, v, f)
    got_h, got_m = acc.Traverse(rays, flags=api.TRAVERSE_FAST)
    res = compare_hits(port, v, f, rays, got_h, got_m, want_h, want_m)
    assert_parity(res)


@pytest.mark.parametrize("conformance", [False, True])
def test_compact_ray32_records_give_the_same_hits(conformance):
    """NRT_TRAVERSE_RAY32: the 32-byte ray record {org, dir, min_t, max_t} (nanort::Ray without `type`, nanort.h:474-496)
    and hit_mask == NULL give bit-identical hit records (misses: prim_id == 0xFFFFFFFF, t == max_t)."""
    import torch
    from nanort_b200 import api, scenes as S

    v, f = S.make_scene("sphere_grid", nx=3, nz=3)
    rays = _rays(S, "sphere_grid", v, f)
    rays[::7]["max_t"] = 3.0  # some rays end before their hit
    acc = api.BVHAccel()
    acc.Build(len(f), v, f)
    flags = api.TRAVERSE_CONFORMANCE if conformance else api.TRAVERSE_FAST
    want_h, want_m = acc.Traverse(rays, flags=flags)
    r32 = np.ascontiguousarray(rays.view(np.uint8).reshape(-1, 36)[:, :32]).view(np.float32).reshape(-1, 8)
    # odd ray counts and more than one pipeline chunk's worth of bytes are covered by the sizes below
    for n in (len(r32), 1, 33):
        got_h, got_m = acc.Traverse(r32[:n].view(np.dtype((np.void, 32))).reshape(-1), flags=flags | api.TRAVERSE_RAY32, mask=False)
        assert got_m is None
        assert np.array_equal(got_h.view(np.uint32), want_h[:n].view(np.uint32))
        assert np.array_equal(got_h["prim_id"] != 0xFFFFFFFF, want_m[:n].astype(bool))
    # device-pointer form
    d_r = torch.as_tensor(r32, device="cuda")
    d_h = torch.zeros(len(r32) * 4, dtype=torch.int32, device="cuda")
    acc.TraverseDevice(d_r.data_ptr(), len(r32), d_h.data_ptr(), flags=flags | api.TRAVERSE_RAY32)
    assert np.array_equal(d_h.cpu().numpy().view(np.uint32).reshape(-1, 4), want_h.view(np.uint32).reshape(-1, 4))
    # misaligned rays are refused, not mis-read
    with pytest.raises(Exception):
        acc.TraverseDevice(d_r.data_ptr() + 4, 8, d_h.data_ptr(), flags=flags | api.TRAVERSE_RAY32)


def test_small_calls_take_the_zero_copy_path_and_agree_with_the_batch():
    """nrt_traverse with <= 64 rays (the facade's per-ray Traverse) reads the rays from / writes the records to a pinned
    host slot directly and is not serialised with other host threads: same records as the batched call, for every small n,
    both kernels, with and without hit flags, from several threads at once."""
    import threading
    from nanort_b200 import api, scenes as S

    v, f = S.make_scene("sphere_grid", nx=3, nz=3)
    rays = _rays(S, "sphere_grid", v, f)[:4096]
    acc = api.BVHAccel()
    acc.Build(len(f), v, f)
    for flags in (api.TRAVERSE_FAST, api.TRAVERSE_CONFORMANCE):
        want_h, want_m = acc.Traverse(rays, flags=flags)
        for n in (1, 2, 31, 32, 33, 63, 64, 65):
            for off in (0, 777):
                h, m = acc.Traverse(rays[off:off + n], flags=flags)
                assert np.array_equal(m, want_m[off:off + n])
                assert np.array_equal(h.view(np.uint32), want_h[off:off + n].view(np.uint32))
        h, m = acc.Traverse(rays[5:6], flags=flags, mask=False)
        assert m is None and np.array_equal(h.view(np.uint32), want_h[5:6].view(np.uint32))
    want_h, want_m = acc.Traverse(rays)
    errors = []

    def worker(k):
        try:
            for i in range(k, len(rays), 8):
                h, m = acc.Traverse(rays[i:i + 1])
                if m[0] != want_m[i] or h.view(np.uint32).tolist() != want_h[i:i + 1].view(np.uint32).tolist():
                    errors.append(i)
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]
