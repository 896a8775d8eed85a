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
