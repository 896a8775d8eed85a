"""Conformance build (NRT_BUILD_REFERENCE_TREE): the device must write exactly the arrays CPU nanort writes --
node boxes, flags, axis labels, child / leaf data, node ORDER (serial pre-order, or the C++11 build's shallow tree +
joined sub-arrays) and the indices_ permutation (std::partition's element order)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _same_tree(got_nodes, got_idx, want_nodes, want_idx):
    assert len(got_nodes) == len(want_nodes), (len(got_nodes), len(want_nodes))
    assert np.array_equal(got_idx, want_idx), "indices_ (std::partition order)"
    assert np.array_equal(got_nodes["flag"], want_nodes["flag"])
    assert np.array_equal(got_nodes["data"], want_nodes["data"])
    br = want_nodes["flag"] == 0
    assert np.array_equal(got_nodes["axis"][br], want_nodes["axis"][br])
    # boxes: exact float min/max (numerically equal; -0.0 vs +0.0 may differ in sign only)
    assert np.array_equal(got_nodes["bmin"], want_nodes["bmin"]) and np.array_equal(got_nodes["bmax"], want_nodes["bmax"])


def _degenerate(kind):
    if kind == "identical":
        v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
        return v, np.tile(np.array([[0, 1, 2]], np.uint32), (3000, 1))
    k = 700
    x = np.arange(k, dtype=np.float32)
    v = np.stack([np.stack([x, 0 * x, 0 * x], 1), np.stack([x + 0.5, 0 * x, 0 * x + 1], 1),
                  np.stack([x, 0 * x + 1, 0 * x], 1)], 1).reshape(-1, 3)
    return v.astype(np.float32), np.arange(3 * k, dtype=np.uint32).reshape(k, 3)


CASES = [
    ("cornell", {}, {}),
    ("sphere_grid", dict(nx=2, nz=2), {}),
    ("sphere_grid", dict(nx=3, nz=3), {}),                       # 9002 prims: joined order in C++11 mode
    ("sphere_grid", dict(nx=3, nz=3), dict(min_leaf_primitives=1)),
    ("sphere_grid", dict(nx=3, nz=3), dict(bin_size=8, min_leaf_primitives=8)),
    ("sphere_grid", dict(nx=3, nz=3), dict(max_tree_depth=6)),
    ("sphere_grid", dict(nx=3, nz=3), dict(shallow_depth=2)),
    ("sphere_grid", dict(nx=3, nz=3), dict(min_primitives_for_parallel_build=100000)),
    ("terrain", dict(n=64), {}),                                   # 8192 prims: not above the threshold
    ("terrain", dict(n=80), {}),                                   # 12800 prims
    ("deg:identical", {}, {}),
    ("deg:line", {}, dict(min_leaf_primitives=2)),
]


@pytest.mark.parametrize("name,kw,okw", CASES)
@pytest.mark.parametrize("cpp11", [True, False])
def test_reference_exact_build(port, name, kw, okw, cpp11):
    from oracle import orc
    from nanort_b200 import api, scenes as S

    v, f = _degenerate(name[4:]) if name.startswith("deg:") else S.make_scene(name, **kw)
    want_nodes, want_idx, want_stats = port.build(v, f, orc.build_options(**okw), mode=orc.MODE_CPP11 if cpp11 else 0)
    acc = api.BVHAccel()
    flags = api.BUILD_REFERENCE_TREE | (0 if cpp11 else api.BUILD_REFERENCE_CPP03_ORDER)
    assert acc.Build(len(f), v, f, api.BVHBuildOptions(**okw), flags=flags)
    _same_tree(acc.GetNodes(), acc.GetIndices(), want_nodes, want_idx)
    st = acc.GetStatistics()
    assert {k: st[k] for k in want_stats} == want_stats


def test_reference_exact_build_on_the_bench_scene_and_traversal(port):
    """100,002 triangles: the device reproduces the reference's 146,295-node tree of depth 66, and the conformance
    walk over it equals the oracle bit for bit (ties included) -- Build and Traverse both conformant, no CPU build."""
    from oracle import orc
    from nanort_b200 import api, scenes as S

    v, f = S.make_scene("sphere_grid")
    want_nodes, want_idx, want_stats = port.build(v, f, mode=orc.MODE_CPP11)
    acc = api.BVHAccel()
    assert acc.Build(len(f), v, f, flags=api.BUILD_REFERENCE_TREE)
    _same_tree(acc.GetNodes(), acc.GetIndices(), want_nodes, want_idx)
    cam = S.scene_camera("sphere_grid", 320, 180)
    rays = np.concatenate([S.primary_rays(cam, 320, 180, spp=1, seed=2),
                           S.incoherent_rays(v.min(axis=0), v.max(axis=0), 40000, seed=4)])
    want_h, want_m = port.traverse(want_nodes, want_idx, v, f, rays, threads=16)
    h, m = acc.Traverse(rays, flags=api.TRAVERSE_CONFORMANCE)
    hit = want_m.astype(bool)
    assert np.array_equal(m, want_m) and np.array_equal(h[hit].view(np.uint32), want_h[hit].view(np.uint32))
