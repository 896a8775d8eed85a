"""Edge cases on the GPU against the oracle: non-finite / zero / denormal / huge ray components, inverted ranges,
degenerate triangles, non-default vertex strides.  The conformance kernel must stay bit-identical (NaN payloads
included), the fast kernel must report the same hits."""
import numpy as np
import pytest

from edge_cases import degenerate_mesh, hostile_rays
from helpers import assert_parity, compare_hits

pytestmark = pytest.mark.gpu


def _bits_equal(a, b, mask):
    return a[mask].tobytes() == b[mask].tobytes()


@pytest.mark.parametrize("cpp11", [True, False])
def test_hostile_rays_conformance_bit_exact_and_fast_same_hits(port, cpp11):
    from nanort_b200 import api
    from oracle import orc

    v, f = degenerate_mesh()
    nodes, idx, _ = port.build(v, f, None, orc.MODE_CPP11 if cpp11 else 0)
    rays = hostile_rays(v[:34 * 3].min(axis=0) - 1, v[:34 * 3].max(axis=0) + 1)
    wh, wm = port.traverse(nodes, idx, v, f, rays, cpp11=cpp11)
    acc = api.BVHAccel()
    acc.Adopt(nodes, idx, v, f)
    inv = 0 if cpp11 else api.TRAVERSE_CPP03_INVERSE
    gh, gm = acc.Traverse(rays, flags=api.TRAVERSE_CONFORMANCE | inv)
    assert np.array_equal(gm, wm)
    assert _bits_equal(gh, wh, wm == 1)
    assert wm.sum() > 1000 and (wm[::8] == 1).any()  # some specials hit, too
    # fast kernel over its own tree, against the oracle walking THAT tree: with min_t < 0 the reference finds or
    # misses hits behind the origin depending on the tree (its far-plane widening shrinks boxes at negative t)
    fast = api.BVHAccel()
    fast.Build(len(f), v, f)
    fh, fm = fast.Traverse(rays, flags=inv)
    oh, om = port.traverse(fast.GetNodes(), fast.GetIndices(), v, f, rays, cpp11=cpp11)
    assert np.array_equal(fm, om)
    finite = np.isfinite(oh["t"]) | (om == 0)
    res = compare_hits(port, v, f, rays[finite], fh[finite], fm[finite], oh[finite], om[finite], cpp11=cpp11)
    assert_parity(res)


def test_vertex_stride_and_unused_vertices(port):
    """TriangleMesh takes a byte stride (nanort.h:925-930): 20- and 32-byte vertices must build the same tree and
    return the same hits as the packed array."""
    from nanort_b200 import api, scenes as S

    v, f = S.sphere_grid(nx=2, nz=2)
    rays = S.incoherent_rays(v.min(axis=0), v.max(axis=0), 20000, seed=2)
    ref = api.BVHAccel()
    ref.Build(len(f), v, f)
    rn, ri = ref.GetNodes(), ref.GetIndices()
    rh, rm = ref.Traverse(rays)
    for stride_floats in (5, 8):
        wide = np.full((len(v), stride_floats), np.float32(7.5e8), np.float32)  # junk in the padding
        wide[:, :3] = v
        acc = api.BVHAccel()
        acc.Build(len(f), wide, f, vertex_stride_bytes=4 * stride_floats)
        assert acc.GetNodes().tobytes() == rn.tobytes() and acc.GetIndices().tobytes() == ri.tobytes()
        h, m = acc.Traverse(rays)
        assert np.array_equal(m, rm) and h.tobytes() == rh.tobytes()


def test_hostile_rays_two_level_scene_conformance(port):
    """The exact scene kernel on non-finite and non-unit rays: bit-identical to the oracle's Scene::Traverse."""
    from nanort_b200 import api, scenes as S
    from oracle import orc
    import test_gpu_scene as T

    insts = S.instances_mixed(12)
    scene_port = orc.PortScene(insts, cpp11=True)
    sc = T._gpu_scene(insts, api.BUILD_REFERENCE_TREE, api.BUILD_REFERENCE_TREE)
    rays = hostile_rays(np.float32([-8, -4, -8]), np.float32([8, 4, 8]), n=8000, seed=5)
    rays["min_t"] = np.where(np.isnan(rays["min_t"]), rays["min_t"], 0.0)
    ph, pm = scene_port.traverse(rays, threads=4)
    gh, gm = sc.Traverse(rays, flags=api.TRAVERSE_CONFORMANCE)
    assert np.array_equal(pm, gm) and _bits_equal(gh, ph, pm == 1)
    # the production path hands non-unit rays to the exact kernel and must agree on everything but exact ties
    fh, fm = sc.Traverse(rays)
    assert np.array_equal(fm, pm)
    same = (pm == 1) & (fh["node_id"] == ph["node_id"]) & (fh["prim_id"] == ph["prim_id"])
    assert _bits_equal(fh, ph, same) and same.sum() >= 0.99 * pm.sum()
