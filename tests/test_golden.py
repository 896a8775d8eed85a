"""CPU: the oracle port against the committed golden vectors produced by the unmodified reference
(tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _same_hits(h, m, gh, gm):
    assert np.array_equal(m, gm)
    hit = gm.astype(bool)
    assert np.array_equal(h[hit].view(np.uint32), gh[hit].view(np.uint32))


@pytest.mark.parametrize("tag,cpp11", [("11", True), ("03", False)])
def test_kat_intersect_branches(port, tag, cpp11):
    from oracle import orc

    d = np.load(os.path.join(G, "kat_intersect.npz"))
    v, f, rays, topts = d["verts"], d["faces"], d["rays"], d["topts"]
    nodes, idx, _ = port.build(v, f, mode=orc.MODE_CPP11 if cpp11 else 0)
    for i, name in enumerate(d["names"]):
        h, m = port.traverse(nodes, idx, v, f, rays[i:i + 1], topts=topts[i:i + 1], cpp11=cpp11)
        assert m[0] == d[f"mask_cpp{tag}"][i], name
        if m[0]:
            assert h[0].tobytes() == d[f"hits_cpp{tag}"][i].tobytes(), name


def test_regression30_known_answer(port):
    """test/regression/possible-accuracy-problem-30: must hit with u=0.68, v=0.131201 (fp64 original);
    the fp32 restatement run through the port must equal the reference's fp32 answer bit for bit."""
    d = np.load(os.path.join(G, "regression30.npz"))
    for k in ("plain", "bug"):
        r64 = d[f"f64_{k}"]
        assert r64[0] == 1 and abs(r64[2] - 0.68) < 1e-9 and abs(r64[3] - 0.131201) < 1e-6
        v, f = d["verts"].astype(np.float32), d["faces"]
        nodes, idx, _ = port.build(v, f)
        h, m = port.traverse(nodes, idx, v, f, d[f"f32_{k}_ray"])
        _same_hits(h, m, d[f"f32_{k}_hit"], d[f"f32_{k}_mask"])


@pytest.mark.parametrize("scene", ["cornell", "sphere_grid"])
@pytest.mark.parametrize("tag,cpp11", [("11", True), ("03", False)])
def test_scene_tree_and_hits_match_reference(port, scene, tag, cpp11):
    from oracle import orc

    d = np.load(os.path.join(G, f"scene_{scene}.npz"))
    v, f = d["verts"], d["faces"]
    nodes, idx, st = port.build(v, f, mode=orc.MODE_CPP11 if cpp11 else 0)
    gn = d[f"nodes_cpp{tag}"]
    nodes = nodes.copy()
    nodes["axis"][nodes["flag"] != 0] = 0
    assert nodes.tobytes() == gn.tobytes(), "node array bit-identical to the reference's"
    assert np.array_equal(idx, d[f"indices_cpp{tag}"])
    assert list(st.values()) == list(d[f"stats_cpp{tag}"])
    h, m = port.traverse(nodes, idx, v, f, d[f"rays_cpp{tag}"], cpp11=cpp11, threads=4)
    _same_hits(h, m, d[f"hits_cpp{tag}"], d[f"mask_cpp{tag}"])
