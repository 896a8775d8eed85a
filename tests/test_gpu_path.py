"""Wavefront path tracer (the reference path_tracer's loop, diffuse + emissive) against a numpy restatement whose
Traverse calls go through the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_path_tracer_matches_oracle_driven_restatement(port):
    import torch
    from oracle import orc
    from nanort_b200 import api, scenes as S
    import pt_reference

    v, f = S.make_scene("cornell")
    v, f, l0, ln = S.with_area_light(v, f, (0.0, 9.99, 0.0), 1.5, 1.5)
    W, H, spp, bounces = 64, 48, 4, 6
    cam = S.scene_camera("cornell", W, H)
    acc = api.BVHAccel()
    acc.Build(len(f), v, f)
    p = api.PathParams()
    for i in range(12):
        p.cam[i] = float(cam[i])
    p.width, p.height, p.spp, p.sample0, p.seed = W, H, spp, 0, 5
    p.tile_w, p.tile_h, p.shard, p.n_shards = 64, 8, 0, 1
    p.max_bounces = bounces
    p.ray_min_t, p.ray_max_t = 1e-3, 1e30
    for k, (a, e) in enumerate(zip((0.7, 0.6, 0.5), (12.0, 11.0, 9.0))):
        p.albedo[k], p.emission[k] = a, e
    p.light_first_face, p.light_n_faces, p.flags = l0, ln, 0
    accum = torch.zeros(W * H * 3, dtype=torch.float32, device="cuda")
    r = acc.RenderPath(p, accum.data_ptr())
    got = accum.cpu().numpy().reshape(H, W, 3).astype(np.float64)

    nodes, idx, _ = port.build(v, f, mode=orc.MODE_CPP11)
    want, counts = pt_reference.render(port, nodes, idx, v, f, cam, W, H, spp, 5, bounces, (0.7, 0.6, 0.5),
                                       (12.0, 11.0, 9.0), l0, ln)
    assert r.camera_rays == counts["camera"] == W * H * spp
    # identical random numbers; float differences (sincosf, contraction) may flip a handful of paths
    assert abs(r.radiance_rays - counts["radiance"]) <= 0.002 * counts["radiance"], (r.radiance_rays, counts)
    assert abs(r.shadow_rays - counts["shadow"]) <= 0.002 * counts["shadow"], (r.shadow_rays, counts)
    assert got.mean() > 0.05
    assert abs(got.mean() - want.mean()) <= 0.005 * want.mean()
    bad = np.abs(got - want) > 1e-3 * np.maximum(1.0, np.abs(want))
    assert bad.mean() < 0.02, bad.mean()
    assert r.traverse_launches == 2 * bounces and r.launches == 1 + 4 * bounces
