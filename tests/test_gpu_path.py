"""Wavefront path tracer (the reference path_tracer's bounce loop with its tinyobj material model) against a
numpy restatement whose Traverse calls go through the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run_gpu(api, S, torch, v, f, mats, ids, emissive, cam, W, H, spp, bounces, seed):
    acc = api.BVHAccel()
    acc.Build(len(f), v, f)
    d_m = torch.as_tensor(mats.view(np.float32).reshape(-1), device="cuda")
    d_i = torch.as_tensor(ids.astype(np.int32), device="cuda")
    d_e = torch.as_tensor(emissive.astype(np.int32), device="cuda")
    p = api.PathParams()
    for i in range(12):
        p.cam[i] = float(cam[i])
    p.width, p.height, p.spp, p.sample0, p.seed = W, H, spp, 0, seed
    p.tile_w, p.tile_h, p.shard, p.n_shards = 64, 8, 0, 1
    p.max_bounces, p.ray_min_t, p.ray_max_t = bounces, 1e-3, 1e30
    p.n_materials, p.n_emissive = len(mats), len(emissive)
    p.d_materials, p.d_material_ids, p.d_emissive_faces = d_m.data_ptr(), d_i.data_ptr(), d_e.data_ptr()
    p.d_facevarying_normals, p.flags = None, 0
    accum = torch.zeros(W * H * 3, dtype=torch.float32, device="cuda")
    r = acc.RenderPath(p, accum.data_ptr())
    return accum.cpu().numpy().reshape(H, W, 3).astype(np.float64), r


def test_path_tracer_matches_oracle_driven_restatement(port):
    import torch
    from oracle import orc
    from nanort_b200 import api, scenes as S
    import pt_reference

    v, f, mats, ids, emissive = S.cornell_with_materials()
    W, H, spp, bounces, seed = 64, 48, 6, 7, 5
    cam = S.scene_camera("cornell", W, H)
    got, r = _run_gpu(api, S, torch, v, f, mats, ids, emissive, cam, W, H, spp, bounces, seed)
    nodes, idx, _ = port.build(v, f, mode=orc.MODE_CPP11)
    want, counts = pt_reference.render(port, nodes, idx, v, f, cam, W, H, spp, seed, bounces, mats, ids, emissive)
    assert r.camera_rays == counts["camera"] == W * H * spp
    # identical random numbers; float differences (sincosf, pow chains) may flip a handful of paths
    assert abs(r.radiance_rays - counts["radiance"]) <= 0.01 * counts["radiance"], (r.radiance_rays, counts)
    assert abs(r.shadow_rays - counts["shadow"]) <= 0.01 * counts["shadow"], (r.shadow_rays, counts)
    assert got.mean() > 0.05 and np.isfinite(got).all()
    assert abs(got.mean() - want.mean()) <= 0.01 * want.mean(), (got.mean(), want.mean())
    bad = np.abs(got - want) > 2e-3 * np.maximum(1.0, np.abs(want))
    assert bad.mean() < 0.03, bad.mean()
    assert r.traverse_launches == 2 * bounces and r.launches == 1 + 4 * bounces


def test_path_tracer_diffuse_only_energy_is_bounded():
    """White furnace-ish sanity: a closed diffuse box with albedo a and an emitter can never return more than
    Le * cos per camera ray; no NaNs; more bounces never darken the image."""
    import torch
    from nanort_b200 import api, scenes as S

    v, f = S.make_scene("cornell")
    v, f, l0, ln = S.with_area_light(v, f, (0.0, 9.99, 0.0), 2.0, 2.0)
    mats = np.concatenate([S.material(diffuse=(0.7, 0.7, 0.7)), S.material(emission=(10, 10, 10))])
    ids = np.zeros(len(f), np.uint32)
    ids[l0:] = 1
    emissive = np.arange(l0, l0 + ln, dtype=np.uint32)
    cam = S.scene_camera("cornell", 64, 48)
    imgs = []
    for bounces in (1, 2, 6):
        got, r = _run_gpu(api, S, torch, v, f, mats, ids, emissive, cam, 64, 48, 8, bounces, 9)
        assert np.isfinite(got).all() and got.min() >= 0
        imgs.append(got.mean())
    assert imgs[0] <= imgs[1] + 1e-6 <= imgs[2] + 2e-6
