"""Parity at BASELINE.json's full sizes (configs[1]: 1920x1080x16 spp primary + AO) through size-independent
properties: (1) the fast kernel and the conformance walk agree on every one of the ~51 M rays, (2) a strided
sample is checked against the oracle, (3) the accumulated frame equals what the hit records imply."""
import numpy as np
import pytest

from helpers import assert_parity, compare_hits

pytestmark = pytest.mark.gpu

W, H, SPP = 1920, 1080, 16


def test_full_size_pass_fast_equals_conformance_and_oracle_sample(port):
    import torch
    from oracle import orc
    from nanort_b200 import api, scenes as S

    v, f = S.make_scene("sphere_grid")
    acc = api.BVHAccel()
    acc.Build(len(f), v, f)
    cam = S.scene_camera("sphere_grid", W, H)
    bmin, bmax = acc.BoundingBox()
    p = api.AoParams()
    for i in range(12):
        p.cam[i] = float(cam[i])
    p.width, p.height, p.spp, p.sample0, p.seed = W, H, SPP, 0, 1
    p.tile_w, p.tile_h, p.shard, p.n_shards = 64, 8, 0, 1
    p.ray_min_t, p.ray_max_t, p.ao_min_t, p.ao_max_t = 1e-3, 1e30, 1e-3, 0.25 * float(np.linalg.norm(bmax - bmin))
    n = W * H * SPP
    accum = torch.zeros(W * H, dtype=torch.float32, device="cuda")
    d_p = torch.empty(n * 36, dtype=torch.uint8, device="cuda")
    d_a = torch.empty(n * 36, dtype=torch.uint8, device="cuda")
    n_p, n_a = acc.ExportAOWorkload(p, accum.data_ptr(), d_p.data_ptr(), d_a.data_ptr())
    assert n_p == n and 0 < n_a < n
    frame_sum = float(accum.double().sum().item())

    rn, ri, _ = port.build(v, f, mode=orc.MODE_CPP11)  # the reference's own (x-only binned) tree
    occluded = 0
    for d_r, cnt in ((d_p, n_p), (d_a, n_a)):
        h_fast = torch.empty(cnt * 16, dtype=torch.uint8, device="cuda")
        h_conf = torch.empty(cnt * 16, dtype=torch.uint8, device="cuda")
        acc.TraverseDevice(d_r.data_ptr(), cnt, h_fast.data_ptr(), flags=api.TRAVERSE_FAST)
        acc.TraverseDevice(d_r.data_ptr(), cnt, h_conf.data_ptr(), flags=api.TRAVERSE_CONFORMANCE)
        torch.cuda.synchronize()
        a, b = h_fast.view(torch.int32).view(-1, 4), h_conf.view(torch.int32).view(-1, 4)
        diff = (a != b).any(dim=1)
        n_diff = int(diff.sum().item())
        # generic jittered rays: no ties expected; tolerate a handful and classify them below
        assert n_diff <= 16, n_diff
        hits_np = h_fast.cpu().numpy().view(S.HIT_DTYPE)
        mask_np = (hits_np["prim_id"] != 0xFFFFFFFF).astype(np.uint8)
        if d_r is d_a:
            occluded = int(mask_np.sum())
        else:
            assert int(mask_np.sum()) == n_a, "one AO ray per primary hit"
        # oracle on a strided sample plus every ray on which the two GPU kernels disagree
        idx = np.unique(np.concatenate([np.arange(0, cnt, 257), np.nonzero(diff.cpu().numpy())[0]]))
        rays_np = d_r[: cnt * 36].cpu().numpy().view(S.RAY_DTYPE)[idx]
        want_h, want_m = port.traverse(rn, ri, v, f, rays_np, threads=32)
        res = compare_hits(port, v, f, rays_np, hits_np[idx], mask_np[idx], want_h, want_m)
        assert_parity(res, max_near_ties=8)
    # frame = primary misses + unoccluded AO rays (each contributes exactly 1.0)
    assert frame_sum == float((n_p - n_a) + (n_a - occluded))


@pytest.mark.parametrize("scene,width,height", [("terrain", 1920, 1080), ("instanced", 3840, 2160)])
def test_large_scenes_fast_equals_conformance_and_oracle_sample(port, scene, width, height):
    """configs[2] (1,002,528 triangles) and configs[3] (10,000,200 triangles) at their image sizes, one sample per
    pixel: the production build's tree walked by the fast kernel and by the conformance kernel must agree ray for
    ray, and a strided sample must agree with the oracle walking the same GPU-built arrays (hits do not depend on
    topology; a CPU build of the 10 M-triangle scene would take longer than the whole suite)."""
    import torch
    from nanort_b200 import api, scenes as S

    v, f = S.make_scene(scene)
    acc = api.BVHAccel()
    acc.Build(len(f), v, f)
    st = acc.GetStatistics()
    assert st["num_leaf_nodes"] == st["num_branch_nodes"] + 1 and st["max_tree_depth"] < 64
    cam = S.scene_camera(scene, width, height)
    rays = S.primary_rays(cam, width, height, spp=1, seed=5)
    n = len(rays)
    d_r = torch.from_numpy(rays.view(np.uint8).reshape(-1, 36)).cuda()
    h_fast = torch.empty(n * 16, dtype=torch.uint8, device="cuda")
    h_conf = torch.empty(n * 16, dtype=torch.uint8, device="cuda")
    acc.TraverseDevice(d_r.data_ptr(), n, h_fast.data_ptr(), flags=api.TRAVERSE_FAST)
    acc.TraverseDevice(d_r.data_ptr(), n, h_conf.data_ptr(), flags=api.TRAVERSE_CONFORMANCE)
    torch.cuda.synchronize()
    a, b = h_fast.view(torch.int32).view(-1, 4), h_conf.view(torch.int32).view(-1, 4)
    diff = (a != b).any(dim=1)
    assert int(diff.sum().item()) <= 64  # shared edges of the tessellation: exact ties only, classified below
    hits_np = h_fast.cpu().numpy().view(S.HIT_DTYPE)
    mask_np = (hits_np["prim_id"] != 0xFFFFFFFF).astype(np.uint8)
    assert mask_np.mean() > 0.3
    idx = np.unique(np.concatenate([np.arange(0, n, 97), np.nonzero(diff.cpu().numpy())[0]]))
    want_h, want_m = port.traverse(acc.GetNodes(), acc.GetIndices(), v, f, rays[idx], threads=32)
    res = compare_hits(port, v, f, rays[idx], hits_np[idx], mask_np[idx], want_h, want_m)
    assert_parity(res, max_near_ties=8)
