"""CPU: deterministic scene / ray generators of BASELINE.json's configs."""
import numpy as np


def test_scene_sizes_and_determinism():
    from nanort_b200 import scenes as S

    v, f = S.make_scene("cornell")
    assert f.shape == (34, 3) and v.dtype == np.float32 and f.dtype == np.uint32
    v, f = S.make_scene("sphere_grid")
    assert f.shape == (100002, 3)
    v2, f2 = S.make_scene("sphere_grid")
    assert np.array_equal(v, v2) and np.array_equal(f, f2)
    v, f = S.make_scene("terrain", n=64)
    assert f.shape == (2 * 64 * 64, 3)
    assert f.max() < len(v)


def test_primary_rays_are_normalised_and_jittered():
    from nanort_b200 import scenes as S

    cam = S.scene_camera("sphere_grid", 64, 36)
    r = S.primary_rays(cam, 64, 36, spp=2, seed=1)
    assert len(r) == 64 * 36 * 2
    n = np.linalg.norm(r["dir"].astype(np.float64), axis=1)
    assert np.allclose(n, 1.0, atol=1e-6)
    r2 = S.primary_rays(cam, 64, 36, spp=2, seed=1)
    assert r.tobytes() == r2.tobytes()
    r3 = S.primary_rays(cam, 64, 36, spp=2, seed=2)
    assert r.tobytes() != r3.tobytes()


def test_shard_pixel_partition_is_exact():
    from nanort_b200 import dist

    W, H, tw, th = 200, 100, 64, 8
    seen = np.zeros(W * H, np.int32)
    for shard in range(3):
        pix = dist.shard_pixels(W, H, tw, th, shard, 3)
        seen[pix] += 1
    assert np.all(seen == 1)
