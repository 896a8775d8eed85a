"""Two ranks on two GPUs (NCCL): tile-sharded AO pass + framebuffer all_gather == the single-GPU frame.
Skipped on boxes with fewer than two GPUs (the gather logic itself is covered on CPU by test_dist_gloo.py)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from nanort_b200 import api, dist as nd, scenes as S

    W, H, spp = 200, 104, 2
    v, f = S.make_scene("sphere_grid", nx=4, nz=4)
    acc = api.BVHAccel(device=rank)
    acc.Build(len(f), v, f)  # every rank rebuilds the identical tree
    cam = S.scene_camera("sphere_grid", W, H)
    bmin, bmax = acc.BoundingBox()
    p = api.AoParams()
    for i in range(12):
        p.cam[i] = float(cam[i])
    p.width, p.height, p.spp, p.sample0, p.seed = W, H, spp, 0, 1
    p.tile_w, p.tile_h, p.shard, p.n_shards = 64, 8, rank, world
    p.ray_min_t, p.ray_max_t, p.ao_min_t, p.ao_max_t = 1e-3, 1e30, 1e-3, 0.25 * float(np.linalg.norm(bmax - bmin))
    accum = torch.zeros(W * H, dtype=torch.float32, device=dev)
    r = acc.RenderAO(p, accum.data_ptr())
    g = nd.FramebufferGather(W, H, 64, 8, world, rank, dev)
    frame = g.gather(accum)
    rays = torch.tensor([r.primary_rays + r.ao_rays], dtype=torch.int64, device=dev)
    dist.all_reduce(rays)
    nodes = acc.GetNodes()
    np.save(os.path.join(out_dir, f"frame{rank}.npy"), frame.cpu().numpy())
    np.save(os.path.join(out_dir, f"nodes{rank}.npy"), nodes.view(np.uint8))
    np.save(os.path.join(out_dir, f"rays{rank}.npy"), rays.cpu().numpy())
    dist.destroy_process_group()


def test_two_gpu_sharded_pass_equals_single_gpu(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from nanort_b200 import api, scenes as S

    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    W, H, spp = 200, 104, 2
    v, f = S.make_scene("sphere_grid", nx=4, nz=4)
    acc = api.BVHAccel(device=0)
    acc.Build(len(f), v, f)
    cam = S.scene_camera("sphere_grid", W, H)
    bmin, bmax = acc.BoundingBox()
    p = api.AoParams()
    for i in range(12):
        p.cam[i] = float(cam[i])
    p.width, p.height, p.spp, p.sample0, p.seed = W, H, spp, 0, 1
    p.tile_w, p.tile_h, p.shard, p.n_shards = 64, 8, 0, 1
    p.ray_min_t, p.ray_max_t, p.ao_min_t, p.ao_max_t = 1e-3, 1e30, 1e-3, 0.25 * float(np.linalg.norm(bmax - bmin))
    accum = torch.zeros(W * H, dtype=torch.float32, device="cuda:0")
    r = acc.RenderAO(p, accum.data_ptr())
    want = accum.cpu().numpy()
    for rank in range(2):
        assert np.array_equal(np.load(tmp_path / f"frame{rank}.npy"), want)
        assert int(np.load(tmp_path / f"rays{rank}.npy")[0]) == r.primary_rays + r.ao_rays
    assert np.array_equal(np.load(tmp_path / "nodes0.npy"), np.load(tmp_path / "nodes1.npy")), "replicated BVH"
