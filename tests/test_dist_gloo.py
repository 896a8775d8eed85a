"""CPU, world_size 2, gloo: the tile-sharded framebuffer gather (the only collective of the path)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, tw, th, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nanort_b200 import dist as nd

    g = nd.FramebufferGather(W, H, tw, th, world, rank, "cpu")
    full = torch.arange(W * H, dtype=torch.float32) * 0.5 + 1.0  # what a 1-rank render would hold
    local = torch.zeros(W * H)
    mine = torch.as_tensor(nd.shard_pixels(W, H, tw, th, rank, world))
    local[mine] = full[mine]  # this rank only rendered its own tiles
    got = g.gather(local)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), got.numpy())
    # ray counts of the shards add up to the whole image
    cnt = torch.tensor([nd.shard_ray_count(W, H, tw, th, rank, world, 3)], dtype=torch.int64)
    dist.all_reduce(cnt)
    assert int(cnt[0]) == W * H * 3
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_rank_frame(tmp_path):
    W, H, tw, th = 200, 100, 64, 8  # partial tiles on both edges
    port = _free_port()
    mp.spawn(_worker, args=(2, port, W, H, tw, th, str(tmp_path)), nprocs=2, join=True)
    want = np.arange(W * H, dtype=np.float32) * 0.5 + 1.0
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"r{r}.npy"), want)


def _worker_packed(rank, world, port, W, H, tw, th, out_dir):
    """The C-ABI path's data movement (csrc/comm.cu) on the host: tile-major slot per rank, ONE equal-count all_gather
    (gloo here, ncclAllGather there), unpack to the row-major frame."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nanort_b200 import dist as nd

    full = np.arange(W * H, dtype=np.float32) * 0.25 + 3.0
    mine = torch.as_tensor(nd.pack_own_tiles(full, W, H, tw, th, rank, world))
    slot = nd.packed_slot_floats(W, H, tw, th, world)
    assert len(mine) == slot
    gathered = torch.zeros(slot * world)
    dist.all_gather_into_tensor(gathered, mine)
    np.save(os.path.join(out_dir, f"p{rank}.npy"), nd.unpack_gathered(gathered.numpy(), W, H, tw, th, world))
    dist.destroy_process_group()


def test_two_rank_packed_tile_gather_equals_single_rank_frame(tmp_path):
    W, H, tw, th = 200, 100, 64, 8  # partial tiles on both edges, odd tile count
    port = _free_port()
    mp.spawn(_worker_packed, args=(2, port, W, H, tw, th, str(tmp_path)), nprocs=2, join=True)
    want = np.arange(W * H, dtype=np.float32) * 0.25 + 3.0
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"p{r}.npy"), want)
