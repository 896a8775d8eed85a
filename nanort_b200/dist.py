"""Multi-GPU plumbing: rays are sharded by image tile, the BVH is replicated, and the only exchange is
the final framebuffer gather.  One process per GPU; torch.distributed carries the
collective (NCCL on GPUs, gloo in the CPU tests).

Tile k (row-major over the tile grid) belongs to shard k % n_shards -- the same rule as
csrc/render.cu:slot_to_pixel, which this module mirrors on the host for the gather."""
from __future__ import annotations

import numpy as np


def shard_tiles(width, height, tile_w, tile_h, shard, n_shards):
    tiles_x = (width + tile_w - 1) // tile_w
    tiles_y = (height + tile_h - 1) // tile_h
    return np.arange(shard, tiles_x * tiles_y, n_shards, dtype=np.int64), tiles_x


def shard_pixels(width, height, tile_w, tile_h, shard, n_shards):
    """Flat pixel indices (y*width+x) owned by `shard`, tile-major / row-major inside the tile."""
    tiles, tiles_x = shard_tiles(width, height, tile_w, tile_h, shard, n_shards)
    ly, lx = np.meshgrid(np.arange(tile_h), np.arange(tile_w), indexing="ij")
    x = (tiles % tiles_x)[:, None] * tile_w + lx.reshape(-1)[None, :]
    y = (tiles // tiles_x)[:, None] * tile_h + ly.reshape(-1)[None, :]
    ok = (x < width) & (y < height)
    return (y * width + x)[ok]


def shard_ray_count(width, height, tile_w, tile_h, shard, n_shards, spp):
    return int(len(shard_pixels(width, height, tile_w, tile_h, shard, n_shards))) * int(spp)


class FramebufferGather:
    """all_gather of every rank's own pixels (packed, padded to the largest shard) + scatter into the frame.
    Built once per (image, tiling, world); `gather(local_frame)` returns the full frame on every rank."""

    def __init__(self, width, height, tile_w, tile_h, world_size, rank, device):
        import torch

        self.world, self.rank = world_size, rank
        self.n_pix = width * height
        px = [shard_pixels(width, height, tile_w, tile_h, r, world_size) for r in range(world_size)]
        self.pad = max(len(p) for p in px)
        self.mine = torch.as_tensor(px[rank], dtype=torch.int64, device=device)
        self.all_idx = [torch.as_tensor(p, dtype=torch.int64, device=device) for p in px]
        self.send = torch.zeros(self.pad, dtype=torch.float32, device=device)
        self.recv = torch.zeros(self.pad * world_size, dtype=torch.float32, device=device)
        self.bytes_per_rank = self.pad * 4

    def gather(self, local_frame):
        import torch
        import torch.distributed as dist

        flat = local_frame.reshape(-1)
        self.send[: len(self.mine)] = flat[self.mine]
        if self.world > 1:
            dist.all_gather_into_tensor(self.recv, self.send)
        else:
            self.recv.copy_(self.send)
        out = torch.zeros(self.n_pix, dtype=torch.float32, device=flat.device)
        for r in range(self.world):
            out[self.all_idx[r]] = self.recv[r * self.pad: r * self.pad + len(self.all_idx[r])]
        return out


def slot_pixels(width, height, tile_w, tile_h, shard, n_shards, spp):
    """Host mirror of csrc/wavefront.cuh:slot_to_pixel: for every primary slot of `shard` (tile-major, inside
    a tile sample-major over 8x4 pixel blocks) the pixel index, or -1 for slots outside the image, and the
    sample index."""
    tiles, tiles_x = shard_tiles(width, height, tile_w, tile_h, shard, n_shards)
    tile_pix = tile_w * tile_h
    q = np.arange(tile_pix, dtype=np.int64)
    bw = tile_w // 8
    blk, inb = q // 32, q % 32
    lx = (blk % bw) * 8 + (inb & 7)
    ly = (blk // bw) * 4 + (inb >> 3)
    x = (tiles % tiles_x)[:, None, None] * tile_w + lx[None, None, :]
    y = (tiles // tiles_x)[:, None, None] * tile_h + ly[None, None, :]
    x = np.broadcast_to(x, (len(tiles), spp, tile_pix))
    y = np.broadcast_to(y, (len(tiles), spp, tile_pix))
    pix = np.where((x < width) & (y < height), y * width + x, -1)
    smp = np.broadcast_to(np.arange(spp, dtype=np.int64)[None, :, None], pix.shape)
    return pix.reshape(-1), smp.reshape(-1)


# ---- host model of the C-ABI multi-GPU path (csrc/comm.cu: nrt_render_ao_sharded) --------------------------------
# A rank accumulates into a TILE-MAJOR buffer of its own tiles (NRT_AO_PACKED_TILES, csrc/wavefront.cuh): its k-th
# tile (image tile k * world + rank) occupies floats [k * tile_w * tile_h, (k + 1) * tile_w * tile_h), rows of tile_w.
# All ranks' buffers have the same size (ceil(n_tiles / world) tiles), so ONE equal-count all-gather moves them, and the
# unpack (comm.cu:unpack_tiles_kernel) reads pixel (x, y) from rank (tile % world), tile slot (tile // world).
def packed_slot_floats(width, height, tile_w, tile_h, world):
    tiles_x = (width + tile_w - 1) // tile_w
    tiles_y = (height + tile_h - 1) // tile_h
    return ((tiles_x * tiles_y + world - 1) // world) * tile_w * tile_h


def pack_own_tiles(frame, width, height, tile_w, tile_h, shard, n_shards):
    """What the device pass leaves in a rank's slot: `frame` (row-major, only this rank's pixels are read) tile-major."""
    out = np.zeros(packed_slot_floats(width, height, tile_w, tile_h, n_shards), dtype=frame.dtype)
    tiles, tiles_x = shard_tiles(width, height, tile_w, tile_h, shard, n_shards)
    img = frame.reshape(height, width)
    for k, t in enumerate(tiles):
        x0, y0 = int(t % tiles_x) * tile_w, int(t // tiles_x) * tile_h
        blk = img[y0:y0 + tile_h, x0:x0 + tile_w]
        dst = out[k * tile_w * tile_h:(k + 1) * tile_w * tile_h].reshape(tile_h, tile_w)
        dst[: blk.shape[0], : blk.shape[1]] = blk
    return out


def unpack_gathered(gathered, width, height, tile_w, tile_h, world):
    """comm.cu:unpack_tiles_kernel on the host: gathered = world slots of packed_slot_floats() floats each."""
    slot = packed_slot_floats(width, height, tile_w, tile_h, world)
    tiles_x = (width + tile_w - 1) // tile_w
    y, x = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    tx, ty = x // tile_w, y // tile_h
    tile = ty * tiles_x + tx
    src = (tile % world) * slot + (tile // world) * tile_w * tile_h + (y - ty * tile_h) * tile_w + (x - tx * tile_w)
    return np.asarray(gathered).reshape(-1)[src.reshape(-1)]
