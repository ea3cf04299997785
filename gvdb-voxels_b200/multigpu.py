"""Image-space tiling across ranks (one process per GPU, torch.distributed for the plumbing).

The volume is replicated on every GPU; a frame is cut into tile_size x tile_size tiles numbered row-major and dealt
round-robin: rank r renders tiles r, r+world, r+2*world, ... into a packed buffer of `slots` tiles
(gvdbx_render_tiles).  Rays that miss the volume are nearly free, so small interleaved tiles balance the load where
contiguous bands would not.  The only exchange step is the gather of finished tiles on rank 0 (NCCL over NVLink),
enqueued directly behind the render kernel; rank 0 then scatters [world][slots][tile] back into a row-major frame
(gvdbx_assemble_tiles).  The output bytes are a pure function of the pixel, hence identical for 1/2/4/8 GPUs.

The reference has no multi-GPU path at all; this file is the north-star extension around Render().
"""
import numpy as np


def tile_grid(width, height, tile_size):
    tx = (width + tile_size - 1) // tile_size
    ty = (height + tile_size - 1) // tile_size
    return tx, ty, tx * ty


def slots_per_rank(width, height, tile_size, world):
    return (tile_grid(width, height, tile_size)[2] + world - 1) // world


def tile_ids_for_rank(width, height, tile_size, rank, world):
    return list(range(rank, tile_grid(width, height, tile_size)[2], world))


def pack_tiles_host(frame, tile_size, rank, world):
    """numpy model of what gvdbx_render_tiles writes for `rank` given the full frame [h,w,4] (edge tiles zero padded)."""
    h, w, c = frame.shape
    tx, ty, n = tile_grid(w, h, tile_size)
    slots = slots_per_rank(w, h, tile_size, world)
    out = np.zeros((slots, tile_size, tile_size, c), frame.dtype)
    for k, t in enumerate(range(rank, n, world)):
        x0, y0 = (t % tx) * tile_size, (t // tx) * tile_size
        blk = frame[y0:y0 + tile_size, x0:x0 + tile_size]
        out[k, :blk.shape[0], :blk.shape[1]] = blk
    return out


def assemble_tiles_host(gathered, width, height, tile_size, world):
    """numpy model of gvdbx_assemble_tiles: gathered [world, slots, ts, ts, 4] -> frame [h, w, 4]."""
    tx, ty, n = tile_grid(width, height, tile_size)
    frame = np.zeros((height, width, gathered.shape[-1]), gathered.dtype)
    for t in range(n):
        r, k = t % world, t // world
        x0, y0 = (t % tx) * tile_size, (t // tx) * tile_size
        hh, ww = min(tile_size, height - y0), min(tile_size, width - x0)
        frame[y0:y0 + hh, x0:x0 + ww] = gathered[r, k, :hh, :ww]
    return frame


class TiledFrame:
    """Per-rank state for rendering one frame size across `world` ranks."""

    def __init__(self, renderer, width, height, tile_size, rank, world, device):
        import torch
        self.r, self.w, self.h, self.ts, self.rank, self.world = renderer, width, height, tile_size, rank, world
        self.slots = slots_per_rank(width, height, tile_size, world)
        self.packed = torch.zeros((self.slots, tile_size, tile_size, 4), dtype=torch.uint8, device=device)
        self.gathered = None
        self.frame = None
        if rank == 0:
            self.gathered = torch.zeros((world, self.slots, tile_size, tile_size, 4), dtype=torch.uint8, device=device)
            self.frame = torch.zeros((height, width, 4), dtype=torch.uint8, device=device)

    def render(self, scninfo, shade):
        self.r.render_tiles(scninfo, shade, self.packed.data_ptr(), self.ts, self.rank, self.world)

    def gather(self):
        """tiles -> rank 0 (NCCL gather on the current stream), then un-permute on rank 0."""
        import torch.distributed as dist
        if self.world == 1:
            src = self.packed
        else:
            dist.gather(self.packed, list(self.gathered.unbind(0)) if self.rank == 0 else None, dst=0)
            src = self.gathered
        if self.rank == 0:
            self.r.assemble_tiles(src.data_ptr(), self.frame.data_ptr(), self.w, self.h, self.ts, self.world)
        return self.frame
