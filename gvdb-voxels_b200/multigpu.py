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
    """Per-rank state for rendering frames of one size across `world` ranks.

    Frames are pipelined: the tile gather of frame j (NCCL stream) overlaps the render of frame j+1 (render stream);
    packed / gathered buffers are double-buffered so that a buffer is never rewritten while its gather is in flight."""

    def __init__(self, renderer, width, height, tile_size, rank, world, device):
        import torch
        self.r, self.w, self.h, self.ts, self.rank, self.world = renderer, width, height, tile_size, rank, world
        self.slots = slots_per_rank(width, height, tile_size, world)
        self.packed = [torch.zeros((self.slots, tile_size, tile_size, 4), dtype=torch.uint8, device=device) for _ in range(2)]
        self.gathered = [None, None]
        self.frame = None
        if rank == 0:
            self.gathered = [torch.zeros((world, self.slots, tile_size, tile_size, 4), dtype=torch.uint8, device=device)
                             for _ in range(2)]
            self.frame = torch.zeros((height, width, 4), dtype=torch.uint8, device=device)
        self._pending = None          # (buffer index, work handle) of the gather in flight

    def _assemble(self, b):
        src = self.packed[b] if self.world == 1 else self.gathered[b]
        self.r.assemble_tiles(src.data_ptr(), self.frame.data_ptr(), self.w, self.h, self.ts, self.world)

    def _finish_pending(self):
        if self._pending is not None:
            b, work = self._pending
            if work is not None:
                work.wait()           # render stream waits for the collective (stream-ordered, no host sync)
            if self.rank == 0:
                self._assemble(b)
            self._pending = None

    def render_frames(self, scninfos, shade, on_frame=None):
        """Render a sequence of frames; on rank 0 `on_frame(j, frame_tensor)` is called when frame j is assembled
        (stream-ordered: consume it on the current stream)."""
        import torch.distributed as dist
        for j, scn in enumerate(scninfos):
            b = j & 1
            self.r.render_tiles(scn, shade, self.packed[b].data_ptr(), self.ts, self.rank, self.world)
            work = None
            if self.world > 1:
                work = dist.gather(self.packed[b], list(self.gathered[b].unbind(0)) if self.rank == 0 else None, dst=0, async_op=True)
            prev = self._pending
            if prev is not None:      # finish frame j-1 while frame j's gather runs
                self._finish_pending()
                if on_frame is not None and self.rank == 0:
                    on_frame(j - 1, self.frame)
            self._pending = (b, work)
        self._finish_pending()
        if on_frame is not None and self.rank == 0 and len(scninfos):
            on_frame(len(scninfos) - 1, self.frame)
        return self.frame

    # single-frame convenience (no overlap)
    def render(self, scninfo, shade):
        self.render_frames([scninfo], shade)
        return self.frame
