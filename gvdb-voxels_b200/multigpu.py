"""Image-space tiling across ranks (one process per GPU, torch.distributed only for bootstrap blobs, barriers and the
replication of the volume).  Thin ctypes callers of the C entry points in csrc/gvdbx_multi.cu — the protocols live there.

The volume is replicated on every GPU; output bytes are a pure function of the pixel, hence identical for 1/2/4/8 GPUs.

* PeerFrameRing (default, frames wanted on rank 0's DEVICE): tile_size x tile_size tiles dealt round-robin; every rank's
  tile-list kernel stores its pixels straight into rank 0's frame over NVLink (CUDA IPC mapping); two 4-byte flags per frame
  order producers and consumer, all waits / signals are stream-ordered device operations (gvdbx_ring_*).
* HostFrameRing (frames wanted on the HOST): full-width bands dealt round-robin; every rank copies ITS bands over ITS OWN PCIe
  link into a shared page-locked POSIX shm frame, one pitched copy + a one-thread flag kernel per frame (gvdbx_hostring_*).
* TiledFrame (`bench.py --exchange nccl`, the first implementation, kept as the NCCL path north_star names): packed tile slots
  (gvdbx_render_tiles) + an NCCL gather on rank 0 + gvdbx_assemble_tiles.

The reference has no multi-GPU path at all; this file is the north-star extension around Render().
"""
import os

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


# ------------------------------------------------------------------------------------------------ peer frame ring
def ring_slot(q, nslots):
    """frame sequence number q = 1, 2, ... -> (slot, how many times the slot has been used up to and including q)"""
    return (q - 1) % nslots, (q - 1) // nslots + 1


class CudaBuffer:
    """raw device memory as a __cuda_array_interface__ object (wrap with torch.as_tensor(buf, device=...))"""

    def __init__(self, ptr, shape, typestr="|u1"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


RING_EXPORT_BYTES = 160


class PeerFrameRing:
    """Multi-GPU output without a gather (C entry points gvdbx_ring_*, csrc/gvdbx_multi.cu): rank 0 owns a ring of `nslots`
    row-major frames; every rank's render kernel stores the pixels of its tiles straight into the current slot — over NVLink
    for ranks != 0 (CUDA IPC peer mapping) — so render and "gather" are ONE kernel and the only other traffic is two 4-byte
    flags per frame:

      done[slot]      (rank 0's memory)  += 1 by every rank behind its render kernel (system-scope release);
                      rank 0's consumer stream waits for world * uses(slot) before it touches the slot
      released        (every rank's own memory) = q, written by rank 0 once frame q has been consumed;
                      a rank waits for released >= q - nslots before it renders frame q into the same slot again

    All waits and signals are stream-ordered device-side operations: no host synchronisation, no NCCL on the data path.
    This class only ships the bootstrap blobs: `exchange(blob) -> list of every rank's blob in any order`
    (torch.distributed.all_gather_object by default); connect=False defers gvdbx_ring_connect (ranks sharing one process)."""

    def __init__(self, renderer, width, height, tile_size, rank, world, nslots=4, exchange=None, connect=True):
        import ctypes as C
        from . import api
        self.r, self.w, self.h, self.ts, self.rank, self.world, self.nslots = renderer, width, height, tile_size, rank, world, nslots
        self._L = api.lib()
        self.seq = 0
        g = C.c_void_p()
        blob = (C.c_uint8 * RING_EXPORT_BYTES)()
        renderer._ck(self._L.gvdbx_ring_create(renderer._h, width, height, tile_size, rank, world, nslots, C.byref(g), blob), "gvdbx_ring_create")
        self._g = g
        self.blob = bytes(blob)
        if connect:
            if exchange is None:
                def exchange(obj):
                    import torch.distributed as dist
                    out = [None] * world
                    dist.all_gather_object(out, obj)
                    return out
            self.connect(exchange(self.blob))

    def connect(self, blobs):
        """blobs: every rank's export blob (any order: the rank is the first int32 of a blob)"""
        blobs = sorted(blobs, key=lambda b: int.from_bytes(b[0:4], "little", signed=True))
        self.r._ck(self._L.gvdbx_ring_connect(self._g, b"".join(blobs)), "gvdbx_ring_connect")

    def check(self):
        """raises GvdbxError if any stream-ordered wait of this rank ran into its ~20 s timeout (lost peer, stalled consumer):
        the polling kernel then lets the stream continue, so frames after it may be incomplete or overwritten — the context
        keeps a sticky error bit that gvdbx_sync reports"""
        self.r.sync()

    def close(self):
        if self._g:
            g, self._g = self._g, None
            self.r._ck(self._L.gvdbx_ring_destroy(g), "gvdbx_ring_destroy")

    # ---- every rank
    def submit(self, scninfo, shade, chan=0):
        """enqueue this rank's share of the next frame on the renderer's stream / lane; returns the frame's sequence number"""
        import ctypes as C
        from .api import _buf
        p, keep = _buf(scninfo)
        q = C.c_uint32()
        self.r._ck(self._L.gvdbx_ring_submit(self._g, p, shade, chan, C.byref(q)), "gvdbx_ring_submit")
        self.seq = int(q.value)
        return self.seq

    # ---- rank 0 (consumer)
    def acquire(self, q, stream=None):
        """make `stream` wait until every rank has delivered frame q; returns the device pointer of the finished frame"""
        import ctypes as C
        d = C.c_uint64()
        self.r._ck(self._L.gvdbx_ring_acquire(self._g, int(q), C.c_void_p(stream or 0), C.byref(d)), "gvdbx_ring_acquire")
        return int(d.value)

    def release(self, q, stream=None):
        """behind the consumer's work on `stream`: hand the slot of frame q back to the producers"""
        import ctypes as C
        self.r._ck(self._L.gvdbx_ring_release(self._g, int(q), C.c_void_p(stream or 0)), "gvdbx_ring_release")

    def frame_ptr_of(self, q):
        import ctypes as C
        d = C.c_uint64()
        self.r._ck(self._L.gvdbx_ring_frame(self._g, int(q), C.byref(d)), "gvdbx_ring_frame")
        return int(d.value)

    def frame_tensor(self, q, torch, device):
        return torch.as_tensor(CudaBuffer(self.frame_ptr_of(q), (self.h, self.w, 4)), device=device)


def bands_for_rank(height, band_rows, rank, world):
    """[(first row, row count)] of the full-width bands rank `rank` renders and copies in the host frame ring (band b of the
    frame belongs to rank b % world) — the partition gvdbx_hostring_submit / gvdbx_render_bands use"""
    nb = (height + band_rows - 1) // band_rows
    return [(b * band_rows, min(band_rows, height - b * band_rows)) for b in range(rank, nb, world)]


class HostFrameRing:
    """Frames of a multi-GPU render delivered to the HOST (C entry points gvdbx_hostring_*): a ring of row-major frames in a
    POSIX shared-memory segment page-locked by every process; each rank renders full-width bands of `band_rows` rows and copies
    ITS bands over ITS OWN PCIe link.  submit(): every rank; wait() / release(): the consumer (rank 0)."""

    def __init__(self, renderer, name, width, height, rank, world, nslots=4, band_rows=32):
        import ctypes as C
        from . import api
        self.r, self.w, self.h, self.rank, self.world, self.nslots = renderer, width, height, rank, world, nslots
        self._L = api.lib()
        g = C.c_void_p()
        renderer._ck(self._L.gvdbx_hostring_create(renderer._h, name.encode(), width, height, band_rows, rank, world, nslots, C.byref(g)),
                     "gvdbx_hostring_create")
        self._g = g
        self.seq = 0

    def submit(self, scninfo, shade, chan=0):
        import ctypes as C
        from .api import _buf
        p, keep = _buf(scninfo)
        q = C.c_uint32()
        self.r._ck(self._L.gvdbx_hostring_submit(self._g, p, shade, chan, C.byref(q)), "gvdbx_hostring_submit")
        self.seq = int(q.value)
        return self.seq

    def wait(self, q, timeout_ms=30000):
        """-> numpy view [h, w, 4] uint8 of the finished frame in the shared segment (valid until release(q))"""
        import ctypes as C
        ptr = C.c_void_p()
        self.r._ck(self._L.gvdbx_hostring_wait(self._g, int(q), C.byref(ptr), int(timeout_ms)), "gvdbx_hostring_wait")
        buf = (C.c_uint8 * (self.w * self.h * 4)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=np.uint8).reshape(self.h, self.w, 4)

    def release(self, q):
        self.r._ck(self._L.gvdbx_hostring_release(self._g, int(q)), "gvdbx_hostring_release")

    def close(self):
        if self._g:
            g, self._g = self._g, None
            self._L.gvdbx_hostring_destroy(g)


class PeerFrameRingPy:
    """Executable model of the peer frame ring protocol that csrc/gvdbx_multi.cu implements (gvdbx_ring_*), written over the
    low-level entry points (peer_alloc / peer_open / stream_wait / stream_signal_* / render_tiles_direct).  The CPU tests run
    it with a fake renderer on shared memory (tests/test_peer_ring_gloo.py: two real processes, gloo); the product path is
    PeerFrameRing above.

    Multi-GPU output without a gather: rank 0 owns a ring of `nslots` row-major frames; every rank's render kernel
    stores the pixels of its tiles straight into the current slot — over NVLink for ranks != 0 (CUDA IPC peer mapping,
    gvdbx_peer_open) — so render and "gather" are ONE kernel and the only other traffic is two 4-byte flags per frame:

      done[slot]      (rank 0's memory)  += 1 by every rank behind its render kernel (system-scope release);
                      rank 0's consumer stream waits for world * uses(slot) before it touches the slot
      released        (every rank's own memory) = q, written by rank 0 once frame q has been consumed;
                      a rank waits for released >= q - nslots before it renders frame q into the same slot again

    All waits and signals are stream-ordered device-side operations: no host synchronisation, no NCCL on the data path.
    `exchange(obj) -> list of every rank's obj` is the bootstrap (torch.distributed.all_gather_object by default)."""

    ALIGN = 256

    def __init__(self, renderer, width, height, tile_size, rank, world, nslots=4, exchange=None):
        self.r, self.w, self.h, self.ts, self.rank, self.world, self.nslots = renderer, width, height, tile_size, rank, world, nslots
        self.frame_bytes = (width * height * 4 + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.seq = 0
        self._own, self._opened = [], []
        if exchange is None:
            def exchange(obj):
                import torch.distributed as dist
                out = [None] * world
                dist.all_gather_object(out, obj)
                return out
        # every rank: a `released` flag of its own; rank 0 additionally: the frame slots followed by the done counters
        rel_ptr, rel_handle = renderer.peer_alloc(self.ALIGN)
        self._own.append(rel_ptr)
        self.released_local = rel_ptr
        ring_handle = None
        if rank == 0:
            ring_ptr, ring_handle = renderer.peer_alloc(self.frame_bytes * nslots + self.ALIGN * nslots)
            self._own.append(ring_ptr)
        # (an IPC handle cannot be opened by the process that made it: ranks living in one process — tests — pass raw pointers)
        handles = exchange({"rank": rank, "pid": os.getpid(), "released": rel_handle, "released_ptr": rel_ptr, "ring": ring_handle,
                            "ring_ptr": ring_ptr if rank == 0 else None})
        handles = sorted(handles, key=lambda d: d["rank"])
        if rank == 0:
            self.ring_base = ring_ptr
            self.released_all = [rel_ptr] + [self._open(handles[k], "released") for k in range(1, world)]
        else:
            self.ring_base = self._open(handles[0], "ring")
            self.released_all = None
        self.frame_ptr = [self.ring_base + s * self.frame_bytes for s in range(nslots)]
        self.done_ptr = [self.ring_base + nslots * self.frame_bytes + s * self.ALIGN for s in range(nslots)]

    def _open(self, entry, key):
        if entry["pid"] == os.getpid():
            return entry[key + "_ptr"]
        p = self.r.peer_open(entry[key])
        self._opened.append(p)
        return p

    def close(self):
        for p in self._opened:
            self.r.peer_close(p)
        for p in self._own:
            self.r.peer_free(p)
        self._opened, self._own = [], []

    # ---- every rank
    def submit(self, scninfo, shade):
        """enqueue this rank's share of the next frame on the renderer's stream; returns the frame's sequence number"""
        self.seq += 1
        q = self.seq
        slot, _ = ring_slot(q, self.nslots)
        if getattr(self.r, "nlanes", 0):                      # consecutive frames on alternating internal streams
            self.r.lane_select((q - 1) % self.r.nlanes)
        # the slot's previous frame (q - nslots) must have been consumed before its pixels are overwritten
        wait = (self.released_local, q - self.nslots) if q > self.nslots else (0, 0)
        if hasattr(self.r, "render_tiles_ring"):
            self.r.render_tiles_ring(scninfo, shade, self.frame_ptr[slot], self.ts, self.rank, self.world, wait[0], wait[1], self.done_ptr[slot])
        else:
            if wait[0]:
                self.r.stream_wait(*wait)
            self.r.render_tiles_direct(scninfo, shade, self.frame_ptr[slot], self.ts, self.rank, self.world)
            self.r.stream_signal_add(self.done_ptr[slot], 1)
        return q

    # ---- rank 0 (consumer)
    def acquire(self, q, stream=None):
        """make `stream` wait until every rank has delivered frame q; returns the device pointer of the finished frame"""
        slot, uses = ring_slot(q, self.nslots)
        self.r.stream_wait(self.done_ptr[slot], self.world * uses, stream)
        return self.frame_ptr[slot]

    def release(self, q, stream=None):
        """behind the consumer's work on `stream`: hand the slot of frame q back to the producers"""
        self.r.stream_signal_many(self.released_all, q, stream)

    def frame_tensor(self, q, torch, device):
        slot, _ = ring_slot(q, self.nslots)
        return torch.as_tensor(CudaBuffer(self.frame_ptr[slot], (self.h, self.w, 4)), device=device)


# ------------------------------------------------------------------------------------------------ volume replication
def replicate_volume(renderer, vol, rank, world, device, src=0):
    """Replicate the volume of rank `src` on every rank: pools and atlas travel GPU to GPU (torch.distributed broadcast,
    NCCL over NVLink) and are imported from device memory (gvdbx_import_topology with the VDBInfo pointers patched to the
    received pools, gvdbx_import_atlas_device) — instead of every process rebuilding the topology and pushing the atlas
    through its own PCIe link.  `vol` = {"vdbinfo", "pool0", "pool1", "atlas"} on rank `src`, ignored elsewhere.
    Returns the metadata dict {"atlas_shape", "bricks", ...} on every rank."""
    import torch
    import torch.distributed as dist
    meta = [None]
    if rank == src:
        meta[0] = {"vdbinfo": bytes(vol["vdbinfo"]), "atlas_shape": tuple(vol["atlas"].shape),
                   "pool0": {int(l): int(len(b)) for l, b in vol["pool0"].items() if len(b)},
                   "pool1": {int(l): int(len(b)) for l, b in vol["pool1"].items() if len(b)}}
    if world > 1:
        dist.broadcast_object_list(meta, src=src)
    m = meta[0]

    def bcast(host_array, nbytes, dtype):
        if rank == src:
            t = torch.from_numpy(np.ascontiguousarray(host_array)).view(dtype).to(device)
        else:
            t = torch.empty(nbytes // torch.empty(0, dtype=dtype).element_size(), dtype=dtype, device=device)
        if world > 1:
            dist.broadcast(t, src=src)
        return t
    pools = {}
    for grp in ("pool0", "pool1"):
        for lev, n in m[grp].items():
            pools[(grp, lev)] = bcast(vol[grp][lev] if rank == src else None, n, torch.uint8)
    # VDBInfo: nodelist[] at byte 440, childlist[] at byte 520 (8 bytes per level) hold the DEVICE pointers of the pools
    vb = bytearray(m["vdbinfo"])
    for lev in range(10):
        p0 = pools.get(("pool0", lev))
        p1 = pools.get(("pool1", lev))
        vb[440 + 8 * lev: 448 + 8 * lev] = int(p0.data_ptr() if p0 is not None else 0).to_bytes(8, "little")
        vb[520 + 8 * lev: 528 + 8 * lev] = int(p1.data_ptr() if p1 is not None else 0).to_bytes(8, "little")
    renderer.import_topology(bytes(vb))
    rz, ry, rx = m["atlas_shape"]
    atlas = bcast(vol["atlas"].reshape(-1) if rank == src else None, rz * ry * rx * 4, torch.float32)
    renderer.import_atlas_device(atlas.data_ptr(), (rx, ry, rz))
    renderer.sync()
    del atlas, pools
    return m
