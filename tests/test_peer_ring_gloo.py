"""World-size-2 CPU test (gloo) of the peer frame ring protocol (gvdb-voxels_b200/multigpu.py::PeerFrameRingPy, the executable model of
csrc/gvdbx_multi.cu::gvdbx_ring_*): two real
processes, "device memory" = POSIX shared memory, the stream-ordered flag operations executed synchronously.  Checks the
host-side logic the multi-GPU path depends on: handle exchange, slot / use arithmetic, back-pressure (a slot is never
rewritten before rank 0 released it) and that every finished frame holds every rank's tiles of THAT frame."""
import os
import socket
import sys
import time

import numpy as np
import pytest

from common import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class FakeRenderer:
    """stands in for api.Renderer: peer_alloc / peer_open on shared memory, flags and tile writes done immediately"""

    def __init__(self, rank):
        from multiprocessing import shared_memory
        self.shm = shared_memory
        self.rank, self.blocks, self.base = rank, {}, {}
        self.next = 1 << 20

    def _map(self, blk):
        ptr = self.next
        self.next += (blk.size + 4095) // 4096 * 4096 + 4096
        self.blocks[ptr] = blk
        return ptr

    def peer_alloc(self, nbytes):
        blk = self.shm.SharedMemory(create=True, size=nbytes)
        blk.buf[:nbytes] = bytes(nbytes)
        return self._map(blk), blk.name.encode().ljust(64, b"\0")

    def peer_open(self, handle):
        return self._map(self.shm.SharedMemory(name=bytes(handle).rstrip(b"\0").decode()))

    def peer_close(self, ptr):
        self.blocks.pop(ptr).close()

    def peer_free(self, ptr):
        blk = self.blocks.pop(ptr)
        blk.close()
        blk.unlink()

    def view(self, ptr, nbytes, dtype=np.uint8):
        for base, blk in self.blocks.items():
            if base <= ptr < base + blk.size:
                return np.frombuffer(blk.buf, dtype=np.uint8, count=nbytes, offset=ptr - base).view(dtype)
        raise KeyError(ptr)

    def stream_wait(self, ptr, value, stream=None):
        f = self.view(ptr, 4, np.uint32)
        t0 = time.time()
        while np.int32(f[0] - np.uint32(value)) < 0:
            assert time.time() - t0 < 60, "flag wait timed out"
            time.sleep(0.0005)

    def stream_signal_add(self, ptr, inc, stream=None):
        # one writer at a time per counter in this test would be a race between two processes: serialise with a lock file
        import fcntl
        with open(os.path.join("/tmp", "gvdbx_ring_test.lock"), "w") as lk:
            fcntl.flock(lk, fcntl.LOCK_EX)
            f = self.view(ptr, 4, np.uint32)
            f[0] = f[0] + np.uint32(inc)

    def stream_signal_many(self, ptrs, value, stream=None):
        for p in ptrs:
            self.view(p, 4, np.uint32)[0] = np.uint32(value)

    def render_tiles_direct(self, scninfo, shade, frame_ptr, ts, rank, nranks, chan=0):
        w, h, frame_no = scninfo
        fr = self.view(frame_ptr, w * h * 4).reshape(h, w, 4)
        tx = (w + ts - 1) // ts
        ntiles = tx * ((h + ts - 1) // ts)
        for t in range(rank, ntiles, nranks):
            x0, y0 = (t % tx) * ts, (t // tx) * ts
            fr[y0:y0 + ts, x0:x0 + ts] = (rank + 1, frame_no % 251, t % 256, 255)
        time.sleep(0.002 * (rank + 1))          # ranks finish at different times


def _worker(rank, world, port, w, h, ts, nslots, nframes, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    load_package()
    from gvdb_voxels_b200 import multigpu as mg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r = FakeRenderer(rank)
    ring = mg.PeerFrameRingPy(r, w, h, ts, rank, world, nslots=nslots)
    ok = True
    tx = (w + ts - 1) // ts
    for f in range(nframes):
        qn = ring.submit((w, h, f), 0)
        if rank == 0:
            ptr = ring.acquire(qn)
            fr = r.view(ptr, w * h * 4).reshape(h, w, 4).copy()
            for t in range(tx * ((h + ts - 1) // ts)):
                x0, y0 = (t % tx) * ts, (t // tx) * ts
                blk = fr[y0:y0 + ts, x0:x0 + ts]
                ok &= bool((blk == np.array([t % world + 1, f % 251, t % 256, 255], np.uint8)).all())
            ring.release(qn)
    dist.barrier()
    ring.close()
    dist.destroy_process_group()
    q.put((rank, ok, ring.seq))


@pytest.mark.parametrize("cfg", [(96, 54, 32, 2, 9), (64, 64, 16, 4, 6), (50, 30, 8, 1, 5)])
def test_peer_ring_two_ranks(cfg):
    import torch.multiprocessing as mp
    w, h, ts, nslots, nframes = cfg
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, w, h, ts, nslots, nframes, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(seq == nframes for _, _, seq in res)


def test_ring_slot_arithmetic():
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    load_package()
    from gvdb_voxels_b200 import multigpu as mg
    for nslots in (1, 2, 4, 5):
        uses = [0] * nslots
        for q in range(1, 40):
            slot, u = mg.ring_slot(q, nslots)
            uses[slot] += 1
            assert u == uses[slot] and slot == (q - 1) % nslots
