"""World-size-2 CPU test (gloo) of the host-side tiling logic used by the multi-GPU path
(gvdb-voxels_b200/multigpu.py): partition -> gather on rank 0 -> assemble reproduces the frame."""
import os
import socket
import sys

import numpy as np
import pytest

from common import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, w, h, ts, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    load_package()
    from gvdb_voxels_b200 import multigpu as mg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    yy, xx = np.mgrid[0:h, 0:w]
    frame = np.stack([xx % 251, yy % 241, (xx * 7 + yy * 13) % 256, np.full_like(xx, 255)], axis=2).astype(np.uint8)
    packed = torch.from_numpy(mg.pack_tiles_host(frame, ts, rank, world))
    slots = mg.slots_per_rank(w, h, ts, world)
    assert packed.shape[0] == slots
    gathered = [torch.zeros_like(packed) for _ in range(world)] if rank == 0 else None
    dist.gather(packed, gathered, dst=0)
    ok = True
    if rank == 0:
        g = torch.stack(gathered).numpy()
        out = mg.assemble_tiles_host(g, w, h, ts, world)
        ok = bool(np.array_equal(out, frame))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok))


@pytest.mark.parametrize("size", [(96, 54, 32), (200, 120, 16), (64, 64, 32)])
def test_tile_partition_gather_assemble_two_ranks(size):
    import torch.multiprocessing as mp
    w, h, ts = size
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, w, h, ts, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_tile_ids_cover_frame_exactly_once():
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    load_package()
    from gvdb_voxels_b200 import multigpu as mg
    for w, h, ts, world in [(3840, 2160, 32, 8), (1920, 1080, 32, 4), (97, 55, 16, 3)]:
        ids = sorted(t for r in range(world) for t in mg.tile_ids_for_rank(w, h, ts, r, world))
        assert ids == list(range(mg.tile_grid(w, h, ts)[2]))
        assert max(len(mg.tile_ids_for_rank(w, h, ts, r, world)) for r in range(world)) == mg.slots_per_rank(w, h, ts, world)
