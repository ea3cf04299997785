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


# ------------------------------------------------------------------------------------------------ volume replication
class _FakeImportRenderer:
    """records what multigpu.replicate_volume hands to the import calls (CPU tensors: the 'device pointers' are host
    addresses, read back with ctypes)"""

    def __init__(self):
        self.pools, self.atlas = {}, None

    def import_topology(self, vdbinfo):
        import ctypes
        vb = np.frombuffer(vdbinfo, np.uint8)
        assert vb.size == 1232
        cnt, wid, cw = vb[320:360].view(np.int32), vb[360:400].view(np.int32), vb[400:440].view(np.int32)
        nl, cl = vb[440:520].view(np.uint64), vb[520:600].view(np.uint64)
        for lev in range(5):
            if nl[lev]:
                self.pools[("pool0", lev)] = ctypes.string_at(int(nl[lev]), int(cnt[lev]) * int(wid[lev]))
            if cl[lev]:
                self.pools[("pool1", lev)] = ctypes.string_at(int(cl[lev]), int(self.n1[lev]))
        self.vdbinfo_masked = bytes(vb[:440]) + bytes(vb[600:])

    def import_atlas_device(self, ptr, res_xyz):
        import ctypes
        rx, ry, rz = res_xyz
        self.atlas = np.frombuffer(ctypes.string_at(int(ptr), rx * ry * rz * 4), np.float32).reshape(rz, ry, rx).copy()

    def sync(self):
        pass


def _replicate_worker(rank, world, port, q):
    import hashlib
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from __graft_entry__ import load_package
    load_package()
    from gvdb_voxels_b200 import multigpu as mg
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p, vol = oracle.scene_volume("cfg3_tiny")          # every rank can build it: the received copy is compared with it
    r = _FakeImportRenderer()
    r.n1 = {lev: len(b) for lev, b in vol["pool1"].items()}
    meta = mg.replicate_volume(r, vol if rank == 0 else None, rank, world, torch.device("cpu"))
    ok = tuple(meta["atlas_shape"]) == vol["atlas"].shape and np.array_equal(r.atlas, vol["atlas"])
    for (grp, lev), b in r.pools.items():
        ok &= b == bytes(np.ascontiguousarray(vol[grp][lev]).tobytes())
    ok &= len(r.pools) == sum(1 for g in ("pool0", "pool1") for b in vol[g].values() if len(b))
    vb = np.frombuffer(vol["vdbinfo"], np.uint8)
    ok &= r.vdbinfo_masked == bytes(vb[:440]) + bytes(vb[600:])
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok)))


def test_replicate_volume_two_ranks():
    """rank 0's pools / atlas / VDBInfo arrive bit-identical on rank 1 and are handed to the import calls by pointer"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_replicate_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_band_partition_covers_every_row_once(pkg):
    from gvdb_voxels_b200 import multigpu as mg
    for h, rows, world in ((2160, 32, 8), (1080, 32, 8), (768, 16, 3), (270, 32, 4), (54, 8, 7)):
        seen = np.zeros(h, np.int32)
        for rank in range(world):
            for y0, n in mg.bands_for_rank(h, rows, rank, world):
                assert n > 0 and y0 % rows == 0 and (y0 // rows) % world == rank
                seen[y0:y0 + n] += 1
        assert (seen == 1).all()


def test_shim_header_compiles_against_reference_headers(tmp_path):
    """include/gvdbx_shim.h (the Level-A drop-in a GVDB maintainer adds) is valid C++ against the reference's own gvdb.h —
    checked wherever the reference sources are present (the GPU box runs it for real: ref_harness_x)."""
    import os
    import subprocess
    from common import ROOT
    ref = "/root/reference/source/gvdb_library"
    stub = os.path.join(ROOT, "oracle", "_ref", "obj", "stubinc")
    if not os.path.isdir(ref) or not os.path.isdir(stub):
        pytest.skip("reference sources / GL stub headers not present")
    src = tmp_path / "shim.cpp"
    src.write_text('#include "gvdbx_shim.h"\nint main() { VolumeGVDBX v; (void)v; return 0; }\n')
    subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-w", "-DBUILD_OPENGL", "-DGLEW_STATIC", "-DGLEW_NO_GLU", "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(ref, "src"), "-I", os.path.join(ref, "glew", "include"), "-I", stub, "-I", "/usr/local/cuda/include", str(src)], check=True)
