"""check_multi.py — run under `gpurun --gpus 2` (or more): the multi-GPU entry points that need more than one device.

  1. gvdbx_render_multi: N contexts of ONE process (one per device) render one frame into context 0's buffer over NVLink;
     compared byte for byte with the single-context render, all four core modes.
  2. torchrun world N: the C peer frame ring (gvdbx_ring_*, CUDA IPC between processes) and the host frame ring
     (gvdbx_hostring_*, shared page-locked segment): every delivered frame equals the single-GPU render.
Prints one JSON line per part; exit code 0 only if everything matched."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def part1():
    import numpy as np
    import torch
    import oracle
    from __graft_entry__ import load_package
    pkg = load_package()
    n = torch.cuda.device_count()
    p, vol = oracle.scene_volume("cfg4_small")
    res = {"part": "gvdbx_render_multi", "devices": n, "modes": {}}
    rs = []
    for d in range(n):
        r = pkg.Renderer(d)
        r.import_topology_host(vol["vdbinfo"], vol["pool0"], vol["pool1"])
        r.import_atlas_host(vol["atlas"])
        rs.append(r)
    ok = True
    w, h = 640, 360
    for mode, shade in (("voxel", 0), ("trilinear", 4), ("levelset", 6), ("deep", 7)):
        scn, table = oracle.scninfo_for(pkg, p, shade=shade, size=(w, h))
        for r in rs:
            r.set_transfer(table)
        with torch.cuda.device(0):
            ref = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda:0")
            out = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda:0")
            rs[0].render(scn, shade, ref.data_ptr())
            rs[0].sync()
            pkg.render_multi(rs, scn, shade, out.data_ptr(), 32)
            rs[0].sync()
        same = bool(torch.equal(ref, out))
        res["modes"][mode] = {"identical": same, "nonbackground": int((ref != ref[0, 0]).any(dim=2).sum())}
        ok &= same
    for r in rs:
        r.close()
    res["ok"] = ok
    print(json.dumps(res), flush=True)
    return ok


def part2():
    import numpy as np
    import torch
    import torch.distributed as dist
    import oracle
    from __graft_entry__ import load_package
    pkg = load_package()
    from gvdb_voxels_b200 import multigpu as mg
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    p, vol = oracle.scene_volume("cfg4_small")
    w, h = 500, 300
    r = pkg.Renderer(local)
    r.import_topology_host(vol["vdbinfo"], vol["pool0"], vol["pool1"])
    r.import_atlas_host(vol["atlas"])
    nframes, nslots = 9, 3
    scns = []
    for j in range(nframes):
        s, table = oracle.scninfo_for(pkg, p, shade=7, size=(w, h), cam_angs=(p.cam_angs[0] + 40.0 * j, p.cam_angs[1], p.cam_angs[2]))
        scns.append(s)
    r.set_transfer(table)
    r.lanes(2)
    ok = True
    # ---- peer frame ring (device frames on rank 0)
    ring = mg.PeerFrameRing(r, w, h, 32, rank, world, nslots=nslots)
    cons = torch.cuda.Stream(device=dev)
    got = []
    for j in range(nframes):
        q = ring.submit(scns[j], 7)
        if rank == 0:
            ring.acquire(q, cons.cuda_stream)
            with torch.cuda.stream(cons):
                got.append(ring.frame_tensor(q, torch, dev).clone())
            ring.release(q, cons.cuda_stream)
    torch.cuda.synchronize()
    dist.barrier()
    ring.close()
    # ---- host frame ring
    hr = mg.HostFrameRing(r, f"/gvdbx_check_{os.environ.get('MASTER_PORT', '0')}", w, h, rank, world, nslots=nslots, band_rows=16)
    hgot = []
    for j in range(nframes):
        q = j + 1
        if rank == 0 and q > nslots:
            hgot.append(hr.wait(q - nslots).copy())
            hr.release(q - nslots)
        hr.submit(scns[j], 7)
    if rank == 0:
        for q in range(nframes - nslots + 1, nframes + 1):
            hgot.append(hr.wait(q).copy())
            hr.release(q)
    torch.cuda.synchronize()
    dist.barrier()
    hr.close()
    if rank == 0:
        r.lanes(0)
        bad_ring = bad_host = 0
        for j in range(nframes):
            ref = torch.zeros((h, w, 4), dtype=torch.uint8, device=dev)
            r.render(scns[j], 7, ref.data_ptr())
            r.sync()
            bad_ring += int(not torch.equal(ref, got[j]))
            bad_host += int(not np.array_equal(ref.cpu().numpy(), hgot[j]))
        ok = bad_ring == 0 and bad_host == 0
        print(json.dumps({"part": "rings", "world": world, "frames": nframes, "peer_ring_bad_frames": bad_ring, "host_ring_bad_frames": bad_host, "ok": ok}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    r.close()
    return ok


if __name__ == "__main__":
    sys.exit(0 if (part2() if "RANK" in os.environ else part1()) else 1)
