"""debug_ring.py — torchrun --nproc-per-node N tests/debug_ring.py : every ring frame checked against a local render."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    import bench
    pkg = bench.load_pkg()
    from gvdb_voxels_b200 import multigpu as mg
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    p, vol = bench.build_workload(sys.argv[1] if len(sys.argv) > 1 else "cfg2_small")
    shade = p.shade
    nframes = 12
    scns, table = bench.frame_scninfos(pkg, p, shade, nframes)
    r = pkg.Renderer(local)
    r.import_topology_host(vol["vdbinfo"], vol["pool0"], vol["pool1"])
    r.import_atlas_host(vol["atlas"])
    r.set_transfer(table)
    w, h, ts = p.width, p.height, 32
    want = []
    for s in scns:
        o = torch.zeros((h, w, 4), dtype=torch.uint8, device=dev)
        r.render(s, shade, o.data_ptr())
        want.append(o)
    r.sync()
    # cross-GPU determinism of the plain render
    mine = want[3].clone()
    dist.broadcast(mine, src=0)
    print(f"[rank {rank}] plain render equal to rank 0's: {bool(torch.equal(mine, want[3]))}", flush=True)
    for lanes in (0, 4):
        r.lanes(lanes)
        ring = mg.PeerFrameRing(r, w, h, ts, rank, world, nslots=4)
        consumer = torch.cuda.Stream(device=dev)
        got = []
        r.lanes_fork()
        for j, s in enumerate(scns):
            q = ring.submit(s, shade)
            if rank == 0:
                ring.acquire(q, consumer.cuda_stream)
                with torch.cuda.stream(consumer):
                    got.append(ring.frame_tensor(q, torch, dev).clone())
                ring.release(q, consumer.cuda_stream)
        r.lanes_join()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        if rank == 0:
            tx = (w + ts - 1) // ts
            for j in range(nframes):
                bad = (got[j] != want[j]).any(dim=2).cpu().numpy()
                ys, xs = np.nonzero(bad)
                owner = ((ys // ts) * tx + xs // ts) % world
                final = (ring.frame_tensor(j + 1, torch, dev) != want[j]).any(dim=2).sum().item() if j >= nframes - 4 else -1
                print(f"lanes {lanes} frame {j}: bad {int(bad.sum())} by owner {np.bincount(owner, minlength=world).tolist()} final-state bad {final}", flush=True)
        dist.barrier()
        # bench-like phase: several steps of 8 frames, fork/join per step, no per-frame consumer work
        for step in range(6):
            r.lanes_fork()
            for j in range(8):
                q = ring.submit(scns[j], shade)
                if rank == 0:
                    ring.acquire(q, consumer.cuda_stream)
                    ring.release(q, consumer.cuda_stream)
            r.lanes_join()
        torch.cuda.current_stream().wait_stream(consumer)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        if rank == 0:
            for j in range(4, 8):
                q = ring.seq - (7 - j)
                bad = (ring.frame_tensor(q, torch, dev) != want[j]).any(dim=2).sum().item()
                print(f"lanes {lanes} bench-like final state of frame {j} (q {q}): bad {bad}", flush=True)
        dist.barrier()
        ring.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
