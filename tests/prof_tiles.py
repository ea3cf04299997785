"""prof_tiles.py — one GPU: cost of the tile-list kernels against the plain full-frame kernel (what a rank of an N-GPU run
executes), per tile size, plus frames alternating between two streams.  CUDA events, 8 orbit frames x reps.
  python tests/prof_tiles.py --workload cfg2 --worlds 1,2,8 --tiles 16,32,64
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import bench
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--mode", default="")
    ap.add_argument("--worlds", default="1,2,8")
    ap.add_argument("--tiles", default="16,32,64")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--bands", default="", help="also time full-width bands of these heights (the host frame ring's partition)")
    ap.add_argument("--size", default="")
    a = ap.parse_args()
    pkg = bench.load_pkg()
    size = tuple(int(x) for x in a.size.split("x")) if a.size else None
    p, vol = bench.build_workload(a.workload, size)
    shade, dshadow = bench.MODE[a.mode] if a.mode else (p.shade, 0)
    scns, table = bench.frame_scninfos(pkg, p, shade, 8)
    r = pkg.Renderer(0)
    r.import_topology_host(vol["vdbinfo"], vol["pool0"], vol["pool1"])
    r.import_atlas_host(vol["atlas"])
    r.set_transfer(table)
    r.set_deep_shadow(dshadow)
    w, h = p.width, p.height
    out = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.reps / len(scns)

    base = timed(lambda: [r.render(s, shade, out.data_ptr()) for s in scns])
    print(f"plain full frame: {base:.4f} ms/frame")
    for world in [int(x) for x in a.worlds.split(",")]:
        for ts in [int(x) for x in a.tiles.split(",")]:
            per = []
            for rank in range(world):
                per.append(timed(lambda: [r.render_tiles_direct(s, shade, out.data_ptr(), ts, rank, world) for s in scns]))
            print(f"world {world} tile {ts}: max {max(per):.4f} min {min(per):.4f} sum {sum(per):.4f} ms/frame; ideal {base / world:.4f}; "
                  f"efficiency {base / world / max(per):.3f}")
    for world in [int(x) for x in a.worlds.split(",")]:
        for rows in [int(x) for x in a.bands.split(",") if x]:
            nb = (h + rows - 1) // rows
            mine = (nb + world - 1) // world
            packed = torch.zeros((mine * rows, w, 4), dtype=torch.uint8, device="cuda")
            for K in (1, 4):
                streams = [torch.cuda.Stream() for _ in range(K)]
                per = []
                for rank in range(world):
                    def fn():
                        for j, s in enumerate(scns):
                            if K > 1:
                                r.set_stream(streams[j % K].cuda_stream)
                            r.render_bands(s, shade, packed.data_ptr(), rows, rank, world)
                        if K > 1:
                            r.set_stream(None)
                            for st in streams:
                                torch.cuda.current_stream().wait_stream(st)
                    per.append(timed(fn))
                print(f"world {world} bands of {rows} rows, {K} stream(s): max {max(per):.4f} min {min(per):.4f} ms/frame; ideal {base / world:.4f}; "
                      f"efficiency {base / world / max(per):.3f}")
    # K streams alternating frames (tails of frame j overlap the head of frame j+1)
    for K in (2, 3, 4):
        streams = [torch.cuda.Stream() for _ in range(K)]
        outs = [torch.zeros_like(out) for _ in range(K)]

        def lanes(world, rank=0, ts=32):
            def fn():
                for j, s in enumerate(scns):
                    r.set_stream(streams[j % K].cuda_stream)
                    if world == 1:
                        r.render(s, shade, outs[j % K].data_ptr())
                    else:
                        r.render_tiles_direct(s, shade, outs[j % K].data_ptr(), ts, rank, world)
                r.set_stream(None)
                for st in streams:
                    torch.cuda.current_stream().wait_stream(st)
            return fn
        line = [f"{K} streams:"]
        for world in [int(x) for x in a.worlds.split(",")]:
            t = timed(lanes(world))
            line.append(f"world {world}: {t:.4f} ms/frame (eff {base / world / t:.3f})")
        print("  ".join(line))


if __name__ == "__main__":
    main()
