"""prof_frame.py — renders a few frames of one workload so that ncu can capture the render kernel.
  ncu --set full --clock-control none --import-source on -k regex:gx_render -s 2 -c 2 -o gpurun_out/prof \
      python tests/prof_frame.py --workload cfg2 --sampler tex --frames 4
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
    ap.add_argument("--sampler", default="tex")
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--block", default="8x8")
    ap.add_argument("--size", default="")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--option", action="append", default=[], help="gvdbx option id=value (repeatable)")
    a = ap.parse_args()
    pkg = bench.load_pkg()
    size = tuple(int(x) for x in a.size.split("x")) if a.size else None
    p, vol = bench.build_workload(a.workload, size)
    mode = a.mode or {0: "voxel", 4: "trilinear", 6: "levelset", 7: "deep"}[p.shade]
    shade, dshadow = bench.MODE[mode]
    scns, table = bench.frame_scninfos(pkg, p, shade, a.frames)
    r = pkg.Renderer(0)
    r.import_topology_host(vol["vdbinfo"], vol["pool0"], vol["pool1"])
    r.import_atlas_host(vol["atlas"])
    r.set_transfer(table)
    r.set_sampler(0 if a.sampler == "tex" else 1)
    bw, bh = (int(x) for x in a.block.split("x"))
    r.set_block(bw, bh)
    r.set_deep_shadow(dshadow)
    for o in a.option:
        k, v = o.split("=")
        r.set_option(int(k), int(v))
    out = torch.zeros((p.height, p.width, 4), dtype=torch.uint8, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.frames + 1)]
    ev[0].record()
    for i, scn in enumerate(scns):
        r.render(scn, shade, out.data_ptr())
        ev[i + 1].record()
    torch.cuda.synchronize()
    print("ms/frame:", [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(a.frames)])


if __name__ == "__main__":
    main()
