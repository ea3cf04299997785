"""strict_ab.py — the strict drop-in sequence (Render + synchronous ReadRenderBuf into pageable memory) with the frame rendered
in 1 (= one launch, copy in series), automatic, 2, 4, 8, 16 bands (gvdbx_render_banded / gvdbx_read_banded).
  python tests/strict_ab.py cfg3:voxel cfg4:deepshadow ..."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import bench

    class A:
        sampler, block, traversal, spp, frames = "tex", "8x8", "default", 1, 8
    pkg = bench.load_pkg()
    for wm in sys.argv[1:] or ["cfg1:trilinear", "cfg2:levelset", "cfg3:voxel", "cfg4:deepshadow"]:
        wl, mode = wm.split(":")
        shade, dshadow = bench.MODE[mode]
        p, vol = bench.build_workload(wl)
        _, table = bench.frame_scninfos(pkg, p, shade, 8)
        vol["transfer"] = table
        row = {}
        for nb in (1, 0, 2, 4, 8, 16):
            e = bench.time_e2e(torch, pkg, p, vol, 0, A, shade, dshadow, 8, 3, 0, bands=nb)
            row["auto" if nb == 0 else str(nb)] = round(e["ms_per_frame"], 4)
        print(wm, "ms per frame by bands:", json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
