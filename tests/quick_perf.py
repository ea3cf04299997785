"""quick_perf.py — ms/frame of every BASELINE workload on one GPU (1 stream and 4 frame lanes), no parity / reference legs:
the inner loop of kernel work.   python tests/quick_perf.py [cfg4:deepshadow cfg3:voxel ...] [--option id=value ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import bench
    pkg = bench.load_pkg()
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    opts = [a.split("=") for a in sys.argv[1:] if a.startswith("--") and "=" in a]
    work = args or ["cfg1:trilinear", "cfg2:levelset", "cfg3:voxel", "cfg4:deep", "cfg4:deepshadow"]
    out = {}
    vols = {}
    for wm in work:
        wl, mode = wm.split(":")
        shade, dshadow = bench.MODE[mode]
        if wl not in vols:
            vols = {wl: bench.build_workload(wl)}
        p, vol = vols[wl]
        scns, table = bench.frame_scninfos(pkg, p, shade, 8)
        r = pkg.Renderer(0)
        r.import_topology_host(vol["vdbinfo"], vol["pool0"], vol["pool1"])
        r.import_atlas_host(vol["atlas"])
        r.set_transfer(table)
        r.set_deep_shadow(dshadow)
        for k, v in opts:
            r.set_option(int(k.lstrip("-")), int(v))
        w, h = p.width, p.height
        frames_d = [torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda") for _ in range(4)]
        lat = bench.time_latency(torch, r, scns, shade, frames_d[0], reps=2)
        r.lanes(4)
        ms, _ = bench.time_resident(torch, r, scns, shade, frames_d, 4, 3, 2)
        r.lanes(0)
        out[wm] = {"latency_ms_1lane": round(lat["mean"], 4), "ms_per_frame_4lanes": round(ms / 3 / 8, 4),
                   "mrays_1lane": round(w * h / lat["mean"] / 1e3, 1), "mrays_4lanes": round(w * h / (ms / 3 / 8) / 1e3, 1)}
        print(wm, json.dumps(out[wm]), flush=True)
        r.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
