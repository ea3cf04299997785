"""sampler_ab.py — Gsamples/s of the four ways of reading a brick (csrc/gvdbx_microbench.cuh) on a BASELINE workload's atlas.
  python tests/sampler_ab.py [cfg4] [spacing ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import bench
    pkg = bench.load_pkg()
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
    spacings = [float(x) for x in sys.argv[2:]] or [0.2, 0.86]
    p, vol = bench.build_workload(wl)
    r = pkg.Renderer(0)
    r.import_topology_host(vol["vdbinfo"], vol["pool0"], vol["pool1"])
    r.import_atlas_host(vol["atlas"])
    _, table = bench.frame_scninfos(pkg, p, 7, 1)
    r.set_transfer(table)
    for s in spacings:
        print(json.dumps({"workload": wl, "lane_spacing_voxels": s, **{k: round(v, 1) for k, v in r.measure_sampler_ab(s).items()},
                          "deep_sample_loop": round(r.measure_deep_loop_peak(s), 1),
                          "deep_sample_loop_table_through_texture": round(r.measure_deep_loop_peak(s, True), 1)}), flush=True)
    r.close()


if __name__ == "__main__":
    main()
