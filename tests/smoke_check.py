"""smoke_check.py — one small invocation of the hot path on cuda:0 checked against the CPU oracle and the reference
golden (called by __graft_entry__.smoke())."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run():
    import torch
    import oracle
    from __graft_entry__ import load_package
    from common import golden, psnr
    pkg = load_package()
    assert torch.cuda.is_available(), "smoke() needs cuda:0"
    name = "cfg3_tiny"
    p, vol = oracle.scene_volume(name)
    g = golden(name)
    r = pkg.Renderer(0)
    r.import_topology_host(vol["vdbinfo"], vol["pool0"], vol["pool1"])
    r.import_atlas_host(vol["atlas"])
    _, table = oracle.scninfo_for(pkg, p)
    r.set_transfer(table)
    vol["transfer"] = table
    out = torch.zeros((p.height, p.width, 4), dtype=torch.uint8, device="cuda")
    for mode, shade in (("voxel", 0), ("trilinear", 4), ("levelset", 6), ("deep", 7)):
        scn = g[f"scn_{mode}"].tobytes()
        for sampler in (0, 1):
            r.set_sampler(sampler)
            r.render(scn, shade, out.data_ptr())
            r.sync()
            img = out.cpu().numpy()
            cpu = oracle.render(vol, scn, shade)
            assert psnr(img, cpu) >= 60.0, (mode, sampler, psnr(img, cpu))
            if sampler == 0 or mode == "voxel":
                assert np.array_equal(img, g[f"rgba_{mode}"]), (mode, sampler)
    r.close()
    print("smoke ok: 4 modes x 2 samplers on cuda:0 match the CPU oracle; texture path bit-exact vs reference golden")


if __name__ == "__main__":
    run()
