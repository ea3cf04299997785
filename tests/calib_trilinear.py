"""calib_trilinear.py — identifies the texture unit's trilinear filter so that the software model
(GxSampler<LINEAR>, gvdbx_device.cuh) can be fitted / verified offline.

Atlases holding the multilinear monomials x, y, z, xy, xz, yz, xyz (texel-centre coordinates) are sampled with
the hardware path at the same points: trilinear filtering reproduces those functions exactly, so the outputs give
the effective per-axis weights and their products directly.  A random atlas is sampled as an end-to-end check.

  python tests/calib_trilinear.py gpurun_out/calib.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(out):
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    import torch
    pkg = load_package()
    rng = np.random.default_rng(1234)
    R = 40
    zz, yy, xx = np.meshgrid(np.arange(R), np.arange(R), np.arange(R), indexing="ij")
    atlases = {
        "rand": rng.standard_normal((R, R, R)).astype(np.float32),
        "x": xx.astype(np.float32), "y": yy.astype(np.float32), "z": zz.astype(np.float32),
        "xy": (xx * yy).astype(np.float32), "xz": (xx * zz).astype(np.float32), "yz": (yy * zz).astype(np.float32),
        "xyz": (xx * yy * zz).astype(np.float32),
    }
    n = 1 << 18
    slot = rng.integers(0, 4, size=(n, 3))
    loc = rng.uniform(0.5, 9.5, size=(n, 3))
    k = n // 2
    loc[:k] = 0.5 + rng.integers(0, 9 * 512, size=(k, 3)) / 512.0        # half of the points on a 1/512 grid (ties)
    loc[: k // 4, 1:] = np.floor(loc[: k // 4, 1:]) + 0.5                  # 1-D cases: y, z on texel centres
    xyz = (slot * 10 + loc).astype(np.float32)
    d_xyz = torch.from_numpy(xyz).cuda()
    res = {"xyz": xyz}
    r = pkg.Renderer(0)
    for name, a in atlases.items():
        r.import_atlas_host(a)
        d_tex = torch.zeros(n, dtype=torch.float32, device="cuda")
        d_lin = torch.zeros(n, dtype=torch.float32, device="cuda")
        r.sample_points(d_xyz.data_ptr(), n, d_tex.data_ptr(), d_lin.data_ptr())
        r.sync()
        tex, lin = d_tex.cpu().numpy(), d_lin.cpu().numpy()
        res["tex_" + name] = tex
        res["lin_" + name] = lin
        d = np.abs(tex - lin)
        print(f"[calib] {name:5s} max|tex-lin|={d.max():.3e} mean={d.mean():.3e} "
              f"bit-exact={np.mean(tex.view(np.uint32) == lin.view(np.uint32)):.4f}", flush=True)
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    np.savez_compressed(out, **res)
    r.close()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "calib.npz"))
