"""calib_trilinear.py — dumps hardware-filtered texture samples of a known random atlas so that the software
model of the texture unit (GxSampler<LINEAR>, gvdbx_device.cuh) can be fitted / verified offline.

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
    atlas = rng.standard_normal((40, 40, 40)).astype(np.float32)          # 4x4x4 brick slots
    r = pkg.Renderer(0)
    r.import_atlas_host(atlas)
    n = 1 << 20
    # points well inside bricks' filter footprint: brick-local [0.5, 9.5)
    slot = rng.integers(0, 4, size=(n, 3))
    loc = rng.uniform(0.5, 9.5, size=(n, 3))
    # a share of points on exact 1/256 and 1/512 grids to probe rounding ties
    k = n // 4
    loc[:k] = 0.5 + rng.integers(0, 9 * 512, size=(k, 3)) / 512.0
    xyz = (slot * 10 + loc).astype(np.float32)
    d_xyz = torch.from_numpy(xyz).cuda()
    d_tex = torch.zeros(n, dtype=torch.float32, device="cuda")
    d_lin = torch.zeros(n, dtype=torch.float32, device="cuda")
    r.sample_points(d_xyz.data_ptr(), n, d_tex.data_ptr(), d_lin.data_ptr())
    r.sync()
    tex, lin = d_tex.cpu().numpy(), d_lin.cpu().numpy()
    d = np.abs(tex - lin)
    print(f"[calib] n={n} max|tex-lin|={d.max():.3e} mean={d.mean():.3e} bit-exact={np.mean(tex.view(np.uint32) == lin.view(np.uint32)):.4f}")
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    m = 1 << 18
    np.savez_compressed(out, xyz=xyz[:m], tex=tex[:m], lin=lin[:m], xyz_grid=xyz[:k][:m], seed=1234)
    r.close()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "calib.npz"))
