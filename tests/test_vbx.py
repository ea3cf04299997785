"""VBX files (GVDB_FILESPEC.txt): the product's LoadVBX / SaveVBX against a file written by the UNMODIFIED reference's own
SaveVBX (tests/golden/ref_vbx_cfg1_tiny.npz, made by tests/make_golden_ref.py --vbx-only on the GPU box)."""
import os

import numpy as np
import pytest

from common import GOLDEN, MODES, mask_vdbinfo

GRID_NAME = slice(62, 62 + 256)      # the reference writes this field uninitialised (gvdb_volume_gvdb.cpp:1644, :1693)


def _golden():
    return np.load(os.path.join(GOLDEN, "ref_vbx_cfg1_tiny.npz"))


def test_vbx_header_parse_reproduces_reference_vdbinfo(pkg, tmp_path):
    """host-only: parsing the reference's file yields the VDBInfo block the reference itself built for that volume
    (FinishTopology + ComputeBounds + PrepareVDB restated in the product's host mirror), and the stored transform."""
    g = _golden()
    path = tmp_path / "ref.vbx"
    g["vbx"].tofile(path)
    v = pkg.Volume(-1)
    v.LoadVBX(path, parse_only=True)
    mine, ref = mask_vdbinfo(v.vdbinfo()), mask_vdbinfo(g["vdbinfo"].tobytes())
    assert np.array_equal(mine, ref), np.nonzero(mine != ref)[0][:20]
    # the transform stored in the file went through SetTransform: ScnInfo carries the same matrices as the reference's
    p = pkg.Volume(-1)
    xf = g["xform"]
    p.SetTransform(xf[0:3], xf[3:6], xf[6:9], xf[9:12])
    a = np.frombuffer(v.PrepareRender(8, 8, 0), np.uint8)[128:320]
    b = np.frombuffer(p.PrepareRender(8, 8, 0), np.uint8)[128:320]
    r = g["scn_voxel"][128:320]
    assert np.array_equal(a, b) and np.array_equal(a, r)
    v.close(); p.close()


def test_vbx_bad_files_are_errors_not_crashes(pkg, tmp_path):
    g = _golden()
    v = pkg.Volume(-1)
    with pytest.raises(pkg.GvdbxError):
        v.LoadVBX(tmp_path / "missing.vbx", parse_only=True)
    for n in (1, 40, 500, 5000, g["vbx"].size - 7):
        path = tmp_path / f"cut{n}.vbx"
        g["vbx"][:n].tofile(path)
        with pytest.raises(pkg.GvdbxError):
            v.LoadVBX(path, parse_only=True)
    bad = g["vbx"].copy()
    bad[0] = 1; bad[1] = 0                                     # GVDB 1.0: bitmask child lists
    bad.tofile(tmp_path / "v10.vbx")
    with pytest.raises(pkg.GvdbxError):
        v.LoadVBX(tmp_path / "v10.vbx", parse_only=True)
    v.close()


@pytest.mark.gpu
def test_vbx_load_render_save_roundtrip(pkg, ora, tmp_path):
    """LoadVBX of the reference's file -> all four core modes bit-exact against the reference's render of that volume
    (grid transform from the file); SaveVBX writes the reference's bytes back (except the uninitialised name field)."""
    g = _golden()
    path = tmp_path / "ref.vbx"
    g["vbx"].tofile(path)
    p = ora.preset(str(g["preset"]))
    v = pkg.Volume(0)
    v.LoadVBX(path)
    v.SetSceneParams(list(p.steps), list(p.extinct), list(p.thresh), list(p.cutoff), list(p.backclr), list(p.shadow))
    if p.transfer == 1:
        v.LinearTransferFunc(0.00, 0.25, (0, 0, 0, 0), (1, 1, 0, 0.1))
        v.LinearTransferFunc(0.25, 0.50, (1, 1, 0, 0.4), (1, 0, 0, 0.3))
        v.LinearTransferFunc(0.50, 0.75, (1, 0, 0, 0.3), (.2, .2, 0.2, 0.1))
        v.LinearTransferFunc(0.75, 1.00, (.2, .2, 0.2, 0.1), (0, 0, 0, 0.0))
    v.CommitTransferFunc()
    v.SetCamera(p.fov, list(p.cam_angs), list(p.cam_target), p.cam_dist)
    v.SetLight(list(p.light_angs), list(p.light_target), p.light_dist)
    w, h = int(g["width"]), int(g["height"])
    v.AddRenderBuf(0, w, h, 4)
    for m, sh in MODES.items():
        v.Render(sh, 0, 0)
        img = v.ReadRenderBuf(0)
        assert np.array_equal(img, g[f"rgba_{m}"]), (m, int((img != g[f"rgba_{m}"]).any(axis=2).sum()))
    assert (g["rgba_trilinear"] != g["rgba_trilinear"][0, 0]).any()
    out = tmp_path / "mine.vbx"
    v.SaveVBX(out)
    mine, ref = np.fromfile(out, dtype=np.uint8), g["vbx"].copy()
    assert mine.size == ref.size
    mine[GRID_NAME] = 0; ref[GRID_NAME] = 0
    assert np.array_equal(mine, ref), np.nonzero(mine != ref)[0][:20]
    # and the file we wrote loads again
    v2 = pkg.Volume(-1)
    v2.LoadVBX(out, parse_only=True)
    assert np.array_equal(mask_vdbinfo(v2.vdbinfo()), mask_vdbinfo(g["vdbinfo"].tobytes()))
    v.close(); v2.close()
