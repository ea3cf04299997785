"""Pins the product's host mirror (gvdb-voxels_b200/csrc/gvdbx_host.cpp: Camera3D, Matrix4, Scene, PrepareRender)
against the reference's own host code: (a) Camera3D/Matrix4F compiled from /root/reference (tests/golden/hoststate.json
via oracle/_ref/ref_hostdump), (b) the ScnInfo bytes the real libgvdb.so produced on the GPU box (ref_*.npz)."""
import json
import os

import numpy as np
import pytest

from common import GOLDEN, MODES, SMALL, TINY, golden, mask_scninfo


def _bits(scn, off, n):
    return ["%08x" % v for v in np.frombuffer(scn[off:off + 4 * n], np.uint32)]


def test_camera_corner_rays_bit_exact(pkg):
    gold = json.load(open(os.path.join(GOLDEN, "hoststate.json")))
    for case in gold["cam"]:
        fov, w, h, ax, ay, az, tx, ty, tz, dist = case["in"]
        v = pkg.Volume(-1)
        v.SetCamera(fov, (ax, ay, az), (tx, ty, tz), dist)
        v.SetRes(int(w), int(h))
        scn = v.PrepareRender(int(w), int(h), 0)
        v.close()
        f = np.frombuffer(scn, np.float32)
        campos, cams, camu, camv = f[4:7], f[7:10], f[10:13], f[13:16]
        exp = np.array([int(x, 16) for x in case["out"]], np.uint32).view(np.float32)
        assert np.array_equal(campos.view(np.uint32), exp[0:3].view(np.uint32)), case["in"]
        assert np.array_equal(cams.view(np.uint32), exp[3:6].view(np.uint32)), case["in"]
        assert np.array_equal(camu.view(np.uint32), (exp[6:9] - exp[3:6]).view(np.uint32)), case["in"]
        assert np.array_equal(camv.view(np.uint32), (exp[9:12] - exp[3:6]).view(np.uint32)), case["in"]


def test_set_transform_matrices_bit_exact(pkg):
    gold = json.load(open(os.path.join(GOLDEN, "hoststate.json")))
    for case in gold["xfm"]:
        a = case["in"]
        v = pkg.Volume(-1)
        v.SetTransform(a[0:3], a[3:6], a[6:9], a[9:12])
        scn = v.PrepareRender(64, 64, 0)
        v.close()
        assert _bits(scn, 128, 48) == case["out"], a


@pytest.mark.parametrize("preset", TINY + SMALL)
def test_scninfo_bytes_equal_reference_library(ora, pkg, preset):
    g = golden(preset)
    p = ora.preset(preset)
    for m, sh in MODES.items():
        scn, _ = ora.scninfo_for(pkg, p, shade=sh)
        assert np.array_equal(mask_scninfo(scn), mask_scninfo(g[f"scn_{m}"].tobytes())), (preset, m)


def test_default_and_rendertofile_transfer_tables(ora, pkg):
    from common import sha
    for preset in ("cfg1_tiny", "cfg4_tiny"):
        _, table = ora.scninfo_for(pkg, ora.preset(preset))
        assert sha(table) == str(golden(preset)["transfer_sha"])


def test_scninfo_with_set_transform_equals_reference(ora, pkg):
    """PrepareRender after SetTransform: xform / invxform / invxrot and the light position transformed by mInvXform."""
    g = np.load(os.path.join(GOLDEN, "ref_extra_xform.npz"))
    p = ora.preset(str(g["preset"]))
    xf = [float(v) for v in g["xform"]]
    v = pkg.Volume(-1)
    v.SetTransform(xf[0:3], xf[3:6], xf[6:9], xf[9:12])
    v.SetSceneParams(list(p.steps), list(p.extinct), list(p.thresh), list(p.cutoff), list(p.backclr), list(p.shadow))
    v.SetCamera(p.fov, list(p.cam_angs), list(p.cam_target), p.cam_dist)
    v.SetRes(p.width, p.height)
    v.SetLight(list(p.light_angs), list(p.light_target), p.light_dist)
    for m, sh in MODES.items():
        scn = v.PrepareRender(p.width, p.height, sh)
        assert np.array_equal(mask_scninfo(scn), mask_scninfo(g[f"scn_{m}"].tobytes())), m
    v.close()
