"""Pins the CPU oracle (oracle/gvdb_oracle.c) against outputs of the UNMODIFIED reference run on a B200
(tests/golden/ref_*.npz, produced by tests/make_golden_ref.py through oracle/_ref/ref_harness)."""
import numpy as np
import pytest

from common import MODES, SMALL, TINY, golden, mask_vdbinfo, psnr, sha


@pytest.mark.parametrize("preset", TINY + SMALL)
def test_topology_pools_atlas_byte_identical(ora, preset):
    """Configure/ActivateSpace/FinishTopology/UpdateAtlas + UpdateApron: pools, atlas and VDBInfo match the
    reference library byte for byte."""
    g = golden(preset)
    p, vol = ora.scene_volume(preset)
    assert vol["meta"]["bricks"] == int(g["bricks"])
    assert tuple(vol["meta"]["atlas_res"]) == tuple(int(x) for x in g["atlas_res"])
    for grp, pools in ((0, vol["pool0"]), (1, vol["pool1"])):
        for lev in range(5):
            key = f"pool{grp}_L{lev}_sha"
            if key in g.files:
                assert len(pools[lev]) == int(g[f"pool{grp}_L{lev}_bytes"]), key
                assert sha(pools[lev]) == str(g[key]), key
    assert sha(vol["atlas"]) == str(g["atlas_sha"])
    assert np.array_equal(mask_vdbinfo(vol["vdbinfo"]), mask_vdbinfo(g["vdbinfo"].tobytes()))


@pytest.mark.parametrize("preset", TINY + ["cfg3_small"])
@pytest.mark.parametrize("mode", list(MODES))
def test_cpu_raycaster_within_tolerance_of_reference(ora, pkg, preset, mode):
    """The CPU restatement cannot reproduce MUFU approximations bit for bit; it must stay within the north-star
    tolerance of the reference's image (PSNR >= 60 dB) and agree exactly on almost every pixel."""
    g = golden(preset)
    p, vol = ora.scene_volume(preset)
    scn, table = ora.scninfo_for(pkg, p, shade=MODES[mode])
    assert sha(table) == str(g["transfer_sha"])
    vol["transfer"] = table
    img = ora.render(vol, g[f"scn_{mode}"].tobytes(), MODES[mode])
    ref = g[f"rgba_{mode}"]
    d = np.abs(img.astype(int) - ref.astype(int)).max(axis=2)
    assert psnr(img, ref) >= 60.0, (psnr(img, ref), int((d > 0).sum()))
    assert (d > 1).mean() < 2e-3


@pytest.mark.parametrize("preset", ["cfg1_tiny", "cfg4_tiny", "cfg1_small"])
@pytest.mark.parametrize("mode", ["tricubic", "emptyskip", "section2d", "section3d", "deepshadow", "deepspp"])
def test_cpu_remaining_modes_within_tolerance_of_reference(ora, pkg, preset, mode):
    """The rest of Render()'s switch (tricubic, empty skip, 2-D / 3-D sections incl. the getNodeAtPoint descent) and the two
    composed BASELINE modes (deep + rayShadowBrick march, 4 rays per pixel) restated on the CPU, against the reference's
    images: PSNR >= 60 dB, at most 0.2 % of the pixels off by more than 1/255."""
    import os
    import refcmp
    from common import GOLDEN
    g = np.load(os.path.join(GOLDEN, f"ref_modes2_{preset}.npz"))
    p, vol = ora.scene_volume(preset)
    _, table = ora.scninfo_for(pkg, p)
    vol["transfer"] = table
    shade, dshadow, spp = refcmp.MODES2[mode]
    img = ora.render(vol, g[f"scn_{mode}"].tobytes(), shade, deep_shadow=bool(dshadow), spp=spp)
    ref = g[f"rgba_{mode}"]
    d = np.abs(img.astype(int) - ref.astype(int)).max(axis=2)
    assert psnr(img, ref) >= 60.0, (psnr(img, ref), int((d > 0).sum()))
    assert (d > 1).mean() < 2e-3
    assert (ref != ref[0, 0]).any()


def test_cpu_hit_points_close_to_reference(ora, pkg):
    g = golden("cfg3_tiny")
    p, vol = ora.scene_volume("cfg3_tiny")
    img, hn = ora.render(vol, g["scn_voxel"].tobytes(), MODES["voxel"], want_hits=True)
    ref = g["hit_voxel"]
    hit_ref = ref[:, :, 2] != np.float32(1e10)
    hit_me = hn[:, :, 2] != np.float32(1e10)
    assert (hit_ref != hit_me).mean() < 1e-3
    both = hit_ref & hit_me
    assert np.abs(hn[both][:, 0:3] - ref[both][:, 0:3]).max() < 1e-2
    assert np.array_equal(hn[both][:, 4:7], ref[both][:, 4:7])       # cube-face normals are exact integers


def test_texture_model_reproduces_integer_texels(ora):
    p, vol = ora.scene_volume("cfg1_tiny")
    a = vol["atlas"]
    pts = [(x + 0.5, y + 0.5, z + 0.5) for x, y, z in [(3, 4, 5), (11, 12, 3), (25, 31, 7)]]
    got = ora.tex3d(vol, pts)
    want = np.array([a[5, 4, 3], a[3, 12, 11], a[7, 31, 25]], np.float32)
    assert np.array_equal(got, want)
