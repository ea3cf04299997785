"""make_golden_ref.py — generates tests/golden/ref_<preset>.npz on the GPU box by running the UNMODIFIED reference
(oracle/_ref/ref_harness) and condensing its dumps: VDBInfo, ScnInfo per mode, RGBA per mode, hit/normal buffers
(tiny presets), SHA-256 of every pool and of the atlas after the reference's own UpdateApron.

  python tests/make_golden_ref.py gpurun_out/golden          (then copy *.npz into tests/golden/ and commit)
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refcmp  # noqa: E402

PRESETS = ["cfg1_tiny", "cfg2_tiny", "cfg3_tiny", "cfg4_tiny", "cfg1_small", "cfg2_small", "cfg3_small", "cfg4_small"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    for preset in PRESETS:
        d = tempfile.mkdtemp(prefix="refdump_")
        timing = refcmp.run_ref(preset, d, modes=list(refcmp.MODES))
        dump = refcmp.load_dump(d)
        out = {"vdbinfo": np.frombuffer(dump["vdbinfo"], np.uint8), "bricks": dump["meta"]["bricks"],
               "atlas_res": np.array(dump["meta"]["atlas_res"]), "atlas_sha": sha(dump["atlas"]),
               "transfer_sha": sha(dump["transfer"]), "width": dump["meta"]["width"], "height": dump["meta"]["height"],
               "ms_per_frame": np.array([timing["render"][m]["ms_median"] for m in refcmp.MODES])}
        for lev, b in dump["pool0"].items():
            out[f"pool0_L{lev}_sha"] = sha(b)
            out[f"pool0_L{lev}_bytes"] = len(b)
        for lev, b in dump["pool1"].items():
            out[f"pool1_L{lev}_sha"] = sha(b)
            out[f"pool1_L{lev}_bytes"] = len(b)
        for m in refcmp.MODES:
            out[f"scn_{m}"] = np.frombuffer(dump["scn"][m], np.uint8)
            out[f"rgba_{m}"] = dump["rgba"][m]
            if "tiny" in preset and m in dump["hit"]:
                out[f"hit_{m}"] = dump["hit"][m]
        np.savez_compressed(os.path.join(outdir, f"ref_{preset}.npz"), **out)
        print("[golden]", preset, "bricks", out["bricks"], "atlas", out["atlas_res"], flush=True)


def extras(outdir):
    """golden cases beyond the plain presets: SetTransform, depth-buffer compositing, explicit ray bundles"""
    xf = (3.0, -2.0, 1.5, 1.25, 0.8, 1.1, 20.0, -35.0, 10.0, 5.0, 7.0, -4.0)
    for tag, preset, kw in (("xform", "cfg1_small", {"xform": xf}), ("dbuf", "cfg4_small", {"dbuf": True}),
                            ("rays", "cfg1_small", {"raytrace": 20000})):
        d = tempfile.mkdtemp(prefix="refdump_")
        refcmp.run_ref(preset, d, modes=list(refcmp.MODES), hits=(tag != "rays"), **kw)
        dump = refcmp.load_dump(d)
        out = {"preset": preset, "width": dump["meta"]["width"], "height": dump["meta"]["height"],
               "vdbinfo": np.frombuffer(dump["vdbinfo"], np.uint8)}
        if tag == "xform":
            out["xform"] = np.array(xf, np.float32)
        for m in refcmp.MODES:
            out[f"scn_{m}"] = np.frombuffer(dump["scn"][m], np.uint8)
            out[f"rgba_{m}"] = dump["rgba"][m]
        if tag == "dbuf":
            out["dbuf"] = dump["dbuf"]
        if tag == "rays":
            out["rays_in"] = dump["rays_in"]
            out["rays_out"] = dump["rays_out"]
            out["scn_raytrace"] = np.frombuffer(dump["scn_raytrace"], np.uint8)
            for m in refcmp.MODES:
                del out[f"rgba_{m}"]
        np.savez_compressed(os.path.join(outdir, f"ref_extra_{tag}.npz"), **out)
        print("[golden] extra", tag, preset, flush=True)


def modes2(outdir):
    """the rest of Render()'s switch (tricubic, empty skip, sections) and the two composed BASELINE modes (deep + shadow,
    deep at 4 rays per pixel): ScnInfo + RGBA per mode, hit / normal buffers for the tiny presets"""
    for preset in ("cfg1_tiny", "cfg4_tiny", "cfg1_small", "cfg4_small"):
        d = tempfile.mkdtemp(prefix="refdump_")
        refcmp.run_ref(preset, d, modes=["deep", refcmp.CUSTOM] + list(refcmp.MODES2))
        dump = refcmp.load_dump(d)
        out = {"preset": preset, "width": dump["meta"]["width"], "height": dump["meta"]["height"], "spp": 4,
               "vdbinfo": np.frombuffer(dump["vdbinfo"], np.uint8)}
        for m in list(refcmp.MODES2) + [refcmp.CUSTOM]:
            out[f"scn_{m}"] = np.frombuffer(dump["scn"][m], np.uint8)
            out[f"rgba_{m}"] = dump["rgba"][m]
            if "tiny" in preset and m in dump["hit"]:
                out[f"hit_{m}"] = dump["hit"][m]
        np.savez_compressed(os.path.join(outdir, f"ref_modes2_{preset}.npz"), **out)
        print("[golden] modes2", preset, flush=True)


def vbx(outdir):
    """a VBX file written by the reference's own SaveVBX (with a grid transform), + the images it renders from it"""
    xf = (3.0, -2.0, 1.5, 1.25, 0.8, 1.1, 20.0, -35.0, 10.0, 5.0, 7.0, -4.0)
    preset = "cfg1_tiny"
    d = tempfile.mkdtemp(prefix="refdump_")
    path = os.path.join(d, "ref.vbx")
    refcmp.run_ref(preset, d, modes=list(refcmp.MODES), xform=xf, savevbx=path)
    dump = refcmp.load_dump(d)
    out = {"preset": preset, "width": dump["meta"]["width"], "height": dump["meta"]["height"], "xform": np.array(xf, np.float32),
           "vbx": np.fromfile(path, dtype=np.uint8), "vdbinfo": np.frombuffer(dump["vdbinfo"], np.uint8)}
    for m in refcmp.MODES:
        out[f"scn_{m}"] = np.frombuffer(dump["scn"][m], np.uint8)
        out[f"rgba_{m}"] = dump["rgba"][m]
    np.savez_compressed(os.path.join(outdir, "ref_vbx_cfg1_tiny.npz"), **out)
    print("[golden] vbx", preset, out["vbx"].size, "bytes", flush=True)


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    if "--vbx-only" in sys.argv:
        vbx(out)
        sys.exit(0)
    if "--modes2-only" in sys.argv:
        modes2(out)
        sys.exit(0)
    if "--extras-only" not in sys.argv:
        main(out)
    extras(out)
    modes2(out)
    vbx(out)
