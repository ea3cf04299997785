"""refcmp.py — parity harness: UNMODIFIED reference (oracle/_ref) vs libgvdbx on identical bytes.

Test infrastructure (may execute oracle/): runs oracle/_ref/ref_harness as a subprocess (it has its own CUDA
context), loads its dumps (the reference's own VDBInfo / ScnInfo / pools / atlas-after-UpdateApron / transfer
table) and renders the same inputs through the C ABI of libgvdbx.so.  Nothing here reads /root/reference.

CLI:  python tests/refcmp.py --presets cfg1_small,cfg3_small --out gpurun_out/refcmp.json
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
MODES = {"voxel": 0, "trilinear": 4, "levelset": 6, "deep": 7}
# the rest of Render()'s switch + the two composed BASELINE modes: name -> (shade, deep_shadow option, spp option)
MODES2 = {"tricubic": (5, 0, 1), "emptyskip": (3, 0, 1), "section2d": (1, 0, 1), "section3d": (2, 0, 1),
          "deepshadow": (7, 1, 1), "deepspp": (7, 0, 4)}
CUSTOM = "custom"          # the reference's gRenderKernel sample kernel (user kernel through RenderKernel)
VOXELID = "voxelid"        # SHADE_VOXEL through the instrumented brick function of oracle_kernels.cu: voxel int3, depth t, leaf id
ALL_SHADE = dict(MODES, **{k: v[0] for k, v in MODES2.items()})
ALL_SHADE[VOXELID] = 0


def have_ref():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in
               ("ref_harness", "libgvdb.so", "cuda_gvdb_module.cubin", "cuda_gvdb_copydata.ptx"))


def have_ref_x():
    return have_ref() and os.path.exists(os.path.join(REF_DIR, "ref_harness_x"))


def run_ref(preset, outdir, modes=None, size=None, frames=1, warmup=0, nodump=False, shadow=None, hits=True, timeout=1800,
            xform=None, dbuf=False, raytrace=0, savevbx=None, config=None, module=None, color=False, gvdbx=False, lightdump=False,
            extra=None):
    """Run the reference harness; returns its timing dict.  gvdbx=True: ref_harness_x, the same harness with the product's
    Level-A shim (include/gvdbx_shim.h) linked in — every mode is rendered by Render() AND RenderX() in one process."""
    cmd = ["./ref_harness_x" if gvdbx else "./ref_harness", preset, os.path.abspath(outdir)]
    if modes:
        cmd += ["--modes", ",".join(modes)]
    if size:
        cmd += ["--size", f"{size[0]}x{size[1]}"]
    cmd += ["--frames", str(frames), "--warmup", str(warmup), "--hits", "1" if hits else "0"]
    if nodump:
        cmd.append("--nodump")
    if shadow is not None:
        cmd += ["--shadow", str(int(shadow))]
    if xform is not None:
        cmd += ["--xform", ",".join(repr(float(v)) for v in xform)]
    if dbuf:
        cmd.append("--dbuf")
    if raytrace:
        cmd += ["--raytrace", str(int(raytrace))]
    if savevbx:
        cmd += ["--savevbx", os.path.abspath(savevbx)]
    if config:
        cmd += ["--config", ",".join(str(int(c)) for c in config)]
    if module:
        cmd += ["--module", os.path.abspath(module)]
    if color:
        cmd.append("--color")
    if gvdbx:
        cmd.append("--gvdbx")
    if lightdump:
        cmd.append("--lightdump")
    if extra:
        cmd += list(extra)
    r = subprocess.run(cmd, cwd=REF_DIR, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"ref_harness failed ({r.returncode}): {r.stderr[-2000:]}")
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    return json.loads(line)


def load_dump(d):
    meta = {"pools": {}}
    for line in open(os.path.join(d, "meta.txt")):
        t = line.split()
        if t[0] == "pool":
            meta["pools"][(int(t[1]), int(t[2]))] = (int(t[3]), int(t[4]))
        elif t[0] == "atlas_res":
            meta["atlas_res"] = tuple(int(x) for x in t[1:4])
        elif t[0] in ("bricks", "levels", "width", "height"):
            meta[t[0]] = int(t[1])
        else:
            meta[t[0]] = t[1]
    out = {"meta": meta, "dir": d}
    out["vdbinfo"] = open(os.path.join(d, "vdbinfo.bin"), "rb").read()
    rx, ry, rz = meta["atlas_res"]
    out["atlas"] = np.fromfile(os.path.join(d, "atlas.bin"), dtype=np.float32).reshape(rz, ry, rx)
    out["transfer"] = np.fromfile(os.path.join(d, "transfer.bin"), dtype=np.float32)
    out["pool0"], out["pool1"] = {}, {}
    for (g, l), (cnt, wid) in meta["pools"].items():
        p = os.path.join(d, f"pool{g}_L{l}.bin")
        b = np.fromfile(p, dtype=np.uint8) if os.path.getsize(p) else np.zeros(0, np.uint8)
        (out["pool0"] if g == 0 else out["pool1"])[l] = b
    out["scn"], out["rgba"], out["hit"] = {}, {}, {}
    w, h = meta["width"], meta["height"]
    for m in list(MODES) + list(MODES2) + [CUSTOM, VOXELID]:
        p = os.path.join(d, f"scninfo_{m}.bin")
        if os.path.exists(p):
            out["scn"][m] = open(p, "rb").read()
            out["rgba"][m] = np.fromfile(os.path.join(d, f"out_{m}.rgba"), dtype=np.uint8).reshape(h, w, 4)
            hp = os.path.join(d, f"hit_{m}.f32")
            if os.path.exists(hp):
                out["hit"][m] = np.fromfile(hp, dtype=np.float32).reshape(h, w, 8)
    q = os.path.join(d, "color.bin")
    if os.path.exists(q):
        out["color"] = np.fromfile(q, dtype=np.uint8).reshape(rz, ry, rx, 4)
    for name in ("dbuf", "rays_in", "rays_out"):
        q = os.path.join(d, name + ".bin")
        if os.path.exists(q):
            out[name] = np.fromfile(q, dtype=np.float32)
    q = os.path.join(d, "scninfo_raytrace.bin")
    if os.path.exists(q):
        out["scn_raytrace"] = open(q, "rb").read()
    return out


def load_lightdump(d):
    """dump of a --lightdump run (large volumes): VDBInfo, ScnInfo + image (+ hit buffer) per mode, and CHECKSUMS of the pools
    and of the atlas instead of their bytes (tests/common.py::cksum32)."""
    meta = {"cksum_pool": {}}
    for line in open(os.path.join(d, "meta.txt")):
        t = line.split()
        if t[0] == "cksum_pool":
            meta["cksum_pool"][(int(t[1]), int(t[2]))] = (int(t[3]), int(t[4]), int(t[5]))
        elif t[0] == "cksum_atlas":
            meta["cksum_atlas"] = (int(t[1]), int(t[2]), int(t[3]))
        elif t[0] == "atlas_res":
            meta["atlas_res"] = tuple(int(x) for x in t[1:4])
        elif t[0] in ("bricks", "levels", "width", "height"):
            meta[t[0]] = int(t[1])
        else:
            meta[t[0]] = t[1]
    out = {"meta": meta, "dir": d, "vdbinfo": open(os.path.join(d, "vdbinfo.bin"), "rb").read(), "scn": {}, "rgba": {}, "hit": {}}
    out["transfer"] = np.fromfile(os.path.join(d, "transfer.bin"), dtype=np.float32)
    w, h = meta["width"], meta["height"]
    for m in list(MODES) + list(MODES2) + [CUSTOM, VOXELID]:
        p = os.path.join(d, f"scninfo_{m}.bin")
        if os.path.exists(p):
            out["scn"][m] = open(p, "rb").read()
            out["rgba"][m] = np.fromfile(os.path.join(d, f"out_{m}.rgba"), dtype=np.uint8).reshape(h, w, 4)
            hp = os.path.join(d, f"hit_{m}.f32")
            if os.path.exists(hp):
                out["hit"][m] = np.fromfile(hp, dtype=np.float32).reshape(h, w, 8)
    return out


def same_volume(light, vol):
    """the volume the reference built itself (checksums of a --lightdump run) == the volume `vol` the product imports"""
    from common import cksum32, mask_vdbinfo
    bad = []
    lv = light["meta"]["levels"]
    if not np.array_equal(mask_vdbinfo(light["vdbinfo"], lv), mask_vdbinfo(bytes(vol["vdbinfo"]), lv)):
        bad.append("vdbinfo")
    for (g, l), (nbytes, s1, s2) in light["meta"]["cksum_pool"].items():
        a = (vol["pool0"] if g == 0 else vol["pool1"]).get(l, np.zeros(0, np.uint8))
        if a.nbytes != nbytes or cksum32(a) != (s1, s2):
            bad.append(f"pool{g}_L{l}")
    nbytes, s1, s2 = light["meta"]["cksum_atlas"]
    if vol["atlas"].nbytes != nbytes or cksum32(vol["atlas"]) != (s1, s2):
        bad.append("atlas")
    return bad


def make_renderer(dump, pkg, device=0):
    r = pkg.Renderer(device)
    r.import_topology_host(dump["vdbinfo"], dump["pool0"], dump["pool1"])
    r.import_atlas_host(dump["atlas"])
    if "color" in dump:
        r.import_color_host(dump["color"], linear=False)       # the harness creates the channel with F_POINT
    r.set_transfer(dump["transfer"])
    return r


def patch_dbuf(scninfo, ptr):
    """ScnInfo.dbuf (offset 400) holds a device pointer of the reference process: replace it with ours."""
    b = bytearray(scninfo)
    b[400:408] = int(ptr).to_bytes(8, "little")
    return bytes(b)


def render_mine(r, dump, mode, sampler, debug=False):
    import torch
    w, h = dump["meta"]["width"], dump["meta"]["height"]
    r.set_sampler(sampler)
    if "dbuf" in dump:
        if "_dbuf_dev" not in dump:
            dump["_dbuf_dev"] = torch.from_numpy(dump["dbuf"]).cuda()
        dump = dict(dump)
        dump["scn"] = {m: patch_dbuf(s, dump["_dbuf_dev"].data_ptr()) for m, s in dump["scn"].items()}
    out = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    _, dshadow, spp = MODES2.get(mode, (0, 0, 1))
    r.set_deep_shadow(dshadow)
    r.set_spp(spp)
    try:
        if debug:
            dbg = torch.zeros((h, w, 12), dtype=torch.float32, device="cuda")
            r.render_debug(dump["scn"][mode], ALL_SHADE[mode], out.data_ptr(), dbg.data_ptr())
            r.sync()
            return out.cpu().numpy(), dbg.cpu().numpy(), r.counters()
        r.render(dump["scn"][mode], ALL_SHADE[mode], out.data_ptr())
        r.sync()
        return out.cpu().numpy(), None, None
    finally:
        r.set_deep_shadow(0)
        r.set_spp(1)


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else float(10 * np.log10(255.0 ** 2 / mse))


def compare_voxel_ids(dbg, ref):
    """north_star: voxel-hit IDs and depths bit-exact.  `dbg` = gvdbx_render_debug output [h, w, 12] (float view), `ref` =
    oracleVoxelId output [h, w, 8]: {hit.xyz, t} {voxel.xyz, leaf} (integers as bit patterns).  Returns mismatch counts."""
    d, r = dbg.view(np.uint32), ref.view(np.uint32)
    hitpix = ref[:, :, 2] != np.float32(1.0e10)
    return {"hit_pixels": int(hitpix.sum()),
            "hit_mismatch": int((d[:, :, 0:3] != r[:, :, 0:3]).any(axis=2).sum()),
            "depth_mismatch": int((d[:, :, 3] != r[:, :, 3]).sum()),
            "voxel_mismatch": int((d[:, :, 8:11] != r[:, :, 4:7]).any(axis=2).sum()),
            "leaf_mismatch": int((d[:, :, 7] != r[:, :, 7]).sum())}


def compare(dump, pkg, modes=None, device=0, verbose=True):
    """Returns {mode: {sampler: stats}}."""
    res = {}
    r = make_renderer(dump, pkg, device)
    for m in (modes or dump["scn"].keys()):
        ref = dump["rgba"][m]
        res[m] = {}
        for sname, s in (("tex", 0), ("linear", 1)):
            mine, dbg, cnt = render_mine(r, dump, m, s, debug=True)
            plain, _, _ = render_mine(r, dump, m, s, debug=False)
            diff = np.abs(mine.astype(np.int32) - ref.astype(np.int32))
            st = {
                "pixels": int(ref.shape[0] * ref.shape[1]),
                "rgba_mismatch_pixels": int((diff.max(axis=2) > 0).sum()),
                "rgba_max_abs": int(diff.max()),
                "rgba_over1_pixels": int((diff.max(axis=2) > 1).sum()),
                "psnr": psnr(mine, ref),
                "plain_equals_debug": bool(np.array_equal(plain, mine)),
                "counters": cnt,
            }
            if m in dump["hit"]:
                rh = dump["hit"][m]
                if m == "deep":
                    a = dbg[:, :, 0:4].view(np.uint32)
                    b = rh[:, :, 0:4].view(np.uint32)
                    st["raw_clr_mismatch_pixels"] = int((a != b).any(axis=2).sum())
                    st["raw_clr_max_abs"] = float(np.abs(dbg[:, :, 0:4] - rh[:, :, 0:4]).max())
                else:
                    a = dbg[:, :, 0:3].view(np.uint32)
                    b = rh[:, :, 0:3].view(np.uint32)
                    bad_hit = (a != b).any(axis=2)
                    an = dbg[:, :, 4:7].view(np.uint32)
                    bn = rh[:, :, 4:7].view(np.uint32)
                    bad_norm = (an != bn).any(axis=2)
                    st["hit_mismatch_pixels"] = int(bad_hit.sum())
                    st["norm_mismatch_pixels"] = int(bad_norm.sum())
                    st["hit_pixels"] = int((rh[:, :, 2] != np.float32(1.0e10)).sum())
                    if bad_hit.any():
                        ys, xs = np.nonzero(bad_hit)
                        st["hit_examples"] = [
                            {"x": int(x), "y": int(y), "ref": rh[y, x, 0:3].tolist(), "mine": dbg[y, x, 0:3].tolist()}
                            for y, x in list(zip(ys, xs))[:5]]
            res[m][sname] = st
            if verbose:
                print(f"[refcmp] {dump['meta'].get('preset')} {m:9s} {sname:6s} " +
                      json.dumps({k: v for k, v in st.items() if k not in ("counters", "hit_examples")}), flush=True)
    r.close()
    return res


def compare2(dump, pkg, modes=None, device=0, verbose=True):
    """The rest of Render()'s switch + the composed modes (texture sampler; deep variants also with linear loads)."""
    res = {}
    r = make_renderer(dump, pkg, device)
    for m in (modes or [k for k in MODES2 if k in dump["scn"]]):
        ref = dump["rgba"][m]
        res[m] = {}
        for sname, s in (("tex", 0), ("linear", 1)):
            if s == 1 and not m.startswith("deep"):
                continue
            mine, dbg, cnt = render_mine(r, dump, m, s, debug=(m != "deepspp"))
            plain, _, _ = render_mine(r, dump, m, s, debug=False)
            diff = np.abs(plain.astype(np.int32) - ref.astype(np.int32))
            st = {"pixels": int(ref.shape[0] * ref.shape[1]),
                  "rgba_mismatch_pixels": int((diff.max(axis=2) > 0).sum()), "rgba_max_abs": int(diff.max()),
                  "rgba_over1_pixels": int((diff.max(axis=2) > 1).sum()), "psnr": psnr(plain, ref),
                  "plain_equals_debug": bool(np.array_equal(plain, mine)), "nonbackground": int((ref != ref[0, 0]).any(axis=2).sum())}
            if m in dump["hit"] and dbg is not None:
                rh = dump["hit"][m]
                bad_hit = (dbg[:, :, 0:3].view(np.uint32) != rh[:, :, 0:3].view(np.uint32)).any(axis=2)
                st["hit_mismatch_pixels"] = int(bad_hit.sum())
                if m == "tricubic":
                    st["norm_mismatch_pixels"] = int((dbg[:, :, 4:7].view(np.uint32) != rh[:, :, 4:7].view(np.uint32)).any(axis=2).sum())
                st["hit_pixels"] = int((rh[:, :, 2] != np.float32(1.0e10)).sum())
            res[m][sname] = st
            if verbose:
                print(f"[refcmp2] {dump['meta'].get('preset')} {m:10s} {sname:6s} " + json.dumps(st), flush=True)
    r.close()
    return res


def main():
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    pkg = load_package()
    ap = argparse.ArgumentParser()
    ap.add_argument("--presets", default="cfg1_tiny,cfg2_tiny,cfg3_tiny,cfg4_tiny")
    ap.add_argument("--modes", default="voxel,trilinear,levelset,deep")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "refcmp.json"))
    ap.add_argument("--keep", default="")
    ap.add_argument("--modes2", default="", help="comma list out of " + ",".join(MODES2) + " or 'all'")
    a = ap.parse_args()
    m2 = list(MODES2) if a.modes2 == "all" else [m for m in a.modes2.split(",") if m]
    allres = {}
    for preset in a.presets.split(","):
        d = os.path.join(a.keep, preset) if a.keep else tempfile.mkdtemp(prefix="refdump_")
        os.makedirs(d, exist_ok=True)
        timing = run_ref(preset, d, modes=a.modes.split(",") + m2)
        dump = load_dump(d)
        allres[preset] = {"ref_timing": timing, "cmp": compare(dump, pkg, a.modes.split(","))}
        if m2:
            allres[preset]["cmp2"] = compare2(dump, pkg, m2)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(allres, open(a.out, "w"), indent=1)
    print("[refcmp] wrote", a.out)


if __name__ == "__main__":
    main()
