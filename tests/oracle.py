"""oracle.py — ctypes front end of oracle/liboracle.so (CPU restatement; TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
SHADE = {"voxel": 0, "trilinear": 4, "levelset": 6, "deep": 7}


class Preset(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("kind", C.c_int), ("N", C.c_int), ("a", C.c_float), ("b", C.c_float), ("c", C.c_float),
                ("width", C.c_int), ("height", C.c_int), ("shade", C.c_int), ("fov", C.c_float),
                ("cam_angs", C.c_float * 3), ("cam_target", C.c_float * 3), ("cam_dist", C.c_float),
                ("light_angs", C.c_float * 3), ("light_target", C.c_float * 3), ("light_dist", C.c_float),
                ("steps", C.c_float * 3), ("extinct", C.c_float * 3), ("thresh", C.c_float * 3), ("cutoff", C.c_float * 3),
                ("backclr", C.c_float * 4), ("shadow", C.c_float * 3), ("epsilon", C.c_float), ("transfer", C.c_int)]


class OraVolume(C.Structure):
    _fields_ = [("vdbinfo", C.c_void_p), ("pool0", C.c_void_p * 10), ("pool1", C.c_void_p * 10), ("atlas", C.c_void_p),
                ("atlas_res", C.c_int * 3), ("transfer", C.c_void_p)]


_L = None


def lib():
    global _L
    if _L is not None:
        return _L
    so = os.path.join(ODIR, "liboracle.so")
    srcs = [os.path.join(ODIR, f) for f in ("gvdb_oracle.c", "gvdb_oracle.h", "scenes.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["make", "oracle"], cwd=ODIR, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    L = C.CDLL(so)
    L.ora_tree_create.restype = C.c_void_p
    L.ora_tree_create.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int]
    L.ora_tree_destroy.argtypes = [C.c_void_p]
    L.ora_activate_space.restype = C.c_int64
    L.ora_activate_space.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.ora_activate_bricks.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.ora_finish_topology.argtypes = [C.c_void_p]
    L.ora_update_atlas.argtypes = [C.c_void_p]
    L.ora_set_epsilon.argtypes = [C.c_void_p, C.c_float, C.c_int]
    L.ora_pool_count.restype = C.c_uint64
    L.ora_pool_count.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ora_pool_width.restype = C.c_uint64
    L.ora_pool_width.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ora_pool_data.restype = C.c_void_p
    L.ora_pool_data.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ora_num_levels.argtypes = [C.c_void_p]
    L.ora_atlas_res.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.ora_fill_vdbinfo.argtypes = [C.c_void_p, C.c_void_p]
    L.ora_fill_atlas.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float]
    L.ora_render.argtypes = [C.POINTER(OraVolume), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.ora_render_ex.argtypes = [C.POINTER(OraVolume), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.ora_tex3d.restype = C.c_float
    L.ora_tex3d.argtypes = [C.POINTER(OraVolume), C.c_float, C.c_float, C.c_float]
    L.ora_scene_preset.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t]
    L.ora_scene_generate.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.POINTER(C.c_float))]
    L.ora_free.argtypes = [C.c_void_p]
    _L = L
    return L


def preset(name):
    p = Preset()
    if lib().ora_scene_preset(name.encode(), C.byref(p), C.sizeof(p)) != 0:
        raise KeyError(name)
    return p


def generate(p):
    """-> (brick_pos [n,3] int32, values [n,512] float32)"""
    n = C.c_int()
    bp = C.POINTER(C.c_int32)()
    vals = C.POINTER(C.c_float)()
    lib().ora_scene_generate(C.byref(p), C.byref(n), C.byref(bp), C.byref(vals))
    nb = n.value
    pos = np.ctypeslib.as_array(bp, shape=(nb, 3)).copy() if nb else np.zeros((0, 3), np.int32)
    v = np.ctypeslib.as_array(vals, shape=(nb, 512)).copy() if nb else np.zeros((0, 512), np.float32)
    lib().ora_free(bp)
    lib().ora_free(vals)
    return pos, v


def build_volume(brick_pos, values, epsilon=0.001, transfer=None, timing=None):
    """CPU topology build (Configure(3,3,3,3,3) + ActivateSpace + FinishTopology + UpdateAtlas) + atlas fill with
    aprons.  Returns the dict refcmp.load_dump returns for a reference dump (vdbinfo / pool0 / pool1 / atlas)."""
    import time
    L = lib()
    logdim = (C.c_int * 5)(3, 3, 3, 3, 3)
    initcnt = (C.c_int * 5)(4, 4, 2, 1, 1)          # Configure(q4..q0), gvdb_volume_gvdb.cpp:2364-2377
    t0 = time.perf_counter()
    # SetChannelDefault(16, 16, 1) like the reference samples; above ~400 k bricks the 16 x 16 x N slot grid would exceed the
    # 16384-texel limit of a 3-D array along z, so the large volume uses 128 x 128 x N (cfg 5)
    cxy = 128 if len(brick_pos) > 400000 else 16
    t = L.ora_tree_create(5, logdim, initcnt, cxy, cxy, 1, 1)
    bp = np.ascontiguousarray(brick_pos, np.int32)
    L.ora_activate_bricks(t, bp.ctypes.data_as(C.c_void_p), len(bp))
    L.ora_finish_topology(t)
    L.ora_update_atlas(t)
    if timing is not None:
        timing["topology_build_s"] = time.perf_counter() - t0
    L.ora_set_epsilon(t, C.c_float(epsilon), 256)
    out = {"pool0": {}, "pool1": {}, "meta": {"bricks": len(brick_pos)}}
    for g in (0, 1):
        for lev in range(5):
            cnt, wid = L.ora_pool_count(t, g, lev), L.ora_pool_width(t, g, lev)
            nbytes = cnt * wid
            buf = np.zeros(nbytes, np.uint8)
            if nbytes:
                C.memmove(buf.ctypes.data, L.ora_pool_data(t, g, lev), nbytes)
            (out["pool0"] if g == 0 else out["pool1"])[lev] = buf
    vb = (C.c_uint8 * 1232)()
    L.ora_fill_vdbinfo(t, vb)
    out["vdbinfo"] = bytes(vb)
    res = (C.c_int * 3)()
    L.ora_atlas_res(t, res)
    rx, ry, rz = res[0], res[1], res[2]
    atlas = np.zeros((rz, ry, rx), np.float32)
    vals = np.ascontiguousarray(values, np.float32)
    L.ora_fill_atlas(t, vals.ctypes.data_as(C.c_void_p), atlas.ctypes.data_as(C.c_void_p), C.c_float(0.0))
    out["atlas"] = atlas
    out["meta"]["atlas_res"] = (rx, ry, rz)
    out["transfer"] = transfer
    L.ora_tree_destroy(t)
    return out


def _ora_volume(vol):
    v = OraVolume()
    keep = []
    vb = np.frombuffer(vol["vdbinfo"], np.uint8).copy()
    keep.append(vb)
    v.vdbinfo = vb.ctypes.data
    for lev in range(10):
        for arr, src in ((v.pool0, vol["pool0"]), (v.pool1, vol["pool1"])):
            b = src.get(lev)
            if b is not None and len(b):
                b = np.ascontiguousarray(b)
                keep.append(b)
                arr[lev] = b.ctypes.data
    a = np.ascontiguousarray(vol["atlas"], np.float32)
    keep.append(a)
    v.atlas = a.ctypes.data
    v.atlas_res[0], v.atlas_res[1], v.atlas_res[2] = a.shape[2], a.shape[1], a.shape[0]
    if vol.get("transfer") is not None:
        t = np.ascontiguousarray(vol["transfer"], np.float32)
        keep.append(t)
        v.transfer = t.ctypes.data
    return v, keep


def render(vol, scninfo, shade, rows=None, threads=0, want_hits=False, deep_shadow=False, spp=1):
    """CPU ray caster (every shade mode of Render()'s switch; deep_shadow / spp = BASELINE configs 4 / 5).
    Returns rgba [h,w,4] (+ hit/norm [h,w,8])."""
    s = np.frombuffer(scninfo, np.uint8).copy()
    w, h = int(s[0:4].view(np.int32)[0]), int(s[4:8].view(np.int32)[0])
    v, keep = _ora_volume(vol)
    out = np.zeros((h, w, 4), np.uint8)
    hn = np.zeros((h, w, 8), np.float32) if want_hits else None
    y0, y1 = rows if rows else (0, h)
    rc = lib().ora_render_ex(C.byref(v), s.ctypes.data_as(C.c_void_p), int(shade), y0, y1, out.ctypes.data_as(C.c_void_p),
                             hn.ctypes.data_as(C.c_void_p) if want_hits else None, int(threads), 1 if deep_shadow else 0, int(spp))
    if rc != 0:
        raise RuntimeError("ora_render: unsupported input")
    return (out, hn) if want_hits else out


def tex3d(vol, pts):
    v, keep = _ora_volume(vol)
    return np.array([lib().ora_tex3d(C.byref(v), float(x), float(y), float(z)) for x, y, z in pts], np.float32)


def scene_volume(name, timing=None):
    """preset name -> (Preset, volume dict incl. transfer table built through the product's host mirror)."""
    p = preset(name)
    pos, vals = generate(p)
    vol = build_volume(pos, vals, epsilon=p.epsilon, timing=timing)
    vol["meta"].update(width=p.width, height=p.height, preset=name)
    vol["brick_pos"] = pos
    return p, vol


def scninfo_for(pkg, p, shade=None, size=None, cam_angs=None):
    """ScnInfo bytes for a preset through the PRODUCT's host mirror (gvdbx_host.cpp) + the transfer table."""
    v = pkg.Volume(-1)
    v.SetSceneParams(list(p.steps), list(p.extinct), list(p.thresh), list(p.cutoff), list(p.backclr), list(p.shadow))
    if p.transfer == 1:         # gRenderToFile ramps, main_rendertofile.cpp:48-51
        v.LinearTransferFunc(0.00, 0.25, (0, 0, 0, 0), (1, 1, 0, 0.1))
        v.LinearTransferFunc(0.25, 0.50, (1, 1, 0, 0.4), (1, 0, 0, 0.3))
        v.LinearTransferFunc(0.50, 0.75, (1, 0, 0, 0.3), (.2, .2, 0.2, 0.1))
        v.LinearTransferFunc(0.75, 1.00, (.2, .2, 0.2, 0.1), (0, 0, 0, 0.0))
    v.SetCamera(p.fov, list(cam_angs) if cam_angs is not None else list(p.cam_angs), list(p.cam_target), p.cam_dist)
    w, h = size if size else (p.width, p.height)
    v.SetRes(w, h)
    v.SetLight(list(p.light_angs), list(p.light_target), p.light_dist)
    scn = v.PrepareRender(w, h, p.shade if shade is None else shade)
    table = v.transfer_table()
    v.close()
    return scn, table
