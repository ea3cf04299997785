"""ctypes bindings for libgvdbx.so (include/gvdbx.h).

Mirrors the reference call sequence of VolumeGVDB (src/gvdb_volume_gvdb.cpp):
    PrepareVDB        :3946  -> Renderer.import_topology / import_topology_host
    SetupAtlasAccess  :720   -> Renderer.import_atlas_host / import_atlas_array
    CommitTransferFunc:4892  -> Renderer.set_transfer
    Render            :4336  -> Renderer.render(scninfo, shade, out_ptr)
    ReadRenderBuf     :4241  -> Renderer.read_buffer
"""
import ctypes as C
import os

import numpy as np

SHADE_VOXEL, SHADE_TRILINEAR, SHADE_LEVELSET, SHADE_VOLUME, SHADE_OFF = 0, 4, 6, 7, 100
SHADE_SECTION2D, SHADE_SECTION3D, SHADE_EMPTYSKIP, SHADE_TRICUBIC = 1, 2, 3, 5
SAMPLER_TEX, SAMPLER_LINEAR = 0, 1
OPT_SAMPLER, OPT_BLOCK_W, OPT_BLOCK_H, OPT_COUNTERS, OPT_TRAVERSAL, OPT_CULL, OPT_SPP, OPT_DEEP_SHADOW = 1, 2, 3, 4, 5, 6, 7, 8
VDBINFO_BYTES, SCNINFO_BYTES = 1232, 416

EXPORTED_SYMBOLS = [
    "gvdbx_create", "gvdbx_destroy", "gvdbx_last_error", "gvdbx_set_option",
    "gvdbx_import_topology", "gvdbx_import_topology_host",
    "gvdbx_import_atlas_array", "gvdbx_import_atlas_host", "gvdbx_import_atlas_device", "gvdbx_import_color_array", "gvdbx_import_color_host", "gvdbx_clear_color", "gvdbx_set_transfer",
    "gvdbx_render", "gvdbx_render_tiles", "gvdbx_tiles_per_rank", "gvdbx_assemble_tiles",
    "gvdbx_render_debug", "gvdbx_raytrace", "gvdbx_read_buffer", "gvdbx_sync", "gvdbx_get_counters",
    "gvdbx_sample_points", "gvdbx_render_banded", "gvdbx_read_banded", "gvdbx_measure_tex_peak", "gvdbx_measure_sampler_ab", "gvdbx_measure_deep_loop_peak", "gvdbx_kernel_params", "gvdbx_update_apron", "gvdbx_update_apron_faces", "gvdbx_export_atlas_host", "gvdbx_render_tiles_direct", "gvdbx_render_tiles_ring", "gvdbx_peer_alloc", "gvdbx_peer_free", "gvdbx_peer_open",
    "gvdbx_peer_close", "gvdbx_stream_signal", "gvdbx_stream_signal_add", "gvdbx_stream_signal_many", "gvdbx_stream_wait", "gvdbx_set_stream",
    "gvdbx_render_bands", "gvdbx_render_multi", "gvdbx_ring_create", "gvdbx_ring_connect", "gvdbx_ring_submit", "gvdbx_ring_acquire",
    "gvdbx_ring_release", "gvdbx_ring_frame", "gvdbx_ring_destroy", "gvdbx_hostring_create", "gvdbx_hostring_submit", "gvdbx_hostring_wait",
    "gvdbx_hostring_release", "gvdbx_hostring_destroy",
    "gvdbx_read_buffer_async", "gvdbx_lanes", "gvdbx_lane_select", "gvdbx_lane_stream", "gvdbx_lanes_fork", "gvdbx_lanes_join",
]


class GvdbxError(RuntimeError):
    pass


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("s_tri", "s_pt", "n_dda", "n_desc", "s_lut", "rays")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def lib_path():
    """in-tree libgvdbx.so; GVDBX_LIB points experiments (A/B builds of the kernels) at another build of the same ABI"""
    return os.environ.get("GVDBX_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgvdbx.so")


_LIB = None


def lib():
    """Load libgvdbx.so; fails loudly when the CUDA extension has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise GvdbxError(f"{p} is missing: build it with `make -C gvdb-voxels_b200` "
                         "(there is no CPU fallback for the render path)")
    L = C.CDLL(p)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    L.gvdbx_create.argtypes = [C.POINTER(vp), i32, vp]
    L.gvdbx_destroy.argtypes = [vp]
    L.gvdbx_last_error.argtypes = [vp]
    L.gvdbx_last_error.restype = C.c_char_p
    L.gvdbx_set_option.argtypes = [vp, i32, i32]
    L.gvdbx_import_topology.argtypes = [vp, vp]
    L.gvdbx_import_topology_host.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)]
    L.gvdbx_import_atlas_array.argtypes = [vp, i32, vp, i32, i32, i32]
    L.gvdbx_import_atlas_host.argtypes = [vp, i32, vp, i32, i32, i32]
    L.gvdbx_import_atlas_device.argtypes = [vp, i32, u64, i32, i32, i32]
    L.gvdbx_import_color_array.argtypes = [vp, vp, i32]
    L.gvdbx_import_color_host.argtypes = [vp, vp, i32, i32, i32, i32]
    L.gvdbx_clear_color.argtypes = [vp]
    L.gvdbx_set_transfer.argtypes = [vp, vp]
    L.gvdbx_render.argtypes = [vp, vp, i32, i32, u64, i32, i32, i32, i32]
    L.gvdbx_render_tiles.argtypes = [vp, vp, i32, i32, u64, i32, i32, i32]
    L.gvdbx_tiles_per_rank.argtypes = [i32, i32, i32, i32]
    L.gvdbx_assemble_tiles.argtypes = [vp, u64, u64, i32, i32, i32, i32]
    L.gvdbx_render_debug.argtypes = [vp, vp, i32, i32, u64, u64]
    L.gvdbx_raytrace.argtypes = [vp, vp, i32, u64, i32, C.c_float]
    L.gvdbx_read_buffer.argtypes = [vp, u64, vp, C.c_size_t]
    L.gvdbx_sync.argtypes = [vp]
    L.gvdbx_get_counters.argtypes = [vp, C.POINTER(Counters)]
    L.gvdbx_sample_points.argtypes = [vp, i32, u64, i32, u64, u64]
    L.gvdbx_measure_tex_peak.argtypes = [vp, C.c_float, C.POINTER(C.c_double)]
    L.gvdbx_measure_sampler_ab.argtypes = [vp, C.c_float, C.POINTER(C.c_double)]
    L.gvdbx_measure_deep_loop_peak.argtypes = [vp, C.c_float, i32, C.POINTER(C.c_double)]
    L.gvdbx_render_tiles_direct.argtypes = [vp, vp, i32, i32, u64, i32, i32, i32]
    L.gvdbx_kernel_params.argtypes = [vp, vp, i32, i32, u64, vp, C.c_size_t]
    L.gvdbx_update_apron.argtypes = [vp, i32, C.c_float]
    L.gvdbx_update_apron_faces.argtypes = [vp, i32]
    L.gvdbx_export_atlas_host.argtypes = [vp, i32, vp, i32, i32, i32]
    L.gvdbx_render_tiles_ring.argtypes = [vp, vp, i32, i32, u64, i32, i32, i32, u64, C.c_uint32, u64]
    L.gvdbx_peer_alloc.argtypes = [vp, C.c_size_t, C.POINTER(u64), vp]
    L.gvdbx_peer_free.argtypes = [vp, u64]
    L.gvdbx_peer_open.argtypes = [vp, vp, C.POINTER(u64)]
    L.gvdbx_peer_close.argtypes = [vp, u64]
    L.gvdbx_stream_signal.argtypes = [vp, vp, u64, C.c_uint32]
    L.gvdbx_stream_wait.argtypes = [vp, vp, u64, C.c_uint32]
    L.gvdbx_stream_signal_add.argtypes = [vp, vp, u64, C.c_uint32]
    L.gvdbx_stream_signal_many.argtypes = [vp, vp, C.POINTER(u64), i32, C.c_uint32]
    L.gvdbx_set_stream.argtypes = [vp, vp]
    L.gvdbx_read_buffer_async.argtypes = [vp, u64, vp, C.c_size_t]
    L.gvdbx_render_banded.argtypes = [vp, vp, i32, i32, u64, i32]
    L.gvdbx_read_banded.argtypes = [vp, u64, vp, C.c_size_t]
    u32 = C.c_uint32
    L.gvdbx_render_bands.argtypes = [vp, vp, i32, i32, u64, i32, i32, i32]
    L.gvdbx_render_multi.argtypes = [C.POINTER(vp), i32, vp, i32, i32, u64, i32]
    L.gvdbx_ring_create.argtypes = [vp, i32, i32, i32, i32, i32, i32, C.POINTER(vp), vp]
    L.gvdbx_ring_connect.argtypes = [vp, vp]
    L.gvdbx_ring_submit.argtypes = [vp, vp, i32, i32, C.POINTER(u32)]
    L.gvdbx_ring_acquire.argtypes = [vp, u32, vp, C.POINTER(u64)]
    L.gvdbx_ring_release.argtypes = [vp, u32, vp]
    L.gvdbx_ring_frame.argtypes = [vp, u32, C.POINTER(u64)]
    L.gvdbx_ring_destroy.argtypes = [vp]
    L.gvdbx_hostring_create.argtypes = [vp, C.c_char_p, i32, i32, i32, i32, i32, i32, C.POINTER(vp)]
    L.gvdbx_hostring_submit.argtypes = [vp, vp, i32, i32, C.POINTER(u32)]
    L.gvdbx_hostring_wait.argtypes = [vp, u32, C.POINTER(vp), i32]
    L.gvdbx_hostring_release.argtypes = [vp, u32]
    L.gvdbx_hostring_destroy.argtypes = [vp]
    L.gvdbx_lanes.argtypes = [vp, i32]
    L.gvdbx_lane_select.argtypes = [vp, i32]
    L.gvdbx_lane_stream.argtypes = [vp, i32]
    L.gvdbx_lanes_fork.argtypes = [vp]
    L.gvdbx_lanes_join.argtypes = [vp]
    for s in EXPORTED_SYMBOLS:
        if s not in ("gvdbx_last_error", "gvdbx_lane_stream"):
            getattr(L, s).restype = i32
    L.gvdbx_lane_stream.restype = vp
    _LIB = L
    return L


def _buf(b):
    """bytes / numpy array -> (ctypes pointer, keep-alive object)."""
    if isinstance(b, (bytes, bytearray)):
        a = np.frombuffer(bytes(b), dtype=np.uint8)
    else:
        a = np.ascontiguousarray(b)
    return a.ctypes.data_as(C.c_void_p), a


class Renderer:
    """One render context per CUDA device (the reference: one VolumeGVDB per device, gvdb_volume_gvdb.h:325)."""

    def __init__(self, device=0, stream=None):
        self._L = lib()
        h = C.c_void_p()
        rc = self._L.gvdbx_create(C.byref(h), int(device), C.c_void_p(stream or 0))
        if rc != 0:
            raise GvdbxError(f"gvdbx_create(device={device}) failed with {rc}: no CUDA device / no CPU fallback")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._L.gvdbx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            msg = self._L.gvdbx_last_error(self._h)
            raise GvdbxError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def set_option(self, opt, value):
        self._ck(self._L.gvdbx_set_option(self._h, opt, int(value)), "gvdbx_set_option")

    def set_sampler(self, sampler):
        self.set_option(OPT_SAMPLER, sampler)

    def set_block(self, w, h):
        self.set_option(OPT_BLOCK_W, w)
        self.set_option(OPT_BLOCK_H, h)

    def set_spp(self, n):
        """rays per pixel (1 = the reference's pixel-centre ray)"""
        self.set_option(OPT_SPP, n)

    def set_deep_shadow(self, on):
        """SHADE_VOLUME + one shadow march towards the light (BASELINE.json config 4)"""
        self.set_option(OPT_DEEP_SHADOW, 1 if on else 0)

    def set_counters(self, on):
        """0 = off, 1 / True = count the algorithm's work (no brick culling), 2 = count the production kernel's work"""
        self.set_option(OPT_COUNTERS, int(on))

    def import_topology(self, vdbinfo):
        p, keep = _buf(vdbinfo)
        assert keep.nbytes == VDBINFO_BYTES
        self._ck(self._L.gvdbx_import_topology(self._h, p), "gvdbx_import_topology")

    def import_topology_host(self, vdbinfo, pool0, pool1):
        """pool0 / pool1: dict or list level -> bytes / uint8 array (reference pool dumps)."""
        p, keep = _buf(vdbinfo)
        assert keep.nbytes == VDBINFO_BYTES
        a0 = (C.c_void_p * 10)()
        a1 = (C.c_void_p * 10)()
        n1 = (C.c_uint64 * 10)()
        alive = []
        for lev in range(10):
            for arr, src in ((a0, pool0), (a1, pool1)):
                b = src.get(lev) if isinstance(src, dict) else (src[lev] if lev < len(src) else None)
                if b is None or len(b) == 0:
                    arr[lev] = None
                    continue
                q, k = _buf(b)
                alive.append(k)
                arr[lev] = q
                if arr is a1:
                    n1[lev] = k.nbytes
        self._ck(self._L.gvdbx_import_topology_host(self._h, p, a0, a1, n1), "gvdbx_import_topology_host")

    def import_atlas_host(self, texels, chan=0):
        a = np.ascontiguousarray(texels, dtype=np.float32)
        assert a.ndim == 3, "atlas image must be [z][y][x]"
        rz, ry, rx = a.shape
        self._ck(self._L.gvdbx_import_atlas_host(self._h, chan, a.ctypes.data_as(C.c_void_p), rx, ry, rz),
                 "gvdbx_import_atlas_host")

    def import_atlas_device(self, texels_ptr, res_xyz, chan=0):
        """atlas from a device image [z][y][x] float32 (e.g. received by a broadcast)"""
        self._ck(self._L.gvdbx_import_atlas_device(self._h, chan, int(texels_ptr), *map(int, res_xyz)), "gvdbx_import_atlas_device")

    def import_color_host(self, rgba8, linear=True):
        """colour channel: uint8 [z][y][x][4] atlas image with the slot layout of channel 0; linear = AddChannel's filter"""
        a = np.ascontiguousarray(rgba8, dtype=np.uint8)
        assert a.ndim == 4 and a.shape[3] == 4
        rz, ry, rx, _ = a.shape
        self._ck(self._L.gvdbx_import_color_host(self._h, a.ctypes.data_as(C.c_void_p), rx, ry, rz, 1 if linear else 0),
                 "gvdbx_import_color_host")

    def clear_color(self):
        self._ck(self._L.gvdbx_clear_color(self._h), "gvdbx_clear_color")

    def import_atlas_array(self, cuarray, res_xyz, chan=0):
        self._ck(self._L.gvdbx_import_atlas_array(self._h, chan, C.c_void_p(cuarray), *map(int, res_xyz)),
                 "gvdbx_import_atlas_array")

    def update_apron(self, boundval=0.0, chan=0):
        """VolumeGVDB::UpdateApron(chan, boundval) on the imported atlas"""
        self._ck(self._L.gvdbx_update_apron(self._h, chan, C.c_float(boundval)), "gvdbx_update_apron")

    def update_apron_faces(self, chan=0):
        """VolumeGVDB::UpdateApronFaces(chan): face layers swapped between adjacent bricks"""
        self._ck(self._L.gvdbx_update_apron_faces(self._h, chan), "gvdbx_update_apron_faces")

    def export_atlas_host(self, shape_zyx, chan=0):
        rz, ry, rx = shape_zyx
        out = np.empty((rz, ry, rx), dtype=np.float32)
        self._ck(self._L.gvdbx_export_atlas_host(self._h, chan, out.ctypes.data_as(C.c_void_p), rx, ry, rz), "gvdbx_export_atlas_host")
        return out

    def set_transfer(self, rgba):
        a = np.ascontiguousarray(rgba, dtype=np.float32).reshape(-1)
        assert a.size == 16384 * 4
        self._ck(self._L.gvdbx_set_transfer(self._h, a.ctypes.data_as(C.c_void_p)), "gvdbx_set_transfer")

    def render(self, scninfo, shade, out_ptr, chan=0, tile=None):
        p, keep = _buf(scninfo)
        assert keep.nbytes == SCNINFO_BYTES
        x0, y0, w, h = tile if tile else (0, 0, 0, 0)
        self._ck(self._L.gvdbx_render(self._h, p, shade, chan, int(out_ptr), x0, y0, w, h), "gvdbx_render")

    def render_debug(self, scninfo, shade, out_ptr, dbg_ptr, chan=0):
        p, keep = _buf(scninfo)
        self._ck(self._L.gvdbx_render_debug(self._h, p, shade, chan, int(out_ptr), int(dbg_ptr)), "gvdbx_render_debug")

    def render_tiles(self, scninfo, shade, packed_ptr, tile_size, rank, nranks, chan=0):
        p, keep = _buf(scninfo)
        self._ck(self._L.gvdbx_render_tiles(self._h, p, shade, chan, int(packed_ptr), tile_size, rank, nranks),
                 "gvdbx_render_tiles")

    def render_tiles_direct(self, scninfo, shade, frame_ptr, tile_size, rank, nranks, chan=0):
        """this rank's tiles straight into a row-major frame (local or peer-mapped)"""
        p, keep = _buf(scninfo)
        self._ck(self._L.gvdbx_render_tiles_direct(self._h, p, shade, chan, int(frame_ptr), tile_size, rank, nranks),
                 "gvdbx_render_tiles_direct")

    def render_bands(self, scninfo, shade, packed_ptr, band_rows, rank, nranks, chan=0):
        """full-width bands (band b belongs to rank b % nranks), packed band after band"""
        p, keep = _buf(scninfo)
        self._ck(self._L.gvdbx_render_bands(self._h, p, shade, chan, int(packed_ptr), band_rows, rank, nranks), "gvdbx_render_bands")

    def render_tiles_ring(self, scninfo, shade, frame_ptr, tile_size, rank, nranks, wait_flag, wait_value, done_flag, chan=0):
        """[wait] + this rank's tiles + done += 1 in one call (the per-frame step of PeerFrameRing)"""
        p, keep = _buf(scninfo)
        self._ck(self._L.gvdbx_render_tiles_ring(self._h, p, shade, chan, int(frame_ptr), tile_size, rank, nranks,
                                                 int(wait_flag), int(wait_value) & 0xFFFFFFFF, int(done_flag)), "gvdbx_render_tiles_ring")

    # --- peer memory + stream-ordered flags (multi-GPU without a gather)
    def peer_alloc(self, nbytes):
        d = C.c_uint64()
        hd = (C.c_uint8 * 64)()
        self._ck(self._L.gvdbx_peer_alloc(self._h, nbytes, C.byref(d), hd), "gvdbx_peer_alloc")
        return int(d.value), bytes(hd)

    def peer_free(self, dptr):
        self._ck(self._L.gvdbx_peer_free(self._h, int(dptr)), "gvdbx_peer_free")

    def peer_open(self, handle):
        d = C.c_uint64()
        p, keep = _buf(handle)
        self._ck(self._L.gvdbx_peer_open(self._h, p, C.byref(d)), "gvdbx_peer_open")
        return int(d.value)

    def peer_close(self, dptr):
        self._ck(self._L.gvdbx_peer_close(self._h, int(dptr)), "gvdbx_peer_close")

    def stream_signal(self, flag_ptr, value, stream=None):
        self._ck(self._L.gvdbx_stream_signal(self._h, C.c_void_p(stream or 0), int(flag_ptr), int(value) & 0xFFFFFFFF), "gvdbx_stream_signal")

    def stream_signal_add(self, flag_ptr, inc, stream=None):
        self._ck(self._L.gvdbx_stream_signal_add(self._h, C.c_void_p(stream or 0), int(flag_ptr), int(inc)), "gvdbx_stream_signal_add")

    def stream_signal_many(self, flag_ptrs, value, stream=None):
        a = (C.c_uint64 * len(flag_ptrs))(*[int(p) for p in flag_ptrs])
        self._ck(self._L.gvdbx_stream_signal_many(self._h, C.c_void_p(stream or 0), a, len(flag_ptrs), int(value) & 0xFFFFFFFF),
                 "gvdbx_stream_signal_many")

    def stream_wait(self, flag_ptr, value, stream=None):
        self._ck(self._L.gvdbx_stream_wait(self._h, C.c_void_p(stream or 0), int(flag_ptr), int(value) & 0xFFFFFFFF), "gvdbx_stream_wait")

    def set_stream(self, stream):
        self._ck(self._L.gvdbx_set_stream(self._h, C.c_void_p(stream or 0)), "gvdbx_set_stream")

    def render_custom_example(self, scninfo, out_ptr, chan=0):
        """the reference's gRenderKernel sample kernel built against csrc/gvdbx_plugin.cuh (libgvdbx_custom_example.so)"""
        path = os.path.join(os.path.dirname(lib_path()), "libgvdbx_custom_example.so")
        if not os.path.exists(path):
            raise GvdbxError(f"{path} is missing: build it with `make -C gvdb-voxels_b200`")
        ex = C.CDLL(path)
        ex.gvdbx_example_render_custom.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_void_p]
        ex.gvdbx_example_render_custom.restype = C.c_int
        p, keep = _buf(scninfo)
        self._ck(ex.gvdbx_example_render_custom(self._h, p, chan, int(out_ptr), None), "gvdbx_example_render_custom")

    # --- frame lanes (consecutive frames on alternating internal streams)
    def lanes(self, n):
        self._ck(self._L.gvdbx_lanes(self._h, int(n)), "gvdbx_lanes")
        self.nlanes = int(n)

    def lane_select(self, lane):
        self._ck(self._L.gvdbx_lane_select(self._h, int(lane)), "gvdbx_lane_select")

    def lane_stream(self, lane):
        return int(self._L.gvdbx_lane_stream(self._h, int(lane)) or 0)

    def lanes_fork(self):
        self._ck(self._L.gvdbx_lanes_fork(self._h), "gvdbx_lanes_fork")

    def lanes_join(self):
        self._ck(self._L.gvdbx_lanes_join(self._h), "gvdbx_lanes_join")

    def render_banded(self, scninfo, shade, out_ptr, nbands, chan=0):
        """the frame as nbands horizontal bands on two alternating internal streams (for read_banded's overlapped copy)"""
        p, keep = _buf(scninfo)
        self._ck(self._L.gvdbx_render_banded(self._h, p, shade, chan, int(out_ptr), int(nbands)), "gvdbx_render_banded")

    def read_banded(self, buf_ptr, host_array):
        """synchronous copy of a frame rendered by render_banded, band after band as they finish"""
        self._ck(self._L.gvdbx_read_banded(self._h, int(buf_ptr), host_array.ctypes.data_as(C.c_void_p), host_array.nbytes), "gvdbx_read_banded")

    def read_into_async(self, buf_ptr, host_array):
        self._ck(self._L.gvdbx_read_buffer_async(self._h, int(buf_ptr), host_array.ctypes.data_as(C.c_void_p), host_array.nbytes),
                 "gvdbx_read_buffer_async")

    def tiles_per_rank(self, width, height, tile_size, nranks):
        return self._L.gvdbx_tiles_per_rank(width, height, tile_size, nranks)

    def assemble_tiles(self, gathered_ptr, frame_ptr, width, height, tile_size, nranks):
        self._ck(self._L.gvdbx_assemble_tiles(self._h, int(gathered_ptr), int(frame_ptr), width, height, tile_size, nranks),
                 "gvdbx_assemble_tiles")

    def raytrace(self, scninfo, rays_ptr, num_rays, bias, chan=0):
        """VolumeGVDB::Raytrace on a device array of 64-byte ScnRay records (in place)."""
        p, keep = _buf(scninfo)
        self._ck(self._L.gvdbx_raytrace(self._h, p, chan, int(rays_ptr), int(num_rays), C.c_float(bias)), "gvdbx_raytrace")

    def read_buffer(self, buf_ptr, nbytes):
        out = np.empty(nbytes, dtype=np.uint8)
        self._ck(self._L.gvdbx_read_buffer(self._h, int(buf_ptr), out.ctypes.data_as(C.c_void_p), nbytes), "gvdbx_read_buffer")
        return out

    def read_into(self, buf_ptr, host_array):
        self._ck(self._L.gvdbx_read_buffer(self._h, int(buf_ptr), host_array.ctypes.data_as(C.c_void_p), host_array.nbytes),
                 "gvdbx_read_buffer")

    def sync(self):
        self._ck(self._L.gvdbx_sync(self._h), "gvdbx_sync")

    def counters(self):
        c = Counters()
        self._ck(self._L.gvdbx_get_counters(self._h, C.byref(c)), "gvdbx_get_counters")
        return c.as_dict()

    def measure_tex_peak(self, lane_spacing=0.2):
        """fp32 trilinear Gsamples/s of the texture units on L1-resident bricks of the imported atlas, the 8x4 lanes of a
        warp `lane_spacing` voxels apart"""
        g = C.c_double()
        self._ck(self._L.gvdbx_measure_tex_peak(self._h, C.c_float(lane_spacing), C.byref(g)), "gvdbx_measure_tex_peak")
        return float(g.value)

    def measure_deep_loop_peak(self, lane_spacing=0.2, table_through_texture=False):
        """Gsamples/s of the deep marcher's inner loop alone (fetch + transfer index + table gather + colour update, no traversal);
        table_through_texture: A/B variant that reads the transfer table through a float4 texture object"""
        g = C.c_double()
        self._ck(self._L.gvdbx_measure_deep_loop_peak(self._h, C.c_float(lane_spacing), 1 if table_through_texture else 0, C.byref(g)), "gvdbx_measure_deep_loop_peak")
        return float(g.value)

    def measure_sampler_ab(self, lane_spacing=0.2):
        """Gsamples/s (fetch + filter only) of the four ways of reading a brick: texture unit, brick-major scalar loads,
        x-pair 8-byte loads, TMA-staged shared memory (csrc/gvdbx_microbench.cuh)"""
        g = (C.c_double * 4)()
        self._ck(self._L.gvdbx_measure_sampler_ab(self._h, C.c_float(lane_spacing), g), "gvdbx_measure_sampler_ab")
        return {"tex": g[0], "linear_ldg": g[1], "linear_pairs_ldg64": g[2], "tma_staged_smem": g[3]}

    def sample_points(self, xyz_ptr, n, out_tex_ptr, out_lin_ptr, chan=0):
        self._ck(self._L.gvdbx_sample_points(self._h, chan, int(xyz_ptr), n, int(out_tex_ptr), int(out_lin_ptr)),
                 "gvdbx_sample_points")


def render_multi(renderers, scninfo, shade, out_ptr, tile_size=32, chan=0):
    """gvdbx_render_multi: several contexts of ONE process (one per device) render one frame into renderers[0]'s buffer"""
    L = lib()
    hs = (C.c_void_p * len(renderers))(*[r._h for r in renderers])
    p, keep = _buf(scninfo)
    renderers[0]._ck(L.gvdbx_render_multi(hs, len(renderers), p, shade, chan, int(out_ptr), tile_size), "gvdbx_render_multi")


# ------------------------------------------------------------------------------------------------ host mirror
HOST_SYMBOLS = [
    "gvdbxh_create", "gvdbxh_destroy", "gvdbxh_set_transform", "gvdbxh_camera", "gvdbxh_camera_nearfar", "gvdbxh_light",
    "gvdbxh_scene_params", "gvdbxh_cross_section", "gvdbxh_linear_transfer", "gvdbxh_transfer_table", "gvdbxh_set_res", "gvdbxh_prepare_render",
    "gvdbxh_import_topology_host", "gvdbxh_import_atlas_host", "gvdbxh_commit_transfer",
    "gvdbxh_load_vbx", "gvdbxh_save_vbx", "gvdbxh_vdbinfo", "gvdbxh_set_epsilon", "gvdbxh_add_render_buf",
    "gvdbxh_render", "gvdbxh_read_render_buf", "gvdbxh_set_render_lanes", "gvdbxh_set_readback_bands", "gvdbxh_read_render_buf_async", "gvdbxh_sync_render_buf", "gvdbxh_set_option", "gvdbxh_last_error",
]


def _f(vals):
    return (C.c_float * len(vals))(*[float(v) for v in vals])


class Volume:
    """Mirror of the render-facing part of the reference's VolumeGVDB (gvdb-voxels_b200/csrc/gvdbx_host.h):
    Scene / Camera3D / Light state, SetTransform, CommitTransferFunc, AddRenderBuf, Render, ReadRenderBuf.

    device < 0 creates host state only (ScnInfo bytes can be produced without a GPU)."""

    def __init__(self, device=0):
        L = lib()
        L.gvdbxh_create.restype = C.c_void_p
        L.gvdbxh_create.argtypes = [C.c_int]
        L.gvdbxh_transfer_table.restype = C.POINTER(C.c_float)
        L.gvdbxh_last_error.restype = C.c_char_p
        for s in HOST_SYMBOLS:
            if s not in ("gvdbxh_create", "gvdbxh_transfer_table", "gvdbxh_last_error", "gvdbxh_vdbinfo", "gvdbxh_set_epsilon"):
                getattr(L, s).restype = C.c_int
        self._L = L
        self._h = C.c_void_p(L.gvdbxh_create(int(device)))
        if not self._h:
            raise GvdbxError(f"gvdbxh_create(device={device}) failed: no CUDA device / no CPU fallback")
        self._bufs = {}

    def close(self):
        if getattr(self, "_h", None):
            self._L.gvdbxh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            msg = self._L.gvdbxh_last_error(self._h)
            raise GvdbxError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    # --- Scene / camera state (reference: gvdb_scene.h, gvdb_camera.h)
    def SetTransform(self, pretrans=(0, 0, 0), scal=(1, 1, 1), angs=(0, 0, 0), trans=(0, 0, 0)):
        self._L.gvdbxh_set_transform(self._h, _f(pretrans), _f(scal), _f(angs), _f(trans))

    def SetCamera(self, fov, angs, target, dist, dolly=1.0):
        self._L.gvdbxh_camera(self._h, C.c_float(fov), _f(angs), _f(target), C.c_float(dist), C.c_float(dolly))

    def SetLight(self, angs, target, dist, dolly=1.0):
        self._L.gvdbxh_light(self._h, _f(angs), _f(target), C.c_float(dist), C.c_float(dolly))

    def SetSceneParams(self, steps, extinct, thresh, cutoff, backclr, shadow):
        self._L.gvdbxh_scene_params(self._h, _f(steps), _f(extinct), _f(thresh), _f(cutoff), _f(backclr), _f(shadow))

    def SetCrossSection(self, pnt, norm):
        """Scene::SetCrossSection (gvdb_scene.h:147): plane of SHADE_SECTION3D / origin + per-axis extent of SHADE_SECTION2D"""
        self._L.gvdbxh_cross_section(self._h, _f(pnt), _f(norm))

    def LinearTransferFunc(self, t0, t1, a, b):
        self._L.gvdbxh_linear_transfer(self._h, C.c_float(t0), C.c_float(t1), _f(a), _f(b))

    def transfer_table(self):
        p = self._L.gvdbxh_transfer_table(self._h)
        return np.ctypeslib.as_array(p, shape=(16384 * 4,)).copy()

    def SetRes(self, w, h):
        self._L.gvdbxh_set_res(self._h, int(w), int(h))

    def PrepareRender(self, w, h, shading):
        out = (C.c_uint8 * SCNINFO_BYTES)()
        self._L.gvdbxh_prepare_render(self._h, int(w), int(h), int(shading), out)
        return bytes(out)

    # --- volume import + render (needs a device)
    def ImportTopologyHost(self, vdbinfo, pool0, pool1):
        p, keep = _buf(vdbinfo)
        a0 = (C.c_void_p * 10)()
        a1 = (C.c_void_p * 10)()
        n1 = (C.c_uint64 * 10)()
        alive = []
        for lev in range(10):
            for arr, src in ((a0, pool0), (a1, pool1)):
                b = src.get(lev)
                if b is None or len(b) == 0:
                    arr[lev] = None
                    continue
                q, k = _buf(b)
                alive.append(k)
                arr[lev] = q
                if arr is a1:
                    n1[lev] = k.nbytes
        self._ck(self._L.gvdbxh_import_topology_host(self._h, p, a0, a1, n1), "ImportTopologyHost")

    def ImportAtlasHost(self, texels, chan=0):
        a = np.ascontiguousarray(texels, dtype=np.float32)
        rz, ry, rx = a.shape
        self._ck(self._L.gvdbxh_import_atlas_host(self._h, chan, a.ctypes.data_as(C.c_void_p), rx, ry, rz), "ImportAtlasHost")

    def LoadVBX(self, fname, parse_only=False):
        """VolumeGVDB::LoadVBX (gvdb_volume_gvdb.cpp:507-683): transform, pools and channel-0 atlas of a .vbx file"""
        self._ck(self._L.gvdbxh_load_vbx(self._h, str(fname).encode(), 1 if parse_only else 0), "LoadVBX")

    def SaveVBX(self, fname):
        self._ck(self._L.gvdbxh_save_vbx(self._h, str(fname).encode()), "SaveVBX")

    def SetEpsilon(self, eps, maxiter=256):
        self._L.gvdbxh_set_epsilon(self._h, C.c_float(eps), int(maxiter))

    def vdbinfo(self):
        out = (C.c_uint8 * VDBINFO_BYTES)()
        self._L.gvdbxh_vdbinfo(self._h, out)
        return bytes(out)

    def CommitTransferFunc(self):
        self._ck(self._L.gvdbxh_commit_transfer(self._h), "CommitTransferFunc")

    def AddRenderBuf(self, chan, w, h, bpp):
        self._ck(self._L.gvdbxh_add_render_buf(self._h, chan, w, h, bpp), "AddRenderBuf")
        self._bufs[chan] = (w, h, bpp)

    def Render(self, shading, chan=0, rbuf=0):
        self._ck(self._L.gvdbxh_render(self._h, int(shading), int(chan), int(rbuf)), "Render")

    def ReadRenderBuf(self, chan, out=None):
        w, h, bpp = self._bufs[chan]
        if out is None:
            out = np.empty((h, w, bpp), dtype=np.uint8)
        self._ck(self._L.gvdbxh_read_render_buf(self._h, chan, out.ctypes.data_as(C.c_void_p)), "ReadRenderBuf")
        return out

    def SetReadbackBands(self, n):
        """Render without lanes: bands per frame for the overlapped synchronous read-back (0 = automatic, 1 = off)"""
        self._ck(self._L.gvdbxh_set_readback_bands(self._h, int(n)), "SetReadbackBands")

    def SetRenderLanes(self, n):
        """extension: render buffer j lives on frame lane j % n, so frames in different render buffers overlap"""
        self._ck(self._L.gvdbxh_set_render_lanes(self._h, int(n)), "SetRenderLanes")

    def ReadRenderBufAsync(self, chan, out):
        self._ck(self._L.gvdbxh_read_render_buf_async(self._h, chan, out.ctypes.data_as(C.c_void_p)), "ReadRenderBufAsync")

    def SyncRenderBuf(self, chan):
        self._ck(self._L.gvdbxh_sync_render_buf(self._h, chan), "SyncRenderBuf")

    def set_option(self, opt, value):
        self._ck(self._L.gvdbxh_set_option(self._h, int(opt), int(value)), "set_option")
