"""gvdb-voxels_b200 — B200-native (sm_100a) ray-cast render path of NVIDIA/gvdb-voxels.

Thin ctypes front end over the C ABI in include/gvdbx.h (libgvdbx.so, built in-tree by
`make -C gvdb-voxels_b200` or `__graft_entry__.build()`).  The directory name contains a
hyphen, so it is imported through `importlib` as module ``gvdb_voxels_b200``
(see __graft_entry__.load_package()).

There is no CPU fallback: every call that needs the device raises GvdbxError when the
shared library or a CUDA device is missing.
"""
from .api import (  # noqa: F401
    GvdbxError, Renderer, Volume, lib, lib_path, HOST_SYMBOLS, render_multi,
    SHADE_VOXEL, SHADE_TRILINEAR, SHADE_LEVELSET, SHADE_VOLUME, SHADE_OFF,
    SHADE_SECTION2D, SHADE_SECTION3D, SHADE_EMPTYSKIP, SHADE_TRICUBIC,
    OPT_SAMPLER, OPT_BLOCK_W, OPT_BLOCK_H, OPT_COUNTERS, OPT_TRAVERSAL, OPT_CULL, OPT_SPP, OPT_DEEP_SHADOW,
    SAMPLER_TEX, SAMPLER_LINEAR, VDBINFO_BYTES, SCNINFO_BYTES,
    EXPORTED_SYMBOLS,
)
