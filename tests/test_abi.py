"""C-ABI surface: the shared library loads and exports every symbol include/gvdbx.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from common import ROOT


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.lib()
    hdr = open(os.path.join(ROOT, "include", "gvdbx.h")).read()
    declared = sorted(set(re.findall(r"\b(gvdbx_[a-z_]+)\s*\(", hdr)))
    assert len(declared) >= 18
    for s in declared:
        assert hasattr(L, s), s
    assert sorted(declared) == sorted(pkg.EXPORTED_SYMBOLS)
    for s in pkg.HOST_SYMBOLS:
        assert hasattr(L, s), s


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "gvdbx.h"\nint main(void){ gvdbx_counters c; (void)c; return GVDBX_VDBINFO_BYTES == 1232 ? 0 : 1; }\n')
    import subprocess
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(tmp_path / "t")], check=True)
    subprocess.run([str(tmp_path / "t")], check=True)


def test_tiles_per_rank_arithmetic(pkg):
    L = pkg.lib()
    assert L.gvdbx_tiles_per_rank(1920, 1080, 32, 1) == 60 * 34
    assert L.gvdbx_tiles_per_rank(1920, 1080, 32, 8) == (60 * 34 + 7) // 8
    assert L.gvdbx_tiles_per_rank(3840, 2160, 32, 8) == (120 * 68 + 7) // 8
    assert L.gvdbx_tiles_per_rank(0, 10, 32, 1) < 0


def test_no_cpu_fallback_without_device(pkg):
    """Without a GPU the render context cannot be created (and never silently falls back to the CPU oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.GvdbxError):
        pkg.Renderer(0)
    h = ctypes.c_void_p()
    assert pkg.lib().gvdbx_create(ctypes.byref(h), 0, None) == -2


def test_product_does_not_reference_oracle():
    """The product tree must not import / link / execute anything under oracle/."""
    pkgdir = os.path.join(ROOT, "gvdb-voxels_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "gvdb_oracle" not in text and "oracle/" not in text.replace("oracle/_ref/ref_hostdump", ""), f


def test_python_constants_match_header(pkg):
    """shade modes / options / sizes of the ctypes front end are the header's"""
    hdr = open(os.path.join(ROOT, "include", "gvdbx.h")).read()
    defs = {k: int(v) for k, v in re.findall(r"#define\s+(GVDBX_[A-Z_0-9]+)\s+(-?\d+)\b", hdr)}
    for name in ("VOXEL", "SECTION2D", "SECTION3D", "EMPTYSKIP", "TRILINEAR", "TRICUBIC", "LEVELSET", "VOLUME", "OFF"):
        assert getattr(pkg, f"SHADE_{name}") == defs[f"GVDBX_SHADE_{name}"], name
    for name in ("SAMPLER", "BLOCK_W", "BLOCK_H", "COUNTERS", "TRAVERSAL", "CULL", "SPP", "DEEP_SHADOW"):
        assert getattr(pkg, f"OPT_{name}") == defs[f"GVDBX_OPT_{name}"], name
    assert pkg.VDBINFO_BYTES == defs["GVDBX_VDBINFO_BYTES"] and pkg.SCNINFO_BYTES == defs["GVDBX_SCNINFO_BYTES"]
    # every option number is unique
    opts = [v for k, v in defs.items() if k.startswith("GVDBX_OPT_")]
    assert len(opts) == len(set(opts))
