"""Shared helpers for the test-suite (test infrastructure)."""
import hashlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
MODES = {"voxel": 0, "trilinear": 4, "levelset": 6, "deep": 7}
TINY = ["cfg1_tiny", "cfg2_tiny", "cfg3_tiny", "cfg4_tiny"]
SMALL = ["cfg1_small", "cfg2_small", "cfg3_small", "cfg4_small"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden(preset):
    return np.load(os.path.join(GOLDEN, f"ref_{preset}.npz"))


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else float(10 * np.log10(255.0 ** 2 / mse))


def mask_scninfo(b):
    """zero the bytes of ScnInfo that are uninitialised / pointers in the reference:
    bias (320..323, m_bias is never initialised), struct padding (326..327, 408..415), transfer/outbuf/dbuf pointers."""
    a = np.frombuffer(bytes(b), np.uint8).copy()
    a[320:324] = 0
    a[326:328] = 0
    a[384:416] = 0
    return a


def mask_vdbinfo(b, levels=5):
    """zero pointers / texture handles / update flag and the entries of levels the tree does not have
    (the reference leaves those uninitialised)."""
    a = np.frombuffer(bytes(b), np.uint8).copy()
    a[440:608] = 0          # nodelist, childlist, atlas_map
    a[648:672] = 0          # apron_table[2..7]: never written for apron 1
    a[684] = 0              # update
    a[686:688] = 0          # padding
    a[712:1232] = 0         # volIn / volOut / tail padding
    for off, sz, n in ((0, 4, 10), (40, 4, 10), (80, 12, 10), (200, 12, 10), (320, 4, 10), (360, 4, 10), (400, 4, 10)):
        a[off + sz * levels: off + sz * n] = 0
    return a


def tolerance_ok(mine, ref, max_over1_frac=2e-3, min_psnr=60.0):
    """north_star tolerance for the non-exact modes: <= 1/255 per channel, or PSNR >= 60 dB."""
    d = np.abs(mine.astype(np.int32) - ref.astype(np.int32)).max(axis=2)
    over1 = float((d > 1).mean())
    return (over1 == 0.0) or (psnr(mine, ref) >= min_psnr and over1 <= max_over1_frac), over1, psnr(mine, ref)


def cksum32(a, chunk=1 << 24):
    """(sum of 32-bit words, sum of word_i * (i % 65521 + 1)) mod 2^64 — the checksum oracle/ref_harness.cpp prints for the
    pools and the atlas of a --lightdump run."""
    w = np.ascontiguousarray(a).reshape(-1).view(np.uint8)
    w = w[: w.size // 4 * 4].view(np.uint32)
    s1 = s2 = 0
    for i in range(0, w.size, chunk):
        c = w[i:i + chunk].astype(np.uint64)
        k = (np.arange(i, i + c.size, dtype=np.uint64) % np.uint64(65521)) + np.uint64(1)
        s1 = (s1 + int(c.sum(dtype=np.uint64))) & 0xFFFFFFFFFFFFFFFF
        s2 = (s2 + int((c * k).sum(dtype=np.uint64))) & 0xFFFFFFFFFFFFFFFF
    return s1, s2
