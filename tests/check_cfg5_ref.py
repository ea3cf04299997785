"""check_cfg5_ref.py — BASELINE.json config 5 at FULL size against the UNMODIFIED reference: the reference builds and
renders the 2.05 M-brick volume itself (ref_harness --lightdump: images + ScnInfo only), the product renders the
byte-identical volume built by the CPU restatement with the reference's ScnInfo.  Writes PNGs + a JSON summary."""
import json
import os
import struct
import subprocess
import sys
import tempfile
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def write_png(path, img):
    h, w, _ = img.shape
    raw = b"".join(b"\x00" + img[y, :, :3].tobytes() for y in range(h))

    def chunk(t, d):
        c = struct.pack(">I", len(d)) + t + d
        return c + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    open(path, "wb").write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
                           chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def main():
    import numpy as np
    import torch
    import bench
    import oracle
    import refcmp
    from common import psnr
    preset = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
    size = (480, 270)
    pkg = bench.load_pkg()
    out = {"preset": preset}
    d = tempfile.mkdtemp(prefix="refdump_")
    t0 = time.perf_counter()
    cmd = ["./ref_harness", preset, d, "--modes", "deep,deepspp,deepshadow", "--size", f"{size[0]}x{size[1]}", "--lightdump", "--hits", "0"]
    r = subprocess.run(cmd, cwd=refcmp.REF_DIR, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=1500)
    out["ref_seconds"] = round(time.perf_counter() - t0, 1)
    if r.returncode != 0:
        print("ref_harness failed:", r.stderr[-1500:])
        return 1
    out["ref_timing"] = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])["render"]
    p, vol = bench.build_workload(preset, None, {})
    ren = pkg.Renderer(0)
    ren.import_topology_host(vol["vdbinfo"], vol["pool0"], vol["pool1"])
    ren.import_atlas_host(vol["atlas"])
    _, table = oracle.scninfo_for(pkg, p)
    ren.set_transfer(table)
    vol["transfer"] = table
    w, h = size
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for m in ("deep", "deepspp", "deepshadow"):
        scn = open(os.path.join(d, f"scninfo_{m}.bin"), "rb").read()
        ref = np.fromfile(os.path.join(d, f"out_{m}.rgba"), dtype=np.uint8).reshape(h, w, 4)
        _, dshadow, spp = refcmp.MODES2.get(m, (7, 0, 1))
        ren.set_deep_shadow(dshadow); ren.set_spp(spp)
        img = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
        ren.render(scn, 7, img.data_ptr())
        ren.sync()
        mine = img.cpu().numpy()
        diff = np.abs(mine.astype(int) - ref.astype(int)).max(axis=2)
        out[m] = {"mismatch_pixels": int((diff > 0).sum()), "max_abs": int(diff.max()), "psnr": round(psnr(mine, ref), 1),
                  "nonbackground": int((ref != ref[0, 0]).any(axis=2).sum())}
        write_png(os.path.join(ROOT, "gpurun_out", f"cfg5_{m}_mine.png"), mine)
        write_png(os.path.join(ROOT, "gpurun_out", f"cfg5_{m}_ref.png"), ref)
        if m == "deep":
            ren.set_deep_shadow(0); ren.set_spp(1)
            cpu = oracle.render(vol, scn, 7)
            out["cpu_oracle_vs_ref_psnr"] = round(psnr(cpu, ref), 1)
            write_png(os.path.join(ROOT, "gpurun_out", "cfg5_deep_cpu.png"), cpu)
    print(json.dumps(out))
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{preset}_ref_parity.json"), "w"), indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
