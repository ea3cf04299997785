"""make_golden_hoststate.py — runs oracle/_ref/ref_hostdump (the reference's own Camera3D / Matrix4F compiled from
/root/reference) on seeded random parameter sets and writes tests/golden/hoststate.json (inputs + expected float bit
patterns).  Run in the build container (needs /root/reference); the JSON is committed."""
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rng = np.random.default_rng(20261017)
    cams, xfms = [], []
    fixed = [(50, 1024, 768, 20, 30, 0, 128, 128, 128, 500), (40, 1920, 1080, 35, 25, 0, 512, 512, 512, 1800),
             (40, 3840, 2160, 60, 20, 0, 1024, 1024, 1024, 3600), (40, 800, 600, 0, 45, 0, 0, 0, 0, 120)]
    cams += [list(map(float, f)) for f in fixed]
    for _ in range(60):
        w, h = rng.choice([(640, 480), (1024, 768), (1920, 1080), (3840, 2160), (333, 777)])
        cams.append([float(np.float32(rng.uniform(20, 90))), float(w), float(h), float(np.float32(rng.uniform(-180, 360))),
                     float(np.float32(rng.uniform(-85, 85))), 0.0, *[float(np.float32(x)) for x in rng.uniform(-500, 2500, 3)],
                     float(np.float32(rng.uniform(50, 5000)))])
    xfms.append([0, 0, 0, 1, 1, 1, 0, 0, 0, 0, 0, 0])
    for _ in range(40):
        xfms.append([*[float(np.float32(x)) for x in rng.uniform(-100, 100, 3)], *[float(np.float32(x)) for x in rng.uniform(0.25, 4, 3)],
                     *[float(np.float32(x)) for x in rng.uniform(-180, 180, 3)], *[float(np.float32(x)) for x in rng.uniform(-300, 300, 3)]])
    lines = ["cam " + " ".join(repr(v) for v in c) for c in cams] + ["xfm " + " ".join(repr(v) for v in x) for x in xfms]
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_hostdump")], input="\n".join(lines) + "\n",
                       stdout=subprocess.PIPE, text=True, check=True)
    outs = r.stdout.strip().splitlines()
    assert len(outs) == len(lines)
    gold = {"cam": [{"in": c, "out": o.split()} for c, o in zip(cams, outs[:len(cams)])],
            "xfm": [{"in": x, "out": o.split()} for x, o in zip(xfms, outs[len(cams):])]}
    json.dump(gold, open(os.path.join(ROOT, "tests", "golden", "hoststate.json"), "w"))
    print("wrote", len(cams), "cameras,", len(xfms), "transforms")


if __name__ == "__main__":
    main()
