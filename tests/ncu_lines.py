"""ncu_lines.py — per-source-line share of executed warp instructions / active lanes / stall samples of one kernel:
the SASS page of an ncu report (--import-source on) joined, instruction by instruction, with the line table nvdisasm -g
prints for the same kernel of the object file it was built from (-lineinfo).

  python tests/ncu_lines.py <report.ncu-rep> <object.o> [min_pct]     (runs here, no GPU needed)"""
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    kernel = rows[0][1]
    hdr = {h: i for i, h in enumerate(rows[1])}
    return kernel, hdr, rows[2:]


def mangle(kernel):
    m = re.search(r"gx_render_kernel<\(int\)(\d+), \(int\)(\d+), \(int\)(\d+), \(bool\)(\d)>", kernel)
    return "_Z16gx_render_kernelILi%sELi%sELi%sELb%sEEv8GxParams" % m.groups()


def line_table(obj, sym):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(d, cubin)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    lines, cur, on = [], None, False
    for l in txt.splitlines():
        if l.startswith(".text."):
            on = (l.strip() == f".text.{sym}:")
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            lines.append(cur)
    return lines


def main():
    rep, obj = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.8
    kernel, hdr, rows = sass_rows(rep)
    table = line_table(obj, mangle(kernel))
    print(kernel, "| sass rows", len(rows), "| line table", len(table))
    n = min(len(rows), len(table))
    src = {}
    for fn in set(t[0] for t in table if t):
        for root in ("gvdb-voxels_b200/csrc",):
            p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), root, fn)
            if os.path.exists(p):
                src[fn] = open(p).read().splitlines()
    agg = {}
    for i in range(n):
        r = rows[i]
        inst = float(r[hdr["Instructions Executed"]] or 0); thr = float(r[hdr["Thread Instructions Executed"]] or 0)
        smp = float(r[hdr["# Samples"]] or 0)
        a = agg.setdefault(table[i], [0.0, 0.0, 0.0, 0])
        a[0] += inst; a[1] += thr; a[2] += smp; a[3] += 1
    tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
    print(f"warp instructions {tot:.4g}  lanes/instr {sum(a[1] for a in agg.values()) / tot:.2f}  stall samples {tots:.0f}")
    print("  file:line                     %instr  lanes  %samples  static  source")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][0]):
        if 100 * a[0] / tot < min_pct:
            break
        f, l = k if k else ("?", 0)
        text = src.get(f, [""] * (l + 1))[l - 1].strip()[:100] if f in src and l - 1 < len(src[f]) else ""
        print(f"  {f[:22]:22s}:{l:<5d} {100 * a[0] / tot:6.2f}  {a[1] / max(a[0], 1):5.1f}  {100 * a[2] / max(tots, 1):7.2f}  {a[3]:5d}   {text}")


if __name__ == "__main__":
    main()
