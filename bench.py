#!/usr/bin/env python
"""bench.py — Mrays/s and ms/frame of the GVDB ray-cast render path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]          headline + (N = 1) a table over BASELINE configs 1-4
  python bench.py --impl reference ...                         the reference's own CUDA render of the same workloads (oracle/_ref)
  torchrun --nproc-per-node N bench.py --gpus N ...            one rank per GPU, image tiles, peer frame ring over NVLink

Headline workload (BASELINE.json configs[3], the 4K multi-GPU config): `cfg4` = deep semitransparent density + shadow rays
(gvdbRayDeep + rayShadowBrick) on the 1024^3 noise cloud of SURVEY.md 8d, 3840x2160.  A step = F frames (camera yaw +
360 * j / F around the preset's orbit).  `value` = primary rays / time of the K timed steps, everything resident in HBM
(max over ranks, CUDA events, barrier + synchronize on both sides), frames pipelined over `frame_lanes` streams;
`latency_ms_1lane` = one frame on one stream; `e2e` = the same frames through the reference-facing API (VolumeGVDB mirror:
SetCamera + Render + ReadRenderBuf) with HOST buffers, pipelined (ReadRenderBufAsync into pinned memory); `e2e_strict` = the
strict drop-in sequence Render() + synchronous ReadRenderBuf() into pageable memory, one buffer, one stream.

Before anything is timed, frame 0 of every measured workload is compared byte for byte with the UNMODIFIED reference
rendering the same frame now (oracle/_ref/ref_harness): `parity_checked`.

oracle/ is used here only (a) to synthesise the input volume (the CPU topology build is one of the two reported baselines),
(b) as the `cpu_baseline` leg, (c) by `--impl reference` and (d) for the parity check; never inside a timed GPU region.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

# render and consumer streams should not share a hardware queue (the default of 8 connections is spread over torch's
# pool of 32 streams): must be set before the CUDA context exists
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# mode name -> (reference shade mode, deep + shadow composition, rays per pixel)
MODE = {"voxel": (0, 0), "trilinear": (4, 0), "levelset": (6, 0), "deep": (7, 0), "deepshadow": (7, 1)}
HEADLINE = ("cfg4", "deepshadow")
TABLE = [("cfg1", "trilinear"), ("cfg2", "levelset"), ("cfg3", "voxel"), ("cfg4", "deep")]
WHAT = {"cfg1": "BASELINE configs[0]: 256^3 sphere density, SHADE_TRILINEAR",
        "cfg2": "BASELINE configs[1]: 1024^3 noise-displaced sphere SDF, SHADE_LEVELSET",
        "cfg3": "BASELINE configs[2]: 2048^3 sparse grid of solid balls, SHADE_VOXEL",
        "cfg4": "BASELINE configs[3]: 1024^3 noise cloud, deep emission / absorption with transfer function",
        "cfg5": "BASELINE configs[4]: 4096^3 noise cloud (~8 GB atlas), deep"}
# algorithmic bytes per unit of work
B_TRI, B_PT, B_DDA, B_DESC, B_PIX = 32, 4, 8, 64, 4
N_SM, ISSUE_PER_SM = 148, 4            # B200: 148 SMs x 4 warp schedulers, one warp instruction per scheduler and cycle


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), float(p.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1965.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernel_counters(key):
    """per-launch hardware counters of the final kernels from committed ncu exports (profiles/r02_kernel_counters.json,
    written by tests/ncu_extract.py): warp instructions executed, DRAM bytes, L1 / L2 hit rates, texture pipe"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_kernel_counters.json"))).get(key)
    except Exception:
        return None


def workload_config(workload, mode, w, h, frames, spp):
    """identical in both arms (the driver compares the dicts)"""
    return {"workload": f"{workload} {mode} {w}x{h}", "what": WHAT.get(workload[:4], workload) + (" + one shadow march per pixel" if mode == "deepshadow" else ""),
            "frames_per_step": frames, "spp": spp,
            "l2_policy": "inputs larger than L2 (brick atlas 40 MB - 2.3 GB vs 126 MB L2 for cfg2-cfg4) and the camera moves every frame; cfg1's 40 MB atlas is L2-resident by nature of the config"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu=0):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [c.strip() for c in line.split(",")][1:])

    def stop(self, t0=None, t1=None):
        """summary of the samples taken between host times t0 and t1 (the timed region); when the region is shorter than
        the 100 ms sampling period, of the samples within half a second around it (same load: warm-up / e2e steps)"""
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        rows, window = self.rows, "all"
        if t0 is not None:
            inside = [r for r in self.rows if t0 <= r[0] <= t1]
            near = [r for r in self.rows if t0 - 0.5 <= r[0] <= t1 + 0.5]
            rows, window = (inside, "timed region") if len(inside) >= 2 else (near, "timed region +- 0.5 s")
        sm, mx, pw, reasons = [], [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def load_pkg():
    from __graft_entry__ import load_package
    return load_package()


def build_workload(name, size=None, timing=None):
    import oracle
    t0 = time.perf_counter()
    p = oracle.preset(name)
    pos, vals = oracle.generate(p)
    t_gen = time.perf_counter() - t0
    tm = {}
    vol = oracle.build_volume(pos, vals, epsilon=p.epsilon, timing=tm)
    if timing is not None:
        timing.update(scene_gen_s=t_gen, topology_build_s=tm["topology_build_s"], bricks=int(len(pos)))
    if size:
        p.width, p.height = size
    return p, vol


def frame_scninfos(pkg, p, shade, frames):
    import oracle
    out = []
    table = None
    for j in range(frames):
        angs = (p.cam_angs[0] + 360.0 * j / frames, p.cam_angs[1], p.cam_angs[2])
        scn, table = oracle.scninfo_for(pkg, p, shade=shade, cam_angs=angs)
        out.append(scn)
    return out, table


# ------------------------------------------------------------------------------------------------ reference arm
def ref_bench(workload, mode, w, h, frames, steps, warmup, spp=1):
    """ref_harness --bench: the UNMODIFIED reference's own CUDA kernels through VolumeGVDB::Render (RenderKernel for the
    composed deep + shadow kernel), driven by one host thread; returns Mrays/s for Render() alone and for Render() +
    ReadRenderBuf() per frame, or None when oracle/_ref is not built"""
    refdir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(refdir, "ref_harness")):
        return None
    cmd = ["./ref_harness", workload, tempfile.mkdtemp(prefix="ref_bench_"), "--bench", "--orbit", str(frames), "--steps", str(steps),
           "--warmup", str(warmup), "--mode", mode, "--size", f"{w}x{h}", "--spp", str(spp)]
    r = subprocess.run(cmd, cwd=refdir, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stderr[-2000:])
        return None
    j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    rays = frames * w * h * spp * steps
    return {"value": rays / j["render_s"] / 1e6, "e2e": rays / j["e2e_s"] / 1e6, "ms_per_frame": j["render_s"] / steps / frames * 1e3,
            "e2e_ms_per_frame": j["e2e_s"] / steps / frames * 1e3, "topology_build_s": j["topology_build_s"], "bricks": j["bricks"]}


def run_reference(a):
    """--impl reference: the reference's own CUDA render of the same workloads through its public API (oracle/_ref/libgvdb.so,
    unmodified; GVDB ships no CPU ray marcher, north_star names this as the baseline), or — if oracle/_ref was not built — the
    CPU port of the oracle on a bounded sample.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    import oracle
    p = oracle.preset(a.workload)
    w, h = a.size if a.size else (p.width, p.height)
    shade, dshadow = MODE[a.mode]
    rays = a.frames * w * h * a.spp
    base = {"metric": "Mrays/s", "unit": "Mrays/s", "impl": "reference", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a.workload, a.mode, w, h, a.frames, a.spp)}
    res = ref_bench(a.workload, "deepspp" if (a.spp > 1 and shade == 7) else a.mode, w, h, a.frames, a.steps, a.warmup, a.spp)
    if res is not None:
        sample = ("unmodified reference CUDA kernels (GVDB ships no CPU ray marcher) driven by 1 host thread through VolumeGVDB::Render "
                  "(RenderKernel + kernel composed from the reference's own device functions for deep + shadow) on 1 GPU; e2e adds the reference's "
                  "synchronous ReadRenderBuf into pageable memory per frame")
        base.update(value=res["value"], ms_per_step=res["ms_per_frame"] * a.frames, ms_per_frame=res["ms_per_frame"],
                    e2e={"value": res["e2e"], "unit": "Mrays/s", "h2d_bytes_per_step": a.frames * 416, "d2h_bytes_per_step": a.frames * w * h * 4,
                         "ms_per_frame": res["e2e_ms_per_frame"], "api": "VolumeGVDB::Render + ReadRenderBuf (stock libgvdb)"},
                    cpu_baseline={"value": res["value"], "unit": "Mrays/s", "cores": 1, "kind": "reference", "sample": sample,
                                  "topology_build_s": res["topology_build_s"], "bricks": res["bricks"]},
                    gpu_launches=0)
        if a.table:
            tab = []
            for wl, mode in TABLE:
                q = oracle.preset(wl)
                r = ref_bench(wl, mode, q.width, q.height, a.frames, max(2, a.steps // 4), 1)
                if r is not None:
                    tab.append({"workload": f"{wl} {mode} {q.width}x{q.height}", "value": r["value"], "ms_per_frame": r["ms_per_frame"],
                                "e2e_strict": r["e2e"], "topology_build_s": r["topology_build_s"], "bricks": r["bricks"]})
            base["configs"] = tab
        print(json.dumps(base))
        return 0
    # fallback: CPU port on a bounded sample
    pkg = load_pkg()
    p, vol = build_workload(a.workload, a.size)
    scns, table = frame_scninfos(pkg, p, shade, a.frames)
    vol["transfer"] = table
    rows = max(8, min(h, 64))
    y0 = h // 2 - rows // 2
    t0 = time.perf_counter()
    for s in range(a.steps):
        oracle.render(vol, scns[s % len(scns)], shade, rows=(y0, y0 + rows), deep_shadow=bool(dshadow))
    dt = time.perf_counter() - t0
    v = rows * w * a.steps / dt / 1e6
    base.update(value=v, ms_per_step=dt / a.steps * 1e3,
                e2e={"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                cpu_baseline={"value": v, "unit": "Mrays/s", "cores": oracle.lib().ora_max_threads(), "kind": "port",
                              "sample": f"{rows} centre rows of one frame per step (CPU port, OpenMP)"}, gpu_launches=0)
    print(json.dumps(base))
    return 0


# ------------------------------------------------------------------------------------------------ parity gate
def parity_check(torch, r, workload, mode, w, h, scn0, dev):
    """frame 0 of the orbit (the preset camera): OUR ScnInfo (host mirror) through gvdbx_render vs the bytes the UNMODIFIED
    reference renders for the same preset now.  Also proves the imported volume identical (checksums of pools and atlas)."""
    import numpy as np
    import refcmp
    if not refcmp.have_ref():
        return {"parity_checked": False, "why": "oracle/_ref not built"}
    d = tempfile.mkdtemp(prefix="ref_parity_")
    t0 = time.perf_counter()
    try:
        refcmp.run_ref(workload, d, modes=[mode], size=(w, h), lightdump=True, hits=False, timeout=1500)
    except Exception as e:          # noqa
        return {"parity_checked": False, "why": f"ref_harness failed: {e}"[:300]}
    light = refcmp.load_lightdump(d)
    shade, dshadow = MODE[mode]
    out = torch.zeros((h, w, 4), dtype=torch.uint8, device=dev)
    r.set_deep_shadow(dshadow)
    r.render(scn0, shade, out.data_ptr())
    r.sync()
    mine = out.cpu().numpy()
    ref = light["rgba"][mode]
    diff = int((mine != ref).any(axis=2).sum())
    return {"parity_checked": diff == 0, "pixels_differing": diff, "pixels": w * h, "nonbackground": int((ref != ref[0, 0]).any(axis=2).sum()),
            "against": "oracle/_ref/ref_harness: the unmodified reference building and rendering the same preset itself, same frame",
            "scninfo_identical": bool(np.array_equal(np.frombuffer(scn0, np.uint8)[:320], np.frombuffer(light["scn"][mode], np.uint8)[:320])),
            "seconds": round(time.perf_counter() - t0, 1)}


# ------------------------------------------------------------------------------------------------ single-GPU measurements
def time_resident(torch, r, scns, shade, frames_d, lanes, steps, warmup):
    """K steps of F frames, frames alternating over `lanes` internal streams (0 = the creation stream only); CUDA events on
    the creation stream, which forks / joins the lanes.  Returns (ms_total, launches)."""
    nbuf = len(frames_d)
    kpf = 2 if shade == 7 else 1        # kernels per frame: deep modes build the frame's derived transfer table first

    def step():
        for j, scn in enumerate(scns):
            r.lane_select(j % nbuf if lanes else -1)
            r.render(scn, shade, frames_d[j % nbuf].data_ptr())
    r.lanes_fork()
    for _ in range(warmup):
        step()
    r.lanes_join()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r.lanes_fork()
    for _ in range(steps):
        step()
    r.lanes_join()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), steps * len(scns) * kpf


def time_latency(torch, r, scns, shade, frame_d, reps=3):
    """one frame, one stream: per-frame device time (CUDA events around each frame of the orbit), no overlap of any kind"""
    r.lane_select(-1)
    for scn in scns:
        r.render(scn, shade, frame_d.data_ptr())
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        for scn in scns:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r.render(scn, shade, frame_d.data_ptr())
            e1.record()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
    ms.sort()
    return {"mean": sum(ms) / len(ms), "min": ms[0], "median": ms[len(ms) // 2], "max": ms[-1], "frames": len(ms)}


def make_mirror(pkg, p, vol, dev_index, a, dshadow):
    v = pkg.Volume(dev_index)
    v.ImportTopologyHost(vol["vdbinfo"], vol["pool0"], vol["pool1"])
    v.ImportAtlasHost(vol["atlas"])
    v.SetSceneParams(list(p.steps), list(p.extinct), list(p.thresh), list(p.cutoff), list(p.backclr), list(p.shadow))
    if p.transfer == 1:
        v.LinearTransferFunc(0.00, 0.25, (0, 0, 0, 0), (1, 1, 0, 0.1))
        v.LinearTransferFunc(0.25, 0.50, (1, 1, 0, 0.4), (1, 0, 0, 0.3))
        v.LinearTransferFunc(0.50, 0.75, (1, 0, 0, 0.3), (.2, .2, 0.2, 0.1))
        v.LinearTransferFunc(0.75, 1.00, (.2, .2, 0.2, 0.1), (0, 0, 0, 0.0))
    v.CommitTransferFunc()
    v.SetLight(list(p.light_angs), list(p.light_target), p.light_dist)
    v.set_option(1, 0 if a.sampler == "tex" else 1)
    bw, bh = (int(x) for x in a.block.split("x"))
    v.set_option(2, bw); v.set_option(3, bh)
    v.set_option(7, a.spp)
    v.set_option(8, 1 if dshadow else 0)
    return v


def time_e2e(torch, pkg, p, vol, dev_index, a, shade, dshadow, frames, steps, lanes, bands=0):
    """through the reference-facing API (VolumeGVDB mirror) with HOST buffers.  lanes > 0: pipelined — render buffer k lives on
    frame lane k, ReadRenderBufAsync into pinned host memory, SyncRenderBuf hands the frame to the caller before its buffer is
    reused.  lanes == 0: the strict drop-in sequence — one render buffer, Render() then the synchronous ReadRenderBuf() into
    pageable memory, exactly the calls a reference application makes."""
    import numpy as np
    w, h = p.width, p.height
    v = make_mirror(pkg, p, vol, dev_index, a, dshadow)
    nbuf = max(1, lanes)
    for k in range(nbuf):
        v.AddRenderBuf(k, w, h, 4)
    v.SetRenderLanes(lanes)
    v.SetReadbackBands(bands)          # strict sequence only: 0 = automatic (~2 MB per band), 1 = one launch per frame
    if lanes:
        hosts = [torch.empty((h, w, 4), dtype=torch.uint8).pin_memory().numpy() for _ in range(nbuf)]
    else:
        hosts = [np.empty((h, w, 4), np.uint8)]

    def cam(j):
        v.SetCamera(p.fov, (p.cam_angs[0] + 360.0 * j / frames, p.cam_angs[1], p.cam_angs[2]), list(p.cam_target), p.cam_dist)
        v.SetRes(w, h)

    def step():
        if lanes:
            for j in range(frames):
                k = j % nbuf
                if j >= nbuf:
                    v.SyncRenderBuf(k)
                cam(j)
                v.Render(shade, 0, k)                       # PrepareRender: 416-byte ScnInfo host -> device with the launch
                v.ReadRenderBufAsync(k, hosts[k])           # device -> pinned host behind the kernel, on the buffer's lane
            for k in range(min(nbuf, frames)):
                v.SyncRenderBuf(k)
        else:
            for j in range(frames):
                cam(j)
                v.Render(shade, 0, 0)
                v.ReadRenderBuf(0, hosts[0])                # synchronous, pageable
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    v.close()
    rays = frames * w * h * a.spp
    return {"value": rays * steps / dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": frames * 416, "d2h_bytes_per_step": frames * w * h * 4,
            "ms_per_frame": dt / steps / frames * 1e3,
            "api": (f"VolumeGVDB mirror: SetCamera + Render(rbuf = frame % {nbuf}) + ReadRenderBufAsync into pinned host memory + SyncRenderBuf" if lanes
                    else "VolumeGVDB mirror, strict drop-in sequence: SetCamera + Render(rbuf 0) + synchronous ReadRenderBuf into pageable host memory "
                         "(the library renders the frame in ~2 MB bands on two internal streams and copies each band as it finishes)")}


def count_work(r, scns, shade, frame_d, w, h, mode):
    """SURVEY 8(d) units per frame from the counted kernel variant: mode 1 = the algorithm as the reference runs it (no brick
    culling), mode 2 = what the production kernel really does (culling on)"""
    out = []
    r.set_counters(mode)
    for scn in scns:
        r.render(scn, shade, frame_d.data_ptr())
        out.append(r.counters())
    r.set_counters(0)
    tot = {k: sum(c[k] for c in out) for k in out[0]}
    alg = B_TRI * tot["s_tri"] + B_PT * tot["s_pt"] + B_DDA * tot["n_dda"] + B_DESC * tot["n_desc"] + B_PIX * w * h * len(scns)
    return tot, float(alg)


def tex_peaks(r, p):
    """texture-unit rate (fp32 trilinear Gsamples/s, fetch-only microbenchmark on L1-resident bricks) for lanes 0.2 voxel apart
    (best case) and for this camera's ray spacing (voxels per pixel at the orbit centre)"""
    import math
    spacing = 2.0 * p.cam_dist * math.tan(math.radians(p.fov) / 2.0) / p.height
    out = {"coherent": r.measure_tex_peak(0.2), "at_ray_spacing": r.measure_tex_peak(spacing), "ray_spacing_voxels": spacing}
    try:        # the deep marcher's inner loop alone (needs a transfer function): fetch + table gather + colour update, no traversal
        out["sample_loop"] = r.measure_deep_loop_peak(spacing)
    except Exception:
        out["sample_loop"] = None
    return out


def roofline_block(key, mode, kernel, ms_total, frames_total, n_gpus, sm_mhz, alg_ref, alg_prod, units_ref, units_prod, tex_peak):
    """Which unit binds depends on the mode (DRAM never does: it sits below 1 % of peak because neighbouring rays share bricks
    and 98-99 % of the sectors hit in L1):
      deep modes    : the TEXTURE units — achieved = fp32 trilinear samples the production kernel takes (counted) / live time,
                      peak = gvdbx_measure_tex_peak run in this process on this GPU (L1-resident bricks, fetches only);
      surface modes : instruction ISSUE — achieved = warp instructions per frame (ncu sm__inst_executed of the committed capture
                      of this kernel on this input; deterministic) x frames / live time, peak = 148 SMs x 4 schedulers x SM
                      clock measured during the run.
    The HBM figures the task's contract asks for are kept next to them: DRAM-measured traffic, and the SURVEY 8(d) algorithmic
    bytes as an EFFECTIVE bandwidth (cache-served, can exceed the copy peak)."""
    hbm_peak, sm_max, src = peaks()
    kc = kernel_counters(key)
    t = ms_total * 1e-3
    clock = (sm_mhz or sm_max) * 1e6
    peak_issue = N_SM * ISSUE_PER_SM * clock * n_gpus / 1e9          # G warp instructions / s, whole job
    issue = {"unit": "Ginst/s", "peak": peak_issue, "achieved": None, "frac": None,
             "peak_source": f"{N_SM} SMs x {ISSUE_PER_SM} warp schedulers x {clock / 1e6:.0f} MHz (nvidia-smi during the run) x {n_gpus} GPU(s)"}
    if kc:
        inst = float(kc["inst_executed"])
        issue.update(achieved=inst * frames_total / t / 1e9, inst_executed_per_frame=inst, lanes_per_instruction=kc.get("lanes_per_inst"),
                     ipc_ncu=kc.get("ipc"), tex_pipe_pct_ncu=kc.get("tex_pipe_pct"), l1_hit_pct=kc.get("l1_hit_pct"), l2_hit_pct=kc.get("l2_hit_pct"),
                     counters_from=kc.get("source"))
        issue["frac"] = issue["achieved"] / peak_issue
    tex = {"unit": "Gsamples/s", "peak": (tex_peak["at_ray_spacing"] * n_gpus) if tex_peak else None,
           "achieved": units_prod["s_tri"] / t / 1e9 if units_prod else None,
           "peak_coherent": (tex_peak["coherent"] * n_gpus) if tex_peak else None, "ray_spacing_voxels": tex_peak["ray_spacing_voxels"] if tex_peak else None,
           "tex_pipe_busy_pct_ncu": kc.get("tex_pipe_pct") if kc else None,
           "peak_source": "gvdbx_measure_tex_peak, measured in this run: fp32 trilinear fetches only (8 in flight per thread) on L1-resident bricks of "
                          "this atlas with the 8x4 lanes of a warp as far apart as this camera's rays (voxels per pixel at the orbit centre); "
                          "peak_coherent = the same with lanes 0.2 voxel apart"}
    tex["frac"] = (tex["achieved"] / tex["peak"]) if tex["peak"] and tex["achieved"] else None
    if tex_peak and tex_peak.get("sample_loop"):
        tex["peak_sample_loop"] = tex_peak["sample_loop"] * n_gpus
        tex["frac_of_sample_loop"] = (tex["achieved"] / tex["peak_sample_loop"]) if tex["achieved"] else None
        tex["sample_loop_is"] = ("gvdbx_measure_deep_loop_peak, measured in this run: the deep marcher's four-sample round alone (4 fetches, 4 transfer indices, "
                                 "4 table gathers of 16 bytes, 4 colour updates with their separately rounded products), no traversal, no brick changes, every lane "
                                 "busy, L1-resident bricks: achieved / this = what traversal, brick changes, lane imbalance and cache misses cost")
    dram = (kc["dram_bytes"] * frames_total / t / 1e9) if kc and kc.get("dram_bytes") else None
    hbm = {"peak": hbm_peak * n_gpus, "unit": "GB/s", "peak_source": src,
           "dram_achieved": dram, "dram_frac": (dram / (hbm_peak * n_gpus)) if dram else None,
           "effective": alg_ref / t / 1e9, "effective_frac": alg_ref / t / 1e9 / (hbm_peak * n_gpus),
           "effective_culled": alg_prod / t / 1e9,
           "note": "effective = SURVEY 8(d) algorithmic bytes (32 B per trilinear sample, 4 B per point sample, 8 B per DDA step, 64 B per node "
                   "record, 4 B per pixel) of the algorithm as the reference runs it / time; effective_culled = the same units counted in the "
                   "production kernel (value-range culling and occupancy bits skip work); both are cache-served: only `dram_achieved` reaches HBM",
           "units_timed_region": units_ref, "units_timed_region_culled": units_prod,
           "bytes_per_unit": {"s_tri": B_TRI, "s_pt": B_PT, "n_dda": B_DDA, "n_desc": B_DESC, "pixel": B_PIX}}
    bound = "tex" if mode in ("deep", "deepshadow") else "issue"
    top = tex if bound == "tex" else issue
    return {"bound": bound, "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"], "frac": top["frac"],
            "traffic": kc.get("dram_bytes") if kc else None, "kernel": kernel, "peak_source": top["peak_source"],
            "issue": issue, "tex": tex, "hbm": hbm}


def measure_single(torch, pkg, a, workload, mode, dev, local, steps, warmup, full):
    """everything measured on one GPU for one workload; full = headline (adds pipelined e2e + work counts)"""
    import oracle
    shade, dshadow = MODE[mode]
    timing = {}
    p, vol = build_workload(workload, a.size if full else None, timing)
    w, h = p.width, p.height
    scns, table = frame_scninfos(pkg, p, shade, a.frames)
    vol["transfer"] = table
    t0 = time.perf_counter()
    r = pkg.Renderer(local)
    r.import_topology_host(vol["vdbinfo"], vol["pool0"], vol["pool1"])
    r.import_atlas_host(vol["atlas"])
    r.set_transfer(table)
    r.sync()
    import_s = time.perf_counter() - t0
    free0, total0 = torch.cuda.mem_get_info(dev)
    r.set_sampler(0 if a.sampler == "tex" else 1)
    bw, bh = (int(x) for x in a.block.split("x"))
    r.set_block(bw, bh)
    r.set_option(5, {"default": 0, "literal": 1, "packet": 2}[a.traversal])
    r.set_deep_shadow(dshadow)
    res = {"workload": f"{workload} {mode} {w}x{h}", "bricks": timing["bricks"], "atlas_mb": vol["atlas"].nbytes / 1e6,
           "import_s": import_s, "scene_gen_s": timing["scene_gen_s"], "topology_build_s": timing["topology_build_s"]}
    res["parity"] = parity_check(torch, r, workload, mode, w, h, scns[0], dev) if a.parity_check else {"parity_checked": False, "why": "--no-parity-check"}
    nbuf = max(1, a.lanes)
    frames_d = [torch.zeros((h, w, 4), dtype=torch.uint8, device=dev) for _ in range(nbuf)]
    units_ref = units_prod = None
    alg_ref = alg_prod = 0.0
    if full:
        units_ref, alg_ref = count_work(r, scns, shade, frames_d[0], w, h, 1)
        units_prod, alg_prod = count_work(r, scns, shade, frames_d[0], w, h, 2)
    r.set_spp(a.spp)
    r.lanes(a.lanes)
    clocks = ClockSampler(local)
    clocks.start()
    time.sleep(0.12)                    # nvidia-smi needs ~0.1 s to deliver its first sample
    t_region0 = time.time()
    ms_total, launches = time_resident(torch, r, scns, shade, frames_d, a.lanes, steps, warmup)
    clk = clocks.stop(t_region0, time.time())
    r.lanes(0)
    lat = time_latency(torch, r, scns, shade, frames_d[0])
    rays_step = a.frames * w * h * a.spp
    res.update(value=rays_step * steps / (ms_total * 1e-3) / 1e6, ms_per_frame=ms_total / steps / a.frames, latency_ms_1lane=lat["mean"],
               latency_1lane=lat, value_1lane=w * h * a.spp / (lat["mean"] * 1e-3) / 1e6)
    dev_resident_gb = (total0 - free0) / 1e9
    tex_peak = tex_peaks(r, p) if full else None
    sampler_ab = None
    if full:      # the four ways of reading a brick, fetch + filter only, at this camera's ray spacing (csrc/gvdbx_microbench.cuh)
        sampler_ab = {k: round(v, 2) for k, v in r.measure_sampler_ab(tex_peak["ray_spacing_voxels"]).items()}
        sampler_ab.update(unit="Gsamples/s", what="fetch + filter only on the imported atlas at this camera's ray spacing: texture unit on the caller's array | "
                          "brick-major copy, 8 scalar loads | x-pair layout, 4 loads of 8 bytes | brick staged into shared memory by TMA (cp.async.bulk + mbarrier), "
                          "8 shared-memory loads; the three linear variants run the software model of the unit's filter (~45 instructions per sample)")
    r.close()
    del frames_d
    torch.cuda.empty_cache()
    strict = time_e2e(torch, pkg, p, vol, local, a, shade, dshadow, a.frames, max(2, steps // 2), 0)
    res["e2e_strict"] = strict
    extra = {"ms_total": ms_total, "launches": launches, "p": p, "vol": vol, "scns": scns, "shade": shade, "dshadow": dshadow,
             "clocks": clk, "tex_peak": tex_peak, "sampler_ab": sampler_ab, "units_ref": units_ref, "units_prod": units_prod, "alg_ref": alg_ref * a.spp, "alg_prod": alg_prod * a.spp, "device_gb_after_import": dev_resident_gb}
    return res, extra


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gvdbx")
    ap.add_argument("--workload", default=HEADLINE[0])
    ap.add_argument("--mode", default="", help="voxel | trilinear | levelset | deep | deepshadow (default: the headline's deep + shadow for cfg4, else the preset's mode)")
    ap.add_argument("--frames", type=int, default=8, help="frames per step (orbit positions)")
    ap.add_argument("--size", default="")
    ap.add_argument("--sampler", default="tex", choices=["tex", "linear"])
    ap.add_argument("--tile", type=int, default=32)
    ap.add_argument("--block", default="8x8")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-table", dest="table", action="store_false", help="skip the per-config table (cfg1-cfg4) of the N = 1 run")
    ap.add_argument("--no-parity-check", dest="parity_check", action="store_false")
    ap.add_argument("--traversal", default="default", choices=["default", "literal", "packet"])
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: peer = render kernels store straight into rank 0's frame over NVLink (default); nccl = packed tiles + gather + assemble")
    ap.add_argument("--slots", type=int, default=4, help="frames in flight in the peer ring")
    ap.add_argument("--replicate", default="broadcast", choices=["broadcast", "local"],
                    help="N>1: broadcast = rank 0 builds the volume, pools and atlas travel GPU to GPU (NCCL over NVLink); "
                         "local = every rank builds and uploads its own copy")
    ap.add_argument("--lanes", type=int, default=4, help="frame lanes: consecutive frames alternate between this many internal streams (0 = one stream)")
    ap.add_argument("--spp", type=int, default=1)
    ap.add_argument("--deep-shadow", action="store_true", help="same as --mode deepshadow")
    a = ap.parse_args()
    a.size = tuple(int(x) for x in a.size.split("x")) if a.size else None
    a.warmup = max(a.warmup, 3) if a.impl != "reference" else a.warmup
    if not a.mode:
        if a.workload == HEADLINE[0] or a.deep_shadow:
            a.mode = HEADLINE[1]
        else:
            import oracle as _o
            a.mode = {0: "voxel", 4: "trilinear", 6: "levelset", 7: "deep"}[_o.preset(a.workload).shade]
    if a.impl == "reference":
        return run_reference(a)

    import numpy as np
    import torch
    import torch.distributed as dist
    import oracle
    pkg = load_pkg()
    from gvdb_voxels_b200 import multigpu as mg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    shade, dshadow = MODE[a.mode]
    hbm_peak, sm_max, _ = peaks()

    if world == 1:
        return main_single(torch, pkg, a, dev, local)

    dist.init_process_group("nccl", device_id=dev)
    # ---------------- workload (CPU: scene synthesis + reference-style topology build = reported CPU baseline #1)
    timing = {}
    bcast = a.replicate == "broadcast"
    if bcast and rank != 0:
        p, vol = oracle.preset(a.workload), None          # the volume arrives over NVLink below
        if a.size:
            p.width, p.height = a.size
    else:
        oracle.lib().ora_set_num_threads(max(1, (os.cpu_count() or 1) // (1 if bcast else world)))   # torchrun exports OMP_NUM_THREADS=1
        p, vol = build_workload(a.workload, a.size, timing)
    w, h = p.width, p.height
    scns, table = frame_scninfos(pkg, p, shade, a.frames)
    if vol is not None:
        vol["transfer"] = table
    rays_step = a.frames * w * h * a.spp

    t0 = time.perf_counter()
    r = pkg.Renderer(local)
    if bcast:
        meta = mg.replicate_volume(r, vol, rank, world, dev)
        atlas_bytes = int(np.prod(meta["atlas_shape"])) * 4
    else:
        r.import_topology_host(vol["vdbinfo"], vol["pool0"], vol["pool1"])
        r.import_atlas_host(vol["atlas"])
        atlas_bytes = vol["atlas"].nbytes
    r.set_transfer(table)
    r.sync()
    import_s = time.perf_counter() - t0
    r.set_sampler(0 if a.sampler == "tex" else 1)
    bw, bh = (int(x) for x in a.block.split("x"))
    r.set_block(bw, bh)
    r.set_option(5, {"default": 0, "literal": 1, "packet": 2}[a.traversal])
    r.set_deep_shadow(dshadow)

    parity = None
    if rank == 0:
        parity = parity_check(torch, r, a.workload, a.mode, w, h, scns[0], dev) if a.parity_check else {"parity_checked": False, "why": "--no-parity-check"}
    dist.barrier()

    nbuf = max(1, a.lanes)
    frame = torch.zeros((h, w, 4), dtype=torch.uint8, device=dev)
    units_ref = units_prod = None
    alg_ref = alg_prod = 0.0
    if rank == 0:
        units_ref, alg_ref = count_work(r, scns, shade, frame, w, h, 1)
        units_prod, alg_prod = count_work(r, scns, shade, frame, w, h, 2)
    tex_peak = tex_peaks(r, p) if rank == 0 else None
    r.set_spp(a.spp)
    r.lanes(a.lanes)
    launches = 0

    ring, consumer, consumers, released_ev = None, None, None, None
    if a.exchange == "peer":
        ring = mg.PeerFrameRing(r, w, h, a.tile, rank, world, nslots=a.slots)
        # rank 0 consumes finished frames on two streams in turn (the wait for frame q+1 overlaps the D2H copy of frame q);
        # slots are released in frame order: the release of q waits for the release of q-1 (event)
        consumers = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)] if rank == 0 else None
        released_ev = [torch.cuda.Event(), torch.cuda.Event()] if rank == 0 else None
        consumer = consumers[0] if rank == 0 else None
    tiled = mg.TiledFrame(r, w, h, a.tile, rank, world, dev) if ring is None else None
    kpf = 2 if shade == 7 else 1        # kernels per frame: deep modes build the frame's derived transfer table first

    def step_resident(on_frame=None):
        """one step = all frames of the orbit; every rank renders its tiles of every frame.
        peer exchange: kpf + 1 launches per frame and rank (render + done flag), +1 on rank 0 (release flags)."""
        nonlocal launches
        if ring is not None:
            for scn in scns:
                q = ring.submit(scn, shade)
                launches += 1 + kpf
                if rank == 0:
                    cs = consumers[q & 1]
                    ring.acquire(q, cs.cuda_stream)
                    if on_frame is not None:
                        on_frame(q, cs)
                    if q > 1:
                        cs.wait_event(released_ev[(q - 1) & 1])
                    ring.release(q, cs.cuda_stream)
                    released_ev[q & 1].record(cs)
                    launches += 1
        else:
            tiled.render_frames(scns, shade)
            launches += len(scns) * (2 if rank == 0 else 1)

    def join_consumer():
        """the measuring stream waits for the frame lanes and (rank 0) the consumer streams: stream-ordered, no host sync"""
        r.lanes_join()
        if consumer is not None:
            for cs in consumers:
                torch.cuda.current_stream().wait_stream(cs)

    def sync_all():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()              # before the warm-up: nvidia-smi needs ~0.1 s to deliver its first sample
    r.lanes_fork()
    for _ in range(a.warmup):
        step_resident()
    join_consumer()
    sync_all()
    launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t_region0 = time.time()
    e0.record()
    r.lanes_fork()
    for _ in range(a.steps):
        step_resident()
    join_consumer()
    e1.record()
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    t_region1 = time.time()
    clk = clocks.stop(t_region0, t_region1) if rank == 0 else None
    launches_timed = launches

    # per-rank render-only time of one step (no exchange): shows the load balance of the static tile partition
    sync_all()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    scratch = torch.zeros((mg.slots_per_rank(w, h, a.tile, world), a.tile, a.tile, 4), dtype=torch.uint8, device=dev)
    for scn in scns:
        r.render_tiles(scn, shade, scratch.data_ptr(), a.tile, rank, world)
    r1.record()
    torch.cuda.synchronize()
    mine = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device=dev)
    allms = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allms, mine)
    rank_render_ms = [round(float(x.item()), 3) for x in allms]

    # multi-GPU determinism check: the assembled frame equals a single-GPU render of the same camera
    frame_ok = None
    if rank == 0:
        ref = torch.zeros_like(frame)
        r.lane_select(-1)
        r.render(scns[-1], shade, ref.data_ptr())
        r.sync()
        last = ring.frame_tensor(ring.seq, torch, dev) if ring is not None else tiled.frame
        frame_ok = bool(torch.equal(ref, last))
        if not frame_ok:        # which rank's tiles differ
            bad = (ref != last).any(dim=2).cpu().numpy()
            tx = (w + a.tile - 1) // a.tile
            ys, xs = np.nonzero(bad)
            owner = ((ys // a.tile) * tx + xs // a.tile) % world
            sys.stderr.write(f"[bench] frame mismatch: {int(bad.sum())} pixels, by owning rank {np.bincount(owner, minlength=world).tolist()}\n")
    dist.barrier()      # the other ranks must not start the e2e frames (which reuse the ring slots) while rank 0 still compares

    # ---------------- e2e with host buffers: every finished frame ends row-major in page-locked host memory that rank 0 reads.
    # Host frame ring (gvdbx_hostring_*): each rank renders full-width 16-row bands and copies ITS bands over ITS OWN PCIe
    # link into a shared-memory frame — N links instead of funnelling every frame through rank 0's.
    hring = mg.HostFrameRing(r, f"/gvdbx_bench_{os.environ.get('MASTER_PORT', '0')}", w, h, rank, world, nslots=a.slots, band_rows=16)
    checksum = 0

    def step_e2e():
        nonlocal checksum
        for scn in scns:
            q = hring.seq + 1
            if rank == 0 and q > a.slots:               # the consumer lags `slots` frames behind the producers
                fr = hring.wait(q - a.slots)
                checksum += int(fr[::97, ::89, 1].sum())    # the caller touches the frame it owns now
                hring.release(q - a.slots)
            hring.submit(scn, shade)
    for _ in range(2):
        step_e2e()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_e2e()
    if rank == 0:                                       # drain: every frame of the timed steps has been handed to the caller
        for q in range(max(1, hring.seq - a.slots + 1), hring.seq + 1):
            fr = hring.wait(q)
            checksum += int(fr[::97, ::89, 1].sum())
            hring.release(q)
    sync_all()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.item())
    # the last frame delivered through the host ring equals the single-GPU render of the same camera
    host_ok = None
    if rank == 0:
        ref = torch.zeros_like(frame)
        r.lane_select(-1)
        r.render(scns[-1], shade, ref.data_ptr())
        r.sync()
        last_host = hring.wait(hring.seq)               # still in its slot (released slots are only rewritten by later frames)
        host_ok = bool(np.array_equal(ref.cpu().numpy(), last_host))
    dist.barrier()
    hring.close()
    e2e = {"value": rays_step * a.steps / dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": a.frames * 416 * world,
           "d2h_bytes_per_step": a.frames * w * h * 4, "ms_per_frame": dt / a.steps / a.frames * 1e3, "host_frame_matches_single_gpu": host_ok,
           "api": "gvdbx_hostring_submit per rank (own 16-row bands, one pitched copy per frame -> own PCIe link -> shared page-locked host frame), gvdbx_hostring_wait / _release on rank 0"}
    if ring is not None:
        ring.check()            # no stream-ordered wait ran into its timeout

    if rank != 0:
        dist.barrier()
        dist.destroy_process_group()
        return 0

    key = f"{a.workload}:{a.mode}:{a.sampler}"
    scale = lambda u: {k: v * a.steps * a.spp for k, v in u.items()} if u else None
    roofline = roofline_block(key, a.mode, f"gx_render_kernel<{a.mode},{a.sampler},tiles>", ms_total, a.frames * a.steps * a.spp, world,
                              clk["sm_mhz"] if clk else None, alg_ref * a.spp * a.steps, alg_prod * a.spp * a.steps, scale(units_ref), scale(units_prod), tex_peak)
    value = rays_step * a.steps / (ms_total * 1e-3) / 1e6
    out = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": ms_total / a.steps, "ms_per_frame": ms_total / a.steps / a.frames, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(a.workload, a.mode, w, h, a.frames, a.spp),
           "impl_config": {"bricks": timing.get("bricks"), "atlas_mb": atlas_bytes / 1e6, "replicate": a.replicate, "sampler": a.sampler, "block": a.block,
                           "traversal": a.traversal, "frame_lanes": a.lanes, "ring_slots": a.slots,
                           "parallelism": f"image tiles {a.tile}x{a.tile} round-robin over {world} GPUs, volume replicated, exchange={a.exchange}"},
           "e2e": e2e, "gpu_launches": launches_timed, "roofline": roofline, "clocks": clk, "import_s": import_s,
           "scene_gen_s": timing.get("scene_gen_s"), "multi_gpu_frame_matches_single_gpu": frame_ok, "rank_render_ms_per_step": rank_render_ms}
    out.update(parity or {})
    print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()
    return 0


def main_single(torch, pkg, a, dev, local):
    import oracle
    res, x = measure_single(torch, pkg, a, a.workload, a.mode, dev, local, a.steps, a.warmup, True)
    clk_all = x["clocks"]               # sampled around warm-up + timed steps of the resident measurement
    p, vol, scns, shade, dshadow = x["p"], x["vol"], x["scns"], x["shade"], x["dshadow"]
    w, h = p.width, p.height
    e2e = time_e2e(torch, pkg, p, vol, local, a, shade, dshadow, a.frames, a.steps, a.lanes)

    key = f"{a.workload}:{a.mode}:{a.sampler}"
    scale = lambda u: {k: v * a.steps * a.spp for k, v in u.items()} if u else None     # counted once per orbit -> the timed region
    roofline = roofline_block(key, a.mode, f"gx_render_kernel<{a.mode},{a.sampler}>", x["ms_total"], a.frames * a.steps * a.spp, 1, clk_all["sm_mhz"],
                              x["alg_ref"] * a.steps, x["alg_prod"] * a.steps, scale(x["units_ref"]), scale(x["units_prod"]), x["tex_peak"])

    # ---------------- CPU baseline: oracle port on a bounded sample (whole frames of the orbit, ~15 s of CPU work)
    cpu = None
    if not a.no_cpu_baseline:
        nthreads = oracle.lib().ora_max_threads()
        rows = (h // 2 - 64, h // 2 + 64) if h > 256 else None
        t0 = time.perf_counter()
        oracle.render(vol, scns[0], shade, deep_shadow=bool(dshadow), rows=rows)
        dt1 = time.perf_counter() - t0
        nrow = (rows[1] - rows[0]) if rows else h
        nfr = int(max(1, min(64, round(15.0 / max(dt1, 1e-3)))))
        t0 = time.perf_counter()
        for j in range(nfr):
            oracle.render(vol, scns[j % len(scns)], shade, deep_shadow=bool(dshadow), rows=rows)
        dt = time.perf_counter() - t0
        cpu = {"value": nfr * w * nrow / dt / 1e6, "unit": "Mrays/s", "cores": nthreads, "kind": "port",
               "sample": f"{nrow} centre rows of {nfr} frames of the orbit ({nfr * w * nrow} primary rays, 1 ray per pixel), CPU restatement "
                         f"oracle/gvdb_oracle.c, OpenMP {nthreads} threads, {dt:.1f} s",
               "topology_build": {"seconds": res["topology_build_s"], "bricks": res["bricks"], "threads": 1, "kind": "port",
                                  "what": "Configure + ActivateSpace per brick + FinishTopology + UpdateAtlas (CPU restatement, byte-identical pools)"},
               "host_cores": os.cpu_count()}

    configs = None
    if a.table:
        configs = []
        for wl, mode in TABLE:
            r1, _ = measure_single(torch, pkg, a, wl, mode, dev, local, max(2, a.steps // 4), 3, False)
            configs.append({k: r1[k] for k in ("workload", "value", "ms_per_frame", "latency_ms_1lane", "value_1lane", "bricks", "atlas_mb")} |
                           {"e2e_strict": r1["e2e_strict"]["value"], "parity_checked": r1["parity"].get("parity_checked"),
                            "pixels_differing": r1["parity"].get("pixels_differing")})
            del _

    out = {"metric": "Mrays/s", "value": res["value"], "unit": "Mrays/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": x["ms_total"] / a.steps, "ms_per_frame": res["ms_per_frame"], "latency_ms_1lane": res["latency_ms_1lane"],
           "latency_1lane": res["latency_1lane"], "value_1lane": res["value_1lane"], "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(a.workload, a.mode, w, h, a.frames, a.spp),
           "impl_config": {"bricks": res["bricks"], "atlas_mb": res["atlas_mb"], "sampler": a.sampler, "block": a.block, "traversal": a.traversal,
                           "frame_lanes": a.lanes, "parallelism": "single GPU", "device_gb_after_import": x["device_gb_after_import"],
                           "ms_per_frame_is": "inverse throughput with frame_lanes frames in flight; latency_ms_1lane = one frame on one stream"},
           "e2e": e2e, "e2e_strict": res["e2e_strict"], "gpu_launches": x["launches"], "roofline": roofline, "clocks": clk_all,
           "import_s": res["import_s"], "scene_gen_s": res["scene_gen_s"]}
    out.update(res["parity"])
    if x.get("sampler_ab"):
        out["sampler_ab"] = x["sampler_ab"]
    if cpu:
        out["cpu_baseline"] = cpu
    if configs is not None:
        out["configs"] = configs
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
