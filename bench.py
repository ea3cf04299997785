#!/usr/bin/env python
"""bench.py — Mrays/s and ms/frame of the GVDB ray-cast render path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--frames F] [--sampler tex|linear]
  python bench.py --impl reference ...       the reference's own render of the same workload (oracle/_ref)
  torchrun --nproc-per-node N bench.py --gpus N ...   one rank per GPU, image-space tiles, NCCL gather to rank 0

A step = one pass of the hot path over one batch of synthetic input = F frames (camera yaw + 360*j/F around the
preset's orbit) of the workload.  `value` = primary rays / time of the K timed steps with everything resident in HBM
(max over ranks, CUDA events, barrier + synchronize on both sides).  `e2e` = the same frames through the
reference-facing API (VolumeGVDB mirror: Render + ReadRenderBuf) with HOST buffers: per frame the 416-byte ScnInfo goes
host->device (kernel parameter block) and the RGBA8 frame comes back into pinned host memory.

oracle/ is used here only (a) to synthesise the input volume (the CPU topology build is one of the two reported
baselines), (b) as the `cpu_baseline` leg and (c) by `--impl reference`; never inside the timed GPU region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# render and consumer streams should not share a hardware queue (the default of 8 connections is spread over torch's
# pool of 32 streams): must be set before the CUDA context exists
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MODES = {"voxel": 0, "trilinear": 4, "levelset": 6, "deep": 7}
SHADE_NAME = {v: k for k, v in MODES.items()}
# algorithmic bytes per unit of work
B_TRI, B_PT, B_DDA, B_DESC, B_PIX = 32, 4, 8, 64, 4


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


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


def run_reference(a):
    """--impl reference: the reference's own CUDA render of the same workload through its public API
    (oracle/_ref/libgvdb.so, unmodified), or — if oracle/_ref was not built — the CPU port of the oracle."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    p = oracle.preset(a.workload)
    w, h = a.size if a.size else (p.width, p.height)
    shade = MODES[a.mode] if a.mode else p.shade
    rays = a.frames * w * h * a.spp
    mode_name = SHADE_NAME[shade]
    if shade == 7 and a.deep_shadow:
        mode_name = "deepshadow"        # composed from the reference's own device functions through RenderKernel
    elif shade == 7 and a.spp > 1:
        mode_name = "deepspp"
    base = {"metric": "Mrays/s", "unit": "Mrays/s", "impl": "reference", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{a.workload} {mode_name} {w}x{h}", "frames_per_step": a.frames, "spp": a.spp}}
    refdir = os.path.join(ROOT, "oracle", "_ref")
    if os.path.exists(os.path.join(refdir, "ref_harness")):
        cmd = ["./ref_harness", a.workload, "/tmp/ref_bench", "--bench", "--orbit", str(a.frames), "--steps", str(a.steps),
               "--warmup", str(a.warmup), "--mode", mode_name, "--size", f"{w}x{h}", "--spp", str(a.spp)]
        r = subprocess.run(cmd, cwd=refdir, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        if r.returncode == 0:
            j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
            v = rays * a.steps / j["render_s"] / 1e6
            e = rays * a.steps / j["e2e_s"] / 1e6
            base.update(value=v, ms_per_step=j["render_s"] / a.steps * 1e3, ms_per_frame=j["render_s"] / a.steps / a.frames * 1e3,
                        e2e={"value": e, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                        cpu_baseline={"value": v, "unit": "Mrays/s", "cores": 1, "kind": "reference",
                                      "sample": "unmodified reference CUDA kernels (GVDB ships no CPU ray marcher) driven by 1 host thread "
                                                "through VolumeGVDB::Render on 1 GPU; e2e adds ReadRenderBuf per frame",
                                      "topology_build_s": j["topology_build_s"], "bricks": j["bricks"]},
                        gpu_launches=0)
            print(json.dumps(base))
            return 0
        sys.stderr.write(r.stderr[-2000:])
    # fallback: CPU port on a bounded sample
    pkg = load_pkg()
    p, vol = build_workload(a.workload, a.size)
    scns, table = frame_scninfos(pkg, p, shade, a.frames)
    vol["transfer"] = table
    rows = max(8, min(h, 64))
    y0 = h // 2 - rows // 2
    t0 = time.perf_counter()
    for s in range(a.steps):
        oracle.render(vol, scns[s % len(scns)], shade, rows=(y0, y0 + rows))
    dt = time.perf_counter() - t0
    v = rows * w * a.steps / dt / 1e6
    base.update(value=v, ms_per_step=dt / a.steps * 1e3,
                e2e={"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                cpu_baseline={"value": v, "unit": "Mrays/s", "cores": oracle.lib().ora_max_threads(), "kind": "port",
                              "sample": f"{rows} centre rows of one frame per step (CPU port, OpenMP)"}, gpu_launches=0)
    print(json.dumps(base))
    return 0


def load_pkg():
    from __graft_entry__ import load_package
    return load_package()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gvdbx")
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--mode", default="")
    ap.add_argument("--frames", type=int, default=8, help="frames per step (orbit positions)")
    ap.add_argument("--size", default="")
    ap.add_argument("--sampler", default="tex", choices=["tex", "linear"])
    ap.add_argument("--tile", type=int, default=32)
    ap.add_argument("--block", default="8x8")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--traversal", default="default", choices=["default", "literal", "packet"])
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: peer = render kernels store straight into rank 0's frame over NVLink (default); nccl = packed tiles + gather + assemble")
    ap.add_argument("--slots", type=int, default=4, help="frames in flight in the peer ring")
    ap.add_argument("--replicate", default="broadcast", choices=["broadcast", "local"],
                    help="N>1: broadcast = rank 0 builds the volume, pools and atlas travel GPU to GPU (NCCL over NVLink); "
                         "local = every rank builds and uploads its own copy")
    ap.add_argument("--lanes", type=int, default=4, help="frame lanes: consecutive frames alternate between this many internal streams (0 = one stream)")
    ap.add_argument("--spp", type=int, default=1)
    ap.add_argument("--deep-shadow", action="store_true")
    a = ap.parse_args()
    a.size = tuple(int(x) for x in a.size.split("x")) if a.size else None
    a.warmup = max(a.warmup, 3) if a.impl != "reference" else a.warmup
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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---------------- workload (CPU: scene synthesis + reference-style topology build = reported CPU baseline #1)
    timing = {}
    bcast = world > 1 and a.replicate == "broadcast"
    if bcast and rank != 0:
        p, vol = oracle.preset(a.workload), None          # the volume arrives over NVLink below
        if a.size:
            p.width, p.height = a.size
    else:
        if world > 1:
            oracle.lib().ora_set_num_threads(max(1, (os.cpu_count() or 1) // (1 if bcast else world)))   # torchrun exports OMP_NUM_THREADS=1
        p, vol = build_workload(a.workload, a.size, timing)
    w, h = p.width, p.height
    shade = MODES[a.mode] if a.mode else p.shade
    scns, table = frame_scninfos(pkg, p, shade, a.frames)
    if vol is not None:
        vol["transfer"] = table
    rays_step = a.frames * w * h

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
    r.set_deep_shadow(a.deep_shadow)

    # ---------------- algorithmic bytes per frame (counted render, outside the timed region)
    nbuf = max(1, a.lanes)
    frames_d = [torch.zeros((h, w, 4), dtype=torch.uint8, device=dev) for _ in range(nbuf)]     # one output frame per lane
    frame = frames_d[0]
    bytes_alg = []
    counters = []
    if rank == 0:          # counted with the reference's own semantics (no brick culling): units of the ALGORITHM
        r.set_counters(True)
        for scn in scns:
            r.render(scn, shade, frame.data_ptr())
            c = r.counters()
            counters.append(c)
            bytes_alg.append(B_TRI * c["s_tri"] + B_PT * c["s_pt"] + B_DDA * c["n_dda"] + B_DESC * c["n_desc"] + B_PIX * w * h)
        r.set_counters(False)
    r.set_spp(a.spp)
    r.lanes(a.lanes)
    rays_step *= a.spp
    bytes_alg = [b * a.spp for b in bytes_alg]
    launches = 0

    ring, consumer = None, None
    if world > 1 and a.exchange == "peer":
        ring = mg.PeerFrameRing(r, w, h, a.tile, rank, world, nslots=a.slots)
        # rank 0 consumes finished frames on two streams in turn (the wait for frame q+1 overlaps the D2H copy of frame q);
        # slots are released in frame order: the release of q waits for the release of q-1 (event)
        consumers = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)] if rank == 0 else None
        released_ev = [torch.cuda.Event(), torch.cuda.Event()] if rank == 0 else None
        consumer = consumers[0] if rank == 0 else None
    tiled = mg.TiledFrame(r, w, h, a.tile, rank, world, dev) if (world > 1 and ring is None) else None

    kpf = 2 if shade == 7 else 1        # kernels per frame: deep modes build the frame's derived transfer table first

    def step_resident(on_frame=None):
        """one step = all frames of the orbit; N>1: every rank renders its tiles of every frame.
        peer exchange: 2 launches per frame and rank (render + done flag), +1 on rank 0 (release flags)."""
        nonlocal launches
        if world == 1:
            for j, scn in enumerate(scns):
                r.lane_select(j % nbuf if a.lanes else -1)
                r.render(scn, shade, frames_d[j % nbuf].data_ptr())
                launches += kpf
        elif ring is not None:
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
        """the measuring stream waits for the frame lanes and (rank 0) the consumer stream: stream-ordered, no host sync"""
        r.lanes_join()
        if consumer is not None:
            for cs in consumers:
                torch.cuda.current_stream().wait_stream(cs)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
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
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    t_region1 = time.time()
    clk = clocks.stop(t_region0, t_region1) if rank == 0 else None
    launches_timed = launches

    # per-rank render-only time of one step (no gather): shows the load balance of the static tile partition
    rank_render_ms = None
    if world > 1:
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

    # multi-GPU determinism check: the gathered frame equals a single-GPU render of the same camera
    frame_ok = None
    if world > 1 and rank == 0:
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
            sys.stderr.write(f"[bench] frame mismatch: {int(bad.sum())} pixels, by owning rank {np.bincount(owner, minlength=world).tolist()}, "
                             f"nonzero in ring frame {int((last != 0).any(dim=2).sum())} of {w * h}\n")
            for k in range(len(scns)):
                r.render(scns[k], shade, ref.data_ptr())
                r.sync()
                sys.stderr.write(f"[bench]   vs camera {k}: {int((ref != last).any(dim=2).sum())} pixels differ\n")

    if world > 1:
        dist.barrier()      # the other ranks must not start the e2e frames (which reuse the ring slots) while rank 0 still compares

    # ---------------- e2e through the reference-facing API with host buffers (rank-local at N=1; tiled at N>1)
    e2e = None
    host = torch.empty((h, w, 4), dtype=torch.uint8).pin_memory()
    host_np = host.numpy()
    if world == 1:
        v = pkg.Volume(local)
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
        hosts = [host] + [torch.empty((h, w, 4), dtype=torch.uint8).pin_memory() for _ in range(nbuf - 1)]
        hosts_np = [t.numpy() for t in hosts]
        for k in range(nbuf):
            v.AddRenderBuf(k, w, h, 4)
        v.SetRenderLanes(a.lanes)
        v.set_option(1, 0 if a.sampler == "tex" else 1)
        v.set_option(2, bw); v.set_option(3, bh)
        v.set_option(5, {"default": 0, "literal": 1, "packet": 2}[a.traversal])
        v.set_option(7, a.spp)
        v.set_option(8, 1 if a.deep_shadow else 0)

        def step_e2e():
            # render buffer k = frame lane k: frame j renders into buffer j % nbuf while earlier frames are still being
            # rendered / copied; a host frame is handed to the caller (SyncRenderBuf) before its buffer is reused
            for j in range(a.frames):
                k = j % nbuf
                if j >= nbuf:
                    v.SyncRenderBuf(k)
                v.SetCamera(p.fov, (p.cam_angs[0] + 360.0 * j / a.frames, p.cam_angs[1], p.cam_angs[2]), list(p.cam_target), p.cam_dist)
                v.SetRes(w, h)
                v.Render(shade, 0, k)                       # PrepareRender: 416-byte ScnInfo host -> device with the launch
                v.ReadRenderBufAsync(k, hosts_np[k])        # device -> pinned host behind the kernel, on the buffer's lane
            for k in range(min(nbuf, a.frames)):
                v.SyncRenderBuf(k)
        for _ in range(2):
            step_e2e()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            step_e2e()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e = {"value": rays_step * a.steps / dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": a.frames * 416,
               "d2h_bytes_per_step": a.frames * w * h * 4, "ms_per_frame": dt / a.steps / a.frames * 1e3,
               "api": f"VolumeGVDB mirror: SetCamera + Render(rbuf = frame % {nbuf}) + ReadRenderBufAsync into pinned host memory + SyncRenderBuf"}
        v.close()
    else:
        if ring is not None:
            host_ring = [host] + [torch.empty((h, w, 4), dtype=torch.uint8).pin_memory() for _ in range(a.slots - 1)] if rank == 0 else []

            def to_host(q, cs):
                with torch.cuda.stream(cs):                     # D2H of the finished frame behind the acquire, on its consumer stream
                    host_ring[(q - 1) % a.slots].copy_(ring.frame_tensor(q, torch, dev), non_blocking=True)

            def step_e2e():
                step_resident(on_frame=to_host if rank == 0 else None)
                if rank == 0:
                    for cs in consumers:
                        cs.synchronize()                        # the caller owns the host frames of this step now
                r.lane_select(-1)
        else:
            def to_host(j, fr):
                host.copy_(fr, non_blocking=True)                   # D2H of the assembled frame, stream-ordered
                torch.cuda.current_stream().synchronize()           # the caller owns the host frame now

            def step_e2e():
                tiled.render_frames(scns, shade, on_frame=to_host)
        for _ in range(2):
            step_e2e()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            step_e2e()
        sync_all()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt.item())
        e2e = {"value": rays_step * a.steps / dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": a.frames * 416 * world,
               "d2h_bytes_per_step": a.frames * w * h * 4, "ms_per_frame": dt / a.steps / a.frames * 1e3,
               "api": ("gvdbx_render_tiles_direct per rank into rank 0's frame ring over NVLink + D2H to pinned host on rank 0" if ring is not None
                       else "gvdbx_render_tiles per rank + NCCL gather + gvdbx_assemble_tiles + D2H to pinned host on rank 0")}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---------------- roofline of the dominant kernel (render kernel; N=1: the timed region is only that kernel)
    peak, peak_src = peaks()
    alg_step = float(sum(bytes_alg))
    achieved = alg_step * a.steps / (ms_total * 1e-3) / 1e9          # GB/s, whole job (all ranks together)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"{a.workload}:{SHADE_NAME[shade]}:{a.sampler}")
        except Exception:
            traffic = None
    tot = {k: sum(c[k] for c in counters) for k in counters[0]}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "kernel": f"gx_render_kernel<{SHADE_NAME[shade]},{a.sampler}>",
                "algorithmic_bytes_per_frame": alg_step / a.frames,
                "dram_achieved": (traffic * a.frames * a.steps * a.spp / (ms_total * 1e-3) / 1e9) if traffic else None,
                "note": "achieved = SURVEY 8(d) algorithmic bytes (32 B per trilinear sample, 8 B per DDA step, 64 B per node record, 4 B per pixel) / time: "
                        "an EFFECTIVE bandwidth - neighbouring rays share bricks, 99 % of the sectors hit in L1 and ~78 % of the rest in L2, so measured "
                        "DRAM traffic (`traffic`, bytes per launch; `dram_achieved`, GB/s) is ~2.5 % of it and frac exceeds 1; the kernel is bound by "
                        "instruction issue (IPC 2.8-3.1 of 4 at 14-16 of 32 lanes, profiles/r01_ncu_summary.md), not by HBM",
                "units_per_step": tot, "bytes_per_unit": {"s_tri": B_TRI, "s_pt": B_PT, "n_dda": B_DDA, "n_desc": B_DESC, "pixel": B_PIX}}

    # ---------------- CPU baselines (rank 0, N=1 only): oracle port on a bounded sample
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        nthreads = oracle.lib().ora_max_threads()
        # bounded sample of the same workload: whole frames of the orbit (1 spp), sized from one timed frame to ~15 s of CPU work
        t0 = time.perf_counter()
        oracle.render(vol, scns[0], shade, deep_shadow=a.deep_shadow)
        dt1 = time.perf_counter() - t0
        nfr = int(max(1, min(64, round(15.0 / max(dt1, 1e-3)))))
        t0 = time.perf_counter()
        for j in range(nfr):
            oracle.render(vol, scns[j % len(scns)], shade, deep_shadow=a.deep_shadow)
        dt = time.perf_counter() - t0
        cpu = {"value": nfr * w * h / dt / 1e6, "unit": "Mrays/s", "cores": nthreads, "kind": "port",
               "sample": f"{nfr} full frames of the orbit ({nfr * w * h} primary rays, 1 ray per pixel), CPU restatement oracle/gvdb_oracle.c, "
                         f"OpenMP {nthreads} threads, {dt:.1f} s",
               "topology_build": {"seconds": timing["topology_build_s"], "bricks": timing["bricks"], "threads": 1, "kind": "port",
                                  "what": "Configure + ActivateSpace per brick + FinishTopology + UpdateAtlas (CPU restatement, byte-identical pools)"},
               "host_cores": os.cpu_count()}

    value = rays_step * a.steps / (ms_total * 1e-3) / 1e6
    out = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": ms_total / a.steps, "ms_per_frame": ms_total / a.steps / a.frames, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{a.workload} {SHADE_NAME[shade]}{'+shadow' if a.deep_shadow else ''} {w}x{h}", "frames_per_step": a.frames, "bricks": timing["bricks"],
                      "atlas_mb": atlas_bytes / 1e6, "replicate": a.replicate if world > 1 else None, "sampler": a.sampler, "block": a.block, "traversal": a.traversal, "frame_lanes": a.lanes,
                      "parallelism": (f"image tiles {a.tile}x{a.tile} round-robin over {world} GPU(s), volume replicated, exchange={a.exchange}"
                                      if world > 1 else "single GPU"), "spp": a.spp, "deep_shadow": bool(a.deep_shadow),
                      "l2_policy": "inputs larger than L2 (atlas %.0f MB vs 126 MB L2); camera changes every frame" % (atlas_bytes / 1e6)},
           "e2e": e2e, "gpu_launches": launches_timed, "roofline": roofline, "clocks": clk, "import_s": import_s,
           "scene_gen_s": timing["scene_gen_s"]}
    if cpu:
        out["cpu_baseline"] = cpu
    if frame_ok is not None:
        out["multi_gpu_frame_matches_single_gpu"] = frame_ok
        out["rank_render_ms_per_step"] = rank_render_ms
    print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
