#!/usr/bin/env python
"""bench.py — Mrays/s (primary + shadow [+ bounce]) and ms/frame of the hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--packets off|union|frustum|auto]
                    [--workload soup1m|soup1m_far|niels360|niels1080|heightfield10m|heightfield10m_b4|niels8k16|soup8k16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one frame: init -> primary rays -> nearest hit -> shadow rays -> occlusion -> lighting + composite -> rgba8
(RTB_PASS_FRAME), on BASELINE.json configs[2]: the 1M-triangle random soup at 3840x2160, 1 spp + 1 shadow ray.
With N > 1 the frame's 32x32-pixel blocks are dealt round-robin to the ranks (scene and BVH replicated), each rank
renders its blocks, NCCL gathers the rgba8 tiles on rank 0 and rtb_untile lays the frame out (strong scaling: the
frame is fixed); the gather and lay-out of frame k run on a second stream under the rendering of frame k + 1.
Prints ONE JSON line on rank 0.

The oracle (oracle/) is executed only for the `cpu_baseline` leg and for `--impl reference`; those legs import nothing
from the product package (the synthetic scenes come from the oracle's own generators).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mrays/s (primary+shadow)"
UNIT = "Mrays/s"

WORKLOADS = {
    # BASELINE.json configs[2]; camera placed so that every pixel's ray enters the soup volume (eye 3.9 in front of
    # the soup's z = 10 face with the reference's default fov 70): the incoherent-BVH stress the config names.
    "soup1m": dict(kind="soup", triangles=1_000_000, width=3840, height=2160, eye=(0.0, 0.0, 13.9), samples=1,
                   desc="configs[2]: synthetic 1M-triangle random soup, 3840x2160, 1 spp primary + 1 shadow ray, camera eye (0,0,13.9) fov 70 (all rays enter the soup)"),
    # SURVEY.md §8(d) camera for the same scene: the soup covers ~3.7 % of the frame
    "soup1m_far": dict(kind="soup", triangles=1_000_000, width=3840, height=2160, eye=(0.0, 0.0, 30.0), samples=1,
                       desc="configs[2] geometry with the survey camera eye (0,0,30): soup covers ~4 % of the frame"),
    "niels360": dict(kind="niels", width=640, height=360, eye=(4.0, 2.0, -2.0), samples=1,
                     desc="configs[0]: NielsScene (13 primitives) at 640x360, 1 spp primary + 1 shadow ray (the reference's own CPU-runnable case; the CPU baseline runs the whole frame)"),
    "niels1080": dict(kind="niels", width=1920, height=1080, eye=(4.0, 2.0, -2.0), samples=1,
                      desc="configs[1]: NielsScene (13 primitives) at 1920x1080, 1 spp primary + 1 shadow ray"),
    "heightfield10m": dict(kind="heightfield", grid=2236, width=1920, height=1080, eye=(0.0, 6.0, 13.0), pitch=0.45, samples=1,
                           desc="configs[3] geometry: 10M-triangle displaced height field at 1080p, first hit + 1 shadow ray (no bounces: the reference has none)"),
    # BASELINE.json configs[3] in full: 4 diffuse bounces (SURVEY.md §8d config 4 semantics; the reference has no bounce rays, so
    # this is rtb_path_frame, not RTB_PASS_FRAME), camera above the field looking down so that the mesh fills the frame
    "heightfield10m_b4": dict(kind="heightfield", grid=2236, width=1920, height=1080, eye=(0.0, 7.5, 0.0), pitch=1.5707964, fov=120.0, samples=1, bounces=4,
                              desc="configs[3]: 10M-triangle displaced height field at 1080p, 1 spp, 4 diffuse bounces + 1 shadow ray per vertex (wavefront path tracing, rtb_path_frame), camera at (0,7.5,0) looking straight down, fov parameter 120 (the reference's convention moves the plane: +-46 x +-30 degrees): the mesh fills the frame"),
    # BASELINE.json configs[4]: 8K, 16 spp progressive accumulation (the reference's export path: the command list replayed
    # per sample with USE_SUPERSAMPLING, raytracing_interface.cpp:196-242), screen tiles across the ranks, one gather per frame
    "niels8k16": dict(kind="niels", width=7680, height=4320, eye=(6.0, 5.0, 12.0), samples=1, spp=16,
                      desc="configs[4]: NielsScene at 7680x4320, 16 spp progressive accumulation (16 replays of init..composite per frame), 1 shadow ray per sample"),
    "soup8k16": dict(kind="soup", triangles=1_000_000, width=7680, height=4320, eye=(0.0, 0.0, 13.9), samples=1, spp=16,
                     desc="configs[4] on the configs[2] scene: 1M-triangle soup at 7680x4320, 16 spp progressive accumulation, 1 shadow ray per sample"),
}


def build_scene(rtb, wl):
    """product side: the scene as raw buffers from librtb200's host packing and generators"""
    sun = rtb.niels_scene()["lights"][:32]
    if wl["kind"] == "niels":
        return rtb.niels_scene(0.0), dict()
    if wl["kind"] == "soup":
        n = wl["triangles"]
        tris = rtb.gen_soup(n, 0xB200)
    else:
        n = 2 * wl["grid"] * wl["grid"]
        tris = rtb.gen_heightfield(wl["grid"], 0xB200)
    mat = rtb.pack_material((0.8, 0.8, 0.8), (0.05, 0.05, 0.05), (0, 0, 0), 0.0, 1.0, 1.0)
    scene = dict(triangles=tris, lights=sun, materials=mat, material_indices=np.zeros(n, np.uint32),
                 info=np.array([1, 1, n, 0, 0, 0, 1, 0, 0], np.uint32))
    return scene, dict(max_triangles=n)


def oracle_scene(orc, wl):
    """CPU legs: the same scene from the ORACLE's own packing and generators (byte-equal to build_scene's: tests/test_cpu_host.py);
    nothing of the product is loaded"""
    from oracle.oracle import Scene
    if wl["kind"] == "niels":
        return orc.niels_scene(0.0, None)
    if wl["kind"] == "soup":
        n = wl["triangles"]
        tris = orc.gen_soup(n, 0xB200)
    else:
        n = 2 * wl["grid"] * wl["grid"]
        tris = orc.gen_heightfield(wl["grid"], 0xB200)
    sun = orc.niels_scene(0.0, None).lights[:32]
    mat = orc.material((0.8, 0.8, 0.8), (0.05, 0.05, 0.05), (0, 0, 0), 0.0, 1.0, 1.0)
    return Scene(tris, None, None, None, sun, mat, np.zeros(n, np.uint32), np.array([1, 1, n, 0, 0, 0, 1, 0, 0], np.uint32), None)


def camera_kwargs(wl):
    return dict(eye=wl["eye"], pitch=wl.get("pitch", 0.0), yaw=wl.get("yaw", 0.0), left_fov=wl.get("fov", 70.0), right_fov=wl.get("fov", 70.0),
                flags=2 if wl.get("spp", 1) > 1 else 0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.monotonic()] + [x.strip() for x in line.split(",")])

    def mark(self):
        """Start of the timed region: nvidia-smi was started before the warm-up so that its start-up (and the driver queries it
        makes) does not fall into a timed region that may be only a few milliseconds long."""
        self.t_begin = time.monotonic()

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r[1:] for r in self.rows if len(r) >= 8]
        inside = [r[1:] for r in self.rows if len(r) >= 8 and r[0] >= getattr(self, "t_begin", 0.0)]
        window = "timed region"
        if len(inside) < 2:   # a timed region shorter than the 200 ms sampling period: the warm-up ran the same load just before
            inside, window = rows, "warm-up + timed region"
        sm = [float(r[0]) for r in inside if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in inside if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in inside for i in range(4) if r[3 + i].lower().startswith("active")})
        pw = [float(r[2]) for r in inside if r[2].replace(".", "").isdigit()]
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    power_w_max=max(pw) if pw else None, samples=len(sm), window=window)


def cpu_sample_pixels(w, h, n, seed=1234):
    rng = np.random.default_rng(seed)
    idx = rng.choice(w * h, size=n, replace=False)
    return np.stack([idx % w, idx // w], axis=1).astype(np.uint32)


def host_cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_time_sample(wl, pixels_n, repeats=1):
    """The CPU restatement of the reference shaders (the oracle compiled -O3 -march=native ON THIS MACHINE, all host threads,
    brute force over every primitive as the shaders do) on a pixel sample of the same workload.  Stated baseline, not the
    target.  Returns (Mrays/s, rays, seconds, threads)."""
    from oracle.oracle import Oracle
    orc = Oracle(native=True)
    w, h = wl["width"], wl["height"]
    cam = orc.camera(w, h, **camera_kwargs(wl))
    seed = orc.init_pass(orc.seed((0.0, 0.0)))
    osc = oracle_scene(orc, wl)
    xy = cpu_sample_pixels(w, h, pixels_n)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        rays, _, _, _ = orc.frame_pixels(osc, cam, seed, wl["samples"], xy)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return rays / best / 1e6, rays, best, orc.threads()


def run_reference(args, wl, rank, world):
    """--impl reference: the reference's algorithm on the host cores.  The reference itself is a Windows/OpenGL program; what runs
    here is oracle/ — its shaders restated for the CPU, brute-force loops included, held bit-equal to the reference's own GLSL
    compiled for the host (oracle/_ref, tests/test_oracle_vs_ref.py) — built -O3 -march=native on this machine, all threads.
    Nothing of the product package is imported on this path."""
    if rank != 0:
        return
    from oracle.oracle import Oracle
    orc = Oracle(native=True)
    osc = oracle_scene(orc, wl)
    n_tri = int(osc.info[2])
    w, h = wl["width"], wl["height"]
    per_step = max(64, min(1024, int(1.0e9 / n_tri))) if n_tri > 100_000 else min(w * h, 262_144)   # ~1 s of host work per step
    cam = orc.camera(w, h, **camera_kwargs(wl))
    seed = orc.init_pass(orc.seed((0.0, 0.0)))
    total_rays, total_s = 0, 0.0
    for i in range(args.warmup + args.steps):
        xy = cpu_sample_pixels(w, h, per_step, seed=100 + i)
        t0 = time.perf_counter()
        rays, _, _, _ = orc.frame_pixels(osc, cam, seed, wl["samples"], xy)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            total_rays += rays
            total_s += dt
    value = total_rays / total_s / 1e6
    sample = (f"{per_step} random pixels of the {w}x{h} frame per step, {total_rays} rays in the timed steps (primary + shadow, brute force over "
              f"{n_tri} triangles); first hit + shadow only for workloads with bounces; {host_cpu_model()}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_s / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": wl["desc"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": orc.threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


class CudaArray:
    """Zero-copy torch view of a device buffer owned by librtb200 (__cuda_array_interface__)."""

    def __init__(self, ptr, nwords):
        self.__cuda_array_interface__ = {"shape": (nwords,), "typestr": "<i4", "data": (ptr, False), "version": 2, "strides": None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="soup1m", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--packets", default="auto", choices=["off", "union", "frustum", "auto"],
                    help="camera rays as 8x4-pixel packets (RTB_OPT_PRIMARY_PACKETS); auto = the library's patch-size rule")
    ap.add_argument("--shadow-order", default="queue", choices=["slots", "queue", "sorted", "beams"],
                    help="RTB_OPT_SHADOW_ORDER: occlusion rays in wavefront-slot order, as a queue of live rays, or that queue sorted in light space")
    ap.add_argument("--builder", default="host", choices=["host", "device", "device3"], help="RTB_OPT_ACCEL_BUILDER: who builds the 8-wide tree")
    ap.add_argument("--no-graph", action="store_true", help="RTB_OPT_FRAME_GRAPH = 0: launch every frame directly instead of replaying its CUDA graphs")
    ap.add_argument("--no-overlap", action="store_true", help="RTB_OPT_FRAME_OVERLAP = 0: one recorded frame after the other instead of frame k+1's camera rays under frame k's shadow + shade launches")
    ap.add_argument("--lanes", type=int, default=1, choices=[1, 2], help="RTB_OPT_FRAME_LANES: a frame as one lane or as two half-frame lanes on two streams")
    ap.add_argument("--median-frames", type=int, default=100, help="frames timed one by one for the median (capped to ~10 s)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    from igx_raytracing_b200 import build as rtb_build
    if rank == 0:
        rtb_build.build_library()   # no-op when the in-tree .so is current
    from igx_raytracing_b200 import rtb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the communicator is created; stdout carries the one JSON line, so the
        # C-level stdout is pointed at stderr until the first collective has run
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)

    w, h, samples = wl["width"], wl["height"], wl["samples"]
    spp, bounces = wl.get("spp", 1), wl.get("bounces", 0)
    scene, limits = build_scene(rtb, wl)
    ctx = rtb.Context(device=local_rank, **limits)
    # a real (non-default) torch stream, made current: the library launches on it, so torch.cuda.Event timing and the
    # NCCL gather see the kernels (the legacy default stream has handle 0, which rtb_set_stream reads as "own stream")
    stream = torch.cuda.Stream(device=local_rank)
    side = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_option(rtb.OPT_PRIMARY_PACKETS, {"off": 0, "union": 1, "auto": 2, "frustum": 3}[args.packets])
    ctx.set_option(rtb.OPT_SHADOW_ORDER, {"slots": 0, "queue": 1, "sorted": 2, "beams": 3}[args.shadow_order])
    ctx.set_option(rtb.OPT_ACCEL_BUILDER, {"host": 0, "device": 1, "device3": 2}[args.builder])
    ctx.set_option(rtb.OPT_FRAME_LANES, args.lanes)
    ctx.set_option(rtb.OPT_FRAME_GRAPH, 0 if args.no_graph else 1)
    ctx.set_option(rtb.OPT_FRAME_OVERLAP, 0 if args.no_overlap else 1)
    ctx.resize(w, h, samples)
    ctx.upload_scene(scene, None)
    ctx.build_accel(rtb.ACCEL_BVH)
    cam = rtb.pack_camera(w, h, **camera_kwargs(wl))
    seed0 = rtb.make_seed((0.0, 0.0))
    cam_pin = torch.from_numpy(cam.copy()).pin_memory()
    seed_pin = torch.from_numpy(seed0.copy()).pin_memory()

    def render():
        """one frame's worth of device work on `stream`"""
        if bounces:
            ctx.path_frame(bounces)
        else:
            for _ in range(spp):   # the recorded command list replayed once per sample
                ctx.dispatch(rtb.PASS_FRAME)

    def checksum_frame():
        """the frame every checksum is taken of: seed0, one frame"""
        ctx.upload(rtb.BUF_CAMERA, cam)
        ctx.upload(rtb.BUF_SEED, seed0)
        render()

    # ---- the single-GPU frame, for the N > 1 correctness check (rank 0 alone, before the screen is partitioned) -----------
    crc_single = None
    if rank == 0:
        checksum_frame()
        crc_single = zlib.crc32(ctx.readback(rtb.TGT_RGBA8).tobytes())
    if world > 1:
        ctx.set_option(rtb.OPT_TILE_COUNT, world)
        ctx.set_option(rtb.OPT_TILE_RANK, rank)

    # ---- presentation plumbing ---------------------------------------------------------------------------------------
    frame_pins = [torch.empty(w * h, dtype=torch.int32).pin_memory() for _ in range(2)]   # e2e: the frame being copied out / the one the host may read
    e2e_count = [0]
    tiled = slots = None
    staging = gathered = untiled = ready = done = None
    shm = shared_ptr = shared_np = None
    if world > 1:
        ptr, nbytes = ctx.device_ptr(rtb.TGT_RGBA8_TILED)
        slots = nbytes // 4
        tiled = torch.as_tensor(CudaArray(ptr, slots), device=f"cuda:{local_rank}")
        dev = f"cuda:{local_rank}"
        staging = [torch.empty(slots, dtype=torch.int32, device=dev) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        for e in done:
            e.record(stream)
        if rank == 0:
            gathered = [torch.empty(world * slots, dtype=torch.int32, device=dev) for _ in range(2)]
            untiled = [torch.empty(w * h, dtype=torch.int32, device=dev) for _ in range(2)]
        # one page-locked frame in shared memory, mapped by every rank: each GPU writes its own tiles over its own PCIe link
        from multiprocessing import shared_memory
        names = [None]
        if rank == 0:
            shm = shared_memory.SharedMemory(create=True, size=w * h * 4)
            names[0] = shm.name
        dist.broadcast_object_list(names, src=0)
        if rank != 0:
            shm = shared_memory.SharedMemory(name=names[0])
            try:   # attaching registers the segment with this process's resource tracker too (CPython < 3.13); only rank 0 owns it
                from multiprocessing import resource_tracker
                resource_tracker.unregister(shm._name, "shared_memory")
            except Exception:
                pass
        shared_np = np.frombuffer(shm.buf, dtype=np.uint32, count=w * h)
        shared_ptr = shared_np.ctypes.data
        rc = torch.cuda.cudart().cudaHostRegister(shared_ptr, w * h * 4, 3)   # portable | mapped
        if int(rc) != 0:
            raise SystemExit(f"cudaHostRegister of the shared frame failed: {rc}")

    def present(k, mode):
        """N > 1: presentation of frame k on the side stream (double-buffered), under the rendering of frame k + 1.
        'device' / 'e2e': NCCL gather + rtb_untile on rank 0 (+ one device-to-host copy by the copy engine); 'e2e_host': every rank
        writes its own tiles into the shared host frame (rtb_present_host)."""
        i = k & 1
        stream.wait_event(done[i])                      # buffers i are free again (frame k - 2 has been presented)
        staging[i].copy_(tiled, non_blocking=True)
        ready[i].record(stream)
        with torch.cuda.stream(side):
            side.wait_event(ready[i])
            if mode == "e2e_host":
                ctx.present_host(shared_ptr, staging[i].data_ptr(), side.cuda_stream)
            else:
                dist.gather(staging[i], list(gathered[i].split(slots)) if rank == 0 else None, dst=0)
                if rank == 0:
                    ctx.untile_on(gathered[i].data_ptr(), world, slots, untiled[i].data_ptr(), side.cuda_stream)
                    if mode == "e2e":
                        frame_pins[i].copy_(untiled[i], non_blocking=True)
            done[i].record(side)

    def frame(k, mode):
        """mode: 'device' (inputs resident), 'e2e' (host buffers in, frame in host memory out), 'e2e_host' (N > 1: the frame reaches
        the host through rtb_present_host on every rank instead of the NCCL gather and rank 0's PCIe link)"""
        if mode != "device":   # what the host does per frame in the reference: camera + seed upload (raytracing_interface.cpp:327, composite_task.cpp:243)
            ctx.upload_raw(rtb.BUF_CAMERA, cam_pin.data_ptr(), 144)
            ctx.upload_raw(rtb.BUF_SEED, seed_pin.data_ptr(), 24)
        elif spp > 1:   # a new accumulation starts at sampleCount 0
            ctx.upload_raw(rtb.BUF_SEED, seed_pin.data_ptr(), 24)
        render()
        if world > 1:
            present(k, mode)
        elif mode == "e2e":   # presentToCpu: the rgba8 frame lands in host memory (a copy + fence, like the reference's PBO read-back:
            # frame N is copied out while frame N+1 renders; the timed region ends with the last copy landed)
            ctx.readback_async_into(rtb.TGT_RGBA8, frame_pins[e2e_count[0] & 1].data_ptr(), w * h * 4)
            e2e_count[0] += 1

    def drain():
        if world > 1:
            stream.wait_stream(side)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- instrumented frame: rays and algorithmic bytes (separate kernels, never timed) ----------------------
    ctx.upload(rtb.BUF_CAMERA, cam)
    ctx.upload(rtb.BUF_SEED, seed0)
    path = None
    if bounces:
        ctx.set_option(rtb.OPT_COUNTERS, 1)
        ctx.path_frame(bounces)
        ctx.sync()
        path = ctx.path_stats()
        c = cf = ctx.counters()
        ctx.set_option(rtb.OPT_COUNTERS, 0)
        ctx.path_frame(bounces)
        ctx.sync()
        hits_all = c.primary_hits
    else:
        ctx.set_option(rtb.OPT_COUNTERS, 1)
        ctx.dispatch(rtb.PASS_FRAME)
        ctx.sync()
        c = ctx.counters()
        hits_all = c.primary_hits
        for _ in range(spp - 1):   # every sample has its own jitter: count the hit pixels (= shadow rays) of each
            ctx.dispatch(rtb.PASS_FRAME)
            ctx.sync()
            hits_all += ctx.counters().primary_hits
        ctx.set_option(rtb.OPT_COUNTERS, 2)   # what the kernels in use fetch (a packet fetch serves 32 rays and counts once)
        ctx.dispatch(rtb.PASS_FRAME)
        ctx.sync()
        cf = ctx.counters()
        ctx.set_option(rtb.OPT_COUNTERS, 0)
        ctx.dispatch(rtb.PASS_FRAME)
        ctx.sync()
    hits_local = torch.tensor([hits_all], dtype=torch.int64, device="cuda")
    info = ctx.accel_info()   # after a plain frame: primary_packets says which nearest-hit kernel the timed frames run
    local = torch.tensor([c.primary_rays, c.shadow_rays, c.primary_nodes, c.primary_tris, c.shadow_nodes, c.shadow_tris, c.primary_hits,
                          c.shadow_occluded] + ([path.closest_rays, path.shadow_rays] if path else [0, 0]), dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(local)
        dist.all_reduce(hits_local)
    tot = local.tolist()
    if bounces:   # every ray the wavefront loop traced: camera + bounce rays (nearest hit) and one shadow ray per vertex
        rays_per_frame = tot[8] + tot[9]
    else:         # every pixel's primary ray + one shadow ray per hit pixel and shadow sample, for each of the spp replays
        rays_per_frame = w * h * spp + int(hits_local.item()) * samples

    # ---- the frame is the same frame on N GPUs ----------------------------------------------------------------------------
    checksum = None
    if world > 1:
        checksum_frame()
        present(0, "device")
        present(1, "e2e_host")
        drain()
        barrier()
        if rank == 0:
            crc_n = zlib.crc32(untiled[0].cpu().numpy().tobytes())
            crc_host = zlib.crc32(shared_np.tobytes())
            checksum = {"frame_crc32": f"{crc_n:08x}", "single_gpu_crc32": f"{crc_single:08x}", "host_frame_crc32": f"{crc_host:08x}",
                        "equal": crc_n == crc_single and crc_host == crc_single,
                        "what": "rgba8 frame after seed (0,0): rank 0 alone before partitioning, the NCCL-gathered + rtb_untile'd frame, and the shared host frame written by rtb_present_host on every rank"}
            if not checksum["equal"]:
                raise SystemExit(f"frame of {world} GPUs differs from the single-GPU frame: {checksum}")
    elif rank == 0:
        checksum = {"frame_crc32": f"{crc_single:08x}", "single_gpu_crc32": f"{crc_single:08x}", "equal": True, "what": "rgba8 frame after seed (0,0)"}

    def timed(mode: str, steps: int, warmup: int, sample_clocks: bool):
        sampler = ClockSampler(local_rank) if sample_clocks and rank == 0 else None
        if sampler:
            sampler.start()
            time.sleep(0.3)   # nvidia-smi is up and sampling before any frame is issued
        ctx.upload(rtb.BUF_SEED, seed0)
        for k in range(warmup):
            frame(k, mode)
        drain()
        barrier()
        if sampler:
            sampler.mark()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(steps):
            frame(k, mode)
        drain()
        if mode == "e2e" and world == 1:
            ctx.readback_wait()   # every frame of the timed region is in host memory
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        clocks = sampler.stop() if sampler else None
        return float(ms.item()), clocks

    total_ms, clocks = timed("device", args.steps, args.warmup, True)
    ms_per_step = total_ms / args.steps
    value = rays_per_frame / (ms_per_step * 1e-3) / 1e6

    # ---- frames timed one by one: median and spread (SURVEY.md §8d asks for a median) ----------------------------------------
    n_med = max(20, min(args.median_frames, int(10_000.0 / max(ms_per_step, 1e-3))))
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_med + 1)]
    ctx.upload(rtb.BUF_SEED, seed0)
    barrier()
    evs[0].record(stream)
    for k in range(n_med):
        frame(k, "device")
        evs[k + 1].record(stream)
    drain()
    barrier()
    per_frame = torch.tensor([evs[k].elapsed_time(evs[k + 1]) for k in range(n_med)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(per_frame, op=dist.ReduceOp.MAX)
    per_frame = np.sort(per_frame.cpu().numpy())
    median_ms = float(np.median(per_frame))

    # traversal launches alone (CUDA events recorded by the library on the same stream), averaged over a few frames
    phases = np.zeros(8)
    t_primary = t_shadow = 0.0
    if not bounces:
        ctx.set_option(rtb.OPT_FRAME_LANES, 1)   # one lane, launched directly: the launches run one after the other and can be timed alone
        ctx.set_option(rtb.OPT_FRAME_GRAPH, 0)
        n_ph = min(args.steps, 10)
        for _ in range(n_ph):
            ctx.dispatch(rtb.PASS_FRAME)
            phases += np.array(ctx.last_frame_ms())
        phases /= n_ph
        ctx.set_option(rtb.OPT_FRAME_LANES, args.lanes)
        ctx.set_option(rtb.OPT_FRAME_GRAPH, 0 if args.no_graph else 1)
        t_primary, t_shadow = float(phases[2]) * 1e-3, float(phases[5]) * 1e-3
    else:
        ctx.set_option(rtb.OPT_FRAME_OVERLAP, 0)   # the launches one after the other, so that each can be timed alone (the timed frames run
        for _ in range(3):                         # the occlusion launch of depth d beside the nearest-hit launch of depth d + 1)
            ctx.path_frame(bounces)
        ctx.sync()
        path = ctx.path_stats()
        ctx.set_option(rtb.OPT_FRAME_OVERLAP, 0 if args.no_overlap else 1)
        t_primary, t_shadow = path.closest_ms * 1e-3, path.shadow_ms * 1e-3
    barrier()

    e2e_ms, _ = timed("e2e", args.steps, args.warmup, False)
    e2e_value = rays_per_frame / (e2e_ms / args.steps * 1e-3) / 1e6
    e2e_host_ms = None
    if world > 1:
        e2e_host_ms, _ = timed("e2e_host", args.steps, args.warmup, False)

    if world > 1:
        torch.cuda.cudart().cudaHostUnregister(shared_ptr)
        del shared_np
        shm.close()
        if rank == 0:
            shm.unlink()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ------------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    try:   # second denominator: L2 read bandwidth measured on this GPU by the library's streaming probe (48 MiB, L2-only loads)
        l2_peak = ctx.probe_l2_read_gbs(48 << 20)
    except Exception:
        l2_peak = None
    n_tri = int(scene["info"][2])
    bvh_bytes = info.node_count * info.node_bytes + n_tri * info.tri_record_bytes
    l2_bytes = int(torch.cuda.get_device_properties(local_rank).L2_cache_size)
    l2_resident = bvh_bytes <= l2_bytes and l2_peak is not None
    traffic = {}
    try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures (profiles/)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "trace_traffic.json"))).get(args.workload, {})
    except Exception:
        pass

    def roof(kernel, key, algorithmic, fetched, secs, launches=1):
        """achieved = bytes the launch really FETCHES / its duration.  For a per-ray kernel that is the algorithmic figure (every
        ray fetches what its own traversal needs); a packet kernel fetches each record once for up to 32 rays, so the per-ray
        algorithmic bytes it serves are reported beside it as `amortisation` and never enter `frac`."""
        ach = fetched / secs / 1e9 if secs > 0 else 0.0
        bound, peak = ("l2", l2_peak) if l2_resident else ("hbm", hbm_peak)
        r = {"bound": bound, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if peak else None, "traffic": traffic.get(key),
             "kernel": kernel, "bytes_per_launch_rank0": fetched / launches, "algorithmic_bytes_per_launch_rank0": algorithmic / launches,
             "ms_per_launch": secs * 1e3 / launches, "launches_per_frame": launches,
             "peak_source": "rtb_probe_l2_read_gbs: 48 MiB buffer streamed with ld.global.cg, best of 5, this GPU, this run" if l2_resident else hbm_src,
             "other_peak": {"hbm": hbm_peak, "l2": l2_peak}}
        if algorithmic != fetched and fetched > 0:
            r["amortisation"] = algorithmic / fetched
        return r

    nb, tb = info.node_bytes, info.tri_record_bytes
    alg_primary, alg_shadow = c.primary_nodes * nb + c.primary_tris * tb, c.shadow_nodes * nb + c.shadow_tris * tb
    fet_primary, fet_shadow = cf.primary_nodes * nb + cf.primary_tris * tb, cf.shadow_nodes * nb + cf.shadow_tris * tb
    if bounces:
        nl = path.closest_launches
        nearest = roof("k_trace_cwbvh<MODE_CLOSEST> (camera + bounce rays: one traversal per ray, %d launches per frame)" % nl, "k_trace_cwbvh<0,0>",
                       alg_primary, alg_primary, t_primary, nl)
        occlusion = roof("k_trace_cwbvh<MODE_ANY_BYTES> (one shadow ray per path vertex, %d launches per frame)" % path.shadow_launches, "k_trace_cwbvh<2,0>",
                         alg_shadow, alg_shadow, t_shadow, path.shadow_launches)
    else:
        if info.primary_packets:
            kname = "k_trace_cwbvh_frustum" if info.primary_packets == 3 else "k_trace_cwbvh_packet"
            nearest = roof(kname + " (nearest-hit launch: one warp-cooperative traversal per 8x4-pixel patch, camera rays generated and G-buffer written in the same launch)",
                           kname, alg_primary, fet_primary, t_primary)
        else:
            nearest = roof("k_trace_cwbvh<MODE_CLOSEST> (nearest-hit launch, one traversal per ray)", "k_trace_cwbvh<0,0>", alg_primary, alg_primary, t_primary)
        occlusion = roof("k_trace_cwbvh<MODE_ANY_BITS> (occlusion launch, one traversal per ray)", "k_trace_cwbvh<1,0>", alg_shadow, alg_shadow, t_shadow)
    # the dominant kernel is whichever traversal launch takes longer on this workload; the other one rides along
    if t_shadow > t_primary:
        roofline, other_key, other = occlusion, "nearest_hit_launch", nearest
    else:
        roofline, other_key, other = nearest, "occlusion_launch", occlusion
    roofline["kernel"] += " - the dominant kernel"
    roofline.update({
        "note": "bytes = node_bytes x node records + 48 B x triangle records the launch fetches (instrumented kernels, separate from the timed ones). "
                "The acceleration structure is %.0f MB against %.0f MB of L2: %s. `traffic` = dram bytes read + written by that launch in the committed "
                "ncu --set full capture of this workload (profiles/trace_traffic.json)" % (
                    bvh_bytes / 1e6, l2_bytes / 1e6, "L2-resident, the L2 read bandwidth bounds the traversal" if l2_resident else "HBM-resident, the HBM bandwidth bounds the traversal"),
        "nodes_per_primary_ray": tot[2] / max(tot[0], 1), "tris_per_primary_ray": tot[3] / max(tot[0], 1),
        "nodes_per_shadow_ray": tot[4] / max(tot[1], 1), "tris_per_shadow_ray": tot[5] / max(tot[1], 1),
        other_key: other})

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        # about 15 s of host work: the brute-force loops cost ~ pixels x 1.5 rays x triangles
        n_px = max(256, min(16384, int(1.6e10 / n_tri))) if n_tri > 100_000 else min(w * h, 1 << 20)
        reps = 1 if n_tri > 100_000 else 25   # small scenes: a frame takes milliseconds on the host; best of 25 passes
        v, rays, secs, threads = cpu_time_sample(wl, n_px, repeats=reps)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{n_px} random pixels of the same {w}x{h} frame ({rays} rays, {secs:.3f} s per pass, best of {reps}), brute force over all {n_tri} triangles and the other primitives as the reference "
                         f"shaders do (first hit + shadow ray; no bounces); oracle built -O3 -march=native on this host ({host_cpu_model()}); a stated baseline, not the target"}

    if bounces:
        kernels_per_frame = path.kernel_launches
    else:   # init, fused camera-ray launch (or raygen, nearest hit, finish), shadow-ray set-up, occlusion, lighting + composite
        kernels_per_frame = (5 if info.primary_packets == 3 else 7) * spp
    gather_path = "NCCL gather to rank 0 + rtb_untile + one device-to-host copy (copy engine) over rank 0's PCIe link, all on a second stream under the next frame"
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 168 * world, "d2h_bytes_per_step": w * h * 4, "ms_per_step": e2e_ms / args.steps,
           "path": gather_path if world > 1 else "rtb_readback_async of the rgba8 frame into pinned memory (copy engine, overlaps the next frame)"}
    if e2e_host_ms is not None:
        # two routes from N GPUs to one host frame, both through the public API and timed alike: the faster one is the e2e number
        # (rank 0's single PCIe link bounds the gather route from N = 8 on), the other is listed beside it
        host = {"value": rays_per_frame / (e2e_host_ms / args.steps * 1e-3) / 1e6, "ms_per_step": e2e_host_ms / args.steps,
                "path": "no gather: every rank's rtb_present_host kernel writes its tiles into ONE page-locked host frame in shared memory (%d PCIe links); the kernel takes SM slots between the next frame's persistent launches" % world}
        gather = {"value": e2e["value"], "ms_per_step": e2e["ms_per_step"], "path": gather_path}
        if host["value"] > gather["value"]:
            e2e.update(host)
            e2e["via_nccl_gather"] = gather
        else:
            e2e["via_present_host"] = host
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "ms_per_step_median": median_ms,
        "per_frame_ms": {"frames": int(n_med), "median": median_ms, "p10": float(per_frame[int(0.1 * (n_med - 1))]), "p90": float(per_frame[int(0.9 * (n_med - 1))]),
                         "min": float(per_frame[0]), "max": float(per_frame[-1]), "value_at_median": rays_per_frame / (median_ms * 1e-3) / 1e6},
        "config": {"workload": wl["desc"], "spp": spp, "bounces": bounces, "rays_per_frame": rays_per_frame,
                   "hit_fraction": (path.closest_rays_at_depth[1] / max(path.closest_rays_at_depth[0], 1) if bounces else tot[6] / (w * h)),
                   "l2": "no explicit flush: each frame streams ~0.9 GB of ray / G-buffer data through the 126 MB L2 between traversal launches; the BVH (nodes + triangles) stays resident as it would in steady-state rendering",
                   "partition": f"{world} ranks x interleaved 32x32-pixel blocks, scene+BVH replicated, NCCL gather of rgba8 tiles to rank 0 + rtb_untile on a second stream under the next frame" if world > 1 else "single GPU",
                   "bvh": {"nodes": info.node_count, "node_bytes": info.node_bytes, "leaves": info.leaf_count, "depth": info.max_depth,
                           "sah_cost": info.sah_cost, "build_ms": info.build_ms, "leaf_node_extent": info.leaf_node_extent, "bytes": bvh_bytes,
                           "builder": {0: "host (binned SAH, optimal collapse)", 1: "device (Morton sort, radix tree, greedy collapse)"}.get(info.builder, "?")},
                   "primary_packets": {0: "per ray", 1: "union packets", 3: "frustum packets"}.get(info.primary_packets, str(info.primary_packets)), "packets_option": args.packets,
                   "shadow_order": args.shadow_order, "frame_lanes": args.lanes, "frame_graph": not args.no_graph, "frame_overlap": (not args.no_overlap) if bounces else not (args.no_overlap or args.no_graph or args.lanes == 2),
                   "phase_note": "phase_ms_rank0 is measured with direct launches in one lane (back to back); the timed frames replay the frame's CUDA graphs" if not args.no_graph else "direct launches",
                   "phase_ms_rank0": {k: float(v) for k, v in zip(["init", "raygen", "trace_primary", "finish", "shadowgen", "trace_shadow", "shade", "total"], phases)}},
        "clocks": clocks,
        "checksum": checksum,
        "e2e": e2e,
        "gpu_launches": (kernels_per_frame * world + (1 if world > 1 else 0)) * args.steps,
        "roofline": roofline,
    }
    if path:
        out["config"]["path"] = path.as_dict()
    if cpu:
        out["cpu_baseline"] = cpu
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
