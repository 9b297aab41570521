#!/usr/bin/env python
"""bench.py — frames/s of the sort-middle frame pipeline on the BASELINE.json workloads, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Headline workload (config.workload): BASELINE.json configs[1] — the synthetic Sponza-scale "hall" scene (263 888
triangles in 25 draws, 25 Morton-tiled mip-mapped textures) at 1920x1080 — on the 1024-camera closed path of configs[4].
A unit is ONE FRAME: BeginFrame -> ClearFrameBuffer -> 25 x DrawIndexed -> EndFrame returns with the frame complete in
the tile buffers.  A step is `--frames-per-step` frames; frames are independent, so with N GPUs every
rank renders its own frames (weak scaling, no data-path collective).

  value        frames/s with the scene resident in HBM: only the 25 draw descriptors (MVPs) go host->device per frame.
               Draws carry HOST pointers like the reference's DrawCall; the library mirrors those buffers on the device
               at their first use and finds the mirrors by pointer afterwards.
  e2e          frames/s through the same C-ABI calls with host buffers: per frame the draw table is uploaded and the
               finished colour tiles (tiles*16 KiB) are copied back into pinned host memory, inside the timed region.
               e2e.d2h_ceiling_gbs = the same copies alone, ALL ranks at the same time (barrier first): what the box's
               PCIe + host memory deliver to N GPUs at once.
  roofline     the dominant kernel: algorithmic bytes per launch / its mean duration,
               measured with CUDA events on the library's own stream in a second pass over the same frames; the kernel
               is instruction-issue bound (`bound`), so the issue-slot fraction stands beside the HBM fraction.
  configs      the other BASELINE.json configurations, each with frames/s, one frame at a time, kernel times and the
               reference on the host cores: configs[0] (1280x720 cube grid, one draw and 100 draws), configs[2] (1 M random
               triangles), configs[3] (the hall at 3840x2160: one GPU, and with --gpus N > 1 the screen-tile split).
  cpu_baseline the UNMODIFIED reference (oracle/_ref/libsrref_fast.so, its own flags) on a bounded sample of the same
               frames at the best of a small thread-count sweep, plus the single-threaded parity build — N=1 only.
`--impl reference` times only that CPU arm, per step a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT = 1920, 1080
PATH_FRAMES = 1024
METRIC = "frames/sec at 1920x1080 (hall scene, 263888 tris in 25 textured draws per frame)"
WORKLOAD = "hall_1080p_camera_path (BASELINE.json configs[1] scene on the configs[4] camera path)"


def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def _bind_to_gpu_numa(local: int):
    """Multi-GPU runs: pin this rank to the CPUs NVML reports as local to its GPU BEFORE any pinned host memory is
    allocated, so that the read-back buffer lives on the GPU's own NUMA node.  Returns the CPU list, or None if NVML /
    the topology gives nothing to bind to (one NUMA node: every CPU is local to every GPU)."""
    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        try:
            pr = torch.cuda.get_device_properties(local)
            bus_id = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        allowed = os.sched_getaffinity(0)
        cpus = sorted(c for c in range(ncpu) if (mask[c // 64] >> (c % 64)) & 1 and c in allowed)
        if not cpus or len(cpus) == len(allowed):
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, n in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(n)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ---------------------------------------------------------------------------------------------------------------
# workloads: the five BASELINE.json configurations as concrete scenes
# ---------------------------------------------------------------------------------------------------------------
class Workload:
    def __init__(self, name, scene, mvps, baseline_config):
        self.name, self.scene, self.mvps, self.baseline_config = name, scene, mvps, baseline_config

    def config(self):
        """The part of `config` both arms print identically (what is rendered; how it is run lives elsewhere)."""
        sc = self.scene
        return {"workload": self.name, "baseline_config": self.baseline_config, "width": sc.width, "height": sc.height,
                "tris_per_frame": sc.num_tris, "draws": len(sc.draws), "textures": len(sc.textures)}

    def frames(self, first, n):
        if self.mvps is None:
            return None
        return self.mvps[(first + np.arange(n)) % len(self.mvps)]


def make_workloads(which):
    from softrast_b200 import scenes

    out = {}
    hall = None
    if {"hall", "hall4k"} & set(which):
        hall = scenes.hall_scene(WIDTH, HEIGHT)
        path = scenes.hall_camera_path(hall, PATH_FRAMES)
    if "hall" in which:
        out["hall"] = Workload(WORKLOAD, hall, path, "configs[1] scene, configs[4] camera path")
    if "cubes" in which:
        out["cubes"] = Workload("cubes_720p_1draw (BASELINE.json configs[0]: 100x100 textured cubes, one draw)",
                                scenes.cube_grid(), None, "configs[0]")
    if "cubes100" in which:
        out["cubes100"] = Workload("cubes_720p_100draws (BASELINE.json configs[0]: 100x100 textured cubes, one draw per row)",
                                   scenes.cube_grid(draws=100), None, "configs[0]")
    if "rand" in which:
        out["rand"] = Workload("rand_1m_1080p (BASELINE.json configs[2]: 1 M small random triangles, one draw)",
                               scenes.random_tris(), None, "configs[2]")
    if "hall4k" in which:
        # the same scene at 3840x2160: same aspect ratio, so the same MVPs
        h4 = copy.copy(hall)
        h4.width, h4.height = 3840, 2160
        out["hall4k"] = Workload("hall_4k (BASELINE.json configs[3]: the hall scene at 3840x2160, one frame at a time)",
                                 h4, path, "configs[3]")
    return out


# ---------------------------------------------------------------------------------------------------------------
# the reference on the host cores
# ---------------------------------------------------------------------------------------------------------------
def thread_candidates():
    """Renderer.cpp:141 starts LogicalCoreCount() - 1 workers; on a box with many vCPUs the reference's mutex task queue
    scales negatively, so the baseline is its BEST over a small sweep of logical core counts."""
    n = os.cpu_count() or 1
    return sorted({t for t in (8, 16, n) if 1 <= t <= n})


def reference_fps(wl, threads, frames, warm=2):
    """frames/s of the compiled reference (fast build) with `threads` logical cores on `frames` frames of a workload."""
    from oracle import refharness as rh

    sc = wl.scene
    r = rh.RefRenderer(sc.width, sc.height, threads, "fast")
    r.load_scene(sc)
    r.render_frames(warm, wl.frames(0, warm))
    t0 = time.perf_counter()
    r.render_frames(frames, wl.frames(warm, frames))
    dt = time.perf_counter() - t0
    used = r.threads
    r.close()
    return frames / dt, used


def reference_best(wl, seconds=2.0):
    """Best thread count for this workload: a short probe per candidate.  Returns (threads, {threads: fps})."""
    sweep = {}
    for t in thread_candidates():
        probe, _ = reference_fps(wl, t, 3, warm=1)
        n = int(min(256, max(4, seconds * probe)))
        fps, used = reference_fps(wl, t, n)
        sweep[used] = fps
    best = max(sweep, key=sweep.get)
    return best, sweep


def cpu_baseline_for(wl, seconds, with_st):
    """cpu_baseline object of one workload (N = 1 only): the reference at its best thread count on a bounded sample, and
    (with_st) the single-threaded parity build — BASELINE.md §3.1's 1-core figure."""
    from oracle import refharness as rh

    if not rh.ref_available("fast"):
        from softrast_b200.capi import harvest_rcp_table

        sc = wl.scene
        r = rh.PortRenderer(sc.width, sc.height, harvest_rcp_table(16))
        r.load_scene(sc)
        t0 = time.perf_counter()
        r.render_frames(2, wl.frames(0, 2))
        dt = time.perf_counter() - t0
        r.close()
        return {"value": 2 / dt, "unit": "frames/s", "cores": 1, "kind": "port", "sample": "2 frames (C restatement, one core)"}
    best, sweep = reference_best(wl, seconds=min(2.0, seconds / 4))
    probe = sweep[best]
    n = int(min(1024, max(8, seconds * probe)))
    t0 = time.perf_counter()
    fps, used = reference_fps(wl, best, n)
    dt = time.perf_counter() - t0
    out = {"value": fps, "unit": "frames/s", "cores": used, "kind": "reference",
           "sample": f"{n} consecutive frames of the workload ({dt:.1f} s incl. scene load)",
           "thread_sweep_fps": {str(k): round(v, 2) for k, v in sweep.items()}, "host_cpus": os.cpu_count()}
    if with_st and rh.ref_available("parity"):
        sc = wl.scene
        r = rh.RefRenderer(sc.width, sc.height, 1, "parity")
        r.load_scene(sc)
        r.render_frames(1, wl.frames(0, 1))
        k = 3
        t0 = time.perf_counter()
        r.render_frames(k, wl.frames(1, k))
        out["oracle_st_fps"] = k / (time.perf_counter() - t0)
        out["oracle_st_note"] = "the parity build (-ffp-contract=off), SR_DEBUG_SINGLE_THREADED order: one core"
        r.close()
    return out


def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path on the box's host cores, bounded sample per step."""
    if rank != 0:
        return
    from oracle import refharness as rh

    wls = make_workloads(["hall"] + ([] if args.no_configs else ["cubes", "cubes100", "rand", "hall4k"]))
    wl = wls["hall"]
    sample = args.ref_frames_per_step
    if rh.ref_available("fast"):
        best, sweep = reference_best(wl, seconds=1.5)
        r = rh.RefRenderer(WIDTH, HEIGHT, best, "fast")
        kind, threads = "reference", r.threads
    else:
        from softrast_b200.capi import harvest_rcp_table

        r = rh.PortRenderer(WIDTH, HEIGHT, harvest_rcp_table(16))
        kind, threads, sweep = "port", 1, {}
    r.load_scene(wl.scene)
    f0 = 0
    for _ in range(args.warmup):
        r.render_frames(sample, wl.frames(f0, sample))
        f0 += sample
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r.render_frames(sample, wl.frames(f0, sample))
        f0 += sample
    dt = time.perf_counter() - t0
    fps = args.steps * sample / dt
    r.close()
    configs = {}
    if kind == "reference":
        for key in ("cubes", "cubes100", "rand", "hall4k"):
            if key in wls:
                b, sw = reference_best(wls[key], seconds=1.0)
                configs[key] = {"config": wls[key].config(), "frames_per_s": sw[b], "cores": b,
                                "thread_sweep_fps": {str(k): round(v, 2) for k, v in sw.items()}}
    line = {
        "impl": "reference",
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": wl.config(),
        "run": {"frames_per_step": sample, "threads": threads,
                "thread_sweep_fps": {str(k): round(v, 2) for k, v in sweep.items()},
                "note": "threads = the logical core count handed to the reference (it starts threads - 1 workers, "
                        "Renderer.cpp:141); the best of the sweep is used"},
        "mtris_per_s": fps * wl.scene.num_tris / 1e6,
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                         "sample": f"{sample} frames per step of the same camera path, {args.steps} steps"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "configs": configs,
        "host_cpus": os.cpu_count(),
    }
    print(json.dumps(line), flush=True)


def algorithmic_bytes(scene, counters, winners):
    """SURVEY.md §8d per-kernel algorithmic bytes for one frame (every datum moved once)."""
    T = scene.num_tris
    idx_bytes = sum(d.indices.size * d.indices.dtype.itemsize for d in scene.draws)
    vu = sum(int(np.unique(d.indices).size) for d in scene.draws)
    Ts, R = counters["tris_setup"], counters["tile_refs"]
    P, Pc = scene.tiles[0] * scene.tiles[1] * 4096, counters["pixels_covered"]
    U = 0
    for t in range(winners.shape[0]):
        w = winners[t].ravel()
        U += int(np.unique(w[w != 0xFFFFFFFF]).size)
    X = Pc  # estimate one new texel per covered pixel
    return {
        "setup": idx_bytes + vu * 32 + 164 * Ts,
        "bin_fill": 44 * Ts + 4 * R,
        "raster": 60 * R,
        "shade": 108 * U + 4 * X + 8 * P,
        "counts": {"T": T, "V_u": vu, "T_s": Ts, "R": R, "P": P, "P_c": Pc, "U": U, "X": X},
    }


def kernel_times(r0, wl, frames, first=0):
    """Mean per-kernel CUDA-event durations of `frames` frames on one context, one frame at a time."""
    r0.ctx.set_frames_in_flight_hint(1)
    r0.ctx.set_timing(True)
    acc = {}
    mv = wl.frames(first, frames)
    for f in range(frames):
        r0.render(mvps=None if mv is None else mv[f])
        for k, v in r0.ctx.kernel_times().items():
            acc[k] = acc.get(k, 0.0) + v / frames
    r0.ctx.set_timing(False)
    acc.pop("detile", None)
    if acc.get("tile_scan", 1.0) == 0.0 and "clip" in acc:
        # the tile scan runs in the tail of the clip kernel (no launch of its own)
        acc["clip+tile_scan"] = acc.pop("clip") + acc.pop("tile_scan")
    return acc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=256,
                    help="frames of one step (one srb_render_frames batch); the read-back buffer holds one step")
    ap.add_argument("--in-flight", type=int, default=12, help="contexts (CUDA streams) rendering frames concurrently")
    ap.add_argument("--ref-frames-per-step", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="headline workload only (skip configs[0], [2], [3])")
    ap.add_argument("--resident", action="store_true", help="bind explicit device buffers instead of host pointers")
    ap.add_argument("--no-share", action="store_true", help="every frame in flight gets its own copy of the scene")
    ap.add_argument("--no-geometry-upload", action="store_true", help="skip the e2e leg that re-uploads the geometry every frame")
    ap.add_argument("--no-numa-bind", action="store_true", help="N > 1: do not pin the rank to its GPU's local CPUs")
    args = ap.parse_args()
    rank, world, local = _dist_env()

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    numa_cpus = None
    if world > 1:
        if not args.no_numa_bind:
            numa_cpus = _bind_to_gpu_numa(local)
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from softrast_b200 import capi, sharding

    wls = make_workloads(["hall"] + ([] if args.no_configs else ["cubes", "cubes100", "rand", "hall4k"]))
    wl = wls["hall"]
    scene, mvps_all = wl.scene, wl.mvps
    F = args.frames_per_step
    # one context per frame in flight; they share ONE device copy of the scene (srb_create_shared)
    # Draws carry HOST pointers, like the reference's DrawCall (Renderer.h:112-141): the library mirrors the buffers on
    # the device at the first DrawIndexed and finds them by pointer afterwards (--resident: explicit device buffers).
    res = bool(args.resident)

    def make_renderers(sc, n):
        rs = [capi.SceneRenderer(sc, device=local, resident=res)]
        for _ in range(max(1, n) - 1):
            rs.append(capi.SceneRenderer(sc, device=local, resident=res, share=None if args.no_share else rs[0]))
        return rs

    renderers = make_renderers(scene, args.in_flight)
    colour_bytes = renderers[0].fb.num_tiles * 16384
    # read-back buffer: with several GPUs copying at once the box's host side is what bounds e2e, and page-locked transparent
    # huge pages raise what it takes (8 GPUs: 119 -> 165 GB/s, profiles/README.md); one GPU alone is a little faster into
    # plain cudaHostAlloc memory (51.9 vs 50.5 GB/s)
    pinned, pinned_kind = None, "cudaHostAlloc"
    if world > 1:
        try:
            pinned = capi.host_alloc_ex(F * colour_bytes, capi.HOST_HUGE_PAGES)
            pinned_kind = "2 MiB-aligned, MADV_HUGEPAGE, cudaHostRegister"
        except capi.SrbError:
            pinned = None
    if pinned is None:
        pinned = capi.host_alloc(F * colour_bytes)
    draw_upload_bytes = 64 + 136 * len(scene.draws)  # control block + sizeof(DrawDev) per draw, uploaded every frame

    def frames_of(step):  # every rank walks its own arc of the closed camera path
        return mvps_all[sharding.frames_for_rank(step, F, rank, world, PATH_FRAMES)]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run(steps, first_step, e2e, rs):
        for s in range(steps):
            rs[0].ctx.flush_l2(256 << 20)  # evict L2 between steps (inside the timed region, ~40 us)
            capi.render_frames(rs, F, frames_of(first_step + s), pinned if e2e else None, colour_bytes)

    def timed(e2e, rs=None):
        rs = renderers if rs is None else rs
        run(args.warmup, 0, e2e, rs)
        barrier()
        launches0 = sum(r.ctx.launch_count() for r in rs)
        capi.timer_mark(rs, 0)
        run(args.steps, args.warmup, e2e, rs)
        capi.timer_mark(rs, 1)
        ms = capi.timer_elapsed_ms(rs, 0, 1)
        barrier()
        launches = sum(r.ctx.launch_count() for r in rs) - launches0
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    clocks = ClockSampler(local)
    clocks.start()
    ms_dev, launches = timed(False)
    clock_info = clocks.stop()
    ms_e2e, _ = timed(True)

    # SURVEY 8d's second GPU number: the geometry crosses PCIe EVERY frame as well (the reference reads the application's
    # vertex and index arrays in place, Renderer.h:112-141).  Same calls, contexts created with SRB_FLAG_UPLOAD_ALWAYS:
    # every DrawIndexed re-uploads its index / position / attribute arrays from the (pinned) host copy of the scene.
    geo = None
    if not args.no_geometry_upload:
        import ctypes as C

        from softrast_b200 import scenes

        sc_up = copy.copy(scene)
        sc_up.draws = []
        geo_bytes = 0
        pins = []
        for d in scene.draws:
            arrs = []
            for a in (np.ascontiguousarray(d.vertices, dtype=np.float32), np.ascontiguousarray(d.indices)):
                p = capi.host_alloc(a.nbytes)
                pins.append(p)
                view = np.frombuffer((C.c_char * a.nbytes).from_address(p), dtype=a.dtype).reshape(a.shape)
                view[...] = a
                arrs.append(view)
                geo_bytes += a.nbytes
            sc_up.draws.append(scenes.Draw(arrs[0], arrs[1], d.mvp, d.shader, d.texture, d.uv_offset))
        ups = [capi.SceneRenderer(sc_up, device=local, resident=False, flags=capi.FLAG_UPLOAD_ALWAYS) for _ in range(len(renderers))]
        ms_geo, _ = timed(True, ups)
        ms_geo_in, _ = timed(False, ups)  # the same without the colour read-back: how fast the geometry comes in alone
        geo = {"value": world * args.steps * F / (ms_geo * 1e-3), "unit": "frames/s",
               "without_readback": world * args.steps * F / (ms_geo_in * 1e-3),
               "h2d_gbs_without_readback": (geo_bytes + draw_upload_bytes) * F / (ms_geo_in / args.steps * 1e-3) / 1e9,
               "h2d_bytes_per_step": (geo_bytes + draw_upload_bytes) * F, "d2h_bytes_per_step": colour_bytes * F,
               "ms_per_step": ms_geo / args.steps,
               "note": "e2e with the scene's vertex and index arrays re-uploaded from pinned host memory at every DrawIndexed "
                       "(SRB_FLAG_UPLOAD_ALWAYS), one device copy of the scene per frame in flight"}
        for r in ups:
            r.close()
        for p in pins:
            capi.host_free(p)

    # What the box delivers for the same copies (device -> pinned host, one frame's colour tiles per call): every rank
    # copies AT THE SAME TIME (barrier first), so at N > 1 this is the concurrent ceiling of PCIe + host memory, and alone
    # (before the barrier, ranks one after the other) the ceiling of one link.
    def measure_d2h_gbs(reps=64):
        import ctypes as C

        ctx = capi.RenderContext(local)
        ms = C.c_float()
        capi.lib.srb_debug_d2h_copies(ctx.h, C.c_void_p(pinned), colour_bytes, 4, C.byref(ms))
        barrier()
        capi.lib.srb_debug_d2h_copies(ctx.h, C.c_void_p(pinned), colour_bytes, reps, C.byref(ms))
        gbs = colour_bytes * reps / (ms.value * 1e-3) / 1e9
        ctx.close()
        if dist is not None:
            t = torch.tensor([gbs], device="cuda", dtype=torch.float64)
            lo = t.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return float(lo.item()), float(t.item())
        return gbs, gbs

    d2h_min_gbs, d2h_sum_gbs = measure_d2h_gbs()

    # one frame at a time (BASELINE configs[1] literally: "one frame"): the same calls on ONE context, next frame submitted
    # only after the previous one is complete
    def single_frame_us(rs, w, frames=64):
        one = rs[:1]
        capi.render_frames(one, 16, w.frames(0, 16))
        best = 1e30
        for rep in range(2):
            capi.timer_mark(one, 2)
            capi.render_frames(one, frames, w.frames(16, frames))
            capi.timer_mark(one, 3)
            best = min(best, capi.timer_elapsed_ms(one, 2, 3) / frames * 1e3)
        return best

    single_us = single_frame_us(renderers, wl, F)

    total_frames = world * args.steps * F
    fps = total_frames / (ms_dev * 1e-3)
    fps_e2e = total_frames / (ms_e2e * 1e-3)

    # ---- roofline pass: per-kernel durations with CUDA events on the library's stream, same frames -------------
    r0 = renderers[0]
    kernel_us = kernel_times(r0, wl, min(F, 32), first=args.warmup * F)
    counters = r0.ctx.counters()
    winners = r0.ctx.winners(r0.fb.num_tiles)
    alg = algorithmic_bytes(scene, counters, winners)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # per-launch DRAM traffic and executed warp instructions from the committed ncu captures of this workload
    # (profiles/ncu_counts.json, regenerated by profiles/ncu_counts.sh at the tree it names)
    ncu = {}
    npath = os.path.join(ROOT, "profiles", "ncu_counts.json")
    if os.path.exists(npath):
        ncu = json.load(open(npath))
    sm_clock_hz = (clock_info.get("sm_mhz") or 1965.0) * 1e6
    issue_peak = 148 * 4 * sm_clock_hz  # warp instructions/s: 4 schedulers per SM, one issue per clock
    kernels = {}
    for k in ("setup", "bin_fill", "raster", "shade"):  # clip + tile scan are reported in kernel_us_per_frame
        gbs = alg[k] / (kernel_us[k] * 1e-6) / 1e9 if kernel_us.get(k) else None
        kernels[k] = {"us": kernel_us.get(k), "alg_bytes": alg[k], "achieved_gbs": gbs,
                      "frac": gbs / peak if gbs else None,
                      "traffic": ncu.get(k, {}).get("dram_bytes")}
        wi = ncu.get(k, {}).get("warp_inst")
        if wi and kernel_us.get(k):
            rate = wi / (kernel_us[k] * 1e-6)
            kernels[k]["issue"] = {"warp_inst": wi, "achieved_ginst_s": rate / 1e9, "peak_ginst_s": issue_peak / 1e9,
                                   "frac": rate / issue_peak}
    dom = max(("setup", "bin_fill", "raster", "shade"), key=lambda k: kernel_us.get(k, 0.0))
    roofline = {"kernel": dom, "bound": "issue", "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac"], "traffic": kernels[dom]["traffic"], "peak_source": peak_src,
                "note": "achieved / peak / frac = algorithmic bytes against the measured HBM copy peak, as the contract "
                        "defines them; the kernel is NOT HBM bound (a frame's working set stays in the 126 MB L2): what "
                        "bounds it is instruction issue — `issue`: executed warp instructions per launch (ncu capture "
                        "named in profiles/ncu_counts.json) / measured duration, against 148 SMs x 4 schedulers x SM clock",
                "issue": kernels[dom].get("issue")}
    wi_frame = sum(v.get("warp_inst", 0) for v in ncu.values() if isinstance(v, dict))

    line = {
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": wl.config(),
        "run": {"frames_per_step": F, "frames_in_flight": len(renderers), "parallelism": f"frame-parallel x{world}",
                "l2": "256 MiB device memset between steps (L2 flush)",
                "frame": "one CUDA graph per frame: head upload, set-up, clip + tile scan, bin fill, raster, shade"},
        "mtris_per_s": fps * scene.num_tris / 1e6,
        "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": draw_upload_bytes * F,
                "d2h_bytes_per_step": colour_bytes * F, "ms_per_step": ms_e2e / args.steps, "host_buffer": pinned_kind,
                "d2h_gbs_per_gpu": colour_bytes * F / (ms_e2e / args.steps * 1e-3) / 1e9,
                "d2h_ceiling_gbs_per_gpu": d2h_min_gbs, "d2h_ceiling_gbs_all_gpus": d2h_sum_gbs,
                "frac_of_ceiling": (colour_bytes * F / (ms_e2e / args.steps * 1e-3) / 1e9) / d2h_min_gbs,
                "note": "bound by the PCIe read-back of the finished colour tiles; d2h_ceiling_* = the same copies alone into "
                        "the same buffer, all ranks at the same time after a barrier, measured in this run; per_gpu = the "
                        "slowest rank, which is what a step (max over ranks) can reach: frac_of_ceiling = d2h_gbs_per_gpu / it"},
        "e2e_geometry_upload": geo,
        "issue_frac_whole_frame": (wi_frame * fps / world / issue_peak) if wi_frame else None,
        "single_frame": {"us_per_frame": single_us, "frames_per_s": 1e6 / single_us,
                         "note": "one frame in flight (a frame is submitted when the previous one is complete)"},
        "gpu_launches": launches,
        "numa_bound_cpus": len(numa_cpus) if numa_cpus else None,
        "clocks": clock_info,
        "roofline": roofline,
        "kernels": kernels,
        "kernel_us_per_frame": kernel_us,
        "counters": counters,
        "alg_counts": alg["counts"],
    }
    for r in renderers:
        r.close()
    capi.host_free(pinned)

    # ---- the other BASELINE.json configurations ----------------------------------------------------------------
    configs = {}
    if not args.no_configs:
        for key in ("cubes", "cubes100", "rand", "hall4k"):
            w = wls[key]
            n_fl = 1 if key == "hall4k" else min(8, args.in_flight)
            rs = make_renderers(w.scene, n_fl)
            frames = {"cubes": 256, "cubes100": 256, "rand": 64, "hall4k": 64}[key]
            capi.render_frames(rs, min(32, frames), w.frames(0, min(32, frames)))
            barrier()
            best = 1e30
            for rep in range(2):
                rs[0].ctx.flush_l2(256 << 20)
                capi.timer_mark(rs, 0)
                capi.render_frames(rs, frames, w.frames(32, frames))
                capi.timer_mark(rs, 1)
                best = min(best, capi.timer_elapsed_ms(rs, 0, 1))
            entry = {"config": w.config(), "frames_in_flight": n_fl,
                     "frames_per_s": frames / (best * 1e-3), "mtris_per_s": frames / (best * 1e-3) * w.scene.num_tris / 1e6,
                     "single_frame_us": single_frame_us(rs, w, min(frames, 64)),
                     "kernel_us": {k: round(v, 2) for k, v in kernel_times(rs[0], w, 8).items()},
                     "counters": rs[0].ctx.counters()}
            for r in rs:
                r.close()
            if rank == 0 and world == 1 and not args.no_cpu_baseline:
                entry["cpu_baseline"] = cpu_baseline_for(w, seconds=3.0, with_st=False)
                entry["vs_cpu_baseline"] = entry["frames_per_s"] / entry["cpu_baseline"]["value"]
            configs[key] = entry
        if world > 1:
            configs["hall4k_tile_split"] = tile_split(wls["hall4k"], rank, world, local, dist, torch, capi)
    line["configs"] = configs

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_for(wl, seconds=8.0, with_st=True)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def tile_split(w, rank, world, local, dist, torch, capi, frames=64):
    """BASELINE.json configs[3]: ONE 3840x2160 frame at a time, screen-tile split across the ranks with the composite done
    by peer stores over NVLink (every rank's shade kernel writes its tiles into rank 0's framebuffer; completion is an
    arrival stamp per rank in rank 0's memory that rank 0's stream waits on — no host barrier inside a frame)."""
    scene = w.scene
    blob = [None]
    if rank == 0:
        r = capi.SceneRenderer(scene, device=local)
        blob[0] = r.ctx.export_framebuffer(r.fb)
    dist.broadcast_object_list(blob, src=0)
    if rank != 0:
        r = capi.SceneRenderer(scene, device=local, fb_import=blob[0])
    r.ctx.set_tile_ownership(world, rank)
    dist.barrier()
    torch.cuda.synchronize()
    mv = w.frames(0, frames + 8)
    one = [r]
    capi.render_frames(one, 8, mv[:8])
    dist.barrier()
    torch.cuda.synchronize()
    best = 1e30
    for rep in range(2):
        dist.barrier()
        t0 = time.perf_counter()
        # srb_render_frames on ONE context: frame f + 1 is recorded while frame f runs, and is submitted when f is complete;
        # rank 0's frame is complete when every rank's tiles of that frame are in its framebuffer
        capi.render_frames(one, frames, mv[8:])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
    us = best / frames * 1e6
    ku = kernel_times(r, w, 4)
    per_rank = [None] * world
    dist.all_gather_object(per_rank, {k: round(v, 1) for k, v in ku.items()})
    dist.barrier()
    r.close()
    return {"config": w.config(), "n_gpus": world, "us_per_frame": us, "frames_per_s": 1e6 / us,
            "timing": "host wall clock around 64 frames through srb_render_frames, max over ranks, best of 2 (every frame ends "
                      "with rank 0's device-side wait for all ranks' arrival stamps; no host barrier inside)",
            "kernel_us_per_rank": per_rank}


if __name__ == "__main__":
    main()
