#!/usr/bin/env python
"""bench.py — frames/s of the sort-middle frame pipeline on the BASELINE.json workload, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[1] — the synthetic Sponza-scale "hall" scene (263 888 triangles in 25
draws, 25 Morton-tiled mip-mapped textures) at 1920x1080.  A unit is ONE FRAME: BeginFrame -> ClearFrameBuffer ->
25 x DrawIndexed -> EndFrame returns with the frame complete in the tile buffers.  A step is
`--frames-per-step` frames along the 1024-camera closed path of configs[4]; frames are independent, so with N GPUs
every rank renders its own frames (weak scaling, no data-path collective).

  value        frames/s with the scene resident in HBM: only the 25 draw descriptors (MVPs) go host->device per frame.
               Draws carry HOST pointers like the reference's DrawCall; the library mirrors those buffers on the device
               at their first use and finds the mirrors by pointer afterwards.
  e2e          frames/s through the same C-ABI calls with host buffers: per frame the draw table is uploaded and the
               finished colour tiles (tiles*16 KiB) are copied back into pinned host memory, inside the timed region.
  roofline     the dominant kernel: algorithmic bytes per launch / its mean duration,
               measured with CUDA events on the library's own stream in a second pass over the same frames.
  cpu_baseline the UNMODIFIED reference (oracle/_ref/libsrref_fast.so, its own flags, all host threads) on a bounded
               sample of the same frames — N=1 only.
`--impl reference` times only that CPU arm, per step a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
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


def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def _bind_to_gpu_numa(local: int):
    """Multi-GPU runs: pin this rank to the CPUs NVML reports as local to its GPU BEFORE any pinned host memory is
    allocated, so that the read-back buffer lives on the GPU's own NUMA node (8 ranks copying into one node's memory is
    what bounded e2e at 8 GPUs).  Returns the CPU list, or None if NVML / the topology gives nothing to bind to."""
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


def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path, all host threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle import refharness as rh
    from softrast_b200 import scenes

    scene = scenes.hall_scene(WIDTH, HEIGHT)
    mvps = scenes.hall_camera_path(scene, PATH_FRAMES)
    sample = args.ref_frames_per_step
    if rh.ref_available("fast"):
        r = rh.RefRenderer(WIDTH, HEIGHT, 0, "fast")
        kind, threads = "reference", r.threads
    else:
        from softrast_b200.capi import harvest_rcp_table

        r = rh.PortRenderer(WIDTH, HEIGHT, harvest_rcp_table(16))
        kind, threads = "port", 1
    r.load_scene(scene)
    f0 = 0
    for _ in range(args.warmup):
        r.render_frames(sample, mvps[np.arange(f0, f0 + sample) % PATH_FRAMES])
        f0 += sample
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r.render_frames(sample, mvps[np.arange(f0, f0 + sample) % PATH_FRAMES])
        f0 += sample
    dt = time.perf_counter() - t0
    fps = args.steps * sample / dt
    r.close()
    line = {
        "impl": "reference",
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "hall_1080p_camera_path (BASELINE.json configs[1] scene on the configs[4] camera path)",
                   "width": WIDTH, "height": HEIGHT, "tris_per_frame": scene.num_tris, "draws": len(scene.draws),
                   "frames_per_step": sample},
        "mtris_per_s": fps * scene.num_tris / 1e6,
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                         "sample": f"{sample} frames per step of the same camera path, {args.steps} steps"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=128)
    ap.add_argument("--in-flight", type=int, default=12, help="contexts (CUDA streams) rendering frames concurrently")
    ap.add_argument("--ref-frames-per-step", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    from softrast_b200 import capi, scenes

    scene = scenes.hall_scene(WIDTH, HEIGHT)
    mvps_all = scenes.hall_camera_path(scene, PATH_FRAMES)
    F = args.frames_per_step
    # one context per frame in flight; they share ONE device copy of the scene (srb_create_shared)
    # Draws carry HOST pointers, like the reference's DrawCall (Renderer.h:112-141): the library mirrors the buffers on
    # the device at the first DrawIndexed and finds them by pointer afterwards (--resident: explicit device buffers).
    res = bool(args.resident)
    renderers = [capi.SceneRenderer(scene, device=local, resident=res)]
    for _ in range(max(1, args.in_flight) - 1):
        renderers.append(capi.SceneRenderer(scene, device=local, resident=res, share=None if args.no_share else renderers[0]))
    colour_bytes = renderers[0].fb.num_tiles * 16384
    pinned = capi.host_alloc(F * colour_bytes)
    draw_upload_bytes = 136 * len(scene.draws)  # sizeof(DrawDev) per draw, uploaded every frame

    from softrast_b200 import sharding

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
        import copy
        import ctypes as C

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

    # what the PCIe link of this GPU delivers for the same copy (device -> pinned host, one frame's colour tiles per call)
    def measure_d2h_gbs():
        n = colour_bytes
        dev = torch.empty(n, dtype=torch.uint8, device="cuda")
        host = torch.empty(n, dtype=torch.uint8).pin_memory()
        for _ in range(3):
            host.copy_(dev, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 64
        e0.record()
        for _ in range(reps):
            host.copy_(dev, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        return n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9

    d2h_peak_gbs = measure_d2h_gbs()

    # one frame at a time (BASELINE configs[1] literally: "one frame"): the same calls on ONE context, next frame submitted
    # only after the previous one is complete
    one = renderers[:1]
    capi.render_frames(one, 16, frames_of(0)[:16])
    capi.timer_mark(one, 2)
    capi.render_frames(one, F, frames_of(1))
    capi.timer_mark(one, 3)
    single_frame_us = capi.timer_elapsed_ms(one, 2, 3) / F * 1e3

    total_frames = world * args.steps * F
    fps = total_frames / (ms_dev * 1e-3)
    fps_e2e = total_frames / (ms_e2e * 1e-3)

    # ---- roofline pass: per-kernel durations with CUDA events on the library's stream, same frames -------------
    r0 = renderers[0]
    r0.ctx.set_timing(True)
    acc, nacc = {}, 0
    mv = frames_of(args.warmup)
    for f in range(min(F, 32)):
        r0.render(mvps=mv[f])
        for k, v in r0.ctx.kernel_times().items():
            acc[k] = acc.get(k, 0.0) + v
        nacc += 1
    r0.ctx.set_timing(False)
    kernel_us = {k: v / nacc for k, v in acc.items() if k != "detile"}
    counters = r0.ctx.counters()
    winners = r0.ctx.winners(r0.fb.num_tiles)
    alg = algorithmic_bytes(scene, counters, winners)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # per-launch DRAM traffic and executed warp instructions from the committed ncu --set full captures of this workload
    ncu = {}
    npath = os.path.join(ROOT, "profiles", "ncu_counts.json")
    if os.path.exists(npath):
        ncu = json.load(open(npath))
    sm_clock_hz = (clock_info.get("sm_mhz") or 1965.0) * 1e6
    issue_peak = 148 * 4 * sm_clock_hz  # warp instructions/s: 4 schedulers per SM, one issue per clock
    kernels = {}
    for k in ("setup", "bin_fill", "raster", "shade"):  # clip + tile_scan are reported in kernel_us_per_frame
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
    roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac"], "traffic": kernels[dom]["traffic"], "peak_source": peak_src,
                "note": "the frame's working set stays in the 126 MB L2, so no kernel of this pipeline is HBM bound: the "
                        "rasteriser and the shader are instruction-issue bound (see kernels.*.issue: executed warp "
                        "instructions per launch from the committed ncu capture / measured duration, against "
                        "148 SMs x 4 schedulers x SM clock); DESIGN.md section 5",
                "issue": kernels[dom].get("issue")}
    wi_frame = sum(v.get("warp_inst", 0) for v in ncu.values())

    line = {
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "hall_1080p_camera_path (BASELINE.json configs[1] scene on the configs[4] camera path)",
                   "width": WIDTH, "height": HEIGHT, "tris_per_frame": scene.num_tris, "draws": len(scene.draws),
                   "textures": len(scene.textures), "frames_per_step": F, "frames_in_flight": len(renderers),
                   "parallelism": f"frame-parallel x{world}", "l2": "256 MiB device memset between steps (L2 flush)"},
        "mtris_per_s": fps * scene.num_tris / 1e6,
        "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": draw_upload_bytes * F,
                "d2h_bytes_per_step": colour_bytes * F, "ms_per_step": ms_e2e / args.steps,
                "d2h_gbs_per_gpu": colour_bytes * F / (ms_e2e / args.steps * 1e-3) / 1e9,
                "d2h_link_gbs": d2h_peak_gbs,
                "note": "bound by the PCIe read-back of the finished colour tiles (one link per GPU); d2h_link_gbs = "
                        "the same copy alone, back to back, measured in this run"},
        "e2e_geometry_upload": geo,
        "issue_frac_whole_frame": (wi_frame * fps / world / issue_peak) if wi_frame else None,
        "single_frame": {"us_per_frame": single_frame_us, "frames_per_s": 1e6 / single_frame_us,
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

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import refharness as rh

        if rh.ref_available("fast"):
            ref = rh.RefRenderer(WIDTH, HEIGHT, 0, "fast")
            kind, cores = "reference", ref.threads
        else:
            ref = rh.PortRenderer(WIDTH, HEIGHT, capi.harvest_rcp_table(16))
            kind, cores = "port", 1
        ref.load_scene(scene)
        est = float(np.median(ref.render_frames(6, mvps_all[:6])[2:]))  # ms per frame
        n = int(min(1024, max(32, 12_000.0 / max(est, 1e-3))))
        t0 = time.perf_counter()
        ref.render_frames(n, mvps_all[np.arange(n) % PATH_FRAMES])
        dt = time.perf_counter() - t0
        ref.close()
        line["cpu_baseline"] = {"value": n / dt, "unit": "frames/s", "cores": cores, "kind": kind,
                                "sample": f"{n} consecutive frames of the same camera path ({dt:.1f} s)",
                                "host_cpus": os.cpu_count()}
    for r in renderers:
        r.close()
    capi.host_free(pinned)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
