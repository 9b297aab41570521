"""Throughput sweep on the GPU box: frames/s of one scene vs. frames in flight, plus host-side submit cost per frame.
usage: python profiles/sweep.py [hall|rand|cubes] [in-flight list, e.g. 1,2,4,8] [frames]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from softrast_b200 import capi, scenes

name = sys.argv[1] if len(sys.argv) > 1 else "hall"
flights = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4,8").split(",")]
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 256
sc = {"hall": scenes.hall_scene, "rand": scenes.random_tris, "cubes": scenes.cube_grid,
      "hall_lit": lambda: scenes.hall_scene(lit=True),  # every draw with the Sponza shader (16 point lights)
      "hallmini": lambda: scenes.hall_scene(256, 128, detail=0.05)}[name]()  # hallmini: 25 tiny draws = host submit cost
mvps = None
if name in ("hall", "hall_lit"):
    mvps = scenes.hall_camera_path(sc, 1024)[:frames]
rs = []
for n in flights:
    while len(rs) < n:
        rs.append(capi.SceneRenderer(sc, resident=True, share=rs[0] if rs and not os.environ.get("SWEEP_NO_SHARE") else None))
    use = rs[:n]
    capi.render_frames(use, min(frames, 32), None if mvps is None else mvps[:min(frames, 32)])
    capi.timer_mark(use, 0)
    t0 = time.perf_counter()
    capi.render_frames(use, frames, mvps)
    t1 = time.perf_counter()
    capi.timer_mark(use, 1)
    ms = capi.timer_elapsed_ms(use, 0, 1)
    print(f"{name} in_flight={n:2d} frames={frames} device {ms / frames * 1e3:8.1f} us/frame = {frames / ms * 1e3:9.0f} frames/s"
          f"   host wall {(t1 - t0) / frames * 1e6:8.1f} us/frame", flush=True)
r0 = rs[0]
r0.ctx.set_timing(True)
acc = {}
for f in range(8):
    r0.render(mvps=None if mvps is None else mvps[f])
    for k, v in r0.ctx.kernel_times().items():
        acc[k] = acc.get(k, 0.0) + v / 8
print(name, "kernel us:", {k: round(v, 1) for k, v in acc.items()}, "sum", round(sum(acc.values()), 1))
print(name, r0.ctx.counters())
for r in rs:
    r.close()
if os.environ.get("SWEEP_BLIT"):  # the viewer's loop: EndFrame (synchronous), then Blit of that frame beside the next one
    import ctypes as C
    g = capi.SceneRenderer(sc)
    nb = sc.width * sc.height * 4
    pin = capi.host_alloc(2 * nb)
    n = int(os.environ["SWEEP_BLIT"])
    for phase in ("warm", "timed"):
        t0 = time.perf_counter()
        for f in range(n):
            g.render(mvps=None if mvps is None else mvps[f % len(mvps)])
            capi.lib.srb_blit_linear(g.ctx.h, g.fb.handle, C.c_void_p(pin + (f & 1) * nb), None, None)
        g.ctx.Sync()
        dt = (time.perf_counter() - t0) / n
    print(f"{name} frame + Blit loop, one context: {dt * 1e6:.1f} us/frame = {1.0 / dt:.0f} frames/s (host wall)", flush=True)
    capi.host_free(pin)
    g.close()
if os.environ.get("SWEEP_REF"):  # the reference's own renderer (all host threads) on the same scene, for the ratio
    from oracle import refharness as rh
    n = int(os.environ["SWEEP_REF"])
    ref = rh.RefRenderer(sc.width, sc.height, 0, "fast")
    ref.load_scene(sc)
    ref.render_frames(2, None if mvps is None else mvps[:2])
    t0 = time.perf_counter()
    ref.render_frames(n, None if mvps is None else mvps[:n])
    dt = (time.perf_counter() - t0) / n
    print(f"{name} reference ({ref.threads} host threads): {dt * 1e3:.2f} ms/frame = {1.0 / dt:.1f} frames/s", flush=True)
    ref.close()
