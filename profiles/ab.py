"""A/B measurement of one build / one set of SRB_* tuning knobs on the GPU box (the knobs are read at context creation,
so every configuration is its own process): single-frame latency, throughput with frames in flight, per-kernel times
with one frame in flight.  usage: python profiles/ab.py [hall|hall_lit|rand|cubes|cubes100|hall4k] [frames] [in_flight]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from softrast_b200 import capi, scenes

name = sys.argv[1] if len(sys.argv) > 1 else "hall"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 256
flight = int(sys.argv[3]) if len(sys.argv) > 3 else 12
sc = {"hall": scenes.hall_scene, "rand": scenes.random_tris, "cubes": scenes.cube_grid,
      "cubes100": lambda: scenes.cube_grid(draws=100), "hall_lit": lambda: scenes.hall_scene(lit=True),
      "hall4k": lambda: scenes.hall_scene(3840, 2160)}[name]()
mvps = scenes.hall_camera_path(sc, 1024)[:frames] if name.startswith("hall") else None
rs = [capi.SceneRenderer(sc, resident=True)]
while len(rs) < flight:
    rs.append(capi.SceneRenderer(sc, resident=True, share=rs[0]))
out = {"scene": name, "knobs": {k: v for k, v in os.environ.items() if k.startswith("SRB_")}}
for n in (flight, 1):
    use = rs[:n]
    w = min(frames, 32)
    capi.render_frames(use, w, None if mvps is None else mvps[:w])
    best = 1e9
    for rep in range(3):
        use[0].ctx.flush_l2()
        capi.timer_mark(use, 0)
        capi.render_frames(use, frames, mvps)
        capi.timer_mark(use, 1)
        best = min(best, capi.timer_elapsed_ms(use, 0, 1))
    out[f"us_per_frame_{n}_in_flight"] = round(best / frames * 1e3, 2)
r0 = rs[0]
r0.ctx.set_frames_in_flight_hint(1)
r0.ctx.set_timing(True)
acc = {}
K = 16
for f in range(K):
    r0.render(mvps=None if mvps is None else mvps[f * 7 % len(mvps)])
    for k, v in r0.ctx.kernel_times().items():
        acc[k] = acc.get(k, 0.0) + v / K
out["kernel_us"] = {k: round(v, 1) for k, v in acc.items() if k != "detile"}
out["kernel_sum"] = round(sum(out["kernel_us"].values()), 1)
out["counters"] = r0.ctx.counters()
import json
print(json.dumps(out), flush=True)
for r in rs:
    r.close()
