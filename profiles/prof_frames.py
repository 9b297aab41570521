"""Renders a few frames of one workload through the C ABI on one context — the command ncu wraps.
usage: python profiles/prof_frames.py [hall|hallpath|rand|cubes|hall4k] [frames] [nulltaps]
hallpath = the bench workload: frames 384 + k of the 1024-camera path (what bench.py's kernel pass times);
nulltaps (statistics build only, SRB_LIB=...): every texel tap reads texel 0 (profiles/texel_taps.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from softrast_b200 import scenes
from softrast_b200 import capi
from softrast_b200.capi import SceneRenderer

name = sys.argv[1] if len(sys.argv) > 1 else "hall"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sc = {"hall": scenes.hall_scene, "hallpath": scenes.hall_scene, "rand": scenes.random_tris, "cubes": scenes.cube_grid,
      "hall4k": lambda: scenes.hall_scene(3840, 2160)}[name]()
mvps = scenes.hall_camera_path(sc, 1024) if name == "hallpath" else None
if len(sys.argv) > 3 and sys.argv[3] == "nulltaps":
    capi.lib.srb_debug_null_taps.restype = None
    capi.lib.srb_debug_null_taps(1)
g = SceneRenderer(sc)
for k in range(frames):
    g.render(mvps=None if mvps is None else mvps[384 + k])
print(name, g.ctx.counters())
g.close()
