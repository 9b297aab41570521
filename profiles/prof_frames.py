"""Renders a few frames of one workload through the C ABI on one context — the command ncu wraps.
usage: python profiles/prof_frames.py [hall|rand|cubes] [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from softrast_b200 import scenes
from softrast_b200.capi import SceneRenderer

name = sys.argv[1] if len(sys.argv) > 1 else "hall"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sc = {"hall": scenes.hall_scene, "rand": scenes.random_tris, "cubes": scenes.cube_grid}[name]()
g = SceneRenderer(sc)
for _ in range(frames):
    g.render()
print(name, g.ctx.counters())
g.close()
