"""Renders one frame and blits it a few times — the command ncu wraps to time the de-tile kernel (the one purely
streaming kernel of the path).  usage: python profiles/prof_blit.py [width height]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from softrast_b200 import scenes
from softrast_b200.capi import SceneRenderer

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
g = SceneRenderer(scenes.hall_scene(w, h, detail=0.2))
for _ in range(6):
    g.render()
    g.blit_linear()
g.close()
