"""Counters of the statistics build (make -C softrast_b200/csrc STATS=1) for one frame of a scene.
usage: SRB_LIB=softrast_b200/lib/libsoftrast_b200_stats.so python profiles/stats.py [hall|rand|cubes] [SRB_* knobs in env]"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SRB_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "softrast_b200", "lib", "libsoftrast_b200_stats.so"))
import numpy as np
from softrast_b200 import capi, scenes

name = sys.argv[1] if len(sys.argv) > 1 else "hall"
sc = {"hall": scenes.hall_scene, "rand": scenes.random_tris, "cubes": scenes.cube_grid,
      "hall_lit": lambda: scenes.hall_scene(lit=True), "hall4k": lambda: scenes.hall_scene(3840, 2160), "cubes100": lambda: scenes.cube_grid(draws=100)}[name]()
mvps = scenes.hall_camera_path(sc, 1024) if name.startswith("hall") else None
g = capi.SceneRenderer(sc)
st = (C.c_uint64 * 16)()
capi.lib.srb_debug_stats.restype = None
capi.lib.srb_debug_stats.argtypes = [C.c_void_p, C.c_int]
g.render(mvps=None if mvps is None else mvps[0])
capi.lib.srb_debug_stats(st, 1)
g.render(mvps=None if mvps is None else mvps[100])
capi.lib.srb_debug_stats(st, 1)
v = [int(x) for x in st]
print(json.dumps({"scene": name, "knobs": {k: x for k, x in os.environ.items() if k.startswith("SRB_") and k != "SRB_LIB"},
                  "list_tests": v[0], "candidates_ref_coarse": v[1], "dropped_by_block_reject": v[2],
                  "block_visits": v[3], "visits_no_pixel_inside": v[4], "visits_all_64_pixels_inside": v[9], "visits_block_fully_covered_before": v[10], "visits_hiz_rejectable": v[11], "texture_sample_warps": v[5],
                  "texture_sample_warps_128bit": v[6], "list_walks": v[7], "sum_longest_list": v[8],
                  "counters": g.ctx.counters()}))
g.close()
