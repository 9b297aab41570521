"""Ad-hoc first-light script for gpurun: renders the BASELINE scenes, compares with the reference, prints timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from softrast_b200 import scenes
from softrast_b200.capi import SceneRenderer
from oracle.refharness import RefRenderer

def run(name, sc, ref_threads=0, compare=True):
    g = SceneRenderer(sc)
    g.ctx.set_timing(True)
    for _ in range(3):
        g.render()
    t0 = time.perf_counter()
    for _ in range(20):
        g.render()
    dt = (time.perf_counter() - t0) / 20
    print(name, "tris", sc.num_tris, "gpu ms/frame (host wall, sync per frame)", round(dt * 1e3, 3), g.ctx.counters(), flush=True)
    print("   kernel us", {k: round(v, 1) for k, v in g.ctx.kernel_times().items()}, flush=True)
    if compare:
        r = RefRenderer(sc.width, sc.height, 1, "parity")
        r.load_scene(sc)
        r.render()
        cr, dr = r.read_tiles()
        cg, dg = g.read_tiles()
        print("   counts equal", np.array_equal(r.tile_counts(), g.ctx.tile_counts(g.fb.num_tiles)),
              "depth bad", int((dg.view(np.uint32) != dr.view(np.uint32)).sum()),
              "colour bad", int((cg != cr).sum()),
              "max diff", int(np.abs(cg.view(np.uint8).astype(int) - cr.view(np.uint8).astype(int)).max()), flush=True)
        r.close()
    rm = RefRenderer(sc.width, sc.height, ref_threads, "fast")
    rm.load_scene(sc)
    ms = rm.render_frames(6)
    print("   reference", rm.threads, "threads ms", ms.round(2), flush=True)
    rm.close()
    g.close()

print("cpus", os.cpu_count())
run("parity", scenes.parity_scene())
run("cubes", scenes.cube_grid())
run("hall", scenes.hall_scene())
run("rand1M", scenes.random_tris())
