"""e2e A/B: frames/s with the colour read-back per frame (what bench.py's e2e leg times), for the SRB_* knobs in the environment.
usage: python profiles/e2e_ab.py [frames] [in_flight]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from softrast_b200 import capi, scenes
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 512
fl = int(sys.argv[2]) if len(sys.argv) > 2 else 12
sc = scenes.hall_scene()
mv = scenes.hall_camera_path(sc, 1024)[:frames]
rs = [capi.SceneRenderer(sc, resident=False)]
while len(rs) < fl:
    rs.append(capi.SceneRenderer(sc, resident=False, share=rs[0]))
nb = rs[0].fb.num_tiles * 16384
pin = capi.host_alloc(frames * nb)
capi.render_frames(rs, 64, mv[:64], pin, nb)
best = 1e9
for rep in range(3):
    capi.timer_mark(rs, 0)
    capi.render_frames(rs, frames, mv, pin, nb)
    capi.timer_mark(rs, 1)
    best = min(best, capi.timer_elapsed_ms(rs, 0, 1))
print(json.dumps({"knobs": {k: v for k, v in os.environ.items() if k.startswith("SRB_")}, "in_flight": fl,
                  "e2e_frames_per_s": round(frames / best * 1e3, 1), "d2h_gbs": round(nb * frames / best / 1e6, 2)}))
for r in rs:
    r.close()
capi.host_free(pin)
