"""The drop-in path of an unmodified reference application: SRB_FLAG_UPLOAD_ALWAYS contexts whose draws bind PAGEABLE host
arrays (numpy), every array copied into its device mirror at its first use in each frame.  Frames/s with 1 and 4 contexts.
usage: python profiles/pageable_ab.py [frames]
Measured: one cudaMemcpyAsync per array (what the library does) 1 510 - 1 790 frames/s = 11.6 - 13.7 GB/s of pageable memory;
ONE cudaMemcpyBatchAsync per frame with cudaMemcpySrcAccessOrderDuringApiCall (tried, not kept): 88 - 96 frames/s."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from softrast_b200 import capi, scenes
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 128
sc = scenes.hall_scene()
mv = scenes.hall_camera_path(sc, 1024)[:frames]
geo = sum(d.vertices.nbytes + d.indices.nbytes for d in sc.draws)
out = {"knobs": {k: v for k, v in os.environ.items() if k.startswith("SRB_")}, "geometry_mb_per_frame": round(geo / 1e6, 2)}
for n in (1, 4):
    rs = [capi.SceneRenderer(sc, resident=False, flags=capi.FLAG_UPLOAD_ALWAYS) for _ in range(n)]
    capi.render_frames(rs, 16, mv[:16])
    best = 1e9
    for rep in range(3):
        capi.timer_mark(rs, 0)
        capi.render_frames(rs, frames, mv)
        capi.timer_mark(rs, 1)
        best = min(best, capi.timer_elapsed_ms(rs, 0, 1))
    out[f"frames_per_s_{n}_contexts"] = round(frames / best * 1e3, 1)
    out[f"h2d_gbs_{n}_contexts"] = round(geo * frames / best / 1e6, 2)
    for r in rs:
        r.close()
print(json.dumps(out))
