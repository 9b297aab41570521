"""Device-side texture builder (srb_texture_create_rgba8): wall time per texture next to the host builder, for the sizes of
the hall scene's texture set (10 x 1024^2, 10 x 512^2, 5 x 256^2) and one 4096^2 image.
    python profiles/prof_texbuild.py            # timings
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        python profiles/prof_texbuild.py ncu    # per-kernel durations / DRAM bytes (one 4096^2 build)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from softrast_b200 import capi

rng = np.random.default_rng(0)
ctx = capi.RenderContext(0)
if len(sys.argv) > 1 and sys.argv[1] == "ncu":
    img = rng.integers(0, 256, (4096, 4096, 4)).astype(np.uint8)
    ctx.create_texture_rgba8(img, capi.MIPS_STB)
    ctx.close()
    sys.exit(0)
ctx.create_texture_rgba8(rng.integers(0, 256, (64, 64, 4)).astype(np.uint8), capi.MIPS_STB)  # module load
total_dev = total_host = 0.0
for n, count in ((256, 5), (512, 10), (1024, 10), (4096, 1)):
    img = rng.integers(0, 256, (n, n, 4)).astype(np.uint8)
    t0 = time.perf_counter(); h = ctx.create_texture_rgba8(img, capi.MIPS_STB); first = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(3):
        h = ctx.create_texture_rgba8(img, capi.MIPS_STB)
    dev = (time.perf_counter() - t0) / 3
    t0 = time.perf_counter(); host = capi.build_texture(img, capi.MIPS_STB); hst = time.perf_counter() - t0
    same = np.array_equal(ctx.read_texture(h).texels, host.texels)
    print(f"{n}x{n} + mips: device {dev * 1e3:8.2f} ms (first of this size, tables computed: {first * 1e3:7.2f} ms)   "
          f"host {hst * 1e3:8.1f} ms   identical={same}", flush=True)
    if n != 4096:
        total_dev += dev * count
        total_host += hst * count
print(f"hall scene texture set (25 textures): device {total_dev * 1e3:.1f} ms, host {total_host * 1e3:.0f} ms")
ctx.close()
