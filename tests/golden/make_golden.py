"""Generates the golden fixtures under tests/golden/ from the REFERENCE ITSELF (oracle/_ref/libsrref_parity.so: the
unmodified /root/reference sources compiled in place, -ffp-contract=off, one thread = canonical order).

Run in the build container (needs /root/reference to have been built by `make -C oracle`):
    python tests/golden/make_golden.py
Each fixture is self-contained: the scene's input arrays, the RCPPS (and, for the lit scene, RSQRTPS) table of the CPU
that produced it, and the
reference's outputs (per-tile counts, ordered tile-relative triangle records, pre-depth coverage masks, ordered
fragment streams, depth tiles, colour tiles).  The reference ships no golden vectors of its own."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.refharness import RefRenderer, harvest_rcp_table, harvest_rsqrt_table  # noqa: E402
from softrast_b200 import scenes  # noqa: E402


def scene_to_arrays(sc):
    out = {
        "width": sc.width,
        "height": sc.height,
        "clear_color": sc.clear_color,
        "n_draws": len(sc.draws),
        "n_textures": len(sc.textures),
    }
    for i, d in enumerate(sc.draws):
        out[f"d{i}_vertices"] = d.vertices
        out[f"d{i}_indices"] = d.indices
        out[f"d{i}_mvp"] = d.mvp
        out[f"d{i}_meta"] = np.array([d.shader, d.texture, d.uv_offset], dtype=np.int64)
    if sc.sponza is not None:
        out["sponza"] = sc.sponza
    for i, t in enumerate(sc.textures):
        out[f"t{i}_texels"] = t.texels
        out[f"t{i}_mip_offsets"] = t.mip_offsets
        out[f"t{i}_meta"] = np.array([t.num_mips, t.width_log2, t.height_log2], dtype=np.int64)
    return out


def make(name, sc):
    r = RefRenderer(sc.width, sc.height, 1, "parity")
    r.load_scene(sc)
    r.render()
    colour, depth = r.read_tiles()
    counts = r.tile_counts()
    out = scene_to_arrays(sc)
    out["rcp_table"] = harvest_rcp_table(11)
    if sc.sponza is not None:
        out["rsqrt_table"] = harvest_rsqrt_table(10)  # the Sponza shader also replays RSQRTPS
    out["ref_counts"] = counts
    out["ref_colour"] = colour
    out["ref_depth_bits"] = depth.view(np.uint32)
    tris, cov, frags = [], [], []
    for t in range(r.num_tiles):
        n = int(counts[t])
        tris.append(r.tile_tris(t, n).copy() if n else np.zeros(0, dtype=r.tile_tris(0, 0).dtype))
        cov.append(r.tile_coverage(t, n).copy() if n else np.zeros((0, 64), dtype=np.uint64))
        frags.append(r.tile_fragments(t)[0].copy())
    out["ref_tris"] = np.concatenate(tris).view(np.uint8)
    out["ref_coverage"] = np.concatenate(cov)
    out["ref_frag_counts"] = np.array([f.size for f in frags], dtype=np.uint64)
    out["ref_frags"] = np.concatenate(frags)
    # second frame without a clear (depth test against the first frame)
    r.render(clear=False)
    c2, d2 = r.read_tiles()
    out["ref_colour_noclear"] = c2
    out["ref_depth_bits_noclear"] = d2.view(np.uint32)
    r.close()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, os.path.getsize(path) // 1024, "KiB", "refs", int(counts.sum()), "frags", int(out["ref_frags"].size))


if __name__ == "__main__":
    make("parity_160x120_s31", scenes.parity_scene(160, 120, 31, n_small=90, n_big=10))
    make("parity_200x136_s32", scenes.parity_scene(200, 136, 32, n_small=60, n_big=14))
    make("cubes_192x128", scenes.cube_grid(192, 128, 6, 6, draws=3, tex_size=64))
    make("lit_168x104_s33", scenes.parity_scene(168, 104, 33, n_small=70, n_big=10, lit=True))
