"""Randomised parity sweep (run on a GPU box; not collected by pytest): small scenes of adversarial triangles —
huge coordinates, vertices on / behind the camera plane, zero-area and sub-pixel triangles, extreme UVs — rendered by
the CUDA path and by the reference itself (oracle/_ref), compared bit for bit.

    python tests/fuzz_parity.py [scenes_per_category] [first_seed]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.refharness import RefRenderer  # noqa: E402
from softrast_b200 import scenes  # noqa: E402
from softrast_b200.capi import SceneRenderer  # noqa: E402
from softrast_b200.scenes import Draw, Scene, build_tiled_texture, procedural_rgba  # noqa: E402

W, H = 200, 136


def base_tris(rng, n, spread=6.0, size=1.0, zlo=1.0, zhi=20.0):
    c = np.stack([rng.uniform(-spread, spread, n), rng.uniform(-spread * 0.7, spread * 0.7, n), rng.uniform(zlo, zhi, n)], 1)
    v = np.zeros((n, 3, 8), dtype=np.float32)
    v[:, :, 0:3] = c[:, None, :] + rng.normal(0, size, (n, 3, 3))
    nr = rng.normal(0, 1, (n, 3, 3))
    v[:, :, 3:6] = nr / np.linalg.norm(nr, axis=-1, keepdims=True)
    v[:, :, 6:8] = rng.uniform(-4, 4, (n, 3, 2))
    return v


def category(name, rng):
    n = 96
    if name == "plain":
        v = base_tris(rng, n)
    elif name == "huge":  # coordinates up to 1e6 .. 1e18: edge arithmetic wraps, snapping saturates
        v = base_tris(rng, n)
        k = rng.integers(0, n, n // 2)
        v[k, rng.integers(0, 3, k.size), rng.integers(0, 3, k.size)] *= np.float32(10.0) ** rng.integers(3, 18, k.size).astype(np.float32)
    elif name == "camera_plane":  # vertices at z ~ 0 and behind the camera (w <= 0)
        v = base_tris(rng, n, zlo=-3.0, zhi=3.0, size=2.0)
        k = rng.integers(0, n, n // 3)
        v[k, rng.integers(0, 3, k.size), 2] = rng.choice(np.array([0.0, 1e-6, -1e-6, 1e-20, 0.1], dtype=np.float32), k.size)
    elif name == "tiny":  # sub-pixel and zero-area triangles, repeated vertices
        v = base_tris(rng, n, size=0.01)
        k = rng.integers(0, n, n // 3)
        v[k, 1] = v[k, 0]
        k = rng.integers(0, n, n // 3)
        v[k, 2, 0:3] = (v[k, 0, 0:3] + v[k, 1, 0:3]) * np.float32(0.5)
    elif name == "uv":  # extreme / denormal / negative-zero texture coordinates
        v = base_tris(rng, n)
        v[:, :, 6:8] = rng.choice(np.array([0.0, -0.0, 1.0, -1.0, 0.5, 1e-30, -1e-30, 1e6, -1e6, 127.99, 1e-40, 3.4e38], dtype=np.float32), (n, 3, 2))
    elif name == "flat":  # big screen-aligned triangles at constant depth: exact depth ties everywhere
        v = base_tris(rng, n // 4, size=4.0)
        v[:, :, 2] = np.float32(5.0)
        v = np.concatenate([v, v[::-1].copy(), v.copy()])
    elif name == "nan_inf":  # non-finite positions and attributes (the reference's SIMD compares treat NaN as "false")
        v = base_tris(rng, n)
        k = rng.integers(0, n, n // 4)
        v[k, rng.integers(0, 3, k.size), rng.integers(0, 8, k.size)] = rng.choice(
            np.array([np.nan, np.inf, -np.inf, 3.4e38, -3.4e38], dtype=np.float32), k.size)
    else:
        raise ValueError(name)
    return v.reshape(-1, 8)


CATEGORIES = ["plain", "huge", "camera_plane", "tiny", "uv", "flat", "nan_inf"]


def make_scene(name, seed):
    rng = np.random.default_rng(seed)
    proj = scenes.reverse_z_projection(W, H)
    view = scenes.look_at_lh((0.2, 0.3, -0.5), (0.0, 0.0, 6.0))
    mvp = scenes.to_column_major(proj @ view)
    sc = Scene(f"fuzz_{name}_{seed}", W, H, clear_color=0x40)
    sc.textures.append(build_tiled_texture(procedural_rgba(64, seed)))
    sc.textures.append(build_tiled_texture(procedural_rgba(32, seed + 1), calc_mips=False))
    v = category(name, rng)
    half = (v.shape[0] // 6) * 3
    sc.draws.append(Draw(v[:half], np.arange(half, dtype=np.uint32), mvp, scenes.SHADER_UNLIT_DIFFUSE, 0))
    sc.draws.append(Draw(v[half:], np.arange(v.shape[0] - half, dtype=np.uint32), mvp,
                         [scenes.SHADER_UNLIT_DIFFUSE, scenes.SHADER_VISUALIZE_NORMALS, scenes.SHADER_VISUALIZE_UVS][seed % 3], 1))
    if seed % 4 == 1:  # every fourth scene through the lit shader (RSQRTPS / RCPPS of whatever the interpolants give)
        sc.draws[0].shader = scenes.SHADER_SPONZA
        sc.sponza = scenes.sponza_constants(seed)
    return sc


def compare(sc):
    r = RefRenderer(sc.width, sc.height, 1, "parity")
    g = SceneRenderer(sc)
    try:
        r.load_scene(sc)
        r.render()
        g.render()
        cr, dr = r.read_tiles()
        cg, dg = g.read_tiles()
        counts_ok = np.array_equal(r.tile_counts(), g.ctx.tile_counts(g.fb.num_tiles))
        depth_bad = int((dr.view(np.uint32) != dg.view(np.uint32)).sum())
        colour_bad = int((cr != cg).sum())
        return counts_ok, depth_bad, colour_bad
    finally:
        r.close()
        g.close()


def main():
    per = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    bad = 0
    for name in CATEGORIES:
        stats = []
        for s in range(per):
            sc = make_scene(name, seed0 + s)
            ok, db, cb = compare(sc)
            stats.append((ok, db, cb))
            if not ok or db or cb:
                bad += 1
                print(f"MISMATCH {sc.name}: counts_ok={ok} depth_px={db} colour_px={cb}", flush=True)
        print(f"{name:14s} scenes {per}  mismatching {sum(1 for o, d, c in stats if not o or d or c)}", flush=True)
    print("TOTAL mismatching scenes:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
