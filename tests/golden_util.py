"""Loads a tests/golden/*.npz fixture back into a scenes.Scene + the reference's recorded outputs."""
import glob
import os

import numpy as np

from softrast_b200 import scenes
from softrast_b200._ctypes_defs import TILE_TRI_DTYPE

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        sc = scenes.Scene(name, int(z["width"]), int(z["height"]), clear_color=int(z["clear_color"]))
        for i in range(int(z["n_textures"])):
            nm, wl, hl = (int(v) for v in z[f"t{i}_meta"])
            sc.textures.append(scenes.TiledTexture(z[f"t{i}_texels"], z[f"t{i}_mip_offsets"], nm, wl, hl))
        for i in range(int(z["n_draws"])):
            shader, tex, uvo = (int(v) for v in z[f"d{i}_meta"])
            sc.draws.append(scenes.Draw(z[f"d{i}_vertices"], z[f"d{i}_indices"], z[f"d{i}_mvp"], shader, tex, uvo))
        if "sponza" in z:
            sc.sponza = z["sponza"]
        self.scene = sc
        self.rcp = (z["rcp_table"], 11)
        self.rsqrt = (z["rsqrt_table"], 10) if "rsqrt_table" in z else None
        self.counts = z["ref_counts"]
        self.colour = z["ref_colour"]
        self.depth_bits = z["ref_depth_bits"]
        self.colour_noclear = z["ref_colour_noclear"]
        self.depth_bits_noclear = z["ref_depth_bits_noclear"]
        tris = z["ref_tris"].view(TILE_TRI_DTYPE)
        offs = np.concatenate([[0], np.cumsum(self.counts)]).astype(np.int64)
        self.tris = [tris[offs[t] : offs[t + 1]] for t in range(self.counts.size)]
        self.coverage = [z["ref_coverage"][offs[t] : offs[t + 1]] for t in range(self.counts.size)]
        fo = np.concatenate([[0], np.cumsum(z["ref_frag_counts"])]).astype(np.int64)
        self.frags = [z["ref_frags"][fo[t] : fo[t + 1]] for t in range(self.counts.size)]


def check_against_golden(g: Golden, r, fragments=True, exact_colour=True, check_colour=True):
    """r: any renderer with the RefRenderer interface, already rendered (one cleared frame)."""
    counts = r.tile_counts()
    assert np.array_equal(counts, g.counts), "per-tile counts"
    for t in range(counts.size):
        n = int(counts[t])
        if not n:
            continue
        assert r.tile_tris(t, n).tobytes() == g.tris[t].tobytes(), f"tile {t}: ordered triangle records"
        assert np.array_equal(r.tile_coverage(t, n), g.coverage[t]), f"tile {t}: coverage masks"
        if fragments:
            assert np.array_equal(r.tile_fragments(t)[0], g.frags[t]), f"tile {t}: fragment stream"
    colour, depth = r.read_tiles()
    assert np.array_equal(depth.view(np.uint32), g.depth_bits), "depth tiles"
    if not check_colour:
        return
    d = np.abs(colour.view(np.uint8).astype(np.int32) - g.colour.view(np.uint8).astype(np.int32)).max()
    assert d <= 1, f"colour differs by {d} LSB"
    if exact_colour:
        assert np.array_equal(colour, g.colour), "colour tiles"
