"""Generates tests/golden/obj/*.npz from the REFERENCE'S OWN loader (sr::Obj::Model::Load, Viewer/Obj.cpp:374-560,
compiled in place into oracle/_ref by oracle/ref_build/ref_obj.cpp).

    python tests/golden/make_golden_obj.py

Each fixture holds the input files (OBJ, MTL and image bytes), the load flags, the reference's meshes and materials, and
the `.bin` cache file the reference wrote — so tests/test_obj.py can check srb_model_load (text path and cache path)
against the reference on a box that has neither /root/reference nor oracle/_ref.  The padding bytes of texture levels
smaller than a 32x32 tile are uninitialised in the reference; they are zeroed in the stored texel blobs (not in the raw
`.bin`)."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refharness as rh  # noqa: E402
from tests import objgen  # noqa: E402
from tests.test_obj import _valid_texel_mask  # noqa: E402


def make(name, seed, flags, crlf):
    with tempfile.TemporaryDirectory() as d:
        path = objgen.write_model(d, seed=seed, crlf=crlf)
        files = {f: open(os.path.join(d, f), "rb").read() for f in sorted(os.listdir(d))}
        meshes, mats = rh.ref_load_model(path, flags)
        cache = open(path + ".bin", "rb").read()
    out = {"flags": flags, "n_files": len(files), "n_meshes": len(meshes), "n_materials": len(mats),
           "cache": np.frombuffer(cache, dtype=np.uint8)}
    for i, (fname, data) in enumerate(files.items()):
        out[f"f{i}_name"] = np.frombuffer(fname.encode(), dtype=np.uint8)
        out[f"f{i}_data"] = np.frombuffer(data, dtype=np.uint8)
    for i, m in enumerate(meshes):
        out[f"m{i}_indices"] = m["indices"]
        out[f"m{i}_vertices"] = m["vertices"]
        out[f"m{i}_material"] = m["material"]
    for i, m in enumerate(mats):
        tex = m["texels"].copy()
        if tex.size:
            tex[~_valid_texel_mask(m)] = 0
        out[f"t{i}_name"] = np.frombuffer(m["name"].encode("latin-1"), dtype=np.uint8)
        out[f"t{i}_texels"] = tex
        out[f"t{i}_mip_offsets"] = m["mip_offsets"] if tex.size else np.zeros(14, np.uint32)
        out[f"t{i}_meta"] = np.array([m["num_mips"], m["width_log2"], m["height_log2"], m["bytes_per_pixel"]] if tex.size else [0, 0, 0, 0],
                                     dtype=np.int64)
    os.makedirs(os.path.join(HERE, "obj"), exist_ok=True)
    p = os.path.join(HERE, "obj", name + ".npz")
    np.savez_compressed(p, **out)
    print(name, os.path.getsize(p) // 1024, "KiB", len(meshes), "meshes", len(mats), "materials")


if __name__ == "__main__":
    make("model_s41", 41, 0, False)
    make("model_s42_flipped_crlf", 42, 1 | 4, True)
