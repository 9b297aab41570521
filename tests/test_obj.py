"""Scene ingestion: srb_model_* / srb_image_load_rgba8 against the reference's own sr::Obj::Model::Load
(Viewer/Obj.cpp:374-560) and stbi_load (Texture.cpp:107), compiled in place into oracle/_ref (ref_obj.cpp).  CPU only."""
import os

import numpy as np
import pytest

from oracle import refharness as rh
from softrast_b200 import capi

from . import objgen

needs_ref = pytest.mark.skipif(not rh.ref_available(), reason="oracle/_ref not built")


def _same_models(ours, ref):
    meshes, mats = ref
    assert len(ours.meshes) == len(meshes)
    assert len(ours.materials) == len(mats)
    for a, b in zip(ours.meshes, meshes):
        assert a["index_stride"] == b["index_stride"]
        assert a["material"] == b["material"]
        assert a["indices"].dtype == b["indices"].dtype and np.array_equal(a["indices"], b["indices"])
        assert np.array_equal(a["vertices"].view(np.uint32), b["vertices"].view(np.uint32))
    for a, b in zip(ours.materials, mats):
        assert a["name"] == b["name"]
        for k in ("num_mips", "width_log2", "height_log2", "bytes_per_pixel"):
            if b["texels"].size:  # (the reference leaves these fields untouched when the image did not load)
                assert a[k] == b[k], k
        assert a["texels"].size == b["texels"].size
        if b["texels"].size:
            assert np.array_equal(a["mip_offsets"][: a["num_mips"]], b["mip_offsets"][: b["num_mips"]])
            # levels smaller than the 32x32 storage tile are padded; the reference never initialises the padding
            # (Texture.cpp:177 Resize of a POD array), so only the texels that exist are compared
            valid = _valid_texel_mask(a)
            assert valid.sum() >= 4 << (a["width_log2"] + a["height_log2"])  # (sanity: at least all of level 0)
            assert np.array_equal(a["texels"][valid], b["texels"][valid])


def _valid_texel_mask(mat) -> np.ndarray:
    """Byte mask of the texels that exist in a tiled / Morton / mip blob (Texture.cpp:73-101,159-175)."""
    mask = np.zeros(mat["texels"].size, dtype=bool)
    W, H = 1 << mat["width_log2"], 1 << mat["height_log2"]

    def spread(v):
        out = np.zeros_like(v)
        for b in range(5):
            out |= ((v >> b) & 1) << (2 * b)
        return out

    for k in range(mat["num_mips"]):
        w, h = max(1, W >> k), max(1, H >> k)
        y, x = np.mgrid[0:h, 0:w]
        tiles_x = (w + 31) // 32
        texel = ((y >> 5) * tiles_x + (x >> 5)) * 1024 + (spread(x & 31) | (spread(y & 31) << 1))
        offs = int(mat["mip_offsets"][k]) + 4 * texel.reshape(-1)
        for c in range(4):
            mask[offs + c] = True
    return mask


def _parse_bin(blob: bytes):
    """An independent reader of the `.bin` cache (kt::Serialize: Obj.cpp:15-39, Texture.cpp:18-26, Serialization.inl)."""
    pos = 0

    def u32():
        nonlocal pos
        v = int.from_bytes(blob[pos:pos + 4], "little")
        pos += 4
        return v

    def take(n):
        nonlocal pos
        b = blob[pos:pos + n]
        assert len(b) == n
        pos += n
        return b

    meshes, mats = [], []
    for _ in range(u32()):
        itype = u32()
        idx = take(u32())
        n_idx = u32()
        verts = np.frombuffer(take(u32() * 32), dtype=np.float32).reshape(-1, 8)
        mat = u32()
        stride = 4 if itype else 2
        meshes.append({"indices": np.frombuffer(idx, dtype=np.uint32 if itype else np.uint16)[:n_idx], "vertices": verts,
                       "material": mat, "index_stride": stride})
    for _ in range(u32()):
        texels = np.frombuffer(take(u32()), dtype=np.uint8)
        wl, hl, bpp = u32(), u32(), u32()
        offs = np.frombuffer(take(56), dtype=np.uint32)
        nm = u32()
        name = take(u32()).decode("latin-1")
        mats.append({"name": name, "texels": texels, "mip_offsets": offs, "num_mips": nm, "width_log2": wl, "height_log2": hl,
                     "bytes_per_pixel": bpp})
    assert pos == len(blob)
    return meshes, mats


# ---- images ---------------------------------------------------------------------------------------------------------

def _png_cases(rng):
    w, h = 37, 29  # odd sizes: partial bytes at low bit depths, uneven Adam7 passes
    for interlace in (False, True):
        for depth in (1, 2, 4, 8, 16):
            yield f"grey{depth}", dict(samples=rng.integers(0, 1 << depth, (h, w, 1)), colour_type=0, depth=depth, interlace=interlace)
            key = int(rng.integers(0, 1 << depth))
            yield f"grey{depth}_key", dict(samples=rng.integers(0, 1 << depth, (h, w, 1)), colour_type=0, depth=depth,
                                           trns=bytes([key >> 8, key & 255]), interlace=interlace)
        for depth in (8, 16):
            s = rng.integers(0, 1 << depth, (h, w, 3))
            yield f"rgb{depth}", dict(samples=s, colour_type=2, depth=depth, interlace=interlace)
            k = s[3, 5]
            s2 = s.copy()
            s2[::3, ::2] = k
            yield f"rgb{depth}_key", dict(samples=s2, colour_type=2, depth=depth, interlace=interlace,
                                          trns=b"".join(bytes([int(v) >> 8, int(v) & 255]) for v in k))
            yield f"ga{depth}", dict(samples=rng.integers(0, 1 << depth, (h, w, 2)), colour_type=4, depth=depth, interlace=interlace)
            yield f"rgba{depth}", dict(samples=rng.integers(0, 1 << depth, (h, w, 4)), colour_type=6, depth=depth, interlace=interlace)
        for depth in (1, 2, 4, 8):
            n = 1 << depth
            pal = rng.integers(0, 256, (n, 3))
            yield f"pal{depth}", dict(samples=rng.integers(0, n, (h, w, 1)), colour_type=3, depth=depth, palette=pal, interlace=interlace)
            yield f"pal{depth}_trns", dict(samples=rng.integers(0, n, (h, w, 1)), colour_type=3, depth=depth, palette=pal,
                                           trns=bytes(rng.integers(0, 256, max(1, n // 2)).tolist()), interlace=interlace)


@needs_ref
def test_png_decoder_matches_stb_image(tmp_path):
    rng = np.random.default_rng(11)
    n = 0
    for name, kw in _png_cases(rng):
        p = str(tmp_path / f"{name}_{int(kw['interlace'])}.png")
        objgen.write_png(p, idat_split=97 if n % 3 == 0 else 0, **kw)
        ref = rh.ref_load_image(p)
        assert ref is not None, name
        ours = capi.load_image_rgba8(p)
        assert ours.shape == ref.shape and np.array_equal(ours, ref), name
        n += 1
    assert n == 52
    # a single-filter image for each filter type (the mixed ones above cycle through all five)
    for f in range(5):
        p = str(tmp_path / f"filter{f}.png")
        objgen.write_png(p, rng.integers(0, 256, (16, 23, 4)), 6, 8, filters=(f,))
        assert np.array_equal(capi.load_image_rgba8(p), rh.ref_load_image(p))


@needs_ref
def test_tga_decoder_matches_stb_image(tmp_path):
    rng = np.random.default_rng(12)
    img = rng.integers(0, 256, (21, 34, 4)).astype(np.uint8)
    img[5:12, 3:30] = img[5, 3]  # runs for the RLE packets
    n = 0
    for bits in (32, 24, 8):
        for rle in (False, True):
            for top in (False, True):
                p = str(tmp_path / f"t{bits}_{int(rle)}_{int(top)}.tga")
                objgen.write_tga(p, img, bits=bits, rle=rle, top_down=top, id_bytes=b"id" if n % 2 else b"")
                ref = rh.ref_load_image(p)
                assert ref is not None
                assert np.array_equal(capi.load_image_rgba8(p), ref), (bits, rle, top)
                n += 1


def test_image_errors(tmp_path):
    with pytest.raises(capi.SrbError):
        capi.load_image_rgba8(str(tmp_path / "missing.png"))
    p = tmp_path / "junk.png"
    p.write_bytes(b"\x89PNG\r\n\x1a\n" + b"\0" * 40)
    with pytest.raises(capi.SrbError):
        capi.load_image_rgba8(str(p))
    p = tmp_path / "photo.jpg"
    p.write_bytes(b"\xff\xd8\xff\xe0" + b"\0" * 64)
    with pytest.raises(capi.SrbError, match="decoder"):
        capi.load_image_rgba8(str(p))


# ---- OBJ / MTL / .bin ------------------------------------------------------------------------------------------------

@needs_ref
@pytest.mark.parametrize("flags,crlf", [(0, False), (capi.OBJ_FLIP_WINDING | capi.OBJ_FLIP_UVS, True), (capi.OBJ_GEN_NORMALS, False)])
def test_obj_loader_matches_reference(tmp_path, flags, crlf):
    ours_dir, ref_dir = str(tmp_path / "ours"), str(tmp_path / "ref")
    po = objgen.write_model(ours_dir, seed=21, crlf=crlf)
    pr = objgen.write_model(ref_dir, seed=21, crlf=crlf)
    ours = capi.Model(po, flags)
    ref = rh.ref_load_model(pr, flags)
    assert ref is not None and not ours.from_cache
    assert len(ours.meshes) == 5 and len(ours.materials) == 5
    assert [m["name"] for m in ours.materials] == ["bricks", "plain", "tiles", "grey", "broken"]
    assert [m["material"] for m in ours.meshes] == [0, 2, 3, 4, 1]
    assert ours.materials[0]["num_mips"] == 7 and ours.materials[4]["texels"].size == 0
    assert "does_not_exist.png" in ours.notes
    _same_models(ours, ref)
    # both wrote their cache: identical bytes (kt::Serialize format, Obj.cpp:15-39)
    # (compared through an independent parser of the format: the reference's texture padding bytes are uninitialised)
    bo, br = open(po + ".bin", "rb").read(), open(pr + ".bin", "rb").read()
    assert len(bo) == len(br) and len(bo) > 1000
    _same_models(ours, _parse_bin(bo))
    _same_models(ours, _parse_bin(br))
    ours.close()
    # each side loads the OTHER side's cache (the reference's first, written by the reference itself)
    os.replace(po + ".bin", str(tmp_path / "ours.bin"))
    os.replace(pr + ".bin", po + ".bin")
    os.replace(str(tmp_path / "ours.bin"), pr + ".bin")
    again = capi.Model(po, 0)
    assert again.from_cache
    ref_again = rh.ref_load_model(pr, 0)
    _same_models(again, ref)
    _same_models(again, ref_again)
    again.close()


@needs_ref
def test_obj_loader_32bit_indices(tmp_path):
    po = objgen.write_model(str(tmp_path / "a"), seed=5, big=True)
    pr = objgen.write_model(str(tmp_path / "b"), seed=5, big=True)
    ours = capi.Model(po, capi.OBJ_NO_CACHE_WRITE)
    assert not os.path.exists(po + ".bin")
    ref = rh.ref_load_model(pr, 0)
    assert ours.meshes[-1]["index_stride"] == 4 and ours.meshes[-1]["vertices"].shape[0] == 261 * 261
    assert ours.meshes[0]["index_stride"] == 2
    _same_models(ours, ref)
    ours.close()


@needs_ref
def test_obj_loader_errors_like_reference(tmp_path):
    cases = {
        "bad_pos.obj": "v 1 2\nf 1 1 1\n",
        "bad_uv.obj": "v 0 0 0\nvt 0.5\n",
        "bad_normal.obj": "v 0 0 0\nvn 1 0\n",
        "face_out_of_range.obj": "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 4\n",
        "face_negative_out_of_range.obj": "v 0 0 0\nv 1 0 0\nv 0 1 0\nf -1 -2 -4\n",
        "uv_out_of_range.obj": "v 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nf 1/1 2/2 3/1\n",
    }
    for name, text in cases.items():
        p = tmp_path / name
        p.write_text(text)
        assert rh.ref_load_model(str(p), 0) is None, name
        with pytest.raises(capi.SrbError):
            capi.Model(str(p), capi.OBJ_NO_CACHE_WRITE)
    with pytest.raises(capi.SrbError, match="Failed to open obj file"):
        capi.Model(str(tmp_path / "nothing.obj"))
    # no faces at all is a valid, empty model on both sides
    p = tmp_path / "empty.obj"
    p.write_text("# nothing\nv 0 0 0\n")
    m = capi.Model(str(p), capi.OBJ_NO_CACHE_WRITE)
    ref = rh.ref_load_model(str(p), 0)
    assert m.meshes == [] and ref[0] == []
    m.close()


def test_truncated_cache_is_reparsed(tmp_path):
    po = objgen.write_model(str(tmp_path / "m"), seed=3)
    a = capi.Model(po, 0)
    blob = open(po + ".bin", "rb").read()
    open(po + ".bin", "wb").write(blob[: len(blob) // 2])
    b = capi.Model(po, capi.OBJ_NO_CACHE_WRITE)
    assert not b.from_cache
    _same_models(b, (a.meshes, a.materials))
    # explicit save / load round trip
    b.save_cache(po + ".bin")
    assert open(po + ".bin", "rb").read() == blob
    c = capi.Model(po, 0)
    assert c.from_cache
    _same_models(c, (a.meshes, a.materials))
    d = capi.Model(po, capi.OBJ_NO_CACHE_READ | capi.OBJ_NO_CACHE_WRITE)
    assert not d.from_cache
    for m in (a, b, c, d):
        m.close()


def test_cache_with_an_index_past_the_vertices_is_reparsed(tmp_path):
    """A `.bin` whose arrays are well-formed but whose index values point past the mesh's vertices (corrupt or foreign file)
    must not be accepted: a draw of the resident model would read out of bounds on the device."""
    import struct

    po = objgen.write_model(str(tmp_path / "m"), seed=4)
    a = capi.Model(po, 0)
    blob = bytearray(open(po + ".bin", "rb").read())
    # layout (Obj.cpp:15-39 through kt::Serialize): u32 meshes | per mesh: u32 indexType, u32 indexBytes, indices ...
    num_meshes, index_type, index_bytes = struct.unpack_from("<III", blob, 0)
    assert num_meshes >= 1 and index_bytes >= 6
    blob[12:14] = b"\xff\xff" if index_type == 0 else blob[12:14]
    if index_type == 1:
        blob[12:16] = b"\xff\xff\xff\x7f"
    open(po + ".bin", "wb").write(bytes(blob))
    b = capi.Model(po, capi.OBJ_NO_CACHE_WRITE)
    assert not b.from_cache
    _same_models(b, (a.meshes, a.materials))
    a.close()
    b.close()


def test_model_to_scene_follows_scene_cpp(tmp_path):
    """Viewer/Scene.cpp:35-63: one draw per mesh, uv offset 6, UnlitDiffuse + the material's texture, VisualizeNormals when
    m_matIdx names no material."""
    p = tmp_path / "nomtl.obj"
    p.write_text("v 0 0 1\nv 1 0 1\nv 0 1 1\nvn 0 0 -1\nf 1//1 2//1 3//1\n")
    m = capi.Model(str(p), capi.OBJ_NO_CACHE_WRITE)
    sc = m.to_scene(64, 64, np.eye(4, dtype=np.float32).reshape(-1))
    assert len(sc.draws) == 1 and sc.draws[0].shader == 1 and sc.draws[0].texture == -1 and sc.draws[0].uv_offset == 6
    assert np.array_equal(sc.draws[0].vertices[:, 3:6], np.tile(np.float32([0, 0, -1]), (3, 1)))
    assert np.array_equal(sc.draws[0].vertices[:, 6:8], np.zeros((3, 2), np.float32))
    m.close()
    po = objgen.write_model(str(tmp_path / "m"), seed=3)
    m = capi.Model(po, capi.OBJ_NO_CACHE_WRITE)
    sc = m.to_scene(64, 64, np.eye(4, dtype=np.float32).reshape(-1))
    assert [d.shader for d in sc.draws] == [0] * 5
    assert [d.texture for d in sc.draws] == [0, 1, 2, -1, -1]  # 'broken' and 'plain' have no texels: null texture
    m.close()


def test_shim_obj_example_compiles_against_header():
    """CPU: Viewer/Scene.cpp's OBJ scene written against include/softrast_b200/Obj.h must compile and link."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert os.path.exists(os.path.join(root, "tests", "cpp", "_build", "shim_obj_example")), "run __graft_entry__.build()"


def _fuzz_obj(rng) -> str:
    """Random OBJ text biased towards the parser's decision points (Obj.cpp:160-312,399-546): corner syntax variants, zero
    / negative / out-of-range indices, short and long faces, junk after numbers, blank-prefixed and tab-separated lines,
    group / material lines in any order."""
    lines = []
    n_pos = n_uv = n_norm = 0
    for _ in range(int(rng.integers(5, 40))):
        kind = rng.choice(["v", "v", "v", "vt", "vn", "f", "f", "f", "f", "g", "usemtl", "junk"])
        pre = rng.choice(["", "", " ", "\t", "  "])
        post = rng.choice(["", "", " ", "\t\r", " \t "])
        if kind == "v":
            sep = rng.choice([" ", "  ", "\t"])
            vals = [f"{rng.uniform(-5, 5):.4f}" for _ in range(int(rng.choice([3] * 30 + [4, 4, 2])))]
            lines.append(pre + "v " + sep.join(vals) + post)
            n_pos += len(vals) >= 3
        elif kind == "vt":
            vals = [f"{rng.uniform(-2, 2):.4f}" for _ in range(int(rng.choice([2] * 30 + [3, 3, 1])))]
            lines.append(pre + "vt " + " ".join(vals) + post)
            n_uv += len(vals) >= 2
        elif kind == "vn":
            vals = [f"{rng.uniform(-1, 1):.4f}" for _ in range(int(rng.choice([3] * 30 + [4, 2])))]
            lines.append(pre + "vn " + " ".join(vals) + post)
            n_norm += len(vals) >= 3
        elif kind == "f":
            corners = []
            for _ in range(int(rng.choice([3, 3, 3, 4, 4, 5, 2, 1, 0]))):
                def idx(n):
                    r = rng.random()
                    if r < 0.70 and n:
                        return str(int(rng.integers(1, n + 1)))
                    if r < 0.94 and n:
                        return str(-int(rng.integers(1, n + 1)))
                    if r < 0.985:
                        return "0"
                    if r < 0.993:
                        return str(n + int(rng.integers(1, 4)))
                    return str(-(n + int(rng.integers(1, 4))))
                form = rng.choice(["p", "p/t", "p//n", "p/t/n"] * 8 + ["p/", "p/t/"])
                c = idx(n_pos)
                if form == "p/t":
                    c += "/" + idx(n_uv)
                elif form == "p//n":
                    c += "//" + idx(n_norm)
                elif form == "p/t/n":
                    c += "/" + idx(n_uv) + "/" + idx(n_norm)
                elif form == "p/":
                    c += "/"
                elif form == "p/t/":
                    c += "/" + idx(n_uv) + "/"
                corners.append(c)
            tail = rng.choice(["", "", "", " x", " #c", " 1e3"])
            lines.append(pre + "f " + rng.choice([" ", "  ", "\t"]).join(corners) + tail + post)
        elif kind == "g":
            lines.append(pre + rng.choice(["g", "g a", "group", "g\tb c"]) + post)
        elif kind == "usemtl":
            lines.append(pre + "usemtl " + rng.choice(["a", "b", "nope"]) + post)
        else:
            lines.append(pre + rng.choice(["# c", "", "s 1", "o x", "vp 1 2", "fx", "mtllib", "usemt l", "vx 1 2 3", "\t"]) + post)
    return "\n".join(lines) + "\n"


@needs_ref
def test_obj_parser_fuzz_against_reference(tmp_path):
    """400 random OBJ files: same outcome as the reference's loader — failure, or identical meshes."""
    rng = np.random.default_rng(2024)
    (tmp_path / "m.mtl").write_text("newmtl a\nnewmtl b\n")
    outcomes = {"ok": 0, "fail": 0, "meshes": 0}
    for i in range(400):
        text = ("mtllib m.mtl\n" if i % 2 else "") + _fuzz_obj(rng)
        p = tmp_path / f"fuzz{i}.obj"
        p.write_text(text, newline="")
        flags = int(rng.choice([0, capi.OBJ_FLIP_UVS]))  # (FlipWinding reads past odd index counts in the reference)
        ref = rh.ref_load_model(str(p), flags)
        cache = str(p) + ".bin"
        if os.path.exists(cache):
            os.remove(cache)
        try:
            ours = capi.Model(str(p), flags | capi.OBJ_NO_CACHE_WRITE)
        except capi.SrbError:
            assert ref is None, f"case {i}: the reference loads this file\n{text}"
            outcomes["fail"] += 1
            continue
        assert ref is not None, f"case {i}: the reference rejects this file\n{text}"
        try:
            _same_models(ours, ref)
        except AssertionError:
            print(text)
            raise
        outcomes["ok"] += 1
        outcomes["meshes"] += len(ours.meshes)
        ours.close()
    assert outcomes["ok"] > 80 and outcomes["fail"] > 80 and outcomes["meshes"] > 150, outcomes


def _golden_obj(name):
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "obj", name + ".npz"))
    meshes = [{"indices": z[f"m{i}_indices"], "vertices": z[f"m{i}_vertices"], "material": int(z[f"m{i}_material"]),
               "index_stride": z[f"m{i}_indices"].dtype.itemsize} for i in range(int(z["n_meshes"]))]
    mats = []
    for i in range(int(z["n_materials"])):
        nm, wl, hl, bpp = (int(v) for v in z[f"t{i}_meta"])
        mats.append({"name": z[f"t{i}_name"].tobytes().decode("latin-1"), "texels": z[f"t{i}_texels"], "mip_offsets": z[f"t{i}_mip_offsets"],
                     "num_mips": nm, "width_log2": wl, "height_log2": hl, "bytes_per_pixel": bpp})
    files = {z[f"f{i}_name"].tobytes().decode(): z[f"f{i}_data"].tobytes() for i in range(int(z["n_files"]))}
    return files, int(z["flags"]), (meshes, mats), z["cache"].tobytes()


@pytest.mark.parametrize("name", ["model_s41", "model_s42_flipped_crlf"])
def test_obj_loader_against_golden_fixture(tmp_path, name):
    """tests/golden/obj/*.npz: inputs + what the reference's loader made of them + the cache file it wrote
    (tests/golden/make_golden_obj.py).  Needs neither /root/reference nor oracle/_ref."""
    files, flags, ref, cache = _golden_obj(name)
    for fname, data in files.items():
        (tmp_path / fname).write_bytes(data)
    obj = str(tmp_path / "model.obj")
    ours = capi.Model(obj, flags)
    assert not ours.from_cache
    _same_models(ours, ref)
    ours.close()
    # the cache the REFERENCE wrote loads here and gives the same model
    (tmp_path / "model.obj.bin").write_bytes(cache)
    cached = capi.Model(obj, 0)
    assert cached.from_cache
    _same_models(cached, ref)
    cached.close()


def test_image_header_claiming_a_huge_size_is_rejected(tmp_path):
    """A PNG / TGA header may claim any size; nothing is allocated for an image larger than the largest texture
    (2^14 texels a side), and no C++ exception crosses the C ABI."""
    import struct
    import zlib

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 1 << 20, 1 << 20, 8, 6, 0, 0, 0))
    png += chunk(b"IDAT", zlib.compress(b"\0" * 64)) + chunk(b"IEND", b"")
    p = tmp_path / "huge.png"
    p.write_bytes(png)
    with pytest.raises(capi.SrbError):
        capi.load_image_rgba8(str(p))
    tga = struct.pack("<BBBHHBHHHHBB", 0, 0, 2, 0, 0, 0, 0, 0, 65535, 65535, 32, 8) + b"\0" * 64
    p = tmp_path / "huge.tga"
    p.write_bytes(tga)
    with pytest.raises(capi.SrbError):
        capi.load_image_rgba8(str(p))


def test_loaders_under_address_and_ub_sanitizers(tmp_path):
    """The host-side loaders compiled with -fsanitize=address,undefined (tests/cpp/asan_loader_driver.cpp) over valid,
    fuzzed, truncated and bit-flipped OBJ / PNG / TGA / cache files: files may be rejected, but no sanitizer report."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "asan_driver")
    src = [os.path.join(root, "tests", "cpp", "asan_loader_driver.cpp"), os.path.join(root, "softrast_b200", "csrc", "srb_model.cpp"),
           os.path.join(root, "softrast_b200", "csrc", "srb_host.cpp")]
    res = subprocess.run(["g++", "-std=c++17", "-g", "-O1", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-msse2",
                          "-I" + os.path.join(root, "include")] + src + ["-o", exe, "-lz"], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    rng = np.random.default_rng(77)
    files = []
    d = tmp_path / "files"
    d.mkdir()
    model = objgen.write_model(str(d / "m"), seed=6)
    files.append(model)
    (d / "m.mtl").write_text("newmtl a\nnewmtl b\n")
    for i in range(80):
        p = d / f"f{i}.obj"
        p.write_text(("mtllib m.mtl\n" if i % 2 else "") + _fuzz_obj(rng), newline="")
        files.append(str(p))
    images = []
    for k, (name, kw) in enumerate(_png_cases(rng)):
        if k % 3 == 0:
            p = str(d / f"{name}_{int(kw['interlace'])}.png")
            objgen.write_png(p, **kw)
            images.append(p)
    img = rng.integers(0, 256, (21, 34, 4)).astype(np.uint8)
    for bits in (32, 24, 8):
        p = str(d / f"t{bits}.tga")
        objgen.write_tga(p, img, bits=bits, rle=True)
        images.append(p)
    files += images
    for p in images:  # truncated and bit-flipped copies
        b = bytearray(open(p, "rb").read())
        for k in range(4):
            c = bytearray(b)
            if k < 2:
                c = c[: int(rng.integers(8, len(c)))]
            else:
                for _ in range(int(rng.integers(1, 6))):
                    c[int(rng.integers(0, len(c)))] = int(rng.integers(0, 256))
            q = os.path.join(str(d), f"c{k}_" + os.path.basename(p))
            open(q, "wb").write(c)
            files.append(q)
    m = capi.Model(model, 0)  # writes the cache
    m.close()
    blob = bytearray(open(model + ".bin", "rb").read())
    os.remove(model + ".bin")
    for k in range(24):
        c = bytearray(blob)
        if k % 4 == 0:
            c = c[: int(rng.integers(4, len(c)))]
        else:
            for _ in range(int(rng.integers(1, 8))):
                c[int(rng.integers(0, min(len(c), 400)))] = int(rng.integers(0, 256))
        kd = d / f"k{k}"
        kd.mkdir()
        open(str(kd / "x.obj.bin"), "wb").write(c)
        files.append(str(kd / "x.obj.bin"))
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    res = subprocess.run([exe] + files, capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, (res.stdout + res.stderr)[-3000:]
    assert "ERROR" not in res.stderr and "runtime error" not in res.stderr, res.stderr[-3000:]
    ok, rejected = (int(v) for v in res.stdout.split()[1::2])
    assert ok > 40 and rejected > 40 and ok + rejected == len(files)
