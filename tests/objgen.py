"""Test helpers: writers for the asset formats the OBJ ingestion path reads — PNG (every colour type / bit depth /
filter / interlace the decoder handles), TGA, and seeded Wavefront OBJ + MTL files that exercise the reference parser's
rules (Viewer/Obj.cpp:160-312,399-546).  Everything is generated; nothing is read from /root/reference."""
import os
import struct
import zlib

import numpy as np

_ADAM7 = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]  # x0, y0, dx, dy


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def _pack_rows(samples: np.ndarray, depth: int) -> list:
    """samples: (h, w, channels) integer samples -> list of packed scanline byte strings."""
    h, w, ch = samples.shape
    rows = []
    for y in range(h):
        flat = samples[y].reshape(-1)
        if depth == 16:
            rows.append(flat.astype(">u2").tobytes())
        elif depth == 8:
            rows.append(flat.astype(np.uint8).tobytes())
        else:
            bits = np.zeros(((w * depth + 7) // 8) * 8, dtype=np.uint8)
            for b in range(depth):
                bits[np.arange(w) * depth + b] = (flat >> (depth - 1 - b)) & 1
            rows.append(np.packbits(bits).tobytes())
    return rows


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    if pa <= pb and pa <= pc:
        return a
    return b if pb <= pc else c


def _filter_rows(rows: list, bpp: int, filters) -> bytes:
    out = bytearray()
    prev = bytes(len(rows[0])) if rows else b""
    for y, row in enumerate(rows):
        f = filters[y % len(filters)]
        out.append(f)
        cur = bytearray(len(row))
        for i in range(len(row)):
            a = row[i - bpp] if i >= bpp else 0
            b = prev[i]
            c = prev[i - bpp] if i >= bpp else 0
            pred = (0, a, b, (a + b) >> 1, _paeth(a, b, c))[f]
            cur[i] = (row[i] - pred) & 255
        out += cur
        prev = row
    return bytes(out)


def write_png(path, samples: np.ndarray, colour_type: int, depth: int, *, palette=None, trns: bytes | None = None,
              interlace: bool = False, filters=(0, 1, 2, 3, 4), idat_split: int = 0):
    """samples: (h, w, channels) raw sample values (< 2**depth).  palette: (n, 3) uint8 for colour type 3."""
    h, w, ch = samples.shape
    bits = ch * depth
    bpp = max(1, bits // 8)
    raw = b""
    if interlace:
        for x0, y0, dx, dy in _ADAM7:
            sub = samples[y0::dy, x0::dx]
            if sub.shape[0] and sub.shape[1]:
                raw += _filter_rows(_pack_rows(sub, depth), bpp, filters)
    else:
        raw = _filter_rows(_pack_rows(samples, depth), bpp, filters)
    z = zlib.compress(raw, 6)
    out = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, colour_type, 0, 0, 1 if interlace else 0))
    out += _chunk(b"gAMA", struct.pack(">I", 45455))  # an ancillary chunk decoders must skip
    if palette is not None:
        out += _chunk(b"PLTE", np.asarray(palette, np.uint8).tobytes())
    if trns is not None:
        out += _chunk(b"tRNS", trns)
    if idat_split:
        for i in range(0, len(z), idat_split):
            out += _chunk(b"IDAT", z[i:i + idat_split])
    else:
        out += _chunk(b"IDAT", z)
    out += _chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(out)


def write_png_rgba8(path, rgba: np.ndarray):
    write_png(path, rgba.astype(np.int64), 6, 8)


def write_tga(path, rgba: np.ndarray, *, bits: int = 32, rle: bool = False, top_down: bool = False, id_bytes: bytes = b""):
    """bits: 32 (BGRA), 24 (BGR) or 8 (grey = the red channel)."""
    h, w, _ = rgba.shape
    kind = 3 if bits == 8 else 2
    hdr = struct.pack("<BBBHHBHHHHBB", len(id_bytes), 0, kind + (8 if rle else 0), 0, 0, 0, 0, 0, w, h, bits,
                      (0x20 if top_down else 0) | (8 if bits == 32 else 0))
    rows = rgba if top_down else rgba[::-1]
    if bits == 8:
        px = rows[..., 0:1]
    elif bits == 24:
        px = rows[..., [2, 1, 0]]
    else:
        px = rows[..., [2, 1, 0, 3]]
    px = np.ascontiguousarray(px, dtype=np.uint8).reshape(-1, px.shape[-1])
    body = bytearray()
    if not rle:
        body += px.tobytes()
    else:
        i, n = 0, px.shape[0]
        while i < n:
            run = 1
            while i + run < n and run < 128 and np.array_equal(px[i + run], px[i]):
                run += 1
            if run >= 2:
                body.append(0x80 | (run - 1))
                body += px[i].tobytes()
                i += run
            else:
                lit = 1
                while i + lit < n and lit < 128 and not np.array_equal(px[i + lit], px[i + lit - 1]):
                    lit += 1
                body.append(lit - 1)
                body += px[i:i + lit].tobytes()
                i += lit
    with open(path, "wb") as f:
        f.write(hdr + id_bytes + bytes(body))


def procedural_rgba(size: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:size, 0:size]
    base = np.stack([(x * 255 // (size - 1)), (y * 255 // (size - 1)), ((x // 8 + y // 8) & 1) * 255, np.full_like(x, 255)], -1)
    noise = rng.integers(-24, 25, size=(size, size, 4))
    noise[..., 3] = 0
    return np.clip(base + noise, 0, 255).astype(np.uint8)


def write_model(dirpath: str, name: str = "model", seed: int = 7, *, big: bool = False, crlf: bool = False) -> str:
    """A seeded OBJ + MTL + textures that walks through the reference parser's rules: shared and per-face vertices, quads,
    a pentagon (cut to a quad), p / p/t / p//n / p/t/n corners, negative indices, several groups with materials chosen
    before and after the group line, an unknown material, comment / blank / unknown lines, tabs and trailing blanks.
    big=True adds a grid with more than 65535 vertices (32-bit indices).  Returns the OBJ path."""
    rng = np.random.default_rng(seed)
    os.makedirs(dirpath, exist_ok=True)
    write_png_rgba8(os.path.join(dirpath, "bricks.png"), procedural_rgba(64, seed))
    write_tga(os.path.join(dirpath, "tiles.tga"), procedural_rgba(32, seed + 1), bits=24, rle=True)
    write_png(os.path.join(dirpath, "grey.png"), procedural_rgba(32, seed + 2)[..., :1].astype(np.int64), 0, 8, interlace=True)
    nl = "\r\n" if crlf else "\n"
    mtl = [
        "# materials", "",
        "newmtl bricks", "  Kd 1 1 1", "  map_Kd bricks.png  ",
        "newmtl plain", "Kd 0.5 0.5 0.5",
        "newmtl\ttiles", "\tmap_Kd tiles.tga",
        "newmtl grey", "map_Kd   grey.png",
        "newmtl broken", "map_Kd does_not_exist.png",
    ]
    with open(os.path.join(dirpath, name + ".mtl"), "w", newline="") as f:
        f.write(nl.join(mtl) + nl)

    L = ["# generated", "mtllib " + name + ".mtl", "o thing"]
    # group 1: a tessellated, tilted sheet with shared vertices (p/t/n), material set BEFORE the group line
    n = 9
    L.append("usemtl bricks")
    L.append("g sheet")
    base_p = 0
    for j in range(n + 1):
        for i in range(n + 1):
            s, t = i / n, j / n
            L.append(f"v {-3 + 6 * s:.6f} {-2 + 4 * t:.6f} {5 + 2.5 * s + 0.75 * t:.6f}")
            L.append(f"vt {3 * s:.6f} {2 * t:.6f}")
    L.append("vn 0 0 -1")
    for j in range(n):
        for i in range(n):
            a = j * (n + 1) + i + 1
            b, c, d = a + 1, a + n + 1, a + n + 2
            if (i + j) & 1:
                L.append(f"f {a}/{a}/1 {b}/{b}/1 {d}/{d}/1 {c}/{c}/1")  # quad
            else:
                L.append(f"f {a}/{a}/1 {b}/{b}/1 {d}/{d}/1")
                L.append(f"f  {a}/{a}/1\t{d}/{d}/1 {c}/{c}/1  ")
    base_p += (n + 1) * (n + 1)
    # group 2: loose triangles, negative (relative) indices, material chosen AFTER the group line (it applies to this
    # group's mesh because the mesh is closed by the NEXT group line)
    L.append("g loose")
    L.append("usemtl tiles")
    for k in range(40):
        c = np.array([rng.uniform(-4, 4), rng.uniform(-3, 3), rng.uniform(3, 9)])
        for _ in range(3):
            p = c + rng.uniform(-0.8, 0.8, 3)
            L.append(f"v {p[0]:.6f} {p[1]:.6f} {p[2]:.6f}")
            L.append(f"vt {rng.uniform(-2, 3):.5f} {rng.uniform(-2, 3):.5f}")
            L.append(f"vn {rng.uniform(-1, 1):.4f} {rng.uniform(-1, 1):.4f} {rng.uniform(-1, 1):.4f}")
        form = k % 4
        if form == 0:
            L.append("f -3/-3/-3 -2/-2/-2 -1/-1/-1")
        elif form == 1:
            L.append("f -3//-3 -2//-2 -1//-1")  # no uv index: uv entry 0
        elif form == 2:
            L.append("f -3/-3 -2/-2 -1/-1")  # no normal index: normal entry 0
        else:
            L.append("f -3 -2 -1")
    base_p += 120
    # group 3: a pentagon (only four corners are read), a degenerate two-corner face, an unknown material (ignored)
    L.append("g odd shapes")
    L.append("usemtl no_such_material")
    L.append("usemtl grey")
    for p in [(-1, -1, 4), (1, -1, 4), (1.5, 0.5, 4), (0, 1.5, 4), (-1.5, 0.5, 4)]:
        L.append(f"v {p[0]} {p[1]} {p[2]}")
    L.append("f -5/1/1 -4/2/1 -3/3/1 -2/4/1 -1/5/1")
    L.append("s off")
    L.append("vp 0.1 0.2")
    L.append("f -5/1/1 -4/2/1")
    L.append("f -5/1/1 -3/3/1 -1/5/1 garbage")
    L.append("f -2/4/1")  # one corner: with the two-corner face above the index count is a multiple of 3 again
    # group 4: material index past the list? (plain has no texture; broken has a missing file)
    L.append("g backdrop")
    L.append("usemtl broken")
    L.append("v -8 -6 12")
    L.append("v 8 -6 12")
    L.append("v 8 6 12")
    L.append("v -8 6 12")
    L.append("f -4/1/1 -3/2/1 -2/3/1 -1/4/1")
    L.append("g plainquad")
    L.append("usemtl plain")
    L.append("v -2 -2 3.5")
    L.append("v -1 -2 3.5")
    L.append("v -1 -1 3.5")
    L.append("f -3 -2 -1")
    if big:
        L.append("g biggrid")
        L.append("usemtl bricks")
        m = 260  # 261 * 261 = 68121 vertices > 65535
        for j in range(m + 1):
            for i in range(m + 1):
                L.append(f"v {-6 + 12 * i / m:.5f} {-4.5 + 9 * j / m:.5f} {14 + 0.002 * ((i * 7 + j * 13) % 17):.5f}")
        first = -(m + 1) * (m + 1)
        for j in range(0, m, 2):
            for i in range(0, m, 2):
                a = first + j * (m + 1) + i
                L.append(f"f {a} {a + 2} {a + 2 * (m + 1) + 2} {a + 2 * (m + 1)}")
        # every vertex must be referenced to make the vertex count exceed 16 bits: a strip of thin triangles
        for j in range(m + 1):
            for i in range(0, m - 1, 3):
                a = first + j * (m + 1) + i
                L.append(f"f {a} {a + 1} {a + 2}")
    path = os.path.join(dirpath, name + ".obj")
    with open(path, "w", newline="") as f:
        f.write(nl.join(L) + nl)
    return path


def write_png_rgba8_fast(path, rgba: np.ndarray):
    """RGBA8 PNG with filter type 0 on every row, written with numpy (for large images)."""
    h, w, _ = rgba.shape
    raw = np.concatenate([np.zeros((h, 1), np.uint8), np.ascontiguousarray(rgba, dtype=np.uint8).reshape(h, w * 4)], axis=1).tobytes()
    out = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0))
    out += _chunk(b"IDAT", zlib.compress(raw, 1)) + _chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(out)


def write_scene_as_obj(dirpath: str, scene, images, name: str = "scene") -> str:
    """Writes a scenes.Scene as OBJ + MTL + PNG files: one group and one material per draw, every vertex with its own
    v / vt / vn lines (printed with 9 significant digits, which round-trips float32 exactly), faces `i/i/i`.  `images`:
    one RGBA8 image per scene texture (the files the materials point to)."""
    os.makedirs(dirpath, exist_ok=True)
    for i, img in enumerate(images):
        write_png_rgba8_fast(os.path.join(dirpath, f"tex{i}.png"), img)
    with open(os.path.join(dirpath, name + ".mtl"), "w") as f:
        for i, d in enumerate(scene.draws):
            f.write(f"newmtl m{i}\n")
            if d.texture >= 0:
                f.write(f"map_Kd tex{d.texture}.png\n")
    path = os.path.join(dirpath, name + ".obj")
    with open(path, "w") as f:
        f.write(f"mtllib {name}.mtl\n")
        base = 0
        for i, d in enumerate(scene.draws):
            v = np.asarray(d.vertices, dtype=np.float32)
            f.write(f"usemtl m{i}\ng draw{i}\n")
            lines = []
            for row in v:
                lines.append("v %.9g %.9g %.9g\nvt %.9g %.9g\nvn %.9g %.9g %.9g\n" % (row[0], row[1], row[2], row[6], row[7], row[3], row[4], row[5]))
            f.write("".join(lines))
            idx = np.asarray(d.indices, dtype=np.int64).reshape(-1, 3) + base + 1
            f.write("".join("f %d/%d/%d %d/%d/%d %d/%d/%d\n" % (a, a, a, b, b, b, c, c, c) for a, b, c in idx))
            base += v.shape[0]
    return path
