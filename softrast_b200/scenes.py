"""Synthetic scenes for the BASELINE.json configs and small parity scenes.

The reference ships no meshes (only Viewer/Models/cube/default.png), so every workload is generated here from fixed
seeds.  A scene is plain host data in the reference's own formats: 32-byte `Obj::Vertex` vertices
(pos3, normal3, uv2 — Viewer/Obj.h:16-21) used both as position buffer (stride 32, Vec3 at offset 0) and as attribute
buffer (8 varyings, uvOffset 6) exactly as Viewer/Scene.cpp:43-51 binds them; u32/u16 indices; a column-major MVP built
like Viewer/Scene.cpp:16-29 (PerspectiveLH_ZO, fov 85 deg, near/far swapped for reverse-Z, kt/src/kt/inl/Mat4.inl:300-315);
and RGBA8 textures in the reference's 32x32-tiled, Morton-swizzled, mip-chained layout (SoftRast/Texture.cpp:73-101,
:159-175).  The same arrays feed the CUDA library, the C oracle and the compiled reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

SHADER_UNLIT_DIFFUSE = 0
SHADER_VISUALIZE_NORMALS = 1
SHADER_VISUALIZE_UVS = 2
SHADER_SPONZA = 3  # Viewer/SponzaScene.cpp:13-104 (needs Scene.sponza)


# ----------------------------------------------------------------------------------------------------------------
# textures
# ----------------------------------------------------------------------------------------------------------------
def _morton5(x, y):
    """Bits of x in even positions, y in odd (Texture.cpp:36-41), for 5-bit in-tile coordinates."""
    x = x.astype(np.uint32)
    y = y.astype(np.uint32)
    m = np.zeros_like(x)
    for b in range(5):
        m |= ((x >> b) & 1) << (2 * b)
        m |= ((y >> b) & 1) << (2 * b + 1)
    return m


def _tile_level(img: np.ndarray) -> np.ndarray:
    """TileTexture (Texture.cpp:73-101): linear (h, w, 4) u8 -> tiled bytes padded to 32x32 tiles."""
    h, w, _ = img.shape
    tw = (w + 31) // 32
    th = (h + 31) // 32
    out = np.zeros(tw * th * 1024 * 4, dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    offs = ((yy >> 5) * tw + (xx >> 5)) * 1024 + _morton5(xx & 31, yy & 31)
    out.reshape(-1, 4)[offs.ravel()] = img.reshape(-1, 4)
    return out


def _box_down(img: np.ndarray) -> np.ndarray:
    """2x2 box filter (rounding to nearest) to the next mip; a dimension already at 1 stays 1."""
    h, w, _ = img.shape
    a = img.astype(np.uint32)
    if w > 1:
        a = a[:, 0::2] + a[:, 1::2]
    else:
        a = a * 2
    if h > 1:
        a = a[0::2] + a[1::2]
    else:
        a = a * 2
    return ((a + 2) >> 2).astype(np.uint8)


@dataclass
class TiledTexture:
    """The fields of sr::Tex::TextureData (Texture.h:33-40)."""

    texels: np.ndarray  # u8 blob
    mip_offsets: np.ndarray  # u32[14]
    num_mips: int
    width_log2: int
    height_log2: int


def build_tiled_texture(rgba: np.ndarray, calc_mips: bool = True) -> TiledTexture:
    """Host-side texture build in the reference layout (CreateFromRGBA8, Texture.cpp:119-199) with box-filter mips.
    Must produce the same bytes as the library's srb_texture_build_rgba8 (tested)."""
    h, w, c = rgba.shape
    assert c == 4 and rgba.dtype == np.uint8
    assert w & (w - 1) == 0 and h & (h - 1) == 0 and w % 32 == 0 and h % 32 == 0
    wl, hl = w.bit_length() - 1, h.bit_length() - 1
    n = (max(wl, hl) + 1) if calc_mips else 1
    offsets = np.zeros(14, dtype=np.uint32)
    blobs = []
    cur = 0
    level = np.ascontiguousarray(rgba)
    for m in range(n):
        if m > 0:
            level = _box_down(level)
        offsets[m] = cur
        t = _tile_level(level)
        blobs.append(t)
        cur += t.size
    return TiledTexture(np.concatenate(blobs), offsets, n, wl, hl)


def procedural_rgba(size: int, seed: int) -> np.ndarray:
    """Deterministic colourful texture: low-frequency colour ramps x checker x value noise."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:size, 0:size].astype(np.float32) / size
    base = rng.uniform(0.2, 1.0, size=3).astype(np.float32)
    freq = rng.integers(2, 9, size=3)
    img = np.zeros((size, size, 4), dtype=np.float32)
    for ch in range(3):
        img[..., ch] = base[ch] * (0.55 + 0.45 * np.sin(2 * np.pi * (freq[ch] * x + (ch + 1) * y)))
    cells = int(rng.integers(4, 17))
    checker = (((x * cells).astype(np.int32) + (y * cells).astype(np.int32)) & 1).astype(np.float32)
    img[..., :3] *= (0.6 + 0.4 * checker)[..., None]
    noise = rng.uniform(0.85, 1.0, size=(size, size, 1)).astype(np.float32)
    img[..., :3] *= noise
    img[..., 3] = 1.0
    return np.clip(img * 255.0 + 0.5, 0, 255).astype(np.uint8)


# ----------------------------------------------------------------------------------------------------------------
# matrices (math convention: clip = M @ [x, y, z, 1]; stored column-major like kt::Mat4)
# ----------------------------------------------------------------------------------------------------------------
def perspective_lh_zo(fov: float, aspect: float, near: float, far: float) -> np.ndarray:
    """kt::Mat4::PerspectiveLH_ZO (kt/src/kt/inl/Mat4.inl:300-315) as a 4x4 math matrix, float32 arithmetic."""
    f32 = np.float32
    f = f32(math.tan(math.pi / 2 - fov * 0.5))
    w = f32(f / f32(aspect))
    rng = f32(f32(far) / f32(f32(far) - f32(near)))
    m = np.zeros((4, 4), dtype=np.float32)
    m[0, 0] = w
    m[1, 1] = f
    m[2, 2] = rng
    m[3, 2] = 1.0
    m[2, 3] = f32(-rng * f32(near))
    return m


def reverse_z_projection(width: int, height: int) -> np.ndarray:
    """Viewer/Scene.cpp:16-29: fov 85 deg, near plane 10000, far plane 0.1 (swapped: reverse-Z)."""
    return perspective_lh_zo(math.radians(85.0), width / height, 10000.0, 0.1)


def look_at_lh(eye, target, up=(0.0, 1.0, 0.0)) -> np.ndarray:
    eye = np.asarray(eye, dtype=np.float64)
    fwd = np.asarray(target, dtype=np.float64) - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(np.asarray(up, dtype=np.float64), fwd)
    right /= np.linalg.norm(right)
    upv = np.cross(fwd, right)
    m = np.eye(4)
    m[0, :3], m[1, :3], m[2, :3] = right, upv, fwd
    m[0, 3], m[1, 3], m[2, 3] = -right @ eye, -upv @ eye, -fwd @ eye
    return m.astype(np.float32)


def to_column_major(m: np.ndarray) -> np.ndarray:
    """4x4 math matrix -> 16 floats with out[4*c + r] = m[r, c] (kt::Mat4::m_cols)."""
    return np.ascontiguousarray(m.astype(np.float32).T).reshape(16)


# ----------------------------------------------------------------------------------------------------------------
# scene containers
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class Draw:
    vertices: np.ndarray  # float32 (N, 8): pos3, normal3, uv2  (or (N, k) with k*4 == attribute stride)
    indices: np.ndarray  # uint32 / uint16 / uint8, flat, 3 per triangle
    mvp: np.ndarray  # float32[16], column-major
    shader: int = SHADER_UNLIT_DIFFUSE
    texture: int = -1  # index into Scene.textures, -1 = null texture
    uv_offset: int = 6

    @property
    def num_tris(self) -> int:
        return self.indices.size // 3


@dataclass
class Scene:
    name: str
    width: int
    height: int
    draws: list = field(default_factory=list)
    textures: list = field(default_factory=list)  # list[TiledTexture]
    clear_color: int = 0
    sponza: np.ndarray | None = None  # float32[136] = srb_sponza_constants, for draws with SHADER_SPONZA

    @property
    def num_tris(self) -> int:
        return sum(d.num_tris for d in self.draws)

    @property
    def tiles(self):
        return ((self.width + 63) // 64, (self.height + 63) // 64)


def _orient(verts: np.ndarray, tris: np.ndarray, want_normal: np.ndarray) -> np.ndarray:
    """Order each triangle so that the reference keeps it when seen from the side `want_normal` points to.
    The reference culls area2 <= 0 in y-down raster space (Binning.cpp:305-311); with the left-handed view space that
    keeps triangles whose right-hand cross(v1-v0, v2-v0) points AWAY from the viewer."""
    p = verts[:, :3].astype(np.float64)
    a = p[tris[:, 1]] - p[tris[:, 0]]
    b = p[tris[:, 2]] - p[tris[:, 0]]
    n = np.cross(a, b)
    flip = np.einsum("ij,ij->i", n, want_normal) > 0
    out = tris.copy()
    out[flip, 1], out[flip, 2] = tris[flip, 2], tris[flip, 1]
    return out


def _grid_surface(origin, du, dv, nu, nv, normal, uv_scale=(1.0, 1.0)):
    """A tessellated parallelogram: (nu x nv) quads -> vertices (N, 8) and oriented triangles (T, 3)."""
    origin, du, dv = (np.asarray(v, dtype=np.float64) for v in (origin, du, dv))
    s, t = np.meshgrid(np.linspace(0, 1, nu + 1), np.linspace(0, 1, nv + 1), indexing="xy")
    pos = origin + s[..., None] * du + t[..., None] * dv
    n = np.asarray(normal, dtype=np.float64)
    verts = np.zeros(((nu + 1) * (nv + 1), 8), dtype=np.float32)
    verts[:, 0:3] = pos.reshape(-1, 3)
    verts[:, 3:6] = n
    verts[:, 6] = (s * uv_scale[0]).ravel()
    verts[:, 7] = (t * uv_scale[1]).ravel()
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="xy")
    v00 = (j * (nu + 1) + i).ravel()
    v10, v01, v11 = v00 + 1, v00 + nu + 1, v00 + nu + 2
    tris = np.concatenate([np.stack([v00, v10, v11], 1), np.stack([v00, v11, v01], 1)]).astype(np.int64)
    tris = _orient(verts, tris, np.broadcast_to(n, (tris.shape[0], 3)))
    return verts, tris


def _param_surface(pos, nrm, uv, nu, nv, wrap_u=False):
    """Triangulate a (nv+1, nu+1) grid of positions with per-vertex outward normals."""
    verts = np.zeros((pos.shape[0] * pos.shape[1], 8), dtype=np.float32)
    verts[:, 0:3] = pos.reshape(-1, 3)
    verts[:, 3:6] = nrm.reshape(-1, 3)
    verts[:, 6:8] = uv.reshape(-1, 2)
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="xy")
    v00 = (j * (nu + 1) + i).ravel()
    v10, v01, v11 = v00 + 1, v00 + nu + 1, v00 + nu + 2
    tris = np.concatenate([np.stack([v00, v10, v11], 1), np.stack([v00, v11, v01], 1)]).astype(np.int64)
    face_n = (verts[tris[:, 0], 3:6] + verts[tris[:, 1], 3:6] + verts[tris[:, 2], 3:6]).astype(np.float64)
    tris = _orient(verts, tris, face_n)
    return verts, tris


def _merge(parts):
    vs, ts, base = [], [], 0
    for v, t in parts:
        vs.append(v)
        ts.append(t + base)
        base += v.shape[0]
    return np.concatenate(vs).astype(np.float32), np.concatenate(ts).astype(np.uint32).reshape(-1)


_CUBE_FACES = [  # (normal, u axis, v axis)
    ((0, 0, -1), (1, 0, 0), (0, 1, 0)),
    ((0, 0, 1), (-1, 0, 0), (0, 1, 0)),
    ((-1, 0, 0), (0, 0, -1), (0, 1, 0)),
    ((1, 0, 0), (0, 0, 1), (0, 1, 0)),
    ((0, 1, 0), (1, 0, 0), (0, 0, 1)),
    ((0, -1, 0), (1, 0, 0), (0, 0, -1)),
]


def unit_cube():
    """24 vertices / 12 triangles, outward-facing."""
    parts = []
    for n, ua, va in _CUBE_FACES:
        n, ua, va = (np.asarray(v, dtype=np.float64) for v in (n, ua, va))
        origin = 0.5 * n - 0.5 * ua - 0.5 * va
        parts.append(_grid_surface(origin, ua, va, 1, 1, n))
    v, idx = _merge(parts)
    return v, idx.reshape(-1, 3)


# ----------------------------------------------------------------------------------------------------------------
# BASELINE.json configs
# ----------------------------------------------------------------------------------------------------------------
def cube_grid(width=1280, height=720, nx=100, nz=100, draws=1, tex_size=128, seed=7) -> Scene:
    """Config 1: grid of nx*nz unit cubes on a plane receding from the camera (120 000 triangles at 100x100)."""
    cv, ct = unit_cube()
    ix, iz = np.meshgrid(np.arange(nx), np.arange(nz), indexing="xy")
    rng = np.random.default_rng(seed)
    offs = np.stack(
        [(ix.ravel() - (nx - 1) / 2) * 1.6, rng.uniform(-0.35, 0.35, nx * nz) - 1.8, 2.5 + iz.ravel() * 1.6], 1
    )
    verts = np.repeat(cv[None], nx * nz, 0)
    verts[:, :, 0:3] += offs[:, None, :].astype(np.float32)
    tris = ct[None] + (np.arange(nx * nz) * 24)[:, None, None]
    proj = reverse_z_projection(width, height)
    view = look_at_lh((0.0, 1.2, -1.0), (0.0, -1.0, 12.0))
    mvp = to_column_major(proj @ view)
    sc = Scene(f"cube_grid_{nx}x{nz}_{draws}draws", width, height)
    sc.textures.append(build_tiled_texture(procedural_rgba(tex_size, seed)))
    rows = np.array_split(np.arange(nz), draws)
    for r in rows:
        lo, hi = r[0] * nx, (r[-1] + 1) * nx
        v = verts[lo:hi].reshape(-1, 8)
        t = (tris[lo:hi] - lo * 24).astype(np.uint32).reshape(-1)
        sc.draws.append(Draw(np.ascontiguousarray(v), np.ascontiguousarray(t), mvp, SHADER_UNLIT_DIFFUSE, 0))
    return sc


def _column(cx, cz, radius, y0, y1, seg, rings, uv_tile):
    th = np.linspace(0, 2 * np.pi, seg + 1)
    yy = np.linspace(y0, y1, rings + 1)
    T, Y = np.meshgrid(th, yy, indexing="xy")
    pos = np.stack([cx + radius * np.cos(T), Y, cz + radius * np.sin(T)], -1)
    nrm = np.stack([np.cos(T), np.zeros_like(T), np.sin(T)], -1)
    uv = np.stack([T / (2 * np.pi) * uv_tile, (Y - y0) / (y1 - y0) * uv_tile * 2], -1)
    return _param_surface(pos, nrm, uv, seg, rings)


def _arch(cx, cz0, cz1, y, tube, seg, rings, uv_tile):
    """Half-torus spanning two columns along z."""
    R = 0.5 * (cz1 - cz0)
    zc = 0.5 * (cz0 + cz1)
    a = np.linspace(0, np.pi, seg + 1)
    b = np.linspace(0, 2 * np.pi, rings + 1)
    A, B = np.meshgrid(a, b, indexing="xy")
    ring_r = R + tube * np.cos(B)
    pos = np.stack([cx + tube * np.sin(B), y + ring_r * np.sin(A), zc - ring_r * np.cos(A)], -1)
    nrm = np.stack([np.sin(B), np.cos(B) * np.sin(A), -np.cos(B) * np.cos(A)], -1)
    uv = np.stack([A / np.pi * uv_tile, B / (2 * np.pi) * 2.0], -1)
    return _param_surface(pos, nrm, uv, seg, rings)


def _hall_view(t: float) -> np.ndarray:
    """Camera on a closed path inside the hall, t in [0, 1)."""
    ang = 2 * math.pi * t
    eye = (1.2 * math.sin(ang), 2.2 + 0.8 * math.sin(2 * ang), 8.0 + 6.0 * (1 - math.cos(ang)))
    target = (eye[0] * 0.5 + 0.8 * math.sin(ang * 3), 3.0, eye[2] + 20.0)
    return look_at_lh(eye, target)


def hall_scene(width=1920, height=1080, detail=1.0, seed=11, camera_t=0.0, lit=False) -> Scene:
    """Config 2: synthetic 'Sponza-scale' hall — ~262 k triangles in 25 draws / 25 mip-mapped textures
    (10x1024^2, 10x512^2, 5x256^2), camera inside, UV tiling up to 8x, depth complexity 3-4.
    `camera_t` in [0,1) moves the camera along a closed path (config 5).  lit: every draw uses the Sponza pixel shader
    (the reference viewer's default scene shader, Viewer/SponzaScene.cpp:13-104) with 16 point lights inside the hall."""
    d = lambda n: max(1, int(round(n * math.sqrt(detail))))
    L, W, H = 60.0, 16.0, 10.0  # length (z), width (x), height (y)
    draws = []
    # floor, ceiling, walls: normals point INTO the hall (towards the viewer inside)
    draws.append([_grid_surface((-W / 2, 0, 0), (W, 0, 0), (0, 0, L), d(100), d(152), (0, 1, 0), (8, 8))])
    draws.append([_grid_surface((-W / 2, H, 0), (W, 0, 0), (0, 0, L), d(70), d(92), (0, -1, 0), (4, 8))])
    draws.append([_grid_surface((-W / 2, 0, 0), (0, 0, L), (0, H, 0), d(120), d(36), (1, 0, 0), (8, 2))])
    draws.append([_grid_surface((W / 2, 0, 0), (0, 0, L), (0, H, 0), d(120), d(36), (-1, 0, 0), (8, 2))])
    draws.append([_grid_surface((-W / 2, 0, L), (W, 0, 0), (0, H, 0), d(48), d(32), (0, 0, -1), (2, 2))])
    draws.append([_grid_surface((-W / 2, 0, 0), (W, 0, 0), (0, H, 0), d(48), d(32), (0, 0, 1), (2, 2))])
    # 12 draws of columns: two rows, 30 columns per row, 5 per draw
    col_z = np.linspace(3.0, L - 3.0, 30)
    cols = [(-W / 2 + 3.0, z) for z in col_z] + [(W / 2 - 3.0, z) for z in col_z]
    for g in range(12):
        draws.append([_column(cx, cz, 0.45, 0.0, 6.5, d(32), d(30), 2.0) for cx, cz in cols[g * 5:(g + 1) * 5]])
    # 4 draws of arches joining consecutive columns
    arches = []
    for side in (-W / 2 + 3.0, W / 2 - 3.0):
        for k in range(0, 29):
            arches.append((side, col_z[k], col_z[k + 1]))
    for g in range(4):
        part = arches[g::4]
        draws.append([_arch(cx, z0, z1, 6.5, 0.3, d(24), d(14), 3.0) for cx, z0, z1 in part])
    # 3 draws of hanging banners (double-sided: two opposite-facing sheets)
    rng = np.random.default_rng(seed)
    for g in range(3):
        parts = []
        for k in range(8):
            z = 5.0 + (g * 8 + k) * 2.2
            x = rng.uniform(-2.5, 1.0)
            for nrm, eps in (((0, 0, -1), 0.0), ((0, 0, 1), 0.02)):
                parts.append(_grid_surface((x, 4.0, z + eps), (2.4, 0, 0), (0, 4.5, 0), d(16), d(28), nrm, (1, 2)))
        draws.append(parts)
    assert len(draws) == 25
    proj = reverse_z_projection(width, height)
    mvp = to_column_major(proj @ _hall_view(camera_t))
    sc = Scene("hall", width, height)
    sizes = [1024] * 10 + [512] * 10 + [256] * 5
    order = np.random.default_rng(seed + 1).permutation(25)
    for i, parts in enumerate(draws):
        v, idx = _merge(parts)
        sc.textures.append(build_tiled_texture(procedural_rgba(sizes[order[i]], seed * 100 + i)))
        sc.draws.append(Draw(v, idx, mvp, SHADER_SPONZA if lit else SHADER_UNLIT_DIFFUSE, i))
    if lit:
        sc.sponza = sponza_constants(seed, lo=(-W / 2 + 1.0, 1.0, 2.0), hi=(W / 2 - 1.0, H - 1.0, L - 2.0), intensity=(1.0, 4.0))
        sc.name = "hall_lit"
    return sc


def hall_camera_path(scene: Scene, frames: int) -> np.ndarray:
    """Config 5: `frames` MVPs on the closed camera path, shape (frames, n_draws, 16)."""
    proj = reverse_z_projection(scene.width, scene.height)
    out = np.zeros((frames, len(scene.draws), 16), dtype=np.float32)
    for f in range(frames):
        out[f, :, :] = to_column_major(proj @ _hall_view(f / frames))
    return out


def random_tris(width=1920, height=1080, n=1_000_000, seed=0x12345, min_px=2.0, max_px=16.0) -> Scene:
    """Config 3: n small random triangles, centres uniform over the screen, extents 2-16 px, view-space z uniform in
    [2, 52], random UVs in [0, 4), one draw, identity view."""
    rng = np.random.default_rng(seed)
    proj = reverse_z_projection(width, height)
    z = rng.uniform(2.0, 52.0, n)
    cx = rng.uniform(0, width, n)
    cy = rng.uniform(0, height, n)
    ext = rng.uniform(min_px, max_px, n)
    ang0 = rng.uniform(0, 2 * np.pi, n)
    px = np.zeros((n, 3, 2))
    for k in range(3):
        a = ang0 + k * 2 * np.pi / 3 + rng.uniform(-0.6, 0.6, n)
        r = ext * rng.uniform(0.35, 0.6, n)
        px[:, k, 0] = cx + r * np.cos(a)
        px[:, k, 1] = cy + r * np.sin(a)
    # unproject: raster x = (x*Pw/z)*W/2 + W/2 ; raster y = -(y*Pf/z)*H/2 + H/2
    pw, pf = float(proj[0, 0]), float(proj[1, 1])
    zv = z[:, None] + rng.uniform(-0.2, 0.2, (n, 3))
    vx = (px[:, :, 0] - width / 2) / (width / 2) * zv / pw
    vy = -(px[:, :, 1] - height / 2) / (height / 2) * zv / pf
    verts = np.zeros((n, 3, 8), dtype=np.float32)
    verts[:, :, 0], verts[:, :, 1], verts[:, :, 2] = vx, vy, zv
    verts[:, :, 5] = -1.0
    verts[:, :, 6:8] = rng.uniform(0, 4, (n, 3, 2))
    verts = verts.reshape(-1, 8)
    tris = np.arange(3 * n, dtype=np.int64).reshape(n, 3)
    # make ~all of them front-facing; leave every 16th as generated (about half of those get culled)
    keep = np.broadcast_to(np.array([0.0, 0.0, -1.0]), (n, 3))
    oriented = _orient(verts, tris, keep)
    oriented[::16] = tris[::16]
    sc = Scene(f"random_tris_{n}", width, height)
    sc.textures.append(build_tiled_texture(procedural_rgba(512, seed & 0xFFFF)))
    sc.draws.append(
        Draw(verts, oriented.astype(np.uint32).reshape(-1), to_column_major(proj), SHADER_UNLIT_DIFFUSE, 0)
    )
    return sc


# ----------------------------------------------------------------------------------------------------------------
# small parity scenes (edge cases the domain has: clipping on every plane, shared edges, ties, all shaders,
# u8/u16 indices, fewer varyings, ragged framebuffer sizes, empty draws)
# ----------------------------------------------------------------------------------------------------------------
def sponza_constants(seed=1, lo=(-6.0, -4.0, 1.0), hi=(6.0, 4.0, 20.0), intensity=(0.3, 1.5), phase=0.0) -> np.ndarray:
    """srb_sponza_constants as a float32[136]: sun_dir[3], ambient[3], pad[2], 16 x (pos3, colour3, intensity, falloff).
    Values in the spirit of SponzaScene::Init/Update (Viewer/SponzaScene.cpp:121-187): ambient 0.1, the sun direction
    broadcast from normalize(0.4, 0.7, 0.1).x into all three components (the reference's quirk, :135-137), 16 random
    coloured point lights inside [lo, hi]; colours lerp between two random colours with sin(phase)."""
    rng = np.random.default_rng(seed)
    k = np.zeros(136, dtype=np.float32)
    sun = np.array([0.4, 0.7, 0.1], dtype=np.float32)
    sun = sun / np.float32(np.sqrt(np.float32((sun * sun).sum())))
    k[0:3] = sun[0]
    k[3:6] = np.float32(0.1)
    lights = k[8:].reshape(16, 8)
    lo, hi = np.asarray(lo, dtype=np.float32), np.asarray(hi, dtype=np.float32)
    lights[:, 0:3] = lo + (hi - lo) * rng.random((16, 3), dtype=np.float32)
    ca, cb = rng.random((16, 3), dtype=np.float32), rng.random((16, 3), dtype=np.float32)
    s = np.float32(np.sin(phase) * 0.5 + 0.5)
    lights[:, 3:6] = (np.float32(1.0) - s) * ca + s * cb
    lights[:, 6] = np.float32(intensity[0]) + np.float32(intensity[1] - intensity[0]) * rng.random(16, dtype=np.float32)
    lights[:, 7] = 500.0 + 2000.0 * rng.random(16, dtype=np.float32)
    return k


def parity_scene(width=320, height=200, seed=3, n_small=400, n_big=24, lit=False) -> Scene:
    """lit: the textured draws use the Sponza pixel shader (sun + 16 point lights, RSQRTPS/RCPPS) instead of UnlitDiffuse."""
    rng = np.random.default_rng(seed)
    proj = reverse_z_projection(width, height)
    view = look_at_lh((0.3, 0.4, -0.5), (0.0, 0.0, 6.0))
    mvp = to_column_major(proj @ view)
    sc = Scene(f"parity_{width}x{height}_s{seed}", width, height, clear_color=0x20)
    sc.textures.append(build_tiled_texture(procedural_rgba(64, seed)))
    sc.textures.append(build_tiled_texture(procedural_rgba(128, seed + 1)))
    sc.textures.append(build_tiled_texture(procedural_rgba(32, seed + 2), calc_mips=False))

    def soup(n, zlo, zhi, size, uvs):
        c = np.stack([rng.uniform(-6, 6, n), rng.uniform(-4, 4, n), rng.uniform(zlo, zhi, n)], 1)
        v = np.zeros((n, 3, 8), dtype=np.float32)
        v[:, :, 0:3] = c[:, None, :] + rng.normal(0, size, (n, 3, 3))
        nr = rng.normal(0, 1, (n, 3, 3))
        v[:, :, 3:6] = nr / np.linalg.norm(nr, axis=-1, keepdims=True)
        v[:, :, 6:8] = rng.uniform(-uvs, uvs, (n, 3, 2))
        return v.reshape(-1, 8), np.arange(3 * n, dtype=np.uint32)

    # draw 0: small triangles, u16 indices, texture 0
    v, i = soup(n_small, 1.0, 20.0, 0.35, 3.0)
    sc.draws.append(Draw(v, i.astype(np.uint16), mvp, SHADER_UNLIT_DIFFUSE, 0))
    # draw 1: big triangles crossing every frustum plane incl. the near plane / behind the camera, texture 1
    v, i = soup(n_big, -3.0, 12.0, 5.0, 2.0)
    sc.draws.append(Draw(v, i, mvp, SHADER_UNLIT_DIFFUSE, 1))
    # draw 2: a tessellated wall with shared edges (fill-rule test), normals shader, u8 indices
    gv, gt = _grid_surface((-3, -2, 9), (6, 0.5, 1.0), (0.3, 4, -0.5), 6, 5, (0, 0, -1), (3, 3))
    sc.draws.append(Draw(gv, gt.reshape(-1).astype(np.uint8), mvp, SHADER_VISUALIZE_NORMALS, -1))
    # draw 3: exact duplicates of part of draw 2 (depth ties: first in canonical order must win), UV shader
    sc.draws.append(Draw(gv.copy(), gt[: len(gt) // 2].reshape(-1).astype(np.uint32), mvp, SHADER_VISUALIZE_UVS, -1))
    # draw 4: null texture -> white (Shaders.h:75-79)
    v, i = soup(20, 2.0, 10.0, 0.8, 1.0)
    sc.draws.append(Draw(v, i, mvp, SHADER_UNLIT_DIFFUSE, -1))
    # draw 5: no-mip texture, large UVs + negative UVs
    v, i = soup(60, 1.5, 8.0, 0.6, 9.0)
    sc.draws.append(Draw(v, i, mvp, SHADER_UNLIT_DIFFUSE, 2))
    # draw 6: empty draw
    sc.draws.append(Draw(v[:3].copy(), np.zeros(0, dtype=np.uint32), mvp, SHADER_UNLIT_DIFFUSE, 0))
    # draw 7: a screen-filling pair of triangles far away (many tiles per triangle)
    gv2, gt2 = _grid_surface((-60, -40, 30), (120, 0, 0), (0, 80, 3), 1, 1, (0, 0, -1), (6, 6))
    sc.draws.append(Draw(gv2, gt2.reshape(-1).astype(np.uint32), mvp, SHADER_UNLIT_DIFFUSE, 1))
    if lit:
        for d in sc.draws:
            if d.shader == SHADER_UNLIT_DIFFUSE:
                d.shader = SHADER_SPONZA  # incl. the null-texture draw: white like UnlitDiffuse (SponzaScene.cpp:17-21)
        sc.sponza = sponza_constants(seed)
        sc.name += "_lit"
    return sc
