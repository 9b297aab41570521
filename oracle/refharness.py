"""TEST INFRASTRUCTURE ONLY — ctypes binding of oracle/_ref/libsrref_{parity,fast}.so, the UNMODIFIED reference
renderer compiled in place from /root/reference by oracle/ref_build/Makefile (see ref_harness.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from softrast_b200._ctypes_defs import (
    COLOUR_TILE_BYTES,
    TILE_TRI_DTYPE,
    DrawDesc,
    MaterialView,
    MeshView,
    copy_material_view,
    copy_mesh_view,
    ptr,
)

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")


def ref_available(variant: str = "parity") -> bool:
    return os.path.exists(os.path.join(REF_DIR, f"libsrref_{variant}.so"))


_libs = {}


def _load(variant: str):
    if variant in _libs:
        return _libs[variant]
    path = os.path.join(REF_DIR, f"libsrref_{variant}.so")
    lib = C.CDLL(path, mode=os.RTLD_LOCAL)
    vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
    lib.srref_create.argtypes = [u32, u32, u32, u64, C.POINTER(vp)]
    lib.srref_destroy.argtypes = [vp]
    lib.srref_destroy.restype = None
    lib.srref_threads.argtypes = [vp]
    lib.srref_threads.restype = u32
    lib.srref_texture_create_tiled.argtypes = [vp, vp, u64, vp, u32, u32, u32]
    lib.srref_texture_create_tiled.restype = u64
    lib.srref_texture_create_rgba8.argtypes = [vp, vp, u32, u32, C.c_int]
    lib.srref_texture_create_rgba8.restype = u64
    lib.srref_texture_get.argtypes = [vp, u64, vp, C.POINTER(u64), vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
    lib.srref_begin_frame.argtypes = [vp]
    lib.srref_clear.argtypes = [vp, u32, C.c_int, C.c_int]
    lib.srref_draw_indexed.argtypes = [vp, C.POINTER(DrawDesc)]
    lib.srref_end_frame.argtypes = [vp]
    lib.srref_render_frames.argtypes = [vp, C.POINTER(DrawDesc), u32, vp, u32, u32, vp]
    lib.srref_read_tiles.argtypes = [vp, vp, vp, u64]
    lib.srref_blit_linear.argtypes = [vp, vp]
    lib.srref_dump_tile_counts.argtypes = [vp, vp, u32]
    lib.srref_dump_tile_tris.argtypes = [vp, u32, vp, u32, C.POINTER(u32)]
    lib.srref_dump_tile_coverage.argtypes = [vp, u32, vp, u32, C.POINTER(u32)]
    lib.srref_dump_tile_fragments.argtypes = [vp, u32, vp, u64, C.POINTER(u64), vp]
    lib.srref_rcp.argtypes = [vp, vp, u64]
    lib.srref_rcp.restype = None
    lib.srref_sample.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp, vp, u64]
    lib.srref_set_sponza_constants.argtypes = [vp]
    lib.srref_set_sponza_constants.restype = None
    lib.srref_rsqrt.argtypes = [vp, vp, u64]
    lib.srref_rsqrt.restype = None
    lib.srref_sponza_scene_create.restype = vp
    lib.srref_sponza_scene_destroy.argtypes = [vp]
    lib.srref_sponza_scene_destroy.restype = None
    lib.srref_sponza_scene_update.argtypes = [vp, vp, vp, C.c_float]
    lib.srref_sponza_scene_update.restype = None
    lib.srref_get_sponza_constants.argtypes = [vp]
    lib.srref_get_sponza_constants.restype = None
    lib.srref_raw_objects.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    lib.srref_raw_objects.restype = None
    lib.srref_model_load.argtypes = [C.c_char_p, u32, C.POINTER(vp)]
    lib.srref_model_free.argtypes = [vp]
    lib.srref_model_free.restype = None
    lib.srref_model_info.argtypes = [vp, C.POINTER(u32), C.POINTER(u32)]
    lib.srref_model_mesh.argtypes = [vp, u32, C.POINTER(MeshView)]
    lib.srref_model_material.argtypes = [vp, u32, C.POINTER(MaterialView)]
    lib.srref_image_load_rgba8.argtypes = [C.c_char_p, C.POINTER(vp), C.POINTER(u32), C.POINTER(u32)]
    lib.srref_image_free.argtypes = [vp]
    lib.srref_image_free.restype = None
    _libs[variant] = lib
    return lib


def ref_load_model(path: str, flags: int = 0, variant: str = "parity"):
    """sr::Obj::Model::Load (Viewer/Obj.cpp:374-560) of the compiled reference -> (meshes, materials) as numpy copies in
    the same shape softrast_b200.capi.Model exposes; None if Load returned false."""
    lib = _load(variant)
    h = C.c_void_p()
    if lib.srref_model_load(os.fsencode(path), flags, C.byref(h)) != 0:
        return None
    nm, nmat = C.c_uint32(), C.c_uint32()
    lib.srref_model_info(h, C.byref(nm), C.byref(nmat))
    meshes, mats = [], []
    for i in range(nm.value):
        v = MeshView()
        assert lib.srref_model_mesh(h, i, C.byref(v)) == 0
        meshes.append(copy_mesh_view(v))
    for i in range(nmat.value):
        v = MaterialView()
        assert lib.srref_model_material(h, i, C.byref(v)) == 0
        mats.append(copy_material_view(v))
    lib.srref_model_free(h)
    return meshes, mats


def ref_load_image(path: str, variant: str = "parity"):
    """stbi_load(path, &x, &y, &comp, 4) of the reference's vendored stb_image (Texture.cpp:107); None on failure."""
    lib = _load(variant)
    px, w, h = C.c_void_p(), C.c_uint32(), C.c_uint32()
    if lib.srref_image_load_rgba8(os.fsencode(path), C.byref(px), C.byref(w), C.byref(h)) != 0:
        return None
    out = np.frombuffer(C.string_at(px.value, w.value * h.value * 4), dtype=np.uint8).reshape(h.value, w.value, 4).copy()
    lib.srref_image_free(px)
    return out


def host_rsqrt(x: np.ndarray, variant: str = "parity") -> np.ndarray:
    """The host CPU's RSQRTPS on float32 inputs."""
    lib = _load(variant)
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    lib.srref_rsqrt(ptr(x), ptr(out), x.size)
    return out


def host_rcp(x: np.ndarray, variant: str = "parity") -> np.ndarray:
    """The host CPU's RCPPS on float32 inputs."""
    lib = _load(variant)
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    lib.srref_rcp(ptr(x), ptr(out), x.size)
    return out


def harvest_rsqrt_table(bits: int = 10) -> np.ndarray:
    """T[(p << bits) | i] = bits(RSQRTPS(2^p * (1 + i*2^-bits))), p in {0, 1} (softrast_b200/csrc/srb_host.cpp)."""
    i = np.arange(1 << bits, dtype=np.uint32) << (23 - bits)
    m = np.concatenate([i | np.uint32(127 << 23), i | np.uint32(128 << 23)])
    return host_rsqrt(m.view(np.float32)).view(np.uint32)


def harvest_rcp_table(bits: int = 11) -> np.ndarray:
    """T[i] = bits(RCPPS(1 + i*2^-bits))."""
    m = (np.arange(1 << bits, dtype=np.uint32) << (23 - bits)) | np.uint32(0x3F800000)
    return host_rcp(m.view(np.float32)).view(np.uint32)


def make_draw_descs(scene, tex_handles, keepalive: list, fb_handle: int = 0):
    """srb_draw_desc array for a scenes.Scene with HOST pointers (buffer handles 0)."""
    descs = (DrawDesc * max(1, len(scene.draws)))()
    for i, d in enumerate(scene.draws):
        v = np.ascontiguousarray(d.vertices, dtype=np.float32)
        idx = np.ascontiguousarray(d.indices)
        keepalive += [v, idx]
        e = descs[i]
        e.shader = d.shader
        e.uv_offset = d.uv_offset
        e.texture = tex_handles[d.texture] if d.texture >= 0 else 0
        e.framebuffer = fb_handle
        e.indices.host = idx.ctypes.data
        e.indices.stride = idx.dtype.itemsize
        e.indices.num = idx.size
        e.positions.host = v.ctypes.data
        e.positions.stride = v.shape[1] * 4
        e.positions.num = v.shape[0]
        e.attributes.host = v.ctypes.data
        e.attributes.stride = v.shape[1] * 4
        e.attributes.num = v.shape[0]
        for k in range(16):
            e.mvp[k] = float(d.mvp[k])
    return descs


class RefRenderer:
    """The reference renderer behind the srref_* C ABI.  threads=1 is the canonical (single-threaded) order."""

    def __init__(self, width: int, height: int, threads: int = 1, variant: str = "parity", arena_bytes: int = 0):
        self.lib = _load(variant)
        self.h = C.c_void_p()
        rc = self.lib.srref_create(threads, width, height, arena_bytes, C.byref(self.h))
        if rc != 0:
            raise RuntimeError(f"srref_create failed: {rc}")
        self.width, self.height = width, height
        self.tiles_x, self.tiles_y = (width + 63) // 64, (height + 63) // 64
        self.num_tiles = self.tiles_x * self.tiles_y
        self._keep = []

    @property
    def threads(self) -> int:
        return int(self.lib.srref_threads(self.h))

    def close(self):
        if self.h:
            self.lib.srref_destroy(self.h)
            self.h = C.c_void_p()

    # -- resources
    def create_texture(self, t) -> int:
        off = np.ascontiguousarray(t.mip_offsets, dtype=np.uint32)
        return int(
            self.lib.srref_texture_create_tiled(
                self.h, ptr(t.texels), t.texels.size, ptr(off), t.num_mips, t.width_log2, t.height_log2
            )
        )

    def create_texture_rgba8(self, rgba: np.ndarray, calc_mips=True) -> int:
        rgba = np.ascontiguousarray(rgba)
        return int(self.lib.srref_texture_create_rgba8(self.h, ptr(rgba), rgba.shape[1], rgba.shape[0], int(calc_mips)))

    def get_texture(self, handle: int):
        from softrast_b200.scenes import TiledTexture

        n = C.c_uint64()
        nm, wl, hl = C.c_uint32(), C.c_uint32(), C.c_uint32()
        off = np.zeros(14, dtype=np.uint32)
        self.lib.srref_texture_get(self.h, handle, None, C.byref(n), ptr(off), C.byref(nm), C.byref(wl), C.byref(hl))
        tex = np.zeros(n.value, dtype=np.uint8)
        self.lib.srref_texture_get(self.h, handle, ptr(tex), None, None, None, None, None)
        return TiledTexture(tex, off, nm.value, wl.value, hl.value)

    def load_scene(self, scene):
        self.tex_handles = [self.create_texture(t) for t in scene.textures]
        if getattr(scene, "sponza", None) is not None:
            # the reference keeps these in a file-static block (Viewer/SponzaScene.cpp:11): one set per process
            k = np.ascontiguousarray(scene.sponza, dtype=np.float32)
            self._keep.append(k)
            self.lib.srref_set_sponza_constants(ptr(k))
        self.descs = make_draw_descs(scene, self.tex_handles, self._keep)
        self.n_draws = len(scene.draws)
        self.clear_color = scene.clear_color

    # -- frames
    def render(self, clear=True, draws=None, clear_colour=None, clear_depth=None):
        self.lib.srref_begin_frame(self.h)
        cc = clear if clear_colour is None else clear_colour
        cd = clear if clear_depth is None else clear_depth
        if cc or cd:
            self.lib.srref_clear(self.h, self.clear_color, int(cc), int(cd))
        for i in (range(self.n_draws) if draws is None else draws):
            rc = self.lib.srref_draw_indexed(self.h, C.byref(self.descs[i]))
            assert rc == 0, rc
        self.lib.srref_end_frame(self.h)

    def render_frames(self, frames: int, mvps: np.ndarray | None = None) -> np.ndarray:
        ms = np.zeros(frames, dtype=np.float64)
        if mvps is not None:
            mvps = np.ascontiguousarray(mvps, dtype=np.float32)
            assert mvps.shape == (frames, self.n_draws, 16)
        rc = self.lib.srref_render_frames(self.h, self.descs, self.n_draws, ptr(mvps), frames, self.clear_color, ptr(ms))
        assert rc == 0, rc
        return ms

    def read_tiles(self):
        colour = np.zeros((self.num_tiles, 64, 64), dtype=np.uint32)
        depth = np.zeros((self.num_tiles, 64, 64), dtype=np.float32)
        self.lib.srref_read_tiles(self.h, ptr(colour), ptr(depth), 16384)
        return colour, depth

    def blit_linear(self) -> np.ndarray:
        px = np.zeros((self.height, self.width), dtype=np.uint32)
        self.lib.srref_blit_linear(self.h, ptr(px))
        return px

    # -- parity dumps
    def tile_counts(self) -> np.ndarray:
        out = np.zeros(self.num_tiles, dtype=np.uint32)
        rc = self.lib.srref_dump_tile_counts(self.h, ptr(out), self.num_tiles)
        assert rc == 0
        return out

    def tile_tris(self, tile: int, count: int) -> np.ndarray:
        out = np.zeros(max(1, count), dtype=TILE_TRI_DTYPE)
        n = C.c_uint32()
        rc = self.lib.srref_dump_tile_tris(self.h, tile, ptr(out), out.size, C.byref(n))
        assert rc == 0 and n.value == count, (rc, n.value, count)
        return out[:count]

    def tile_coverage(self, tile: int, count: int) -> np.ndarray:
        out = np.zeros((max(1, count), 64), dtype=np.uint64)
        n = C.c_uint32()
        rc = self.lib.srref_dump_tile_coverage(self.h, tile, ptr(out), out.shape[0], C.byref(n))
        assert rc == 0 and n.value == count
        return out[:count]

    def tile_fragments(self, tile: int, cap: int = 1 << 22):
        out = np.zeros(cap, dtype=np.uint32)
        depth = np.zeros((64, 64), dtype=np.float32)
        n = C.c_uint64()
        rc = self.lib.srref_dump_tile_fragments(self.h, tile, ptr(out), cap, C.byref(n), ptr(depth))
        assert rc == 0
        return out[: n.value], depth

    def sample(self, tex_handle, u, v, dudx, dudy, dvdx, dvdy) -> np.ndarray:
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (u, v, dudx, dudy, dvdx, dvdy)]
        n = arrs[0].size
        assert n % 8 == 0
        out = np.zeros(n, dtype=np.uint32)
        rc = self.lib.srref_sample(self.h, tex_handle, *[ptr(a) for a in arrs], ptr(out), n)
        assert rc == 0
        return out


def detile(tiles: np.ndarray, width: int, height: int) -> np.ndarray:
    """(num_tiles, 64, 64) -> (height, width) like BlitJobFn (Renderer.cpp:319-347)."""
    tx, ty = (width + 63) // 64, (height + 63) // 64
    img = tiles.reshape(ty, tx, 64, 64).transpose(0, 2, 1, 3).reshape(ty * 64, tx * 64)
    return np.ascontiguousarray(img[:height, :width])


# ---------------------------------------------------------------------------------------------------------------
# the plain-C restatement (oracle/sr_oracle.c -> oracle/_build/libsr_oracle.so)
# ---------------------------------------------------------------------------------------------------------------
PORT_PATH = os.path.join(_HERE, "_build", "libsr_oracle.so")
_port = None


def port_available() -> bool:
    return os.path.exists(PORT_PATH)


def _load_port():
    global _port
    if _port is not None:
        return _port
    lib = C.CDLL(PORT_PATH, mode=os.RTLD_LOCAL)
    vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
    lib.sro_create.argtypes = [u32, u32]
    lib.sro_create.restype = vp
    lib.sro_destroy.argtypes = [vp]
    lib.sro_destroy.restype = None
    lib.sro_set_rcp_table.argtypes = [vp, vp, u32]
    lib.sro_texture_create.argtypes = [vp, vp, u64, vp, u32, u32, u32]
    lib.sro_texture_create.restype = u64
    lib.sro_begin_frame.argtypes = [vp]
    lib.sro_clear.argtypes = [vp, u32, C.c_int, C.c_int]
    lib.sro_draw_indexed.argtypes = [vp, C.POINTER(DrawDesc)]
    lib.sro_end_frame.argtypes = [vp]
    lib.sro_render_frames.argtypes = [vp, C.POINTER(DrawDesc), u32, vp, u32, u32, vp]
    lib.sro_read_tiles.argtypes = [vp, vp, vp, u64]
    lib.sro_dump_tile_counts.argtypes = [vp, vp, u32]
    lib.sro_dump_tile_tris.argtypes = [vp, u32, vp, u32, C.POINTER(u32)]
    lib.sro_dump_tile_coverage.argtypes = [vp, u32, vp, u32, C.POINTER(u32)]
    lib.sro_dump_tile_fragments.argtypes = [vp, u32, vp, u64, C.POINTER(u64), vp]
    lib.sro_sample.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp, vp, u64]
    lib.sro_rcp.argtypes = [vp, vp, vp, u64]
    lib.sro_rcp.restype = None
    lib.sro_rsqrt.argtypes = [vp, vp, vp, u64]
    lib.sro_rsqrt.restype = None
    lib.sro_set_rsqrt_table.argtypes = [vp, vp, u32]
    lib.sro_set_sponza_constants.argtypes = [vp, vp]
    _port = lib
    return lib


class PortRenderer(RefRenderer):
    """Same interface as RefRenderer, backed by the plain-C restatement.  `rcp` = (table, bits): the RCPPS table to
    replay (harvested from the host CPU, or taken from a golden fixture); `rsqrt` likewise for RSQRTPS (only the Sponza
    shader needs it; harvested from the host on demand)."""

    def __init__(self, width: int, height: int, rcp, rsqrt=None):
        self.lib = _load_port()
        self.h = C.c_void_p(self.lib.sro_create(width, height))
        self.width, self.height = width, height
        self.tiles_x, self.tiles_y = (width + 63) // 64, (height + 63) // 64
        self.num_tiles = self.tiles_x * self.tiles_y
        self._keep = []
        table, bits = rcp
        table = np.ascontiguousarray(table, dtype=np.uint32)
        assert table.size == 1 << bits
        assert self.lib.sro_set_rcp_table(self.h, ptr(table), bits) == 0
        self._has_rsqrt = False
        if rsqrt is not None:
            self.set_rsqrt_table(*rsqrt)

    threads = 1

    def set_rsqrt_table(self, table, bits):
        table = np.ascontiguousarray(table, dtype=np.uint32)
        assert table.size == 2 << bits
        assert self.lib.sro_set_rsqrt_table(self.h, ptr(table), bits) == 0
        self._has_rsqrt = True

    def load_scene(self, scene):
        if getattr(scene, "sponza", None) is not None:
            if not self._has_rsqrt:
                if ref_available():
                    self.set_rsqrt_table(harvest_rsqrt_table(10), 10)
                else:  # no compiled reference here: the product library's harvest runs the same host instruction
                    from softrast_b200.capi import harvest_rsqrt_table as product_harvest

                    self.set_rsqrt_table(*product_harvest(16))
            k = np.ascontiguousarray(scene.sponza, dtype=np.float32)
            assert self.lib.sro_set_sponza_constants(self.h, ptr(k)) == 0
        self.tex_handles = [self.create_texture(t) for t in scene.textures]
        self.descs = make_draw_descs(scene, self.tex_handles, self._keep)
        self.n_draws = len(scene.draws)
        self.clear_color = scene.clear_color

    def rsqrt(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        self.lib.sro_rsqrt(self.h, ptr(x), ptr(out), x.size)
        return out

    def close(self):
        if self.h:
            self.lib.sro_destroy(self.h)
            self.h = C.c_void_p()

    def create_texture(self, t) -> int:
        off = np.ascontiguousarray(t.mip_offsets, dtype=np.uint32)
        return int(
            self.lib.sro_texture_create(self.h, ptr(t.texels), t.texels.size, ptr(off), t.num_mips, t.width_log2, t.height_log2)
        )

    def render(self, clear=True, draws=None, clear_colour=None, clear_depth=None):
        self.lib.sro_begin_frame(self.h)
        cc = clear if clear_colour is None else clear_colour
        cd = clear if clear_depth is None else clear_depth
        if cc or cd:
            self.lib.sro_clear(self.h, self.clear_color, int(cc), int(cd))
        for i in (range(self.n_draws) if draws is None else draws):
            assert self.lib.sro_draw_indexed(self.h, C.byref(self.descs[i])) == 0
        self.lib.sro_end_frame(self.h)

    def render_frames(self, frames, mvps=None):
        ms = np.zeros(frames, dtype=np.float64)
        if mvps is not None:
            mvps = np.ascontiguousarray(mvps, dtype=np.float32)
        assert self.lib.sro_render_frames(self.h, self.descs, self.n_draws, ptr(mvps), frames, self.clear_color, ptr(ms)) == 0
        return ms

    def read_tiles(self):
        colour = np.zeros((self.num_tiles, 64, 64), dtype=np.uint32)
        depth = np.zeros((self.num_tiles, 64, 64), dtype=np.float32)
        self.lib.sro_read_tiles(self.h, ptr(colour), ptr(depth), 16384)
        return colour, depth

    def blit_linear(self):
        return detile(self.read_tiles()[0], self.width, self.height)

    def tile_counts(self):
        out = np.zeros(self.num_tiles, dtype=np.uint32)
        assert self.lib.sro_dump_tile_counts(self.h, ptr(out), self.num_tiles) == 0
        return out

    def tile_tris(self, tile, count):
        out = np.zeros(max(1, count), dtype=TILE_TRI_DTYPE)
        n = C.c_uint32()
        rc = self.lib.sro_dump_tile_tris(self.h, tile, ptr(out), out.size, C.byref(n))
        assert rc == 0 and n.value == count
        return out[:count]

    def tile_coverage(self, tile, count):
        out = np.zeros((max(1, count), 64), dtype=np.uint64)
        n = C.c_uint32()
        rc = self.lib.sro_dump_tile_coverage(self.h, tile, ptr(out), out.shape[0], C.byref(n))
        assert rc == 0 and n.value == count
        return out[:count]

    def tile_fragments(self, tile, cap=1 << 22):
        out = np.zeros(cap, dtype=np.uint32)
        depth = np.zeros((64, 64), dtype=np.float32)
        n = C.c_uint64()
        assert self.lib.sro_dump_tile_fragments(self.h, tile, ptr(out), cap, C.byref(n), ptr(depth)) == 0
        return out[: n.value], depth

    def sample(self, tex_handle, u, v, dudx, dudy, dvdx, dvdy):
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (u, v, dudx, dudy, dvdx, dvdy)]
        out = np.zeros(arrs[0].size, dtype=np.uint32)
        assert self.lib.sro_sample(self.h, tex_handle, *[ptr(a) for a in arrs], ptr(out), out.size) == 0
        return out

    def rcp(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        self.lib.sro_rcp(self.h, ptr(x), ptr(out), x.size)
        return out


class RefSponzaScene:
    """The reference's own SponzaScene (Viewer/SponzaScene.cpp:105-215) on an empty model: Init seeds the lights, every
    update(dt) runs SponzaScene::Update and returns the constants block (float32[136] = srb_sponza_constants)."""

    def __init__(self, renderer: "RefRenderer"):
        self.lib = renderer.lib
        self.ctx, self.fb = C.c_void_p(), C.c_void_p()
        self.lib.srref_raw_objects(renderer.h, C.byref(self.ctx), C.byref(self.fb))
        self.h = C.c_void_p(self.lib.srref_sponza_scene_create())

    @property
    def constants(self) -> np.ndarray:
        k = np.zeros(136, dtype=np.float32)
        self.lib.srref_get_sponza_constants(ptr(k))
        return k

    def update(self, dt: float) -> np.ndarray:
        self.lib.srref_sponza_scene_update(self.h, self.ctx, self.fb, C.c_float(dt))
        return self.constants

    def close(self):
        if self.h:
            self.lib.srref_sponza_scene_destroy(self.h)
            self.h = C.c_void_p()
