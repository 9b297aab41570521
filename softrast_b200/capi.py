"""ctypes binding of include/softrast_b200.h + a small host-side mirror of the reference's RenderContext flow.

Nothing here computes pixels: every call goes through the C ABI into the CUDA library.  If the library is missing this
module raises at import (no fallback of any kind)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._ctypes_defs import (
    TILE_TRI_DTYPE,
    BufferRef,
    Counters,
    DrawDesc,
    ptr,
    MaterialView,
    MeshView,
    copy_material_view,
    copy_mesh_view,
)

_HERE = os.path.dirname(os.path.abspath(__file__))
# SRB_LIB: another build of the same library (profiles/stats.py loads the statistics build); experiments only
LIB_PATH = os.environ.get("SRB_LIB") or os.path.join(_HERE, "lib", "libsoftrast_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(make -C softrast_b200/csrc). There is no CPU fallback."
    )

lib = C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL)

_vp, _u32, _u64, _int = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
_H = C.POINTER(_u64)

# every symbol include/softrast_b200.h declares (tests/test_abi.py checks this list against the header)
_SIGNATURES = {
    "srb_create": (_int, [_int, _u32, C.POINTER(_vp)]),
    "srb_create_shared": (_int, [_vp, _u32, C.POINTER(_vp)]),
    "srb_set_frames_in_flight_hint": (_int, [_vp, _u32]),
    "srb_destroy": (None, [_vp]),
    "srb_last_error": (C.c_char_p, [_vp]),
    "srb_version": (C.c_char_p, []),
    "srb_set_rcp_table": (_int, [_vp, _vp, _u32]),
    "srb_harvest_rcp_table": (_u32, [_vp, _u32]),
    "srb_set_rsqrt_table": (_int, [_vp, _vp, _u32]),
    "srb_harvest_rsqrt_table": (_u32, [_vp, _u32]),
    "srb_set_sponza_constants": (_int, [_vp, _vp]),
    "srb_texture_create": (_int, [_vp, _vp, _u64, _vp, _u32, _u32, _u32, _H]),
    "srb_texture_destroy": (_int, [_vp, _u64]),
    "srb_texture_build_rgba8": (_int, [_vp, _u32, _u32, _int, _vp, _H, _vp, C.POINTER(_u32)]),
    "srb_buffer_create": (_int, [_vp, _vp, _u64, _H]),
    "srb_buffer_update": (_int, [_vp, _u64, _u64, _vp, _u64]),
    "srb_buffer_destroy": (_int, [_vp, _u64]),
    "srb_invalidate_host": (_int, [_vp, _vp]),
    "srb_framebuffer_create": (_int, [_vp, _u32, _u32, _H]),
    "srb_framebuffer_destroy": (_int, [_vp, _u64]),
    "srb_framebuffer_info": (_int, [_vp, _u64] + [C.POINTER(_u32)] * 4),
    "srb_framebuffer_export": (_int, [_vp, _u64, _vp]),
    "srb_framebuffer_import": (_int, [_vp, _vp, _u32, _u32, _H]),
    "srb_set_tile_ownership": (_int, [_vp, _u32, _u32]),
    "srb_begin_frame": (_int, [_vp]),
    "srb_clear": (_int, [_vp, _u64, _u32, _int, _int]),
    "srb_draw_indexed": (_int, [_vp, C.POINTER(DrawDesc)]),
    "srb_end_frame": (_int, [_vp]),
    "srb_end_frame_async": (_int, [_vp]),
    "srb_sync": (_int, [_vp]),
    "srb_read_tiles": (_int, [_vp, _u64, _vp, _vp, _u64]),
    "srb_blit_linear": (_int, [_vp, _u64, _vp, _vp, _vp]),
    "srb_get_counters": (_int, [_vp, C.POINTER(Counters)]),
    "srb_get_kernel_times": (_int, [_vp, _vp, _vp, _u32, C.POINTER(_u32)]),
    "srb_set_timing": (_int, [_vp, _int]),
    "srb_launch_count": (_u64, [_vp]),
    "srb_dump_tile_counts": (_int, [_vp, _vp, _u32]),
    "srb_dump_tile_tris": (_int, [_vp, _u32, _vp, _u32, C.POINTER(_u32)]),
    "srb_dump_tile_ranks": (_int, [_vp, _u32, _vp, _u32, C.POINTER(_u32)]),
    "srb_dump_tile_coverage": (_int, [_vp, _u32, _vp, _u32, C.POINTER(_u32)]),
    "srb_dump_winners": (_int, [_vp, _vp, _u64]),
    "srb_render_frames": (_int, [_vp, _u32, _u32, _vp, _u32, _u32, _vp, _u64]),
    "srb_host_alloc": (_vp, [_u64]),
    "srb_host_alloc_ex": (_vp, [_u64, _u32]),
    "srb_debug_d2h_copies": (_int, [_vp, _vp, _u64, _u32, C.POINTER(C.c_float)]),
    "srb_host_free": (None, [_vp]),
    "srb_flush_l2": (_int, [_vp, _u64]),
    "srb_timer_mark": (_int, [_vp, _u32]),
    "srb_timer_elapsed": (_int, [_vp, _u32, _vp, _u32, C.POINTER(C.c_float)]),
    "srb_debug_sample": (_int, [_vp, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u32]),
    "srb_debug_rcp": (_int, [_vp, _vp, _vp, _u32]),
    "srb_debug_rsqrt": (_int, [_vp, _vp, _vp, _u32]),
    "srb_sponza_scene_init": (None, [_vp]),
    "srb_sponza_scene_update": (None, [_vp, C.c_float]),
    "srb_texture_create_rgba8": (_int, [_vp, _vp, _u32, _u32, _int, C.POINTER(_u64)]),
    "srb_texture_read": (_int, [_vp, _u64, _vp, _u64, C.POINTER(_u64), _vp, C.POINTER(_u32), C.POINTER(_u32), C.POINTER(_u32)]),
    "srb_model_load": (_int, [C.c_char_p, _u32, C.POINTER(_vp)]),
    "srb_model_load_ex": (_int, [C.c_char_p, _u32, _vp, _vp, C.POINTER(_vp)]),
    "srb_model_load_on": (_int, [_vp, C.c_char_p, _u32, _vp, _vp, C.POINTER(_vp)]),
    "srb_model_free": (None, [_vp]),
    "srb_model_last_error": (C.c_char_p, []),
    "srb_model_info": (_int, [_vp, C.POINTER(_u32), C.POINTER(_u32), C.POINTER(_int)]),
    "srb_model_mesh": (_int, [_vp, _u32, C.POINTER(MeshView)]),
    "srb_model_material": (_int, [_vp, _u32, C.POINTER(MaterialView)]),
    "srb_model_save_cache": (_int, [_vp, C.c_char_p]),
    "srb_image_load_rgba8": (_int, [C.c_char_p, C.POINTER(_vp), C.POINTER(_u32), C.POINTER(_u32)]),
    "srb_image_free": (None, [_vp]),
    "srb_model_make_resident": (_int, [_vp, _vp, C.POINTER(_vp)]),
    "srb_resident_model_free": (None, [_vp]),
    "srb_resident_model_draws": (_int, [_vp, _u64, _vp, _u32, _vp, _u32, C.POINTER(_u32)]),
}
for _name, (_res, _args) in _SIGNATURES.items():
    _f = getattr(lib, _name)  # AttributeError here == the library does not export what the header declares
    _f.restype = _res
    _f.argtypes = _args


FLAG_UPLOAD_ALWAYS = 1  # SRB_FLAG_UPLOAD_ALWAYS
FLAG_FULL_RECORDS = 2  # SRB_FLAG_FULL_RECORDS: set up every varying's plane (needed by tile_tris dumps)


class SrbError(RuntimeError):
    pass


class BatchItem(C.Structure):
    _fields_ = [("ctx", _vp), ("draws", C.POINTER(DrawDesc))]


def render_frames(renderers, frames: int, mvps=None, colour_out=None, colour_stride: int = 0):
    """srb_render_frames over one or more SceneRenderers holding the same scene (frames in flight = len(renderers)).
    colour_out: integer address of host memory (e.g. from host_alloc) or None."""
    items = (BatchItem * len(renderers))()
    for i, r in enumerate(renderers):
        items[i].ctx = r.ctx.h
        items[i].draws = C.cast(r.descs, C.POINTER(DrawDesc))
    n_draws = renderers[0].n_draws
    if mvps is not None:
        mvps = np.ascontiguousarray(mvps, dtype=np.float32)
        assert mvps.shape == (frames, n_draws, 16)
    rc = lib.srb_render_frames(
        items, len(renderers), n_draws, ptr(mvps), frames, renderers[0].scene.clear_color, colour_out, colour_stride
    )
    if rc != 0:
        msgs = [lib.srb_last_error(r.ctx.h).decode() for r in renderers]
        raise SrbError(f"srb_render_frames failed ({rc}): {msgs}")


def timer_mark(renderers, slot: int):
    for r in renderers:
        r.ctx._check(lib.srb_timer_mark(r.ctx.h, slot), "srb_timer_mark")


def timer_elapsed_ms(renderers, slot_a: int, slot_b: int) -> float:
    """Device milliseconds from the first context's mark `slot_a` to the latest context's mark `slot_b`."""
    best = 0.0
    for r in renderers:
        ms = C.c_float()
        r.ctx._check(
            lib.srb_timer_elapsed(renderers[0].ctx.h, slot_a, r.ctx.h, slot_b, C.byref(ms)), "srb_timer_elapsed"
        )
        best = max(best, float(ms.value))
    return best


def host_alloc(nbytes: int) -> int:
    p = lib.srb_host_alloc(nbytes)
    if not p:
        raise SrbError("srb_host_alloc failed")
    return p


HOST_WRITE_COMBINED, HOST_PORTABLE, HOST_HUGE_PAGES = 1, 2, 4  # SRB_HOST_*


def host_alloc_ex(nbytes: int, flags: int) -> int:
    p = lib.srb_host_alloc_ex(nbytes, flags)
    if not p:
        raise SrbError("srb_host_alloc_ex failed")
    return p


def host_free(p: int):
    lib.srb_host_free(p)


MIPS_NONE, MIPS_BOX, MIPS_STB = 0, 1, 2  # SRB_MIPS_*


def build_texture(rgba: np.ndarray, calc_mips=True):
    """srb_texture_build_rgba8 -> scenes.TiledTexture (host side, no GPU needed).  calc_mips: False / MIPS_NONE, True /
    MIPS_BOX (2x2 box filter) or MIPS_STB (the reference's own mips, stb_image_resize's Mitchell down-sampling)."""
    from .scenes import TiledTexture

    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w, _ = rgba.shape
    n = _u64()
    nm = _u32()
    off = np.zeros(14, dtype=np.uint32)
    rc = lib.srb_texture_build_rgba8(None, w, h, int(calc_mips), None, C.byref(n), ptr(off), C.byref(nm))
    if rc != 0:
        raise SrbError(f"srb_texture_build_rgba8: {rc}")
    tex = np.zeros(n.value, dtype=np.uint8)
    rc = lib.srb_texture_build_rgba8(ptr(rgba), w, h, int(calc_mips), ptr(tex), C.byref(n), ptr(off), C.byref(nm))
    if rc != 0:
        raise SrbError(f"srb_texture_build_rgba8: {rc}")
    return TiledTexture(tex, off, nm.value, w.bit_length() - 1, h.bit_length() - 1)


def harvest_rsqrt_table(max_bits: int = 16):
    table = np.zeros(2 << max_bits, dtype=np.uint32)
    bits = lib.srb_harvest_rsqrt_table(ptr(table), max_bits)
    return table[: 2 << bits].copy(), int(bits)


def harvest_rcp_table(max_bits: int = 16):
    table = np.zeros(1 << max_bits, dtype=np.uint32)
    bits = lib.srb_harvest_rcp_table(ptr(table), max_bits)
    return table[: 1 << bits].copy(), int(bits)


class RenderContext:
    """Host-side mirror of sr::RenderContext (reference SoftRast/Renderer.h:153-177) over the C ABI: BeginFrame,
    ClearFrameBuffer, DrawIndexed, EndFrame, Blit — same names, same call order, same meaning."""

    def __init__(self, device: int = 0, flags: int = 0, share: "RenderContext | None" = None):
        """share: another context of the same device whose textures / buffers this one shares (srb_create_shared)."""
        self.h = _vp()
        if share is not None:
            rc = lib.srb_create_shared(share.h, flags, C.byref(self.h))
        else:
            rc = lib.srb_create(device, flags, C.byref(self.h))
        if rc != 0:
            msg = lib.srb_last_error(self.h).decode() if self.h else "no CUDA device (there is no CPU fallback)"
            if self.h:
                lib.srb_destroy(self.h)
            self.h = _vp()
            raise SrbError(f"srb_create failed ({rc}): {msg}")
        self._keep = []

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise SrbError(f"{what} failed ({rc}): {lib.srb_last_error(self.h).decode()}")

    def close(self):
        if self.h:
            lib.srb_destroy(self.h)
            self.h = _vp()

    Shutdown = close

    # -- resources
    def set_sponza_constants(self, k: np.ndarray):
        """k: float32[136] = srb_sponza_constants (scenes.sponza_constants)."""
        k = np.ascontiguousarray(k, dtype=np.float32)
        assert k.size == 136
        self._check(lib.srb_set_sponza_constants(self.h, ptr(k)), "srb_set_sponza_constants")

    def set_rsqrt_table(self, table: np.ndarray, bits: int):
        table = np.ascontiguousarray(table, dtype=np.uint32)
        assert table.size == 2 << bits
        self._check(lib.srb_set_rsqrt_table(self.h, ptr(table), bits), "srb_set_rsqrt_table")

    def set_rcp_table(self, table: np.ndarray, bits: int):
        table = np.ascontiguousarray(table, dtype=np.uint32)
        assert table.size == 1 << bits
        self._check(lib.srb_set_rcp_table(self.h, ptr(table), bits), "srb_set_rcp_table")

    def create_texture(self, t) -> int:
        out = _u64()
        off = np.ascontiguousarray(t.mip_offsets, dtype=np.uint32)
        texels = np.ascontiguousarray(t.texels, dtype=np.uint8)
        self._check(
            lib.srb_texture_create(
                self.h, ptr(texels), texels.size, ptr(off), t.num_mips, t.width_log2, t.height_log2, C.byref(out)
            ),
            "srb_texture_create",
        )
        return int(out.value)

    def create_texture_rgba8(self, rgba: np.ndarray, calc_mips: int = 2) -> int:
        """srb_texture_create_rgba8: TextureData::CreateFromRGBA8 on the device (tiling + the reference's stb mips)."""
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        h, w, _ = rgba.shape
        out = _u64()
        self._check(lib.srb_texture_create_rgba8(self.h, ptr(rgba), w, h, int(calc_mips), C.byref(out)), "srb_texture_create_rgba8")
        return int(out.value)

    def read_texture(self, handle: int):
        """srb_texture_read -> scenes.TiledTexture with the device copy's bytes."""
        from .scenes import TiledTexture

        n, nm, wl, hl = _u64(), _u32(), _u32(), _u32()
        off = np.zeros(14, dtype=np.uint32)
        self._check(lib.srb_texture_read(self.h, handle, None, 0, C.byref(n), ptr(off), C.byref(nm), C.byref(wl), C.byref(hl)),
                    "srb_texture_read")
        tex = np.zeros(n.value, dtype=np.uint8)
        self._check(lib.srb_texture_read(self.h, handle, ptr(tex), tex.size, None, None, None, None, None), "srb_texture_read")
        return TiledTexture(tex, off, nm.value, wl.value, hl.value)

    def create_buffer(self, a: np.ndarray) -> int:
        a = np.ascontiguousarray(a)
        out = _u64()
        self._check(lib.srb_buffer_create(self.h, ptr(a), a.nbytes, C.byref(out)), "srb_buffer_create")
        return int(out.value)

    def create_framebuffer(self, width: int, height: int) -> "FrameBuffer":
        out = _u64()
        self._check(lib.srb_framebuffer_create(self.h, width, height, C.byref(out)), "srb_framebuffer_create")
        return FrameBuffer(self, int(out.value), width, height)

    def export_framebuffer(self, fb: "FrameBuffer") -> bytes:
        buf = C.create_string_buffer(128)
        self._check(lib.srb_framebuffer_export(self.h, fb.handle, buf), "srb_framebuffer_export")
        return buf.raw

    def import_framebuffer(self, blob: bytes, width: int, height: int) -> "FrameBuffer":
        out = _u64()
        buf = C.create_string_buffer(blob, 128)
        self._check(lib.srb_framebuffer_import(self.h, buf, width, height, C.byref(out)), "srb_framebuffer_import")
        return FrameBuffer(self, int(out.value), width, height)

    def set_tile_ownership(self, modulus: int, remainder: int):
        self._check(lib.srb_set_tile_ownership(self.h, modulus, remainder), "srb_set_tile_ownership")

    # -- frame (names follow the reference)
    def BeginFrame(self):
        self._check(lib.srb_begin_frame(self.h), "srb_begin_frame")

    def ClearFrameBuffer(self, fb: "FrameBuffer", color: int = 0, clear_colour=True, clear_depth=True):
        self._check(lib.srb_clear(self.h, fb.handle, color, int(clear_colour), int(clear_depth)), "srb_clear")

    def DrawIndexed(self, desc: DrawDesc):
        self._check(lib.srb_draw_indexed(self.h, C.byref(desc)), "srb_draw_indexed")

    def EndFrame(self, sync: bool = True):
        if sync:
            self._check(lib.srb_end_frame(self.h), "srb_end_frame")
        else:
            self._check(lib.srb_end_frame_async(self.h), "srb_end_frame_async")

    def Sync(self):
        self._check(lib.srb_sync(self.h), "srb_sync")

    def Blit(self, fb: "FrameBuffer", linear: np.ndarray):
        """RenderContext::Blit; returns after the pixels are in `linear` (the C ABI itself is asynchronous)."""
        assert linear.nbytes == fb.width * fb.height * 4 and linear.flags["C_CONTIGUOUS"]
        self._check(lib.srb_blit_linear(self.h, fb.handle, ptr(linear), None, None), "srb_blit_linear")
        self.Sync()

    # -- results / introspection
    def counters(self) -> dict:
        c = Counters()
        self._check(lib.srb_get_counters(self.h, C.byref(c)), "srb_get_counters")
        return c.as_dict()

    def set_timing(self, on: bool):
        lib.srb_set_timing(self.h, int(on))

    def kernel_times(self) -> dict:
        us = (C.c_float * 8)()
        names = (C.c_char_p * 8)()
        n = _u32()
        self._check(lib.srb_get_kernel_times(self.h, us, names, 8, C.byref(n)), "srb_get_kernel_times")
        return {names[i].decode(): float(us[i]) for i in range(n.value)}

    def launch_count(self) -> int:
        return int(lib.srb_launch_count(self.h))

    def tile_counts(self, num_tiles: int) -> np.ndarray:
        out = np.zeros(num_tiles, dtype=np.uint32)
        self._check(lib.srb_dump_tile_counts(self.h, ptr(out), num_tiles), "srb_dump_tile_counts")
        return out

    def tile_tris(self, tile: int, count: int) -> np.ndarray:
        out = np.zeros(max(1, count), dtype=TILE_TRI_DTYPE)
        n = _u32()
        self._check(lib.srb_dump_tile_tris(self.h, tile, ptr(out), out.size, C.byref(n)), "srb_dump_tile_tris")
        assert n.value == count, (n.value, count)
        return out[:count]

    def tile_ranks(self, tile: int, count: int) -> np.ndarray:
        out = np.zeros(max(1, count), dtype=np.uint32)
        n = _u32()
        self._check(lib.srb_dump_tile_ranks(self.h, tile, ptr(out), out.size, C.byref(n)), "srb_dump_tile_ranks")
        return out[: n.value]

    def tile_coverage(self, tile: int, count: int) -> np.ndarray:
        out = np.zeros((max(1, count), 64), dtype=np.uint64)
        n = _u32()
        self._check(
            lib.srb_dump_tile_coverage(self.h, tile, ptr(out), out.shape[0], C.byref(n)), "srb_dump_tile_coverage"
        )
        assert n.value == count
        return out[:count]

    def winners(self, num_tiles: int) -> np.ndarray:
        out = np.zeros((num_tiles, 64, 64), dtype=np.uint32)
        self._check(lib.srb_dump_winners(self.h, ptr(out), out.size), "srb_dump_winners")
        return out

    def set_frames_in_flight_hint(self, n: int):
        self._check(lib.srb_set_frames_in_flight_hint(self.h, n), "srb_set_frames_in_flight_hint")

    def flush_l2(self, nbytes: int = 256 << 20):
        self._check(lib.srb_flush_l2(self.h, nbytes), "srb_flush_l2")

    def debug_sample(self, tex: int, u, v, dudx, dudy, dvdx, dvdy) -> np.ndarray:
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (u, v, dudx, dudy, dvdx, dvdy)]
        out = np.zeros(arrs[0].size, dtype=np.uint32)
        self._check(lib.srb_debug_sample(self.h, tex, *[ptr(a) for a in arrs], ptr(out), out.size), "srb_debug_sample")
        return out

    def debug_rcp(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        self._check(lib.srb_debug_rcp(self.h, ptr(x), ptr(out), x.size), "srb_debug_rcp")
        return out

    def debug_rsqrt(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        self._check(lib.srb_debug_rsqrt(self.h, ptr(x), ptr(out), x.size), "srb_debug_rsqrt")
        return out


class FrameBuffer:
    """sr::FrameBuffer (reference Renderer.h:75-108): 64x64 tiles, two planes, device resident."""

    def __init__(self, ctx: RenderContext, handle: int, width: int, height: int):
        self.ctx, self.handle, self.width, self.height = ctx, handle, width, height
        self.tiles_x, self.tiles_y = (width + 63) // 64, (height + 63) // 64
        self.num_tiles = self.tiles_x * self.tiles_y

    def read_tiles(self):
        colour = np.zeros((self.num_tiles, 64, 64), dtype=np.uint32)
        depth = np.zeros((self.num_tiles, 64, 64), dtype=np.float32)
        self.ctx._check(
            lib.srb_read_tiles(self.ctx.h, self.handle, ptr(colour), ptr(depth), 16384), "srb_read_tiles"
        )
        return colour, depth


class SceneRenderer:
    """Drives a scenes.Scene through the C ABI exactly the way the reference's Scene::Update + main loop do
    (Viewer/Scene.cpp:32-65, Viewer/Main.cpp:50-69).  `resident=True` creates device buffers once (srb_buffer_create);
    otherwise draws carry host pointers and the library mirrors them."""

    def __init__(self, scene, device: int = 0, resident: bool = True, flags: int = 0, rcp=None, fb_import: bytes | None = None,
                 rsqrt=None, share: "SceneRenderer | None" = None):
        """share: a SceneRenderer of the same scene whose device-resident textures and buffers this one uses (one copy of
        the scene for all frames in flight); it gets its own framebuffer, frame state and stream."""
        self.scene = scene
        self.ctx = RenderContext(device, flags, share=share.ctx if share is not None else None)
        if rcp is not None:
            self.ctx.set_rcp_table(*rcp)
        if rsqrt is not None:
            self.ctx.set_rsqrt_table(*rsqrt)
        if fb_import is not None:  # screen-tile split: draw into another process's framebuffer (CUDA IPC)
            self.fb = self.ctx.import_framebuffer(fb_import, scene.width, scene.height)
        else:
            self.fb = self.ctx.create_framebuffer(scene.width, scene.height)
        if share is not None:
            assert share.scene is scene and (share.buf_handles is not None) == resident
            self.tex_handles = share.tex_handles
        else:
            self.tex_handles = [self.ctx.create_texture(t) for t in scene.textures]
        self.buf_handles = [] if resident else None  # (vertex buffer, index buffer) per draw
        if getattr(scene, "sponza", None) is not None:
            self.ctx.set_sponza_constants(scene.sponza)
        self.descs = (DrawDesc * max(1, len(scene.draws)))()
        self._keep = []
        for i, d in enumerate(scene.draws):
            v = np.ascontiguousarray(d.vertices, dtype=np.float32)
            idx = np.ascontiguousarray(d.indices)
            self._keep += [v, idx]
            e = self.descs[i]
            e.shader, e.uv_offset = d.shader, d.uv_offset
            e.texture = self.tex_handles[d.texture] if d.texture >= 0 else 0
            e.framebuffer = self.fb.handle
            stride = v.shape[1] * 4
            if resident:
                vb, ib = share.buf_handles[i] if share is not None else (self.ctx.create_buffer(v), self.ctx.create_buffer(idx))
                self.buf_handles.append((vb, ib))
                e.indices = BufferRef(ib, 0, None, idx.dtype.itemsize, idx.size)
                e.positions = BufferRef(vb, 0, None, stride, v.shape[0])
                e.attributes = BufferRef(vb, 0, None, stride, v.shape[0])
            else:
                e.indices = BufferRef(0, 0, idx.ctypes.data, idx.dtype.itemsize, idx.size)
                e.positions = BufferRef(0, 0, v.ctypes.data, stride, v.shape[0])
                e.attributes = BufferRef(0, 0, v.ctypes.data, stride, v.shape[0])
            for k in range(16):
                e.mvp[k] = float(d.mvp[k])
        self.n_draws = len(scene.draws)

    def render(self, clear: bool = True, sync: bool = True, mvps: np.ndarray | None = None, draws=None,
               clear_colour: bool | None = None, clear_depth: bool | None = None):
        """One frame.  draws: indices of the scene's draws to submit (default: all, in order); clear_colour / clear_depth
        override `clear` separately (ClearFrameBuffer's _clearColour / _clearDepth, Renderer.cpp:168-194)."""
        c = self.ctx
        c.BeginFrame()
        cc = clear if clear_colour is None else clear_colour
        cd = clear if clear_depth is None else clear_depth
        if cc or cd:
            c.ClearFrameBuffer(self.fb, self.scene.clear_color, cc, cd)
        for i in (range(self.n_draws) if draws is None else draws):
            if mvps is not None:
                C.memmove(self.descs[i].mvp, mvps[i].ctypes.data, 64)
            c.DrawIndexed(self.descs[i])
        c.EndFrame(sync)

    def read_tiles(self):
        return self.fb.read_tiles()

    def blit_linear(self) -> np.ndarray:
        px = np.zeros((self.scene.height, self.scene.width), dtype=np.uint32)
        self.ctx.Blit(self.fb, px)
        return px

    def close(self):
        self.ctx.close()


OBJ_FLIP_WINDING, OBJ_GEN_NORMALS, OBJ_FLIP_UVS = 1, 2, 4  # sr::Obj::LoadFlags (Viewer/Obj.h:52-58)
OBJ_NO_CACHE_READ, OBJ_NO_CACHE_WRITE = 0x100, 0x200


def load_image_rgba8(path: str) -> np.ndarray:
    """srb_image_load_rgba8: what stbi_load(path, ..., 4) returns for PNG / TGA, as uint8 (h, w, 4)."""
    px, w, h = _vp(), _u32(), _u32()
    rc = lib.srb_image_load_rgba8(os.fsencode(path), C.byref(px), C.byref(w), C.byref(h))
    if rc != 0:
        raise SrbError(f"srb_image_load_rgba8 ({rc}): {lib.srb_model_last_error().decode()}")
    out = np.frombuffer(C.string_at(px.value, w.value * h.value * 4), dtype=np.uint8).reshape(h.value, w.value, 4).copy()
    lib.srb_image_free(px)
    return out


class Model:
    """Host-side mirror of sr::Obj::Model (reference Viewer/Obj.h:60-71): Load(path, flags) reads the `.bin` cache or
    parses the OBJ/MTL and writes the cache (Viewer/Obj.cpp:374-560).  `meshes` / `materials` are numpy copies of
    m_meshes / m_materials; `notes` holds non-fatal messages (missing MTL / texture)."""

    def __init__(self, path: str, flags: int = 0, ctx: "RenderContext | None" = None):
        """ctx: a render context whose device builds the materials' textures (srb_model_load_on) instead of the host."""
        self.h = _vp()
        if ctx is not None:
            rc = lib.srb_model_load_on(ctx.h, os.fsencode(path), flags, None, None, C.byref(self.h))
        else:
            rc = lib.srb_model_load(os.fsencode(path), flags, C.byref(self.h))
        if rc != 0:
            raise SrbError(f"srb_model_load ({rc}): {lib.srb_model_last_error().decode()}")
        self.notes = lib.srb_model_last_error().decode()
        nm, nmat, fc = _u32(), _u32(), _int()
        lib.srb_model_info(self.h, C.byref(nm), C.byref(nmat), C.byref(fc))
        self.from_cache = bool(fc.value)
        self.meshes, self.materials = [], []
        for i in range(nm.value):
            v = MeshView()
            assert lib.srb_model_mesh(self.h, i, C.byref(v)) == 0
            self.meshes.append(copy_mesh_view(v))
        for i in range(nmat.value):
            v = MaterialView()
            assert lib.srb_model_material(self.h, i, C.byref(v)) == 0
            self.materials.append(copy_material_view(v))

    def save_cache(self, bin_path: str):
        rc = lib.srb_model_save_cache(self.h, os.fsencode(bin_path))
        if rc != 0:
            raise SrbError(f"srb_model_save_cache ({rc}): {lib.srb_model_last_error().decode()}")

    def close(self):
        if self.h:
            lib.srb_model_free(self.h)
            self.h = _vp()

    def to_scene(self, width: int, height: int, mvp: np.ndarray, textured_shader: int = 0, clear_color: int = 0):
        """The draw list Viewer/Scene.cpp:35-63 issues for the model, as a scenes.Scene (one draw per mesh with indices)."""
        from .scenes import Draw, Scene, TiledTexture

        sc = Scene("obj_model", width, height, clear_color=clear_color)
        tex_of = {}
        for i, m in enumerate(self.materials):
            if m["texels"].size and m["num_mips"]:
                tex_of[i] = len(sc.textures)
                sc.textures.append(TiledTexture(m["texels"], m["mip_offsets"], m["num_mips"], m["width_log2"], m["height_log2"]))
        for m in self.meshes:
            if not m["indices"].size:
                continue
            has_mat = m["material"] < len(self.materials)
            shader = textured_shader if (has_mat or textured_shader == 3) else 1
            sc.draws.append(Draw(m["vertices"], m["indices"], np.asarray(mvp, np.float32), shader,
                                 tex_of.get(m["material"], -1) if has_mat else -1, 6))
        return sc


class ResidentModel:
    """srb_model_make_resident + srb_resident_model_draws: the model's buffers and textures on the context's device and
    its draw list (one draw per mesh)."""

    def __init__(self, ctx: "RenderContext", model: Model):
        self.ctx = ctx
        self.h = _vp()
        rc = lib.srb_model_make_resident(ctx.h, model.h, C.byref(self.h))
        if rc != 0:
            raise SrbError(f"srb_model_make_resident ({rc}): {lib.srb_model_last_error().decode()}")

    def draws(self, fb_handle: int, mvp: np.ndarray, textured_shader: int = 0):
        n = _u32()
        lib.srb_resident_model_draws(self.h, fb_handle, None, textured_shader, None, 0, C.byref(n))
        descs = (DrawDesc * max(1, n.value))()
        mvp = np.ascontiguousarray(mvp, dtype=np.float32)
        rc = lib.srb_resident_model_draws(self.h, fb_handle, ptr(mvp), textured_shader, descs, n.value, C.byref(n))
        if rc != 0:
            raise SrbError(f"srb_resident_model_draws ({rc}): {lib.srb_model_last_error().decode()}")
        return descs, n.value

    def close(self):
        if self.h:
            lib.srb_resident_model_free(self.h)
            self.h = _vp()


class SponzaSceneAnim:
    """srb_sponza_scene_*: the frame constants of the viewer's default scene over time (SponzaScene::Init + Update,
    Viewer/SponzaScene.cpp:126-160, :168-187).  `constants` is the float32[136] block set_sponza_constants takes."""

    _STATE_FLOATS = 1 + 16 * 16 + 136  # anim_phase, 16 x PointLightAnim (16 floats), srb_sponza_constants

    def __init__(self):
        self.state = np.zeros(self._STATE_FLOATS, dtype=np.float32)
        lib.srb_sponza_scene_init(ptr(self.state))

    def update(self, dt: float) -> np.ndarray:
        lib.srb_sponza_scene_update(ptr(self.state), C.c_float(dt))
        return self.constants

    @property
    def constants(self) -> np.ndarray:
        return self.state[1 + 16 * 16:].copy()
