"""The device-side texture builder (srb_texture_create_rgba8): TextureData::CreateFromRGBA8
(SoftRast/Texture.cpp:119-199) as CUDA kernels — tiling + stb_image_resize's down-sampled mips — against the host builder
(already pinned to the reference byte for byte) and against the reference itself."""
import time

import numpy as np
import pytest

from softrast_b200 import scenes

pytestmark = pytest.mark.gpu


def _spread5(v):
    out = np.zeros_like(v)
    for b in range(5):
        out |= ((v >> b) & 1) << (2 * b)
    return out


def _image(w, h, seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, 4)).astype(np.uint8)
    # smooth regions, saturated regions and hard edges as well as noise: rounding ties in the encode, clamping
    img[: h // 3] = (np.linspace(0, 255, w)[None, :, None] + np.arange(4)[None, None, :] * 3).astype(np.uint8)
    img[h // 3: h // 2, : w // 2] = 255
    img[h // 3: h // 2, w // 2:] = 0
    return img


@pytest.mark.parametrize("w,h", [(32, 32), (64, 128), (256, 256), (1024, 512), (32, 2048)])
def test_device_texture_builder_matches_host_and_reference(w, h):
    from oracle import refharness as rh
    from softrast_b200 import capi

    img = _image(w, h, w * 7 + h)
    ctx = capi.RenderContext(0)
    try:
        for mips in (capi.MIPS_STB, capi.MIPS_NONE):
            handle = ctx.create_texture_rgba8(img, mips)
            dev = ctx.read_texture(handle)
            host = capi.build_texture(img, mips)
            assert dev.num_mips == host.num_mips and dev.width_log2 == host.width_log2 and dev.height_log2 == host.height_log2
            assert np.array_equal(dev.mip_offsets, host.mip_offsets)
            assert dev.texels.size == host.texels.size
            bad = np.nonzero(dev.texels != host.texels)[0]
            assert bad.size == 0, f"{bad.size} bytes differ from the host builder, first at {bad[:8]}"
        if rh.ref_available():
            r = rh.RefRenderer(64, 64, 1, "parity")
            try:
                ref = r.get_texture(r.create_texture_rgba8(img, True))
            finally:
                r.close()
            handle = ctx.create_texture_rgba8(img, capi.MIPS_STB)
            dev = ctx.read_texture(handle)
            assert dev.num_mips == ref.num_mips and np.array_equal(dev.mip_offsets[: dev.num_mips], ref.mip_offsets[: ref.num_mips])
            W, H = 1 << dev.width_log2, 1 << dev.height_log2
            for k in range(dev.num_mips):  # the texels that exist (the reference leaves the padding uninitialised)
                mw, mh = max(1, W >> k), max(1, H >> k)
                y, x = np.mgrid[0:mh, 0:mw]
                idx = ((y >> 5) * ((mw + 31) // 32) + (x >> 5)) * 1024 + (_spread5(x) | (_spread5(y) << 1))
                offs = int(dev.mip_offsets[k]) + 4 * idx.reshape(-1)
                for c in range(4):
                    assert np.array_equal(dev.texels[offs + c], ref.texels[offs + c]), f"mip {k} channel {c}"
    finally:
        ctx.close()


def test_device_built_textures_render_like_host_built():
    """A frame drawn with textures made by srb_texture_create_rgba8 equals the frame with host-built textures."""
    from softrast_b200 import capi

    scene = scenes.parity_scene(320, 200, 3)
    imgs = [_image(64, 64, 1), _image(128, 128, 2), _image(32, 32, 3)]
    scene.textures = [capi.build_texture(im, capi.MIPS_STB) for im in imgs]
    g = capi.SceneRenderer(scene)
    try:
        g.render()
        c0, d0 = g.read_tiles()
    finally:
        g.close()
    g = capi.SceneRenderer(scene)
    try:
        handles = [g.ctx.create_texture_rgba8(im, capi.MIPS_STB) for im in imgs]
        for i, d in enumerate(scene.draws):
            if d.texture >= 0:
                g.descs[i].texture = handles[d.texture]
        g.render()
        c1, d1 = g.read_tiles()
    finally:
        g.close()
    assert np.array_equal(d0.view(np.uint32), d1.view(np.uint32)) and np.array_equal(c0, c1)


def test_device_texture_builder_errors_and_speed():
    from softrast_b200 import capi

    ctx = capi.RenderContext(0)
    try:
        with pytest.raises(capi.SrbError):
            ctx.create_texture_rgba8(_image(48, 32, 1), capi.MIPS_STB)  # not a power of two
        with pytest.raises(capi.SrbError):
            ctx.create_texture_rgba8(_image(16, 16, 1), capi.MIPS_NONE)  # smaller than a storage tile
        with pytest.raises(capi.SrbError, match="box"):
            ctx.create_texture_rgba8(_image(32, 32, 1), capi.MIPS_BOX)
        img = _image(1024, 1024, 5)
        ctx.create_texture_rgba8(img, capi.MIPS_STB)  # warm-up (module load, allocator)
        t0 = time.perf_counter()
        h = ctx.create_texture_rgba8(img, capi.MIPS_STB)
        dt_dev = time.perf_counter() - t0
        t0 = time.perf_counter()
        host = capi.build_texture(img, capi.MIPS_STB)
        dt_host = time.perf_counter() - t0
        assert np.array_equal(ctx.read_texture(h).texels, host.texels)
        print(f"1024x1024 + 10 mips: device builder {dt_dev * 1e3:.2f} ms (upload, tables and sync included), host builder {dt_host * 1e3:.1f} ms")
        assert dt_dev < dt_host
    finally:
        ctx.close()
