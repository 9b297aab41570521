"""CPU tests (no GPU): the plain-C oracle restatement (oracle/sr_oracle.c) is pinned against
  (1) the golden fixtures generated from the reference itself (tests/golden/*.npz, made by make_golden.py), and
  (2) when oracle/_ref is present, the compiled reference live on larger seeded scenes,
and the compiled reference is re-checked against its own fixtures (guards the fixtures and the build flags)."""
import numpy as np
import pytest

from oracle import refharness as rh
from softrast_b200 import scenes
from tests.golden_util import Golden, check_against_golden, golden_names

needs_port = pytest.mark.skipif(not rh.port_available(), reason="oracle/_build/libsr_oracle.so not built")
needs_ref = pytest.mark.skipif(not rh.ref_available(), reason="oracle/_ref not built (reference sources absent)")


def test_fixtures_exist():
    assert len(golden_names()) >= 3


@needs_port
@pytest.mark.parametrize("name", golden_names())
def test_port_matches_golden(name):
    g = Golden(name)
    p = rh.PortRenderer(g.scene.width, g.scene.height, g.rcp, g.rsqrt)
    try:
        p.load_scene(g.scene)
        p.render()
        check_against_golden(g, p)
        p.render(clear=False)
        c, d = p.read_tiles()
        assert np.array_equal(d.view(np.uint32), g.depth_bits_noclear)
        assert np.array_equal(c, g.colour_noclear)
    finally:
        p.close()


@needs_ref
@pytest.mark.parametrize("name", golden_names())
def test_reference_matches_its_golden(name):
    g = Golden(name)
    host_table = rh.harvest_rcp_table(11)
    r = rh.RefRenderer(g.scene.width, g.scene.height, 1, "parity")
    try:
        r.load_scene(g.scene)
        r.render()
        # the colour of the fixtures depends on the RCPPS (lit scene: and RSQRTPS) table of the CPU that made them
        same_rsqrt = g.rsqrt is None or np.array_equal(rh.harvest_rsqrt_table(10), g.rsqrt[0])
        check_against_golden(g, r, exact_colour=np.array_equal(host_table, g.rcp[0]) and same_rsqrt, check_colour=same_rsqrt)
    finally:
        r.close()


@needs_ref
@needs_port
@pytest.mark.parametrize(
    "make",
    [
        lambda: scenes.parity_scene(320, 200, 3),
        lambda: scenes.parity_scene(257, 131, 4),
        lambda: scenes.cube_grid(640, 360, 24, 24, draws=5),
        lambda: scenes.hall_scene(640, 360, detail=0.05),
        lambda: scenes.random_tris(480, 270, 20000, seed=77),
    ],
)
def test_port_matches_reference_live(make):
    sc = make()
    rcp = (rh.harvest_rcp_table(11), 11)
    r = rh.RefRenderer(sc.width, sc.height, 1, "parity")
    p = rh.PortRenderer(sc.width, sc.height, rcp)
    try:
        for x in (r, p):
            x.load_scene(sc)
            x.render()
        counts = r.tile_counts()
        assert np.array_equal(counts, p.tile_counts())
        for t in np.nonzero(counts)[0][::3]:
            t, n = int(t), int(counts[t])
            assert r.tile_tris(t, n).tobytes() == p.tile_tris(t, n).tobytes()
            assert np.array_equal(r.tile_fragments(t)[0], p.tile_fragments(t)[0])
        (cr, dr), (cp, dp) = r.read_tiles(), p.read_tiles()
        assert np.array_equal(dr.view(np.uint32), dp.view(np.uint32))
        assert np.array_equal(cr, cp)
    finally:
        r.close()
        p.close()


@needs_ref
def test_multithreaded_reference_depth_matches_single_threaded():
    """thread count never changes depth (colour may differ on exact depth ties)."""
    sc = scenes.cube_grid(640, 360, 24, 24, draws=3)
    a = rh.RefRenderer(sc.width, sc.height, 1, "parity")
    b = rh.RefRenderer(sc.width, sc.height, 4, "parity")
    try:
        for x in (a, b):
            x.load_scene(sc)
            x.render()
        assert b.threads == 4
        assert np.array_equal(a.read_tiles()[1].view(np.uint32), b.read_tiles()[1].view(np.uint32))
        assert np.array_equal(a.tile_counts(), b.tile_counts())
    finally:
        a.close()
        b.close()


@needs_ref
@needs_port
def test_rcp_replay_model_matches_host_rcpps():
    rng = np.random.default_rng(5)
    x = rng.integers(0, 1 << 32, 1 << 21, dtype=np.uint64).astype(np.uint32).view(np.float32)
    x = np.concatenate([x, np.array([0.0, -0.0, np.inf, -np.inf, 1.0, 1e-45, 1e-38, 3e38, 1.7e38], np.float32)])
    p = rh.PortRenderer(64, 64, (rh.harvest_rcp_table(11), 11))
    try:
        got, want = p.rcp(x).view(np.uint32), rh.host_rcp(x).view(np.uint32)
        ok = ~np.isnan(x)
        assert np.array_equal(got[ok], want[ok])
    finally:
        p.close()


@needs_ref
@needs_port
def test_rsqrt_replay_model_matches_host_rsqrtps():
    rng = np.random.default_rng(6)
    x = rng.integers(0, 1 << 32, 1 << 21, dtype=np.uint64).astype(np.uint32).view(np.float32)
    x = np.concatenate([x, np.abs(x), np.array([0.0, -0.0, np.inf, -np.inf, 1.0, 2.0, 1e-45, 1e-38, 3e38, -1.0], np.float32)])
    p = rh.PortRenderer(64, 64, (rh.harvest_rcp_table(11), 11), (rh.harvest_rsqrt_table(10), 10))
    try:
        got, want = p.rsqrt(x).view(np.uint32), rh.host_rsqrt(x).view(np.uint32)
        ok = ~np.isnan(want.view(np.float32))
        assert np.array_equal(got[ok], want[ok])
        assert np.all(np.isnan(got[~ok].view(np.float32)))
    finally:
        p.close()


@needs_ref
@needs_port
def test_port_sponza_shader_matches_reference():
    """The C restatement of SponzaShader against the reference's own (Viewer/SponzaScene.cpp:13-104)."""
    scene = scenes.parity_scene(200, 136, 23, lit=True)
    r = rh.RefRenderer(scene.width, scene.height, 1, "parity")
    p = rh.PortRenderer(scene.width, scene.height, (rh.harvest_rcp_table(11), 11))
    try:
        for x in (r, p):
            x.load_scene(scene)
            x.render()
        (cr, dr), (cp, dp) = r.read_tiles(), p.read_tiles()
        assert np.array_equal(dr.view(np.uint32), dp.view(np.uint32))
        assert np.array_equal(cr, cp)
    finally:
        r.close()
        p.close()


@needs_ref
@needs_port
def test_sampler_port_matches_reference():
    rng = np.random.default_rng(6)
    n = 1 << 15
    r = rh.RefRenderer(64, 64, 1, "parity")
    p = rh.PortRenderer(64, 64, (rh.harvest_rcp_table(11), 11))
    try:
        for size, mips in ((256, True), (32, False)):
            t = scenes.build_tiled_texture(scenes.procedural_rgba(size, size), mips)
            hr, hp = r.create_texture(t), p.create_texture(t)
            u, v = (rng.uniform(-3, 3, n).astype(np.float32) for _ in range(2))
            d = [(rng.uniform(-1, 1, n) * 10.0 ** rng.uniform(-5, 0, n)).astype(np.float32) for _ in range(4)]
            assert np.array_equal(r.sample(hr, u, v, *d), p.sample(hp, u, v, *d))
    finally:
        r.close()
        p.close()


@needs_ref
def test_reference_texture_builder_layout_matches_ours_on_mip0():
    """The reference's own CreateFromRGBA8 (stb mips) and our builders agree on layout: mip offsets and mip 0 bytes."""
    rgba = scenes.procedural_rgba(64, 9)
    r = rh.RefRenderer(64, 64, 1, "parity")
    try:
        t_ref = r.get_texture(r.create_texture_rgba8(rgba, True))
        t_own = scenes.build_tiled_texture(rgba, True)
        assert t_ref.num_mips == t_own.num_mips
        assert np.array_equal(t_ref.mip_offsets[: t_ref.num_mips], t_own.mip_offsets[: t_own.num_mips])
        assert t_ref.texels.size == t_own.texels.size
        assert np.array_equal(t_ref.texels[: t_own.mip_offsets[1]], t_own.texels[: t_own.mip_offsets[1]])
    finally:
        r.close()


@needs_ref
@needs_port
@pytest.mark.parametrize("category", __import__("tests.fuzz_parity", fromlist=["x"]).CATEGORIES)
def test_port_matches_reference_on_adversarial_inputs(category):
    """The C restatement against the reference on the adversarial generator of tests/fuzz_parity.py (huge coordinates,
    camera-plane vertices, degenerate triangles, extreme UVs, depth ties, NaN / inf, lit shader)."""
    from tests import fuzz_parity as fz

    for seed in range(8000, 8004):
        sc = fz.make_scene(category, seed)
        r = rh.RefRenderer(sc.width, sc.height, 1, "parity")
        p = rh.PortRenderer(sc.width, sc.height, (rh.harvest_rcp_table(11), 11))
        try:
            for x in (r, p):
                x.load_scene(sc)
                x.render()
            assert np.array_equal(r.tile_counts(), p.tile_counts()), f"{category} {seed}: counts"
            (cr, dr), (cp, dp) = r.read_tiles(), p.read_tiles()
            assert np.array_equal(dr.view(np.uint32), dp.view(np.uint32)), f"{category} {seed}: depth"
            assert np.array_equal(cr, cp), f"{category} {seed}: colour"
        finally:
            r.close()
            p.close()


@needs_ref
@needs_port
def test_port_frame_sequence_with_partial_clears():
    """Frames over one framebuffer with different draws and every ClearFrameBuffer combination (Renderer.cpp:168-194):
    the C restatement carries depth and colour between frames exactly like the reference."""
    a, b = scenes.parity_scene(200, 136, 12), scenes.parity_scene(200, 136, 13)
    sc = scenes.Scene("sequence", 200, 136, clear_color=0x5A)
    sc.textures = a.textures + b.textures
    for d in b.draws:
        if d.texture >= 0:
            d.texture += len(a.textures)
    sc.draws = a.draws + b.draws
    na = len(a.draws)
    da, db = list(range(na)), list(range(na, na + len(b.draws)))
    steps = [dict(draws=da, clear_colour=True, clear_depth=True), dict(draws=db, clear_colour=False, clear_depth=False),
             dict(draws=da, clear_colour=True, clear_depth=False), dict(draws=db, clear_colour=False, clear_depth=True),
             dict(draws=[], clear_colour=True, clear_depth=True)]
    r = rh.RefRenderer(sc.width, sc.height, 1, "parity")
    p = rh.PortRenderer(sc.width, sc.height, (rh.harvest_rcp_table(11), 11))
    try:
        for x in (r, p):
            x.load_scene(sc)
        for k, st in enumerate(steps):
            r.render(**st)
            p.render(**st)
            (cr, dr), (cp, dp) = r.read_tiles(), p.read_tiles()
            assert np.array_equal(dr.view(np.uint32), dp.view(np.uint32)), f"step {k}: depth"
            assert np.array_equal(cr, cp), f"step {k}: colour"
    finally:
        r.close()
        p.close()


def _untile(t, m, w, h):
    """Valid texels of mip m as (h, w, 4) — the padding of a level smaller than 32x32 is never read by the sampler (and
    is uninitialised memory in the reference)."""
    off = int(t.mip_offsets[m])
    tiles_x = (max(w, 32) + 31) // 32
    ys, xs = np.mgrid[0:h, 0:w]
    mor = np.zeros_like(xs)
    for b in range(5):
        mor |= (((xs & 31) >> b) & 1) << (2 * b)
        mor |= (((ys & 31) >> b) & 1) << (2 * b + 1)
    idx = (((ys >> 5) * tiles_x + (xs >> 5)) * 1024 + mor) * 4 + off
    return np.stack([t.texels[idx + c] for c in range(4)], axis=-1)


@needs_ref
@pytest.mark.parametrize("size", [(32, 32), (64, 64), (128, 32), (32, 256), (256, 64), (512, 512)])
def test_texture_builder_stb_mips_match_reference(size):
    """SRB_MIPS_STB: every mip texel equals TextureData::CreateFromRGBA8's (Texture.cpp:122-199), which filters each
    level from the original image with stbir_resize_uint8 (Mitchell, clamped edges) — on noise and on a smooth ramp."""
    from softrast_b200 import capi

    h, w = size
    rng = np.random.default_rng(w * 7 + h)
    r = rh.RefRenderer(64, 64, 1, "parity")
    try:
        for kind in range(2):
            rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
            if kind:
                rgba[:, :, :3] = ((np.add.outer(np.arange(h), np.arange(w)) * 3) % 256)[:, :, None].astype(np.uint8)
            ours = capi.build_texture(rgba, capi.MIPS_STB)
            ref = r.get_texture(r.create_texture_rgba8(rgba, True))
            assert ours.num_mips == ref.num_mips
            assert np.array_equal(ours.mip_offsets[: ref.num_mips], ref.mip_offsets[: ref.num_mips])
            for m in range(ref.num_mips):
                mw, mh = max(1, w >> m), max(1, h >> m)
                assert np.array_equal(_untile(ours, m, mw, mh), _untile(ref, m, mw, mh)), f"mip {m}"
    finally:
        r.close()


@pytest.mark.skipif(not rh.ref_available(), reason="oracle/_ref not built")
def test_sponza_scene_animation_matches_reference():
    """srb_sponza_scene_init / _update against the reference's own SponzaScene::Init / Update (Viewer/SponzaScene.cpp:
    126-160, 168-187, compiled in place): light set-up from kt::XorShift32 (incl. the right-to-left evaluation of the
    Vec3 constructor's arguments) and 400 animated frames with varying time steps, every float bit-identical."""
    from softrast_b200 import capi

    r = rh.RefRenderer(128, 64, 1, "parity")
    ref = rh.RefSponzaScene(r)
    try:
        ours = capi.SponzaSceneAnim()
        a, b = ours.constants, ref.constants
        # Init sets sun, ambient, intensity and falloff; positions and colours are whatever g_constants held before (the
        # reference's block is a file-static that other tests write) until the first Update
        keep = np.ones(136, dtype=bool)
        keep[8:].reshape(16, 8)[:, 0:6] = False
        assert np.array_equal(a.view(np.uint32)[keep], b.view(np.uint32)[keep])
        assert np.all(a[0:3] == a[0]) and a[0] != 0  # the all-x sun direction (SponzaScene.cpp:135-137)
        assert np.allclose(a[3:6], 0.1)
        seen_motion = False
        prev = None
        for f in range(400):
            dt = float(np.float32(0.004 + 0.003 * (f % 11)))
            a, b = ours.update(dt), ref.update(dt)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"frame {f}"
            if prev is not None and not np.array_equal(prev, a):
                seen_motion = True
            prev = a
        assert seen_motion
        lights = a[8:].reshape(16, 8)
        assert np.all(lights[:, 6] >= 150) and np.all(lights[:, 6] <= 350) and np.all(lights[:, 7] >= 500)
    finally:
        ref.close()
        r.close()


def test_detile_helper_matches_reference_blit():
    """oracle.refharness.detile (used by host-side checks) pinned against the reference's own RenderContext::Blit /
    BlitJobFn (Renderer.cpp:319-372) on ragged and tile-aligned sizes."""
    from oracle.refharness import RefRenderer, detile

    for w, h in ((257, 131), (320, 200), (64, 64), (65, 1)):
        sc = scenes.parity_scene(w, h, 5, n_small=60, n_big=6)
        r = RefRenderer(w, h, 1, "parity")
        try:
            r.load_scene(sc)
            r.render()
            colour, _ = r.read_tiles()
            assert np.array_equal(r.blit_linear(), detile(colour, w, h)), (w, h)
        finally:
            r.close()


def test_texel_conversion_identity():
    """The shade kernel converts a texel byte b to k * (float)b (k = 1.0f / 255.0f, Texture.cpp:438-456) without an integer
    conversion: fma(2^23 + b, k, -(2^23 * k)) (srb_raster.cu: texel_to_float).  The exact value of that FMA's argument is
    b * k, so its single rounding is the reference's product: checked here for all 256 bytes in exact arithmetic."""
    k = np.float32(1.0) / np.float32(255.0)
    b = np.arange(256, dtype=np.float64)
    want = k * np.arange(256, dtype=np.float32)  # one rounding of b * k
    bias = np.float64(8388608.0) * np.float64(k)
    assert np.float32(bias) == bias  # 2^23 * k is a float: the kernel's constant is exact
    exact = (np.float64(8388608.0) + b) * np.float64(k) - bias  # 48-bit product and the difference are exact in float64
    assert np.array_equal(exact, b * np.float64(k))
    got = exact.astype(np.float32)  # the FMA's single rounding
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
