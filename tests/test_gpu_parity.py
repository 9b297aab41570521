"""GPU parity tests proper: the CUDA path, called through the C ABI, against the reference renderer itself
(oracle/_ref/libsrref_parity.so = the unmodified reference sources built -ffp-contract=off, single-threaded = canonical
order) on the same seeded inputs.

Bars (BASELINE.json north_star): per-tile triangle order and coverage masks bit-exact; depth within 1 ulp (we require
bit-exact); RGBA8 within 1 LSB per channel (we require bit-exact, and report the max difference if not)."""
import numpy as np
import pytest

from softrast_b200 import scenes

pytestmark = pytest.mark.gpu


def _ref(scene, threads=1):
    from oracle.refharness import RefRenderer

    r = RefRenderer(scene.width, scene.height, threads, "parity")
    r.load_scene(scene)
    r.render()
    return r


def _gpu(scene, **kw):
    from softrast_b200.capi import SceneRenderer

    g = SceneRenderer(scene, **kw)
    g.render()
    return g


def _max_channel_diff(a, b):
    a8 = a.view(np.uint8).astype(np.int32)
    b8 = b.view(np.uint8).astype(np.int32)
    return int(np.abs(a8 - b8).max())


def _compare_frame(scene, g, r, check_lists=True, check_coverage=False):
    counts_r = r.tile_counts()
    counts_g = g.ctx.tile_counts(g.fb.num_tiles)
    assert np.array_equal(counts_g, counts_r), "per-tile reference counts differ"
    if check_lists:
        for t in np.nonzero(counts_r)[0]:
            tr = r.tile_tris(int(t), int(counts_r[t]))
            tg = g.ctx.tile_tris(int(t), int(counts_r[t]))
            assert tr.tobytes() == tg.tobytes(), f"tile {t}: ordered tile-relative triangle records differ"
            ranks = g.ctx.tile_ranks(int(t), int(counts_r[t]))
            assert np.all(np.diff(ranks.astype(np.int64)) > 0), f"tile {t}: ranks not strictly ascending"
            if check_coverage:
                cr = r.tile_coverage(int(t), int(counts_r[t]))
                cg = g.ctx.tile_coverage(int(t), int(counts_r[t]))
                assert np.array_equal(cr, cg), f"tile {t}: coverage masks differ"
    colour_r, depth_r = r.read_tiles()
    colour_g, depth_g = g.read_tiles()
    assert np.array_equal(depth_g.view(np.uint32), depth_r.view(np.uint32)), "depth tiles not bit-exact"
    diff = _max_channel_diff(colour_g, colour_r)
    nbad = int((colour_g != colour_r).sum())
    assert diff <= 1, f"colour differs by {diff} LSB on {nbad} pixels"
    assert nbad == 0, f"colour within 1 LSB but not bit-exact on {nbad} pixels"


@pytest.mark.parametrize("full_records", [True, False])
@pytest.mark.parametrize("size,seed", [((320, 200), 3), ((257, 131), 4), ((64, 64), 5), ((640, 360), 6)])
def test_parity_scene(size, seed, full_records):
    """full_records: every varying's plane is set up and the ordered per-tile records are compared with the
    reference's; otherwise (the default, production mode) only the planes the bound shader reads are set up and the
    per-tile order, coverage, depth and colour are compared."""
    from softrast_b200.capi import FLAG_FULL_RECORDS

    scene = scenes.parity_scene(size[0], size[1], seed)
    r, g = _ref(scene), _gpu(scene, flags=FLAG_FULL_RECORDS if full_records else 0)
    try:
        c = g.ctx.counters()
        assert c["overflow"] == 0
        assert c["tris_clipped"] > 0, "the parity scene must exercise the clipper"
        if full_records:
            _compare_frame(scene, g, r, check_coverage=True)
        else:
            with pytest.raises(Exception):
                g.ctx.tile_tris(0, 1)  # record dumps need FLAG_FULL_RECORDS: fail loudly
            counts_r = r.tile_counts()
            for t in np.nonzero(counts_r)[0]:
                n = int(counts_r[t])
                ranks = g.ctx.tile_ranks(int(t), n)
                assert np.all(np.diff(ranks.astype(np.int64)) > 0), f"tile {t}: ranks not strictly ascending"
                # entry i of the canonically ordered list covers exactly what the reference's i-th triangle covers
                assert np.array_equal(g.ctx.tile_coverage(int(t), n), r.tile_coverage(int(t), n))
            _compare_frame(scene, g, r, check_lists=False)
    finally:
        r.close()
        g.close()


def test_host_pointer_draws_match_resident():
    scene = scenes.parity_scene(320, 200, 9)
    a, b = _gpu(scene, resident=True), _gpu(scene, resident=False)
    try:
        ca, da = a.read_tiles()
        cb, db = b.read_tiles()
        assert np.array_equal(ca, cb) and np.array_equal(da.view(np.uint32), db.view(np.uint32))
    finally:
        a.close()
        b.close()


def test_cube_grid_small():
    from softrast_b200.capi import FLAG_FULL_RECORDS

    scene = scenes.cube_grid(640, 360, 20, 20, draws=4)
    r, g = _ref(scene), _gpu(scene, flags=FLAG_FULL_RECORDS)
    try:
        _compare_frame(scene, g, r)
    finally:
        r.close()
        g.close()


def test_no_clear_accumulates_like_reference():
    """Second frame without ClearFrameBuffer: depth test against the previous frame's depth, colour kept."""
    scene = scenes.parity_scene(320, 200, 12)
    r, g = _ref(scene), _gpu(scene)
    try:
        scene2 = scenes.parity_scene(320, 200, 13)
        # same buffers layout? simply re-render the same scene without clearing: nothing may change
        c0, d0 = g.read_tiles()
        g.render(clear=False)
        r.render(clear=False)
        c1, d1 = g.read_tiles()
        cr, dr = r.read_tiles()
        assert np.array_equal(d1.view(np.uint32), dr.view(np.uint32))
        assert np.array_equal(c1, cr)
        assert np.array_equal(c0, c1)
    finally:
        r.close()
        g.close()


def test_frame_sequence_with_partial_clears():
    """A sequence of frames over ONE framebuffer with different draws and every ClearFrameBuffer combination
    (Renderer.cpp:168-194: colour only, depth only, both, none): depth and colour carry over between frames exactly as
    in the reference (the un-cleared plane is tested against / kept)."""
    a, b = scenes.parity_scene(320, 200, 12), scenes.parity_scene(320, 200, 13)
    sc = scenes.Scene("sequence", 320, 200, clear_color=0x5A)
    sc.textures = a.textures + b.textures
    for d in b.draws:
        if d.texture >= 0:
            d.texture += len(a.textures)
    sc.draws = a.draws + b.draws
    na = len(a.draws)
    da, db = list(range(na)), list(range(na, na + len(b.draws)))
    steps = [
        dict(draws=da, clear_colour=True, clear_depth=True),
        dict(draws=db, clear_colour=False, clear_depth=False),  # b over a: depth test against a's depth, a's colour kept
        dict(draws=da, clear_colour=True, clear_depth=False),   # colour wiped, depth kept: only nearer fragments reappear
        dict(draws=db, clear_colour=False, clear_depth=True),   # depth wiped, colour kept where b draws nothing
        dict(draws=da[:3] + db[:2], clear_colour=False, clear_depth=False),
        dict(draws=[], clear_colour=True, clear_depth=True),     # a frame of nothing but the clear
    ]
    from oracle.refharness import RefRenderer
    from softrast_b200.capi import SceneRenderer

    r = RefRenderer(sc.width, sc.height, 1, "parity")
    r.load_scene(sc)
    g = SceneRenderer(sc)
    try:
        for k, st in enumerate(steps):
            r.render(**st)
            g.render(**st)
            (cr, dr), (cg, dg) = r.read_tiles(), g.read_tiles()
            assert np.array_equal(dg.view(np.uint32), dr.view(np.uint32)), f"step {k}: depth"
            assert np.array_equal(cg, cr), f"step {k}: colour"
    finally:
        r.close()
        g.close()


@pytest.mark.parametrize("size", [(257, 131), (640, 360), (64, 64), (1000, 200), (132, 70)])
def test_blit_linear(size):
    """srb_blit_linear against the REFERENCE's own RenderContext::Blit (Renderer.cpp:319-372) of the same frame."""
    scene = scenes.parity_scene(size[0], size[1], 21)
    g = _gpu(scene)
    r = _ref(scene)
    try:
        want = r.blit_linear()
        px = g.blit_linear()
        assert np.array_equal(px, want)
    finally:
        r.close()
        g.close()


def test_blit_linear_bulk_copy_kernel():
    """The de-tile kernel built on bulk asynchronous copies (SRB_BULK_DETILE=1, cp.async.bulk + mbarrier) gives the same
    image as the reference's Blit; the knob is read when the library first blits, so this runs in a process of its own."""
    import subprocess
    import sys

    code = (
        "import os, sys, numpy as np\n"
        "sys.path.insert(0, os.getcwd())\n"
        "from softrast_b200 import scenes\n"
        "from softrast_b200.capi import SceneRenderer\n"
        "from oracle.refharness import RefRenderer\n"
        "for w, h in ((640, 360), (1000, 200), (64, 64)):\n"
        "    sc = scenes.parity_scene(w, h, 21)\n"
        "    g = SceneRenderer(sc); g.render()\n"
        "    r = RefRenderer(w, h, 1, 'parity'); r.load_scene(sc); r.render()\n"
        "    assert np.array_equal(g.blit_linear(), r.blit_linear()), (w, h)\n"
        "    g.close(); r.close()\n"
        "print('ok')\n"
    )
    import os

    env = dict(os.environ, SRB_BULK_DETILE="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "ok" in res.stdout, res.stdout[-1000:] + res.stderr[-2000:]


def test_blits_overlap_following_frames():
    """Viewer loop (Viewer/Main.cpp:50-69): EndFrame, then an asynchronous Blit that runs beside the next frame (the
    planes are double-buffered, Renderer.h:81-107).  Every blitted image must be its own frame's, whatever the overlap."""
    import ctypes as C

    from softrast_b200 import capi

    scene = scenes.hall_scene(640, 360, detail=0.1)
    frames = 6
    mvps = scenes.hall_camera_path(scene, 64)[::7][:frames].copy()
    g = capi.SceneRenderer(scene)
    nbytes = scene.width * scene.height * 4
    pinned = capi.host_alloc(frames * nbytes)
    try:
        for f in range(frames):
            g.render(mvps=mvps[f])  # synchronous EndFrame
            rc = capi.lib.srb_blit_linear(g.ctx.h, g.fb.handle, C.c_void_p(pinned + f * nbytes), None, None)
            assert rc == 0
        g.ctx.Sync()
        got = np.ctypeslib.as_array(C.cast(pinned, C.POINTER(C.c_uint32)), shape=(frames, scene.height, scene.width)).copy()
        from oracle.refharness import RefRenderer

        ref = RefRenderer(scene.width, scene.height, 1, "parity")
        try:
            ref.load_scene(scene)
            for f in range(frames):
                for i in range(ref.n_draws):
                    for k in range(16):
                        ref.descs[i].mvp[k] = float(mvps[f][i][k])
                ref.render()
                assert np.array_equal(got[f], ref.blit_linear()), f"frame {f}"  # the reference's own Blit of its own frame
        finally:
            ref.close()
    finally:
        capi.host_free(pinned)
        g.close()


def test_rcp_replay_matches_host_rcpps():
    from oracle.refharness import host_rcp
    from softrast_b200.capi import RenderContext

    rng = np.random.default_rng(1)
    x = rng.integers(0, 1 << 32, 1 << 20, dtype=np.uint64).astype(np.uint32).view(np.float32)
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -1.0, 1e-45, 1e-38, 3e38, 1.7e38, -2e38], np.float32)
    x = np.concatenate([x, special])
    ctx = RenderContext()
    try:
        got = ctx.debug_rcp(x).view(np.uint32)
        want = host_rcp(x).view(np.uint32)
        nan = np.isnan(x)
        assert np.array_equal(got[~nan], want[~nan])
        assert np.all(np.isnan(got[nan].view(np.float32)))
    finally:
        ctx.close()


def test_rsqrt_replay_matches_host_rsqrtps():
    """RSQRTPS of the host CPU (Viewer/SponzaScene.cpp:66) replayed on the device from the harvested table."""
    from oracle.refharness import host_rsqrt
    from softrast_b200.capi import RenderContext

    rng = np.random.default_rng(2)
    x = rng.integers(0, 1 << 32, 1 << 20, dtype=np.uint64).astype(np.uint32).view(np.float32)
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -1.0, 1e-45, 1e-38, 3e38, 1.7e38, -2e38, 2.0, 4.0], np.float32)
    x = np.concatenate([x, np.abs(x), special])
    ctx = RenderContext()
    try:
        got = ctx.debug_rsqrt(x).view(np.uint32)
        want = host_rsqrt(x).view(np.uint32)
        nan = np.isnan(want.view(np.float32))
        assert np.array_equal(got[~nan], want[~nan])
        assert np.all(np.isnan(got[nan].view(np.float32)))
    finally:
        ctx.close()


@pytest.mark.parametrize("size,seed", [((320, 200), 21), ((257, 131), 22)])
def test_sponza_shader_parity(size, seed):
    """The lit pixel shader of the reference's default scene (Viewer/SponzaScene.cpp:13-104): sun + 16 point lights with
    RSQRTPS / RCPPS, ambient, times the texture sample — bit-exact colour, incl. clipped triangles and null textures."""
    scene = scenes.parity_scene(size[0], size[1], seed, lit=True)
    r, g = _ref(scene), _gpu(scene)
    try:
        _compare_frame(scene, g, r, check_lists=False)
        colour, _ = g.read_tiles()
        assert len(np.unique(colour)) > 1000, "the lit scene should not be flat"
    finally:
        r.close()
        g.close()


def test_degenerate_plane_equations():
    """Slivers that are collinear in float raster space (K == 0 -> inf / NaN plane equations, Binning.cpp:261-277) but
    survive the fixed-point area cull: the reference's ordered compares make every fragment fail; the rasteriser's
    general row loop (non-finite z planes) must agree.  Identity MVP, w = 1: raster = ndc * 256 + 256 at 512x512."""
    from softrast_b200.scenes import Draw, Scene, build_tiled_texture, procedural_rgba

    def ndc(rx, ry):
        return (rx - 256.0) / 256.0, (256.0 - ry) / 256.0

    verts, idx = [], []

    def tri(p0, p1, p2, z):
        for (rx, ry), uv in zip((p0, p1, p2), ((0, 0), (1, 0), (0, 1))):
            x, y = ndc(rx, ry)
            verts.append([x, y, z, 0, 0, -1, uv[0], uv[1]])
        idx.extend(range(len(verts) - 3, len(verts)))

    for k in range(12):  # slivers at different places and z: area2 >> 8 == 10 after the snap, K == 0 in floats
        ox, oy = 10.0 + 37.0 * k, 10.0 + 29.0 * k
        tri((ox, oy), (ox + 10.0, oy + 2.0**-9), (ox + 20.0, oy + 2.0**-8), 0.25 + 0.05 * k)
    tri((5.0, 5.0), (5.0, 500.0), (500.0, 5.0), 0.125)  # an ordinary triangle behind them, both windings
    tri((5.0, 5.0), (500.0, 5.0), (5.0, 500.0), 0.125)
    sc = Scene("degenerate_planes", 512, 512, clear_color=0x33)
    sc.textures.append(build_tiled_texture(procedural_rgba(64, 77)))
    ident = np.eye(4, dtype=np.float32).reshape(-1)
    sc.draws.append(Draw(np.array(verts, dtype=np.float32), np.array(idx, dtype=np.uint32), ident, scenes.SHADER_UNLIT_DIFFUSE, 0))
    r, g = _ref(sc), _gpu(sc)
    try:
        assert g.ctx.counters()["tris_setup"] >= 13, "the slivers must survive the cull"
        _compare_frame(sc, g, r, check_lists=False)
    finally:
        r.close()
        g.close()


def test_device_capacity_overflow_grows_and_reruns():
    """More clipped fan triangles than the initial slot capacity (4096): the kernels flag the overflow on the device,
    the host grows the buffers and re-runs the frame (srb_api.cu: Finish).  The result must be the reference's."""
    scene = scenes.parity_scene(128, 128, 51, n_small=10, n_big=12000)
    r, g = _ref(scene), _gpu(scene)
    try:
        c = g.ctx.counters()
        assert c["overflow"] == 0, "the re-run must end without overflow"
        assert c["tris_clipped"] > 7000 and c["tris_setup"] > 4096 + 1000, c  # > 4096 fan slots needed
        _compare_frame(scene, g, r, check_lists=False)
        g.render()  # and again, now with the grown buffers
        _compare_frame(scene, g, r, check_lists=False)
    finally:
        r.close()
        g.close()


def test_tile_reference_overflow_grows_and_reruns():
    """More (triangle, tile) references than the initial list capacity (2^20): 280 000 triangles that each cover all four
    tiles of a 128x128 framebuffer.  The first frame overflows on the device and is re-run with grown buffers; it must
    equal the second frame (no overflow) and the analytic answer: every covered pixel ends at the nearest triangle's depth."""
    from softrast_b200.scenes import Draw, Scene

    n = 280_000
    rng = np.random.default_rng(61)
    z = rng.uniform(0.05, 0.95, n).astype(np.float32)  # ndc depth (reverse Z: larger = nearer), all distinct enough
    z[12345] = np.float32(0.99)
    # the lower-left half of the viewport, inside the frustum (no clipping), counter-clockwise after the y flip; its
    # bounding box spans all four tiles, which is what the reference bins by when the box is at most 2 tile rows high
    tri = np.array([[-1.0, -1.0], [1.0, -1.0], [-1.0, 1.0]], dtype=np.float32)
    v = np.zeros((n, 3, 8), dtype=np.float32)
    v[:, :, 0:2] = tri[None]
    v[:, :, 2] = z[:, None]
    v[:, :, 5] = -1.0
    sc = Scene("ref_overflow", 128, 128, clear_color=0)
    sc.draws.append(Draw(v.reshape(-1, 8), np.arange(3 * n, dtype=np.uint32), np.eye(4, dtype=np.float32).reshape(-1),
                         scenes.SHADER_VISUALIZE_NORMALS, -1))
    g = _gpu(sc)
    try:
        c = g.ctx.counters()
        assert c["overflow"] == 0 and c["tile_refs"] > (1 << 20), c
        c0, d0 = g.read_tiles()
        g.render()
        c1, d1 = g.read_tiles()
        assert np.array_equal(c0, c1) and np.array_equal(d0.view(np.uint32), d1.view(np.uint32))
        assert int((d0 > 0).sum()) > 128 * 128 // 2 - 200 and np.all(d0[d0 > 0] == np.float32(0.99))
    finally:
        g.close()


def test_unusual_vertex_layouts():
    """uv_offset != 6 (the mip derivatives come from varyings 3, 4 while the sample coordinates stay 6, 7:
    Rasterizer.cpp:378-399 vs Shaders.h:71-104) and vertices with fewer than 8 floats (6: position + normal, drawn with
    the normals shader)."""
    from softrast_b200.scenes import Draw

    base = scenes.parity_scene(320, 200, 71)
    sc = scenes.Scene("layouts", 320, 200, clear_color=0x21)
    sc.textures = base.textures
    d0, d1 = base.draws[0], base.draws[1]
    sc.draws.append(Draw(d0.vertices, d0.indices, d0.mvp, scenes.SHADER_UNLIT_DIFFUSE, 0, uv_offset=3))
    sc.draws.append(Draw(d1.vertices, d1.indices, d1.mvp, scenes.SHADER_UNLIT_DIFFUSE, 1, uv_offset=0))
    v6 = np.ascontiguousarray(base.draws[5].vertices[:, :6])
    sc.draws.append(Draw(v6, base.draws[5].indices, d0.mvp, scenes.SHADER_VISUALIZE_NORMALS, -1))
    r, g = _ref(sc), _gpu(sc)
    try:
        _compare_frame(sc, g, r, check_lists=False)
    finally:
        r.close()
        g.close()


def test_tile_ownership_renders_only_owned_tiles():
    """srb_set_tile_ownership (the screen-tile split of BASELINE config 4) on one GPU: a context that owns the tiles with
    tile % 3 == r must produce exactly the reference's content in those tiles and leave the others alone."""
    scene = scenes.parity_scene(320, 200, 72)
    r = _ref(scene)
    try:
        cr, dr = r.read_tiles()
        for rem in range(3):
            from softrast_b200.capi import SceneRenderer

            g = SceneRenderer(scene)
            try:
                g.ctx.set_tile_ownership(3, rem)
                g.render()
                cg, dg = g.read_tiles()
                own = (np.arange(cg.shape[0]) % 3) == rem
                assert np.array_equal(cg[own], cr[own]) and np.array_equal(dg[own].view(np.uint32), dr[own].view(np.uint32))
                assert not dg[~own].any(), "tiles of other owners must stay untouched (zero-initialised here)"
            finally:
                g.close()
    finally:
        r.close()


def test_host_buffer_edits_need_invalidation():
    """Borrowed host buffers are mirrored on the device and found by pointer: after an in-place edit
    the mirror is refreshed by srb_invalidate_host, or on every draw with SRB_FLAG_UPLOAD_ALWAYS."""
    from softrast_b200 import capi

    scene = scenes.parity_scene(320, 200, 73)
    g = capi.SceneRenderer(scene, resident=False)
    always = capi.SceneRenderer(scene, resident=False, flags=capi.FLAG_UPLOAD_ALWAYS)
    try:
        g.render()
        always.render()
        v = g._keep[0]  # draw 0's vertex array: the host memory the draw descriptor points to
        assert v is always._keep[0]
        v[:, 2] += np.float32(1.5)  # push draw 0 away from the camera, in place
        r = _ref(scene)  # the reference reads the edited memory
        try:
            cr, dr = r.read_tiles()
            g.render()
            assert not np.array_equal(g.read_tiles()[1].view(np.uint32), dr.view(np.uint32)), "stale mirror expected"
            capi.lib.srb_invalidate_host(g.ctx.h, C_void(v))
            g.render()
            always.render()
            for x in (g, always):
                cg, dg = x.read_tiles()
                assert np.array_equal(dg.view(np.uint32), dr.view(np.uint32)) and np.array_equal(cg, cr)
        finally:
            r.close()
    finally:
        g.close()
        always.close()


def C_void(a):
    import ctypes

    return ctypes.c_void_p(a.ctypes.data)


def _upload_always_pinned_arrays(expect_gather_kernel):
    import copy
    import ctypes

    from softrast_b200 import capi

    scene = scenes.parity_scene(320, 200, 74)
    pinned, sc = [], copy.copy(scene)
    sc.draws = []
    try:
        for k, d in enumerate(scene.draws):
            arrs = []
            for j, a in enumerate((np.ascontiguousarray(d.vertices, dtype=np.float32), np.ascontiguousarray(d.indices))):
                p = capi.host_alloc(a.nbytes + 16)
                pinned.append(p)
                off = (4 * ((k + j) % 4)) if a.dtype.itemsize >= 4 else (k + j) % 3 * a.dtype.itemsize  # misaligned starts too
                view = np.frombuffer((ctypes.c_char * a.nbytes).from_address(p + off), dtype=a.dtype).reshape(a.shape)
                view[...] = a
                arrs.append(view)
            sc.draws.append(scenes.Draw(arrs[0], arrs[1], d.mvp, d.shader, d.texture, d.uv_offset))
        g = capi.SceneRenderer(sc, resident=False, flags=capi.FLAG_UPLOAD_ALWAYS)
        r = _ref(sc)
        try:
            launches0 = g.ctx.launch_count()
            g.render()  # first frame: mirrors are created (plain copies)
            per_frame = g.ctx.launch_count() - launches0
            g.render()  # second frame: everything comes through the batched copy / the gather kernel
            extra = g.ctx.launch_count() - launches0 - 2 * per_frame
            assert extra == (1 if expect_gather_kernel else 0), "one batched copy (no kernel) or one gather launch per frame"
            cr, dr = r.read_tiles()
            cg, dg = g.read_tiles()
            assert np.array_equal(dg.view(np.uint32), dr.view(np.uint32)) and np.array_equal(cg, cr)
            sc.draws[0].vertices[:, 2] += np.float32(1.25)  # edit in place, no invalidation call
            sc.draws[1].vertices[:, 0] -= np.float32(0.5)
            r.render()
            g.render()
            cr, dr = r.read_tiles()
            cg, dg = g.read_tiles()
            assert np.array_equal(dg.view(np.uint32), dr.view(np.uint32)) and np.array_equal(cg, cr)
        finally:
            r.close()
            g.close()
    finally:
        for p in pinned:
            capi.host_free(p)


def test_upload_always_pulls_pinned_arrays():
    """SRB_FLAG_UPLOAD_ALWAYS with the application's arrays in pinned memory: one batched copy per frame pulls them into
    the device mirrors (aligned and misaligned segments), every frame — in-place edits show up without
    srb_invalidate_host, like the reference, which reads the arrays in place."""
    _upload_always_pinned_arrays(expect_gather_kernel=False)


@pytest.mark.parametrize("knob", ["SRB_GATHER_KERNEL", "SRB_GATHER_BULK"])
def test_upload_always_gather_kernels(knob):
    """The same through the gather kernel (what a driver without cudaMemcpyBatchAsync gets) and through its variant on
    bulk asynchronous copies; the knobs are read once per process, so each runs in a process of its own."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import os, sys\nsys.path.insert(0, os.getcwd())\nsys.path.insert(0, os.path.join(os.getcwd(), 'tests'))\n"
            "import test_gpu_parity as t\nt._upload_always_pinned_arrays(True)\nprint('ok')\n")
    res = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, **{knob: "1"}), capture_output=True, text=True,
                         timeout=300)
    assert res.returncode == 0 and "ok" in res.stdout, res.stdout[-1000:] + res.stderr[-2000:]


def test_many_textures_use_global_descriptors():
    """More textures than the shade kernel keeps in shared memory (48): the second instantiation reads the
    descriptors from global memory.  60 draws, one small texture each."""
    from softrast_b200.scenes import Draw, Scene, build_tiled_texture, procedural_rgba

    base = scenes.parity_scene(320, 200, 41, n_small=60, n_big=6)
    sc = Scene("many_textures", 320, 200, clear_color=0x11)
    v, i, mvp = base.draws[0].vertices, base.draws[0].indices, base.draws[0].mvp
    for k in range(60):
        sc.textures.append(build_tiled_texture(procedural_rgba(32 if k % 3 else 64, 500 + k), calc_mips=bool(k % 2)))
        tri = i[3 * k : 3 * k + 3]
        sc.draws.append(Draw(v, np.ascontiguousarray(tri), mvp, scenes.SHADER_UNLIT_DIFFUSE, k))
    sc.draws.append(Draw(base.draws[1].vertices, base.draws[1].indices, mvp, scenes.SHADER_UNLIT_DIFFUSE, 59))
    r, g = _ref(sc), _gpu(sc)
    try:
        _compare_frame(sc, g, r, check_lists=False)
    finally:
        r.close()
        g.close()


def test_sampler_matches_reference():
    from oracle.refharness import RefRenderer
    from softrast_b200.capi import RenderContext

    rng = np.random.default_rng(2)
    n = 1 << 16
    ctx = RenderContext()
    ref = RefRenderer(64, 64, 1, "parity")
    try:
        for size, mips in ((256, True), (64, True), (32, False)):
            t = scenes.build_tiled_texture(scenes.procedural_rgba(size, size), mips)
            hg, hr = ctx.create_texture(t), ref.create_texture(t)
            u = rng.uniform(-3, 3, n).astype(np.float32)
            v = rng.uniform(-3, 3, n).astype(np.float32)
            d = [(rng.uniform(-1, 1, n) * 10.0 ** rng.uniform(-5, 0, n)).astype(np.float32) for _ in range(4)]
            got = ctx.debug_sample(hg, u, v, *d)
            want = ref.sample(hr, u, v, *d)
            assert np.array_equal(got, want), f"{size}: {int((got != want).sum())} of {n} samples differ"
            # every sample on an even texel of mip 0: the 2x2 footprint is one aligned 16-byte group of the Morton order,
            # which the sampler fetches with ONE 128-bit load when a whole warp agrees (it does here)
            k = rng.integers(0, size // 2, (2, n))
            f = rng.uniform(0.02, 0.98, (2, n))
            u = ((2 * k[0] + f[0]) / size + rng.integers(-2, 3, n)).astype(np.float32)
            v = ((2 * k[1] + f[1]) / size + rng.integers(-2, 3, n)).astype(np.float32)
            d = [np.full(n, 1e-7, dtype=np.float32) for _ in range(4)]
            got = ctx.debug_sample(hg, u, v, *d)
            want = ref.sample(hr, u, v, *d)
            assert np.array_equal(got, want), f"{size}, 128-bit taps: {int((got != want).sum())} of {n} samples differ"
    finally:
        ctx.close()
        ref.close()


def test_frame_with_128_bit_texel_taps():
    """SRB_QUAD_TAPS=1 selects the shade instantiation that tests every sample for the 128-bit tap; the knob is read once
    per process, so this runs in a process of its own.  A strongly magnified texture (a quad filling the screen with a
    16th of a 32x32 texture) so that whole warps qualify, and the parity scene."""
    import os
    import subprocess
    import sys

    code = (
        "import os, sys, numpy as np\n"
        "sys.path.insert(0, os.getcwd())\n"
        "sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))\n"
        "import test_gpu_parity as t\n"
        "from softrast_b200 import scenes\n"
        "mag = scenes.Scene('magnified', 320, 200, clear_color=0x10)\n"
        "mag.textures.append(scenes.build_tiled_texture(scenes.procedural_rgba(32, 9)))\n"
        "mvp = scenes.to_column_major(scenes.reverse_z_projection(320, 200) @ scenes.look_at_lh((0.3, 0.4, -0.5), (0.0, 0.0, 6.0)))\n"
        "gv, gt = scenes._grid_surface((-60, -40, 30), (120, 0, 0), (0, 80, 3), 1, 1, (0, 0, -1), (0.0625, 0.0625))\n"
        "mag.draws.append(scenes.Draw(gv, gt.reshape(-1).astype(np.uint32), mvp, scenes.SHADER_UNLIT_DIFFUSE, 0))\n"
        "for sc in (mag, scenes.hall_scene(640, 360, detail=0.3)):\n"
        "    g, r = t._gpu(sc), t._ref(sc)\n"
        "    t._compare_frame(sc, g, r, check_lists=False)\n"
        "    g.close(); r.close()\n"
        "print('ok')\n"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, SRB_QUAD_TAPS="1"), capture_output=True,
                         text=True, timeout=600)
    assert res.returncode == 0 and "ok" in res.stdout, res.stdout[-1000:] + res.stderr[-2000:]


@pytest.mark.parametrize("name", __import__("tests.golden_util", fromlist=["x"]).golden_names())
def test_golden_fixture(name):
    """Committed fixtures generated from the reference (tests/golden/make_golden.py), replaying the RCPPS table of the
    CPU that produced them."""
    from softrast_b200.capi import FLAG_FULL_RECORDS, SceneRenderer
    from tests.golden_util import Golden

    gold = Golden(name)
    g = SceneRenderer(gold.scene, rcp=gold.rcp, rsqrt=gold.rsqrt, flags=FLAG_FULL_RECORDS)
    try:
        g.render()
        counts = g.ctx.tile_counts(g.fb.num_tiles)
        assert np.array_equal(counts, gold.counts)
        for t in np.nonzero(counts)[0]:
            t, n = int(t), int(counts[t])
            assert g.ctx.tile_tris(t, n).tobytes() == gold.tris[t].tobytes()
            assert np.array_equal(g.ctx.tile_coverage(t, n), gold.coverage[t])
        colour, depth = g.read_tiles()
        assert np.array_equal(depth.view(np.uint32), gold.depth_bits)
        assert np.array_equal(colour, gold.colour)
        g.render(clear=False)
        colour, depth = g.read_tiles()
        assert np.array_equal(depth.view(np.uint32), gold.depth_bits_noclear)
        assert np.array_equal(colour, gold.colour_noclear)
    finally:
        g.close()


def test_against_c_oracle_port():
    """The plain-C restatement as checker (it is what travels if oracle/_ref is absent)."""
    from oracle.refharness import PortRenderer, port_available
    from softrast_b200.capi import SceneRenderer, harvest_rcp_table

    if not port_available():
        pytest.skip("oracle/_build/libsr_oracle.so not built")
    table, bits = harvest_rcp_table(16)
    scene = scenes.hall_scene(960, 540, detail=0.1)
    p = PortRenderer(scene.width, scene.height, (table, bits))
    g = SceneRenderer(scene)
    try:
        p.load_scene(scene)
        p.render()
        g.render()
        _compare_frame(scene, g, p, check_lists=False)
    finally:
        p.close()
        g.close()


def test_full_size_configs_properties():
    """BASELINE.json configs 2 and 3 at full size (and config 2 through the lit Sponza shader): size-independent
    properties + bit-exact depth/colour against the compiled reference (single-threaded run takes a few seconds)."""
    for scene in (scenes.hall_scene(), scenes.random_tris(), scenes.hall_scene(lit=True)):
        g = _gpu(scene)
        r = _ref(scene)
        try:
            c = g.ctx.counters()
            assert c["overflow"] == 0 and c["tris_in"] == scene.num_tris
            counts = g.ctx.tile_counts(g.fb.num_tiles)
            assert int(counts.sum()) == c["tile_refs"] and int(counts.max()) == c["max_refs_in_tile"]
            assert np.array_equal(counts, r.tile_counts())
            # idempotence: rendering the same frame again over the finished one without a clear changes nothing
            c0, d0 = g.read_tiles()
            g.render(clear=False)
            c1, d1 = g.read_tiles()
            assert np.array_equal(c0, c1) and np.array_equal(d0.view(np.uint32), d1.view(np.uint32))
            assert int((d0 > 0).sum()) == c["pixels_covered"]
            cr, dr = r.read_tiles()
            assert np.array_equal(d0.view(np.uint32), dr.view(np.uint32))
            assert np.array_equal(c0, cr)
            # ranks of the heaviest tile ascend strictly (canonical order)
            t = int(np.argmax(counts))
            ranks = g.ctx.tile_ranks(t, int(counts[t]))
            assert np.all(np.diff(ranks.astype(np.int64)) > 0)
        finally:
            g.close()
            r.close()


def test_4k_and_many_draws():
    """BASELINE config 4's frame (hall at 3840x2160, 2040 tiles) on one GPU, and config 1's variant with one draw per
    cube row (100 draws): counts, depth and colour bit-exact against the compiled reference."""
    for scene in (scenes.hall_scene(3840, 2160), scenes.cube_grid(1280, 720, 100, 100, draws=100),
                  scenes.cube_grid(1280, 720, 100, 100, draws=1)):  # config 1 as ONE draw of 120 000 triangles
        g, r = _gpu(scene), _ref(scene)
        try:
            assert g.ctx.counters()["overflow"] == 0
            _compare_frame(scene, g, r, check_lists=False)
        finally:
            g.close()
            r.close()


def test_overflow_rerun_keeps_the_batch_readback():
    """srb_render_frames with colour_out: a frame that overflows a device capacity is re-run at the context's next
    end_frame / sync, and the read-back of that frame must be issued again — colour_out must hold the finished frames."""
    import ctypes as C

    from softrast_b200 import capi

    scene = scenes.parity_scene(128, 128, 51, n_small=10, n_big=12000)  # > 4096 clipped fan slots: overflows at first
    rs = [capi.SceneRenderer(scene)]
    rs.append(capi.SceneRenderer(scene, share=rs[0]))
    frames = 4
    nbytes = rs[0].fb.num_tiles * 16384
    pinned = capi.host_alloc(frames * nbytes)
    C.memset(pinned, 0xAB, frames * nbytes)
    r = _ref(scene)
    try:
        capi.render_frames(rs, frames, None, pinned, nbytes)
        assert rs[0].ctx.counters()["overflow"] == 0 and rs[0].ctx.counters()["tris_setup"] > 4096 + 1000
        got = np.ctypeslib.as_array(C.cast(pinned, C.POINTER(C.c_uint32)), shape=(frames, rs[0].fb.num_tiles, 64, 64)).copy()
        want, _ = r.read_tiles()
        for f in range(frames):
            assert np.array_equal(got[f], want), f"frame {f} of the batch is not the finished frame"
    finally:
        capi.host_free(pinned)
        r.close()
        for x in rs:
            x.close()


def test_buffer_bindings_are_validated():
    """Offsets into device buffers: no 64-bit wrap in the range check, and alignment to what the kernels read (a misaligned
    load would fault on the device and poison the CUDA context).  Upload-always contexts cannot share resources."""
    import ctypes as C

    from softrast_b200 import capi

    scene = scenes.parity_scene(64, 64, 5, n_small=20, n_big=2)
    g = capi.SceneRenderer(scene, resident=True)
    try:
        c = g.ctx
        d = g.descs[0]
        for field, bad in (("positions", 2), ("attributes", 6), ("indices", 1 if d.indices.stride > 1 else None),
                           ("positions", (1 << 64) - 4)):
            if bad is None:
                continue
            e = capi.DrawDesc.from_buffer_copy(d)
            getattr(e, field).offset = bad
            c.BeginFrame()
            c.ClearFrameBuffer(g.fb, 0, True, True)
            rc = capi.lib.srb_draw_indexed(c.h, C.byref(e))
            assert rc != 0, (field, bad)
            c.EndFrame()
        g.render()  # the context is still usable
        with pytest.raises(capi.SrbError):
            capi.RenderContext(0, capi.FLAG_UPLOAD_ALWAYS, share=c)
    finally:
        g.close()


def test_one_frame_draws_into_two_framebuffers():
    """DrawCall::SetFrameBuffer is per draw (Renderer.h:129): the draws of one frame may address different framebuffers.
    Each framebuffer must end up as if its draws (in order) had been a frame of their own."""
    from softrast_b200 import capi

    scene = scenes.parity_scene(320, 200, 33)
    half = len(scene.draws) // 2
    g = capi.SceneRenderer(scene)
    fb2 = g.ctx.create_framebuffer(scene.width, scene.height)
    try:
        c = g.ctx
        c.BeginFrame()
        c.ClearFrameBuffer(g.fb, scene.clear_color, True, True)
        c.ClearFrameBuffer(fb2, 0x11, True, True)
        for i in range(g.n_draws):  # interleaved: even draws -> fb, odd draws -> fb2
            g.descs[i].framebuffer = g.fb.handle if i % 2 == 0 else fb2.handle
            c.DrawIndexed(g.descs[i])
        c.EndFrame()
        got = [g.fb.read_tiles(), fb2.read_tiles()]
        for k, (fb, clear) in enumerate(((g.fb, scene.clear_color), (fb2, 0x11))):
            import copy

            sub = copy.copy(scene)
            sub.draws = [d for i, d in enumerate(scene.draws) if i % 2 == k]
            sub.clear_color = clear
            r = _ref(sub)
            try:
                cr, dr = r.read_tiles()
                assert np.array_equal(got[k][1].view(np.uint32), dr.view(np.uint32)), f"framebuffer {k}: depth"
                assert np.array_equal(got[k][0], cr), f"framebuffer {k}: colour"
            finally:
                r.close()
        assert half > 0
    finally:
        g.close()


def test_more_draws_than_shared_memory_holds():
    """6 000 draws of one triangle each: the set-up kernel's per-draw table no longer fits its shared-memory budget and is
    searched in global memory (the reference has no limit on the number of draws per frame, Renderer.cpp:161-166; its
    only limit is 512 chunks per tile and thread, Binning.h:16, so the triangles are spread over the screen)."""
    from softrast_b200.scenes import Draw, Scene

    base = scenes.random_tris(640, 360, n=6000, seed=0x77)
    d0 = base.draws[0]
    tris = np.ascontiguousarray(d0.indices).reshape(-1, 3)
    sc = Scene("many_draws", base.width, base.height, clear_color=base.clear_color)
    sc.textures = base.textures
    for k in range(len(tris)):
        sc.draws.append(Draw(d0.vertices, np.ascontiguousarray(tris[k]), d0.mvp, d0.shader, d0.texture, d0.uv_offset))
    g, r = _gpu(sc, resident=False), _ref(sc)
    try:
        assert g.ctx.counters()["tris_in"] == len(tris) and len(sc.draws) > 4096
        _compare_frame(sc, g, r, check_lists=False)
    finally:
        g.close()
        r.close()


@pytest.mark.parametrize("shared", [True, False])
def test_batched_frames_in_flight_match_single_frames(shared):
    """srb_render_frames (camera-path batch, several contexts = several frames in flight, per-frame D2H into pinned
    memory) must give exactly the frames that one-at-a-time rendering gives, and those must match the reference.
    shared: the contexts share one device copy of the scene (srb_create_shared) instead of one copy each."""
    import ctypes as C

    from softrast_b200 import capi

    scene = scenes.hall_scene(640, 360, detail=0.1)
    frames = 7
    mvps = scenes.hall_camera_path(scene, 64)[::9][:frames].copy()
    rs = [capi.SceneRenderer(scene)]
    for _ in range(2):
        rs.append(capi.SceneRenderer(scene, share=rs[0] if shared else None))
    nbytes = rs[0].fb.num_tiles * 16384
    pinned = capi.host_alloc(frames * nbytes)
    try:
        capi.render_frames(rs, frames, mvps, pinned, nbytes)
        got = np.ctypeslib.as_array(C.cast(pinned, C.POINTER(C.c_uint32)), shape=(frames, rs[0].fb.num_tiles, 64, 64)).copy()
        single = capi.SceneRenderer(scene)
        ref = _ref(scene)
        try:
            for f in range(frames):
                single.render(mvps=mvps[f])
                colour, _ = single.read_tiles()
                assert np.array_equal(got[f], colour), f"frame {f}: batch != single"
            # and one of them against the reference renderer
            import ctypes

            for i in range(ref.n_draws):
                ctypes.memmove(ref.descs[i].mvp, mvps[3, i].ctypes.data, 64)
            ref.render()
            assert np.array_equal(got[3], ref.read_tiles()[0])
        finally:
            single.close()
            ref.close()
    finally:
        capi.host_free(pinned)
        for r in rs:  # the parent first: the shared scene must survive until the last context of the family closes
            r.close()


@pytest.mark.parametrize("category", __import__("tests.fuzz_parity", fromlist=["x"]).CATEGORIES)
def test_fuzz_parity_sweep(category):
    """Adversarial inputs (tests/fuzz_parity.py): huge coordinates, vertices on / behind the camera plane, zero-area and
    sub-pixel triangles, extreme UVs, exact depth ties, NaN / inf vertices, every fourth scene through the lit shader —
    per-tile counts, depth and colour bit-exact against the reference.  The suite runs 24 scenes per category; the log of
    a larger sweep of the same generator on the final kernels of the round is profiles/r02_fuzz.log."""
    from tests import fuzz_parity as fz

    for seed in range(7000, 7024):
        ok, depth_bad, colour_bad = fz.compare(fz.make_scene(category, seed))
        assert ok and depth_bad == 0 and colour_bad == 0, f"{category} seed {seed}: counts_ok={ok} depth={depth_bad} colour={colour_bad}"
