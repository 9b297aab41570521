"""Scene ingestion end to end on the GPU: an OBJ + MTL + textures loaded by srb_model_load, made resident
by srb_model_make_resident and drawn with the draw list of Viewer/Scene.cpp:35-63 gives the pixels the reference renders
from ITS OWN loader's arrays (sr::Obj::Model::Load, compiled in place)."""
import ctypes as C
import os

import numpy as np
import pytest

from softrast_b200 import scenes

from . import objgen

pytestmark = pytest.mark.gpu


def _render_resident(model, width, height, mvp, shader, clear_color, sponza=None):
    from softrast_b200 import capi

    ctx = capi.RenderContext(0)
    fb = ctx.create_framebuffer(width, height)
    if sponza is not None:
        ctx.set_sponza_constants(sponza)
    rm = capi.ResidentModel(ctx, model)
    descs, n = rm.draws(fb.handle, mvp, shader)
    ctx.BeginFrame()
    ctx.ClearFrameBuffer(fb, clear_color, True, True)
    for i in range(n):
        ctx.DrawIndexed(descs[i])
    ctx.EndFrame(True)
    colour, depth = fb.read_tiles()
    counters = ctx.counters()
    rm.close()
    ctx.close()
    return colour, depth, counters, n


@pytest.mark.parametrize("lit", [False, True])
def test_obj_model_renders_like_reference(tmp_path, lit):
    from oracle import refharness as rh
    from softrast_b200 import capi

    W, H = 448, 256
    flags = capi.OBJ_FLIP_UVS
    po = objgen.write_model(str(tmp_path / "ours"), seed=9)
    pr = objgen.write_model(str(tmp_path / "ref"), seed=9)
    proj = scenes.reverse_z_projection(W, H)
    view = scenes.look_at_lh((0.4, 0.3, -1.0), (0.0, 0.0, 6.0))
    mvp = scenes.to_column_major(proj @ view)
    shader = 3 if lit else 0
    sponza = scenes.sponza_constants(5) if lit else None

    model = capi.Model(po, flags)
    colour, depth, counters, n = _render_resident(model, W, H, mvp, shader, 0x18, sponza)
    assert n == 5 and counters["overflow"] == 0 and counters["tris_in"] == sum(m["indices"].size // 3 for m in model.meshes)
    assert counters["pixels_covered"] > W * H // 8

    # the reference: its own loader's arrays through its own renderer
    meshes, mats = rh.ref_load_model(pr, flags)
    sc = scenes.Scene("obj_ref", W, H, clear_color=0x18)
    sc.sponza = sponza
    tex_of = {}
    for i, m in enumerate(mats):
        if m["texels"].size:
            tex_of[i] = len(sc.textures)
            sc.textures.append(scenes.TiledTexture(m["texels"], m["mip_offsets"], m["num_mips"], m["width_log2"], m["height_log2"]))
    for m in meshes:
        has_mat = m["material"] < len(mats)
        sc.draws.append(scenes.Draw(m["vertices"], m["indices"], mvp, shader if (has_mat or lit) else 1,
                                    tex_of.get(m["material"], -1), 6))
    r = rh.RefRenderer(W, H, 1, "parity")
    try:
        r.load_scene(sc)
        r.render()
        rc, rd = r.read_tiles()
    finally:
        r.close()
    assert np.array_equal(depth.view(np.uint32), rd.view(np.uint32))
    assert np.array_equal(colour, rc)

    # the same model from the cache the first load wrote, drawn from host pointers (Model.to_scene): same pixels
    cached = capi.Model(po, flags)
    assert cached.from_cache
    sc2 = cached.to_scene(W, H, mvp, shader, 0x18)
    sc2.sponza = sponza
    g = capi.SceneRenderer(sc2, resident=False)
    try:
        g.render()
        c2, d2 = g.read_tiles()
    finally:
        g.close()
    assert np.array_equal(d2.view(np.uint32), rd.view(np.uint32)) and np.array_equal(c2, rc)
    cached.close()
    model.close()


@pytest.mark.parametrize("example", ["shim_obj_example", "abi_example"])
def test_shim_obj_scene_matches_reference(tmp_path, example):
    """tests/cpp/shim_obj_example.cpp = Viewer/Scene.cpp's SimpleModelScene compiled against the C++ shim (sr::Obj::Model::
    Load + one DrawCall per mesh); tests/c/abi_example.c = the same frame through the C ABI from plain C99
    (srb_model_load, srb_model_make_resident, srb_resident_model_draws): the tiles they dump are the reference's."""
    import os
    import subprocess

    from oracle import refharness as rh

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "_build", example)
    assert os.path.exists(exe), "run __graft_entry__.build()"
    po = objgen.write_model(str(tmp_path / "ours"), seed=14)
    pr = objgen.write_model(str(tmp_path / "ref"), seed=14)
    out = tmp_path / "dump.bin"
    argv = [exe, po, "0", str(out)] if example == "shim_obj_example" else [exe, po, str(out)]
    res = subprocess.run(argv, capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr + res.stdout
    assert ("meshes 5 materials 5" if example == "shim_obj_example" else "draws 5 ") in res.stdout
    raw = out.read_bytes()
    W, H, nt, nm = (int(v) for v in np.frombuffer(raw, np.uint32, 4))
    mvp = np.frombuffer(raw, np.float32, 16, 16).copy()
    colour = np.frombuffer(raw, np.uint32, nt * 4096, 80).reshape(nt, 64, 64)
    depth = np.frombuffer(raw, np.uint32, nt * 4096, 80 + nt * 16384).reshape(nt, 64, 64)

    meshes, mats = rh.ref_load_model(pr, 0)
    sc = scenes.Scene("obj_ref", W, H, clear_color=0)
    tex_of = {}
    for i, m in enumerate(mats):
        if m["texels"].size:
            tex_of[i] = len(sc.textures)
            sc.textures.append(scenes.TiledTexture(m["texels"], m["mip_offsets"], m["num_mips"], m["width_log2"], m["height_log2"]))
    for m in meshes:
        sc.draws.append(scenes.Draw(m["vertices"], m["indices"], mvp, 0 if m["material"] < len(mats) else 1,
                                    tex_of.get(m["material"], -1), 6))
    r = rh.RefRenderer(W, H, 1, "parity")
    try:
        r.load_scene(sc)
        r.render()
        rc, rd = r.read_tiles()
    finally:
        r.close()
    assert (rd > 0).mean() > 0.1
    assert np.array_equal(depth, rd.view(np.uint32))
    assert np.array_equal(colour, rc)


def test_hall_scene_through_obj_files(tmp_path):
    """BASELINE configs[1] geometry (263 888 triangles, 25 draws) written out as OBJ + MTL +
    PNG, loaded by srb_model_load, made resident and rendered at 1920x1080: the frame is the one the reference renders
    from its own loader's arrays."""
    import time

    from oracle import refharness as rh
    from softrast_b200 import capi

    hall = scenes.hall_scene(1920, 1080)
    images = [objgen.procedural_rgba(64 if i % 2 else 32, 100 + i) for i in range(len(hall.textures))]
    po = objgen.write_scene_as_obj(str(tmp_path / "ours"), hall, images)
    pr = objgen.write_scene_as_obj(str(tmp_path / "ref"), hall, images)
    t0 = time.perf_counter()
    model = capi.Model(po, capi.OBJ_NO_CACHE_WRITE)
    t_ours = time.perf_counter() - t0
    t0 = time.perf_counter()
    meshes, mats = rh.ref_load_model(pr, 0)
    t_ref = time.perf_counter() - t0
    print(f"hall as OBJ ({os.path.getsize(po) >> 20} MiB of text): srb_model_load {t_ours:.2f} s, reference Obj::Model::Load {t_ref:.2f} s")
    assert len(model.meshes) == len(hall.draws) == len(meshes)
    assert sum(m["indices"].size for m in model.meshes) == hall.num_tris * 3
    for a, b in zip(model.meshes, meshes):
        assert np.array_equal(a["indices"], b["indices"]) and np.array_equal(a["vertices"].view(np.uint32), b["vertices"].view(np.uint32))

    mvp = hall.draws[0].mvp
    colour, depth, counters, n = _render_resident(model, 1920, 1080, mvp, 0, 0)
    assert n == len(hall.draws) and counters["tris_in"] == hall.num_tris and counters["overflow"] == 0
    colour2, depth2, _, _ = _render_resident(model, 1920, 1080, mvp, 0, 0)  # a fresh context gives the same frame
    assert np.array_equal(colour, colour2) and np.array_equal(depth.view(np.uint32), depth2.view(np.uint32))

    sc = scenes.Scene("hall_obj_ref", 1920, 1080, clear_color=0)
    tex_of = {}
    for i, m in enumerate(mats):
        if m["texels"].size:
            tex_of[i] = len(sc.textures)
            sc.textures.append(scenes.TiledTexture(m["texels"], m["mip_offsets"], m["num_mips"], m["width_log2"], m["height_log2"]))
    for m in meshes:
        sc.draws.append(scenes.Draw(m["vertices"], m["indices"], mvp, 0, tex_of.get(m["material"], -1), 6))
    # one reference thread = the canonical triangle order: the hall has exact depth ties (coplanar triangles where arches
    # meet columns), and the multi-threaded reference's winner on a tie depends on its thread scheduling
    r = rh.RefRenderer(1920, 1080, 1, "parity")
    try:
        r.load_scene(sc)
        r.render()
        rc, rd = r.read_tiles()
    finally:
        r.close()
    assert np.array_equal(depth.view(np.uint32), rd.view(np.uint32))
    same = colour == rc
    assert same.all(), f"{(~same).sum()} pixels differ"
    model.close()


def test_model_load_builds_textures_on_the_device(tmp_path):
    """srb_model_load_on: with a context, the materials' textures are tiled and mip-mapped by the device builder — the
    model (and the cache file it writes) is byte-identical to the host-built one, and a Sponza-sized texture set loads
    much sooner."""
    import time

    from softrast_b200 import capi

    ctx = capi.RenderContext(0)
    try:
        pa = objgen.write_model(str(tmp_path / "a"), seed=17)
        pb = objgen.write_model(str(tmp_path / "b"), seed=17)
        host = capi.Model(pa, 0)
        dev = capi.Model(pb, 0, ctx=ctx)
        assert len(host.materials) == len(dev.materials) == 5
        for a, b in zip(host.materials, dev.materials):
            assert a["name"] == b["name"] and a["num_mips"] == b["num_mips"] and np.array_equal(a["mip_offsets"], b["mip_offsets"])
            assert np.array_equal(a["texels"], b["texels"])
        for a, b in zip(host.meshes, dev.meshes):
            assert np.array_equal(a["indices"], b["indices"]) and np.array_equal(a["vertices"].view(np.uint32), b["vertices"].view(np.uint32))
        assert open(pa + ".bin", "rb").read() == open(pb + ".bin", "rb").read()
        host.close()
        dev.close()

        # eight 1024^2 textures, one quad each
        d = tmp_path / "big"
        d.mkdir()
        rng = np.random.default_rng(3)
        obj, mtl = ["mtllib big.mtl", "v -1 -1 3", "v 1 -1 3", "v 1 1 3", "v -1 1 3", "vt 0 0", "vt 1 0", "vt 1 1", "vt 0 1"], []
        for i in range(8):
            objgen.write_png_rgba8_fast(str(d / f"t{i}.png"), rng.integers(0, 256, (1024, 1024, 4)).astype(np.uint8))
            mtl += [f"newmtl m{i}", f"map_Kd t{i}.png"]
            obj += [f"usemtl m{i}", f"g q{i}", "f 1/1 2/2 3/3 4/4"]
        (d / "big.mtl").write_text("\n".join(mtl) + "\n")
        (d / "big.obj").write_text("\n".join(obj) + "\n")
        t0 = time.perf_counter()
        m_dev = capi.Model(str(d / "big.obj"), capi.OBJ_NO_CACHE_WRITE, ctx=ctx)
        t_dev = time.perf_counter() - t0
        t0 = time.perf_counter()
        m_host = capi.Model(str(d / "big.obj"), capi.OBJ_NO_CACHE_WRITE)
        t_host = time.perf_counter() - t0
        print(f"OBJ with eight 1024x1024 PNG textures: srb_model_load_on (device-built mips) {t_dev:.2f} s, srb_model_load (host) {t_host:.2f} s")
        for a, b in zip(m_host.materials, m_dev.materials):
            assert np.array_equal(a["texels"], b["texels"])
        assert t_dev < t_host
        m_dev.close()
        m_host.close()
    finally:
        ctx.close()
