"""The C++ drop-in shim (include/softrast_b200/Renderer.h): a scene written against the reference's class API
(tests/cpp/shim_example.cpp, modelled on Viewer/Scene.cpp + Viewer/Main.cpp) runs on the GPU and gives the reference's
pixels."""
import os
import subprocess

import numpy as np
import pytest

from softrast_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "shim_example")


def test_shim_example_compiles_against_header():
    """CPU: the example must compile and link against the shim + library (built by __graft_entry__.build())."""
    assert os.path.exists(EXE), "run __graft_entry__.build()"


@pytest.mark.gpu
def test_shim_example_matches_reference(tmp_path):
    from oracle.refharness import RefRenderer
    from softrast_b200.capi import MIPS_STB, build_texture

    out = tmp_path / "dump.bin"
    res = subprocess.run([EXE, str(out)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    assert "flipped 1" in res.stdout
    raw = out.read_bytes()
    W, H, nv, ni, ts, nt = np.frombuffer(raw, np.uint32, 6)
    o = 24
    verts = np.frombuffer(raw, np.float32, nv * 8, o).reshape(nv, 8); o += nv * 32
    idx = np.frombuffer(raw, np.uint16, ni, o); o += ni * 2
    rgba = np.frombuffer(raw, np.uint8, ts * ts * 4, o).reshape(ts, ts, 4); o += ts * ts * 4
    mvp = np.frombuffer(raw, np.float32, 16, o); o += 64
    colour = np.frombuffer(raw, np.uint32, nt * 4096, o).reshape(nt, 64, 64); o += nt * 16384
    depth = np.frombuffer(raw, np.uint32, nt * 4096, o).reshape(nt, 64, 64); o += nt * 16384
    linear = np.frombuffer(raw, np.uint32, W * H, o).reshape(H, W)

    sc = scenes.Scene("shim", int(W), int(H), clear_color=0x30)
    # the shim's TextureData::CreateFromRGBA8(..., calcMips) builds the reference's own (stb) mips
    sc.textures.append(build_texture(rgba, MIPS_STB))
    sc.draws.append(scenes.Draw(verts.copy(), idx.copy(), mvp.copy(), scenes.SHADER_UNLIT_DIFFUSE, 0))
    sc.draws.append(scenes.Draw(verts.copy(), idx[: ni // 2].copy(), mvp.copy(), scenes.SHADER_VISUALIZE_NORMALS, -1))
    r = RefRenderer(sc.width, sc.height, 1, "parity")
    try:
        r.load_scene(sc)
        r.render()
        rc, rd = r.read_tiles()
        assert (rd > 0).mean() > 0.05, "the example scene must cover part of the screen"
        assert np.array_equal(depth, rd.view(np.uint32))
        assert np.array_equal(colour, rc)
        assert np.array_equal(linear, r.blit_linear())  # the reference's own Blit (Renderer.cpp:319-372)
    finally:
        r.close()
