"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/softrast_b200.h declares,
its PODs have the documented layout, the host-only entry points work, and without a GPU it fails loudly."""
import ctypes as C
import os
import re

import numpy as np

from softrast_b200 import _ctypes_defs as D
from softrast_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "softrast_b200.h")).read()
    return sorted(set(re.findall(r"SRB_API\s+[\w\s\*]+?\b(srb_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from softrast_b200 import capi

    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(capi.lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(capi._SIGNATURES), set(names) ^ set(capi._SIGNATURES)


def test_pod_layouts():
    assert C.sizeof(D.BufferRef) == 32
    assert C.sizeof(D.DrawDesc) == 4 + 4 + 8 + 8 + 3 * 32 + 64
    assert C.sizeof(D.Counters) == 64
    assert D.TILE_TRI_DTYPE.itemsize == 168


def test_no_cpu_fallback():
    from tests.conftest import HAS_GPU
    from softrast_b200 import capi

    if HAS_GPU:
        return
    h = C.c_void_p()
    assert capi.lib.srb_create(0, 0, C.byref(h)) == D.SRB_ERR_NO_DEVICE
    assert not h


def test_texture_builder_matches_numpy_layout():
    from softrast_b200 import capi

    for size, mips in ((32, False), (64, True), (256, True)):
        rgba = scenes.procedural_rgba(size, size + 1)
        a, b = capi.build_texture(rgba, mips), scenes.build_tiled_texture(rgba, mips)
        assert a.num_mips == b.num_mips and (a.width_log2, a.height_log2) == (b.width_log2, b.height_log2)
        assert np.array_equal(a.mip_offsets, b.mip_offsets)
        assert np.array_equal(a.texels, b.texels)
    wide = np.ascontiguousarray(np.broadcast_to(scenes.procedural_rgba(64, 3)[:32, :, :], (32, 64, 4)))
    a, b = capi.build_texture(wide, True), scenes.build_tiled_texture(wide, True)
    assert np.array_equal(a.texels, b.texels) and a.num_mips == 7


def test_texture_builder_rejects_bad_sizes():
    from softrast_b200 import capi

    n = C.c_uint64()
    assert capi.lib.srb_texture_build_rgba8(None, 48, 32, 1, None, C.byref(n), None, None) == D.SRB_ERR_INVALID
    assert capi.lib.srb_texture_build_rgba8(None, 16, 16, 1, None, C.byref(n), None, None) == D.SRB_ERR_INVALID


def test_rcp_harvest_is_a_small_table():
    from softrast_b200 import capi

    table, bits = capi.harvest_rcp_table(16)
    assert 11 <= bits <= 16 and table.size == 1 << bits
    # RCPPS(1.0) is close to, but need not be, 1.0
    assert abs(float(table[:1].view(np.float32)[0]) - 1.0) < 1e-3


def test_header_is_plain_c_and_the_c_example_links():
    """include/softrast_b200.h must be C (a cgo / JNI / N-API binding includes it from C): compiled here as strict C99, and
    the C example built by __graft_entry__.build() exists."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c",
                          os.path.join(root, "include", "softrast_b200.h")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert os.path.exists(os.path.join(root, "tests", "cpp", "_build", "abi_example")), "run __graft_entry__.build()"
