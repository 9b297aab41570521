"""softrast_b200 — B200-native sort-middle frame pipeline behind the SoftRast Renderer API.

The product is the C-ABI shared library `softrast_b200/lib/libsoftrast_b200.so` (CUDA, sm_100a; sources in
`softrast_b200/csrc`, interface in `include/softrast_b200.h`, C++ drop-in shim in `include/softrast_b200/Renderer.h`).
This Python package is only a thin ctypes driver over that ABI for tests and benchmarks, plus the synthetic scene
generators.  There is no CPU fallback: importing `softrast_b200.capi` without the built library raises.
"""
