"""ctypes mirrors of the PODs in include/softrast_b200.h (srb_buffer_ref, srb_draw_desc, srb_counters, srb_tile_tri).

Shared by the product binding (softrast_b200.capi) and by the test-only bindings under oracle/ — the structs are the
C ABI, not an implementation."""
import ctypes as C

import numpy as np

SRB_OK = 0
SRB_ERR_CUDA = 1
SRB_ERR_INVALID = 2
SRB_ERR_OVERFLOW = 3
SRB_ERR_UNKNOWN_SHADER = 4
SRB_ERR_NO_DEVICE = 5

SHADER_UNLIT_DIFFUSE = 0
SHADER_VISUALIZE_NORMALS = 1
SHADER_VISUALIZE_UVS = 2

BIN_DIM = 64
COLOUR_TILE_BYTES = 16384
DEPTH_TILE_BYTES = 16416
MAX_VARYINGS = 8
MAX_TEX_DIM_LOG2 = 14


class BufferRef(C.Structure):
    _fields_ = [
        ("buffer", C.c_uint64),
        ("offset", C.c_uint64),
        ("host", C.c_void_p),
        ("stride", C.c_uint32),
        ("num", C.c_uint32),
    ]


class DrawDesc(C.Structure):
    _fields_ = [
        ("shader", C.c_uint32),
        ("uv_offset", C.c_uint32),
        ("texture", C.c_uint64),
        ("framebuffer", C.c_uint64),
        ("indices", BufferRef),
        ("positions", BufferRef),
        ("attributes", BufferRef),
        ("mvp", C.c_float * 16),
    ]


class Counters(C.Structure):
    _fields_ = [
        ("tris_in", C.c_uint64),
        ("tris_setup", C.c_uint64),
        ("tris_clipped", C.c_uint64),
        ("tile_refs", C.c_uint64),
        ("tiles_nonempty", C.c_uint64),
        ("max_refs_in_tile", C.c_uint64),
        ("pixels_covered", C.c_uint64),
        ("overflow", C.c_uint64),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


# numpy view of srb_tile_tri (168 bytes)
TILE_TRI_DTYPE = np.dtype(
    [
        ("c", "<i4", (3,)),
        ("dx", "<i4", (3,)),
        ("dy", "<i4", (3,)),
        ("block", "u1", (4,)),  # min_x, max_x, min_y, max_y
        ("recip_w", "<f4", (3,)),  # c0, dx, dy
        ("z_over_w", "<f4", (3,)),
        ("attr_dx", "<f4", (8,)),
        ("attr_dy", "<f4", (8,)),
        ("attr_c", "<f4", (8,)),
        ("attribs_per_tri", "<u4"),
        ("draw_idx", "<u4"),
    ]
)
assert TILE_TRI_DTYPE.itemsize == 168


def ptr(a):
    """void* of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class MeshView(C.Structure):  # srb_mesh_view
    _fields_ = [
        ("indices", C.c_void_p),
        ("index_stride", C.c_uint32),
        ("num_indices", C.c_uint32),
        ("vertices", C.c_void_p),
        ("num_vertices", C.c_uint32),
        ("material", C.c_uint32),
    ]


class MaterialView(C.Structure):  # srb_material_view
    _fields_ = [
        ("name", C.c_char_p),
        ("texels", C.c_void_p),
        ("texel_bytes", C.c_uint64),
        ("mip_offsets", C.c_uint32 * MAX_TEX_DIM_LOG2),
        ("num_mips", C.c_uint32),
        ("width_log2", C.c_uint32),
        ("height_log2", C.c_uint32),
        ("bytes_per_pixel", C.c_uint32),
    ]


def copy_mesh_view(v: "MeshView") -> dict:
    """Copies the arrays a srb_mesh_view points at into numpy (indices as u16/u32, vertices as float32 (N, 8))."""
    nb = v.num_indices * v.index_stride
    idx = np.frombuffer(C.string_at(v.indices, nb), dtype=np.uint16 if v.index_stride == 2 else np.uint32).copy() if nb else np.zeros(0, np.uint16)
    vb = v.num_vertices * 32
    verts = np.frombuffer(C.string_at(v.vertices, vb), dtype=np.float32).reshape(-1, 8).copy() if vb else np.zeros((0, 8), np.float32)
    return {"indices": idx, "vertices": verts, "material": int(v.material), "index_stride": int(v.index_stride)}


def copy_material_view(v: "MaterialView") -> dict:
    tex = np.frombuffer(C.string_at(v.texels, v.texel_bytes), dtype=np.uint8).copy() if v.texel_bytes else np.zeros(0, np.uint8)
    return {
        "name": (v.name or b"").decode("latin-1"),
        "texels": tex,
        "mip_offsets": np.array(list(v.mip_offsets), dtype=np.uint32),
        "num_mips": int(v.num_mips),
        "width_log2": int(v.width_log2),
        "height_log2": int(v.height_log2),
        "bytes_per_pixel": int(v.bytes_per_pixel),
    }
