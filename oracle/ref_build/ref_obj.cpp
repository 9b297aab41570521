/*
 * ref_obj.cpp — TEST INFRASTRUCTURE ONLY (oracle/_ref build).
 * Compiles the reference's OBJ loader (Viewer/Obj.cpp) and the two kt translation units it needs, in place, and puts a
 * C ABI around sr::Obj::Model so tests can compare the product's srb_model_* against the reference itself: mesh
 * arrays, material textures (stbi_load + CreateFromRGBA8 mips) and the `.bin` cache bytes.
 * Two portability problems are worked around by the preprocessor instead of by editing the files:
 *   - kt/src/kt/Logging.h:4-6: KT_LOG_*(fmt, ...) expand to f(fmt, __VA_ARGS__), which only MSVC accepts with no
 *     variadic arguments; redefined with GNU ##__VA_ARGS__ (Logging.h is #pragma once, so the redefinition sticks).
 *   - kt/src/kt/File.cpp:55-64: the POSIX branch of kt::FileExists tests `stat(...) == 1`, which never holds, so on Linux
 *     the cache would never be read (Obj.cpp:379).  The broken definition is renamed away and one with the Windows
 *     branch's meaning is supplied.
 */
#include <sys/stat.h>

#include <kt/Logging.h>
#undef KT_LOG_ERROR
#undef KT_LOG_WARNING
#undef KT_LOG_INFO
#define KT_LOG_ERROR(fmt, ...) kt::LogError(fmt, ##__VA_ARGS__)
#define KT_LOG_WARNING(fmt, ...) kt::LogWarning(fmt, ##__VA_ARGS__)
#define KT_LOG_INFO(fmt, ...) kt::LogInfo(fmt, ##__VA_ARGS__)

#include <kt/File.h>
#define FileExists FileExists_reference_posix
#include "kt/src/kt/File.cpp"
#undef FileExists
namespace kt
{
bool FileExists(char const* _name)
{
	struct stat buf;
	return stat(_name, &buf) == 0;
}
}
#include "kt/src/kt/FilePath.cpp"

#include "Viewer/Obj.cpp"

#include "SoftRast/stb_image.h"
#include "../../include/softrast_b200.h"

extern "C"
{

SRB_API int srref_model_load(const char* path, uint32_t flags, void** out)
{
	sr::Obj::Model* m = new sr::Obj::Model();
	bool const ok = m->Load(path, kt::GetDefaultAllocator(), flags);
	if (!ok)
	{
		delete m;
		*out = nullptr;
		return SRB_ERR_INVALID;
	}
	*out = m;
	return SRB_OK;
}

SRB_API void srref_model_free(void* h) { delete static_cast<sr::Obj::Model*>(h); }

SRB_API int srref_model_info(void* h, uint32_t* numMeshes, uint32_t* numMaterials)
{
	sr::Obj::Model* m = static_cast<sr::Obj::Model*>(h);
	*numMeshes = m->m_meshes.Size();
	*numMaterials = m->m_materials.Size();
	return SRB_OK;
}

SRB_API int srref_model_mesh(void* h, uint32_t i, srb_mesh_view* out)
{
	sr::Obj::Model* m = static_cast<sr::Obj::Model*>(h);
	if (i >= m->m_meshes.Size()) return SRB_ERR_INVALID;
	sr::Obj::Mesh& mesh = m->m_meshes[i];
	out->indices = mesh.m_indexData.Data();
	out->index_stride = mesh.m_indexType == sr::IndexType::u16 ? 2u : 4u;
	out->num_indices = mesh.m_numIndices;
	out->vertices = mesh.m_vertexData.Data();
	out->num_vertices = mesh.m_vertexData.Size();
	out->material = mesh.m_matIdx;
	return SRB_OK;
}

SRB_API int srref_model_material(void* h, uint32_t i, srb_material_view* out)
{
	sr::Obj::Model* m = static_cast<sr::Obj::Model*>(h);
	if (i >= m->m_materials.Size()) return SRB_ERR_INVALID;
	sr::Obj::Material& mat = m->m_materials[i];
	out->name = mat.m_name.Data();
	out->texels = mat.m_diffuse.m_texels.Data();
	out->texel_bytes = mat.m_diffuse.m_texels.Size();
	memcpy(out->mip_offsets, mat.m_diffuse.m_mipOffsets, sizeof(out->mip_offsets));
	out->num_mips = mat.m_diffuse.m_numMips;
	out->width_log2 = mat.m_diffuse.m_widthLog2;
	out->height_log2 = mat.m_diffuse.m_heightLog2;
	out->bytes_per_pixel = mat.m_diffuse.m_bytesPerPixel;
	return SRB_OK;
}

/* stbi_load(path, &x, &y, &comp, 4) exactly as Texture.cpp:107 calls it; free with srref_image_free */
SRB_API int srref_image_load_rgba8(const char* path, uint8_t** rgba, uint32_t* w, uint32_t* ht)
{
	int x = 0, y = 0, comp = 0;
	uint8_t* px = stbi_load(path, &x, &y, &comp, 4);
	if (!px) return SRB_ERR_INVALID;
	*rgba = px;
	*w = uint32_t(x);
	*ht = uint32_t(y);
	return SRB_OK;
}

SRB_API void srref_image_free(uint8_t* rgba) { stbi_image_free(rgba); }
}
