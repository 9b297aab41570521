// softrast_b200/Obj.h — source-compatible C++ shim of the reference's OBJ model (Viewer/Obj.h:13-71) over the C ABI's
// srb_model_* (scene ingestion).  Same names and members as the reference, so scene code written like
// Viewer/Scene.cpp:8-11,35-63 (m_model.Load(path, allocator, flags); for (Mesh const& mesh : m_model.m_meshes) ...
// mesh.m_vertexData.Data() / .Size(), mesh.m_indexData.Data(), mesh.m_numIndices, mesh.m_indexType, mesh.m_matIdx,
// m_model.m_materials[i].m_diffuse) compiles unchanged.  Parsing, the `.bin` cache and the texture build happen in the
// library (softrast_b200/csrc/srb_model.cpp); this header only copies the results into reference-shaped containers.
#pragma once
#include <string>

#include "Renderer.h"

namespace sr
{

enum class IndexType // SoftRast/SoftRastTypes.h:15-19
{
	u16,
	u32
};

namespace Obj
{

// the subset of kt::Array scene code uses (kt/src/kt/Array.h): Data(), Size(), operator[], range-for
template <typename T>
struct Array : std::vector<T>
{
	T* Data() { return this->data(); }
	T const* Data() const { return this->data(); }
	uint32_t Size() const { return uint32_t(this->size()); }
};

struct Vec3 { float x, y, z; };
struct Vec2 { float x, y; };

struct Vertex // Obj.h:16-21
{
	Vec3 pos;
	Vec3 norm;
	Vec2 uv;
};
static_assert(sizeof(Vertex) == 32, "sr::Obj::Vertex is 32 bytes");

struct Mesh // Obj.h:24-43
{
	void Clear()
	{
		m_vertexData.clear();
		m_indexData.clear();
	}
	Array<uint8_t> m_indexData;
	IndexType m_indexType = IndexType::u16;
	uint32_t m_numIndices = 0;
	Array<Vertex> m_vertexData;
	uint32_t m_matIdx = 0;
};

struct Name : std::string // kt::String128
{
	char const* Data() const { return c_str(); }
	uint32_t Size() const { return uint32_t(size()); }
};

struct Material // Obj.h:45-53
{
	Material() = default;
	Material(Material&&) = default;
	Material& operator=(Material&&) = default;
	Name m_name;
	Tex::TextureData m_diffuse;
};

enum LoadFlags : uint32_t // Obj.h:55-61
{
	None = 0x0,
	FlipWinding = SRB_OBJ_FLIP_WINDING,
	GenNormals = SRB_OBJ_GEN_NORMALS, // todo in the reference; ignored
	FlipUVs = SRB_OBJ_FLIP_UVS
};

struct Model // Obj.h:63-72
{
	// Obj.cpp:374-560.  _tempAllocator (a kt::IAllocator* in the reference) is not needed and ignored.
	bool Load(char const* _path, void* /*_tempAllocator*/, uint32_t const _flags)
	{
		srb_model* m = nullptr;
		if (srb_model_load(_path, _flags, &m) != SRB_OK)
		{
			fprintf(stderr, "softrast_b200: %s\n", srb_model_last_error());
			Clear();
			return false;
		}
		uint32_t numMeshes = 0, numMaterials = 0;
		srb_model_info(m, &numMeshes, &numMaterials, nullptr);
		m_meshes.clear();
		m_meshes.resize(numMeshes);
		for (uint32_t i = 0; i < numMeshes; ++i)
		{
			srb_mesh_view v;
			srb_model_mesh(m, i, &v);
			Mesh& mesh = m_meshes[i];
			mesh.m_indexType = v.index_stride == 2 ? IndexType::u16 : IndexType::u32;
			mesh.m_numIndices = v.num_indices;
			uint8_t const* idx = (uint8_t const*)v.indices;
			mesh.m_indexData.assign(idx, idx + size_t(v.num_indices) * v.index_stride);
			Vertex const* vb = (Vertex const*)v.vertices;
			mesh.m_vertexData.assign(vb, vb + v.num_vertices);
			mesh.m_matIdx = v.material;
		}
		m_materials.clear();
		m_materials.resize(numMaterials);
		for (uint32_t i = 0; i < numMaterials; ++i)
		{
			srb_material_view v;
			srb_model_material(m, i, &v);
			Material& mat = m_materials[i];
			mat.m_name.assign(v.name);
			mat.m_diffuse.Clear();
			mat.m_diffuse.m_texels.assign(v.texels, v.texels + v.texel_bytes);
			memcpy(mat.m_diffuse.m_mipOffsets, v.mip_offsets, sizeof(v.mip_offsets));
			mat.m_diffuse.m_numMips = v.num_mips;
			mat.m_diffuse.m_widthLog2 = v.width_log2;
			mat.m_diffuse.m_heightLog2 = v.height_log2;
			mat.m_diffuse.m_bytesPerPixel = v.bytes_per_pixel;
		}
		srb_model_free(m);
		return true;
	}
	bool Load(char const* _path, uint32_t const _flags = 0) { return Load(_path, nullptr, _flags); }
	void Clear() { m_meshes.clear(); } // Obj.cpp:562-569 (materials stay, as in the reference)

	Array<Mesh> m_meshes;
	Array<Material> m_materials;
};

} // namespace Obj
} // namespace sr
