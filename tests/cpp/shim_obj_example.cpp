// The reference's SimpleModelScene (Viewer/Scene.cpp:8-11 constructor, :32-65 Update) written against the drop-in shim:
// loads an OBJ with sr::Obj::Model::Load, draws one call per mesh exactly as Scene.cpp does, and dumps the colour and
// depth tiles.  tests/test_gpu_obj.py compares them with the reference renderer fed by the reference's own loader.
//   usage: shim_obj_example <model.obj> <load flags> <out.bin>
#include <softrast_b200/Obj.h>
#include <softrast_b200/Renderer.h>

#include <math.h>
#include <stddef.h>
#include <vector>

int main(int argc, char** argv)
{
	if (argc < 4) return 2;
	uint32_t const W = 448, H = 256;
	sr::Obj::Model m_model;
	if (!m_model.Load(argv[1], nullptr, uint32_t(atoi(argv[2])))) return 3;

	// kt::Mat4::PerspectiveLH_ZO(85 deg, W/H, near = 10000, far = 0.1) (Scene.cpp:16-29), camera at the origin
	float mvp[16] = {0};
	float const f = tanf(1.57079632679f - 85.0f * 0.01745329252f * 0.5f);
	float const range = 0.1f / (0.1f - 10000.0f);
	mvp[0] = f / (float(W) / float(H));
	mvp[5] = f;
	mvp[10] = range;
	mvp[11] = 1.0f;
	mvp[14] = -range * 10000.0f;
	struct Mat4 { float m[16]; } viewProj;
	memcpy(viewProj.m, mvp, sizeof(mvp));

	sr::RenderContext _ctx;
	sr::FrameBuffer _fb(W, H);
	_ctx.BeginFrame();
	_ctx.ClearFrameBuffer(_fb, 0);
	for (sr::Obj::Mesh const& mesh : m_model.m_meshes)
	{
		sr::DrawCall call;
		call.SetFrameBuffer(&_fb);
		call.SetMVP(viewProj);

		call.SetAttributeBuffer(mesh.m_vertexData.Data(), sizeof(sr::Obj::Vertex), mesh.m_vertexData.Size(), offsetof(sr::Obj::Vertex, uv) / sizeof(float));

		call.m_positionBuffer.m_ptr = (uint8_t*)mesh.m_vertexData.Data();
		call.m_positionBuffer.m_stride = sizeof(sr::Obj::Vertex);
		call.m_positionBuffer.m_num = mesh.m_vertexData.Size();

		call.m_indexBuffer.m_ptr = mesh.m_indexData.Data();
		call.m_indexBuffer.m_num = mesh.m_numIndices;
		call.m_indexBuffer.m_stride = mesh.m_indexType == sr::IndexType::u16 ? sizeof(uint16_t) : sizeof(uint32_t);

		if (mesh.m_matIdx < m_model.m_materials.Size())
		{
			call.m_pixelUniforms = &m_model.m_materials[mesh.m_matIdx].m_diffuse;
			call.m_pixelShader = sr::shader::UnlitDiffuseShader;
		}
		else
		{
			call.m_pixelShader = sr::shader::VisualizeNormalsShader;
		}

		_ctx.DrawIndexed(call);
	}
	_ctx.EndFrame();

	sr::FrameBufferPlane const* plane = _fb.WritePlane();
	uint32_t const tiles = plane->m_tilesX * plane->m_tilesY;
	FILE* out = fopen(argv[3], "wb");
	if (!out) return 4;
	uint32_t const hdr[4] = {W, H, tiles, m_model.m_meshes.Size()};
	fwrite(hdr, sizeof(hdr), 1, out);
	fwrite(mvp, sizeof(mvp), 1, out);
	for (uint32_t t = 0; t < tiles; ++t) fwrite(plane->m_colourTiles[t].m_colour, 1, 16384, out);
	for (uint32_t t = 0; t < tiles; ++t) fwrite(plane->m_depthTiles[t].m_depth, 1, 16384, out);
	fclose(out);
	printf("meshes %u materials %u tiles %u\n", m_model.m_meshes.Size(), m_model.m_materials.Size(), tiles);
	return 0;
}
