// A scene written against the reference's API exactly like Viewer/Scene.cpp:32-65 (SimpleModelScene::Update) and
// Viewer/Main.cpp:50-69 (BeginFrame / Update / EndFrame / Blit), compiled against the drop-in shim.
// Prints an FNV-1a hash of the colour tiles, the depth tiles and the blitted image; tests/test_gpu_shim.py compares them
// with the same scene rendered by the reference renderer.
#include <softrast_b200/Renderer.h>

#include <atomic>
#include <math.h>
#include <vector>

struct Vertex // sr::Obj::Vertex, Viewer/Obj.h:16-21
{
	float pos[3], norm[3], uv[2];
};

static uint64_t Fnv(const void* p, size_t n, uint64_t h = 1469598103934665603ull)
{
	const uint8_t* b = (const uint8_t*)p;
	for (size_t i = 0; i < n; ++i)
	{
		h = (h ^ b[i]) * 1099511628211ull;
	}
	return h;
}

int main(int argc, char** argv)
{
	uint32_t const W = 320, H = 200;
	// a tessellated, tilted quad + its mirror, textured
	std::vector<Vertex> verts;
	std::vector<uint16_t> idx;
	int const N = 12;
	for (int j = 0; j <= N; ++j)
	{
		for (int i = 0; i <= N; ++i)
		{
			float const s = float(i) / N, t = float(j) / N;
			Vertex v = {{-3.0f + 6.0f * s, -2.0f + 4.0f * t, 4.0f + 3.0f * s + 0.5f * t}, {0.0f, 0.0f, -1.0f}, {3.0f * s, 2.0f * t}};
			verts.push_back(v);
		}
	}
	for (int j = 0; j < N; ++j)
	{
		for (int i = 0; i < N; ++i)
		{
			uint16_t const a = uint16_t(j * (N + 1) + i), b = a + 1, c = uint16_t(a + N + 1), d = c + 1;
			uint16_t const q[6] = {a, b, d, a, d, c};
			idx.insert(idx.end(), q, q + 6);
		}
	}
	std::vector<uint8_t> rgba(64 * 64 * 4);
	for (int y = 0; y < 64; ++y)
	{
		for (int x = 0; x < 64; ++x)
		{
			uint8_t* p = &rgba[(y * 64 + x) * 4];
			p[0] = uint8_t(x * 4);
			p[1] = uint8_t(y * 4);
			p[2] = uint8_t(((x / 8 + y / 8) & 1) * 255);
			p[3] = 255;
		}
	}
	sr::Tex::TextureData diffuse;
	diffuse.CreateFromRGBA8(rgba.data(), 64, 64, true);

	// kt::Mat4::PerspectiveLH_ZO(85 deg, W/H, near = 10000, far = 0.1) — reverse-Z as in Viewer/Scene.cpp:16-29
	float mvp[16] = {0};
	float const f = tanf(1.57079632679f - 85.0f * 0.01745329252f * 0.5f);
	float const range = 0.1f / (0.1f - 10000.0f);
	mvp[0] = f / (float(W) / float(H));
	mvp[5] = f;
	mvp[10] = range;
	mvp[11] = 1.0f;
	mvp[14] = -range * 10000.0f;
	struct Mat4 { float m[16]; } m;
	memcpy(m.m, mvp, sizeof(mvp));

	sr::RenderContext ctx;
	sr::FrameBuffer fb(W, H);
	std::vector<uint8_t> linear(size_t(W) * H * 4);
	std::atomic<int> flipped{0};

	ctx.BeginFrame();
	ctx.ClearFrameBuffer(fb, 0x30);
	for (int pass = 0; pass < 2; ++pass)
	{
		sr::DrawCall call;
		call.SetFrameBuffer(&fb).SetMVP(m);
		call.SetAttributeBuffer(verts.data(), sizeof(Vertex), (uint32_t)verts.size(), offsetof(Vertex, uv) / sizeof(float));
		call.SetPositionBuffer(verts.data(), sizeof(Vertex), (uint32_t)verts.size());
		call.SetIndexBuffer(idx.data(), sizeof(uint16_t), (uint32_t)idx.size());
		if (pass == 0)
		{
			call.SetPixelShader(sr::shader::UnlitDiffuseShader, &diffuse);
		}
		else
		{
			call.SetPixelShader(sr::shader::VisualizeNormalsShader, nullptr);
			call.m_indexBuffer.m_num /= 2; // second draw: first half of the triangles again -> exact depth ties
		}
		ctx.DrawIndexed(call);
	}
	ctx.EndFrame();
	sr::FrameBufferPlane* p = fb.WritePlane();
	size_t const n = size_t(p->m_tilesX) * p->m_tilesY;
	uint64_t hc = 1469598103934665603ull, hd = hc;
	for (size_t i = 0; i < n; ++i)
	{
		hc = Fnv(p->m_colourTiles[i].m_colour, sizeof(p->m_colourTiles[i].m_colour), hc);
		hd = Fnv(p->m_depthTiles[i].m_depth, sizeof(p->m_depthTiles[i].m_depth), hd);
	}
	ctx.Blit(fb, linear.data(), [](void* u) { ((std::atomic<int>*)u)->store(1); }, &flipped);
	srb_sync(ctx.Native());
	printf("colour %016llx depth %016llx linear %016llx flipped %d tiles %zu\n", (unsigned long long)hc, (unsigned long long)hd,
	       (unsigned long long)Fnv(linear.data(), linear.size()), flipped.load(), n);
	if (argc > 1)
	{
		// inputs + outputs for tests/test_gpu_shim.py: u32 header {W, H, nVerts, nIdx, texSize, nTiles}, then the arrays
		FILE* fo = fopen(argv[1], "wb");
		uint32_t const hdr[6] = {W, H, (uint32_t)verts.size(), (uint32_t)idx.size(), 64u, (uint32_t)n};
		fwrite(hdr, sizeof(hdr), 1, fo);
		fwrite(verts.data(), sizeof(Vertex), verts.size(), fo);
		fwrite(idx.data(), sizeof(uint16_t), idx.size(), fo);
		fwrite(rgba.data(), 1, rgba.size(), fo);
		fwrite(mvp, sizeof(mvp), 1, fo);
		for (size_t i = 0; i < n; ++i) fwrite(p->m_colourTiles[i].m_colour, 1, sizeof(p->m_colourTiles[i].m_colour), fo);
		for (size_t i = 0; i < n; ++i) fwrite(p->m_depthTiles[i].m_depth, 1, sizeof(p->m_depthTiles[i].m_depth), fo);
		fwrite(linear.data(), 1, linear.size(), fo);
		fclose(fo);
	}
	ctx.Shutdown();
	return 0;
}
