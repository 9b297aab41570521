/*
 * ref_harness.cpp — TEST INFRASTRUCTURE ONLY.  Builds into oracle/_ref/libsrref_*.so.
 *
 * A C-ABI harness around the UNMODIFIED reference renderer, compiled from the sources where they lie under
 * /root/reference (nothing is copied).  It drives the reference through its own public API
 * (sr::RenderContext / sr::DrawCall / sr::FrameBuffer, SoftRast/Renderer.h:119-177; pixel shaders from
 * Viewer/Shaders.h:71-130; textures via Tex::TextureData, SoftRast/Texture.h:21-41) and exposes
 *   - frame rendering (the parity pin for depth/colour tiles, and the timed CPU baseline),
 *   - the per-tile sorted BinChunk contents (what Rasterizer.cpp:538-553 builds),
 *   - per-triangle coverage masks and the ordered fragment stream, obtained by calling the reference's own
 *     static RasterizeTrisInBin_OutputFragments (Rasterizer.cpp:194-304) — possible because this translation unit
 *     #includes Rasterizer.cpp.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * How the reference is adapted without editing it:
 *   - `#define private public` gives access to RenderContext::m_binner / m_taskSystem so the binner can be re-sized
 *     to the framebuffer at run time (the reference sizes it from the compile-time Config::c_screenWidth/Height,
 *     Renderer.cpp:146-148) and the per-thread arenas can be enlarged (TaskSystem.cpp:34 hard-codes 128 MB, which
 *     overflows on the 1 M-triangle config).
 *   - kt::LogicalCoreCount() has no POSIX branch (kt/src/kt/Concurrency.cpp:187-194); ref_kt_fixups.cpp supplies one
 *     that returns g_srref_logical_cores, so the worker count (Renderer.cpp:141) is chosen at run time:
 *     1 thread == the reference's SR_DEBUG_SINGLE_THREADED build (0 workers) == the canonical triangle order.
 */
#include <vector>
#include <algorithm>
#include <chrono>
#include <thread>
#include <sys/mman.h>
#include <unistd.h>
#include <immintrin.h>

#define private public
#define protected public
#include "SoftRast/Renderer.cpp"
#include "SoftRast/Rasterizer.cpp"
#include "Viewer/Shaders.h"
#undef private
#undef protected

#include "../../include/softrast_b200.h"

extern uint32_t g_srref_logical_cores;
sr::PixelShaderFn* srref_sponza_shader_fn(); // ref_sponza.cpp: the reference's file-static SponzaShader

namespace
{

struct RefTexture
{
	sr::Tex::TextureData tex;
};

struct RefCtx
{
	sr::RenderContext* ctx = nullptr;
	sr::FrameBuffer* fb = nullptr;
	uint32_t threads = 1;
	uint32_t width = 0, height = 0;
	std::vector<RefTexture*> textures;
	std::vector<void*> arenas;
	size_t arenaBytes = 0;
	kt::LinearAllocator scratch;
	void* scratchMem = nullptr;
	size_t scratchBytes = 0;
};

sr::PixelShaderFn* ShaderFromId(uint32_t id)
{
	switch (id)
	{
		case SRB_SHADER_SPONZA: return srref_sponza_shader_fn();
		case SRB_SHADER_UNLIT_DIFFUSE: return sr::shader::UnlitDiffuseShader;
		case SRB_SHADER_VISUALIZE_NORMALS: return sr::shader::VisualizeNormalsShader;
		case SRB_SHADER_VISUALIZE_UVS: return sr::shader::VisualizeUVsShader;
		default: return nullptr;
	}
}

void FillDraw(RefCtx* c, srb_draw_desc const& d, sr::DrawCall& call)
{
	call.SetFrameBuffer(c->fb);
	sr::Tex::TextureData const* tex = nullptr;
	if (d.texture && d.texture <= c->textures.size())
	{
		tex = &c->textures[d.texture - 1]->tex;
	}
	call.SetPixelShader(ShaderFromId(d.shader), tex);
	call.SetIndexBuffer(d.indices.host, d.indices.stride, d.indices.num);
	call.SetPositionBuffer(d.positions.host, d.positions.stride, d.positions.num);
	call.SetAttributeBuffer(d.attributes.host, d.attributes.stride, d.attributes.num, d.uv_offset);
	kt::Mat4 m;
	memcpy(m.Data(), d.mvp, sizeof(float) * 16);
	call.SetMVP(m);
}

// The sorted chunk list of one tile, exactly as RasterAndShadeBin builds it (Rasterizer.cpp:538-553).
void SortedChunks(RefCtx* c, uint32_t tileIdx, std::vector<sr::BinChunk*>& out)
{
	sr::BinContext& b = c->ctx->m_binner;
	uint32_t const tx = tileIdx % b.m_numBinsX;
	uint32_t const ty = tileIdx / b.m_numBinsX;
	out.clear();
	for (uint32_t t = 0; t < b.m_numThreads; ++t)
	{
		sr::ThreadBin& bin = b.LookupThreadBin(t, tx, ty);
		for (uint32_t i = 0; i < bin.m_numChunks; ++i)
		{
			out.push_back(bin.m_binChunks[i]);
		}
	}
	std::stable_sort(out.begin(), out.end(),
	                 [](sr::BinChunk const* a, sr::BinChunk const* b) { return a->m_drawCallIdx < b->m_drawCallIdx; });
}

void CopyTri(sr::BinChunk const& ch, uint32_t t, srb_tile_tri& o)
{
	memset(&o, 0, sizeof(o));
	sr::BinChunk::EdgeEq const& e = ch.m_edgeEq[t];
	for (int i = 0; i < 3; ++i)
	{
		o.c[i] = e.c[i];
		o.dx[i] = e.dx[i];
		o.dy[i] = e.dy[i];
	}
	o.block_min_x = e.blockMinX;
	o.block_max_x = e.blockMaxX;
	o.block_min_y = e.blockMinY;
	o.block_max_y = e.blockMaxY;
	o.recip_w[0] = ch.m_recipW[t].c0;
	o.recip_w[1] = ch.m_recipW[t].dx;
	o.recip_w[2] = ch.m_recipW[t].dy;
	o.z_over_w[0] = ch.m_zOverW[t].c0;
	o.z_over_w[1] = ch.m_zOverW[t].dx;
	o.z_over_w[2] = ch.m_zOverW[t].dy;
	uint32_t const n = ch.m_attribsPerTri;
	for (uint32_t i = 0; i < n && i < SRB_MAX_VARYINGS; ++i)
	{
		o.attr_dx[i] = ch.m_attribsDx[t * n + i];
		o.attr_dy[i] = ch.m_attribsDy[t * n + i];
		o.attr_c[i] = ch.m_attribsC[t * n + i];
	}
	o.attribs_per_tri = n;
	o.draw_idx = ch.m_drawCallIdx;
}

} // namespace

extern "C"
{

SRB_API int srref_create(uint32_t threads, uint32_t width, uint32_t height, uint64_t arena_bytes, void** out)
{
	if (!out || !width || !height)
	{
		return SRB_ERR_INVALID;
	}
	if (threads == 0)
	{
		threads = (uint32_t)sysconf(_SC_NPROCESSORS_ONLN);
	}
	RefCtx* c = new RefCtx;
	c->threads = threads;
	c->width = width;
	c->height = height;
	g_srref_logical_cores = threads; // RenderContext() starts LogicalCoreCount()-1 workers (Renderer.cpp:141)
	c->ctx = new sr::RenderContext();
	c->fb = new sr::FrameBuffer(width, height);

	// Re-size the binner from the framebuffer instead of Config::c_screenWidth/Height (Renderer.cpp:146-148).
	sr::BinContext& b = c->ctx->m_binner;
	kt::Free(b.m_bins);
	b.m_bins = nullptr;
	b.Init(c->ctx->m_taskSystem.TotalThreadsIncludingMainThread(), c->fb->WritePlane()->m_tilesX,
	       c->fb->WritePlane()->m_tilesY);

	// Enlarge the per-thread scratch arenas (TaskSystem.cpp:32-41 gives each thread 128 MB).
	if (arena_bytes == 0)
	{
		arena_bytes = threads == 1 ? (size_t(4) << 30) : (size_t(1) << 30);
	}
	c->arenaBytes = arena_bytes;
	uint32_t const total = c->ctx->m_taskSystem.TotalThreadsIncludingMainThread();
	for (uint32_t i = 0; i < total; ++i)
	{
		void* p = mmap(nullptr, arena_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
		if (p == MAP_FAILED)
		{
			return SRB_ERR_INVALID;
		}
		c->arenas.push_back(p);
		c->ctx->m_taskSystem.m_allocators[i].Init(p, arena_bytes);
	}

	c->scratchBytes = size_t(64) << 20;
	c->scratchMem = malloc(c->scratchBytes);
	c->scratch.Init(c->scratchMem, c->scratchBytes);
	*out = c;
	return SRB_OK;
}

SRB_API void srref_destroy(void* h)
{
	RefCtx* c = (RefCtx*)h;
	if (!c)
	{
		return;
	}
	c->ctx->Shutdown();
	// The reference never frees its arenas/threads; leak the context objects rather than risk its destructors.
	for (RefTexture* t : c->textures)
	{
		delete t;
	}
	delete c->fb;
	free(c->scratchMem);
	for (void* p : c->arenas)
	{
		munmap(p, c->arenaBytes);
	}
	delete c;
}

/* the reference objects behind a handle (for ref_sponza.cpp's SponzaScene::Update, which takes them by reference) */
SRB_API void srref_raw_objects(void* h, void** renderContext, void** frameBuffer)
{
	RefCtx* c = static_cast<RefCtx*>(h);
	*renderContext = c->ctx;
	*frameBuffer = c->fb;
}

SRB_API uint32_t srref_threads(void* h)
{
	return ((RefCtx*)h)->ctx->m_taskSystem.TotalThreadsIncludingMainThread();
}

// Texture from an already tiled blob (same bytes the device library gets).
SRB_API uint64_t srref_texture_create_tiled(void* h, const uint8_t* texels, uint64_t bytes, const uint32_t* mipOffsets,
                                            uint32_t numMips, uint32_t wLog2, uint32_t hLog2)
{
	RefCtx* c = (RefCtx*)h;
	RefTexture* t = new RefTexture;
	t->tex.m_texels.Resize((uint32_t)bytes);
	memcpy(t->tex.m_texels.Data(), texels, bytes);
	memset(t->tex.m_mipOffsets, 0, sizeof(t->tex.m_mipOffsets));
	for (uint32_t i = 0; i < numMips && i < sr::Config::c_maxTexDimLog2; ++i)
	{
		t->tex.m_mipOffsets[i] = mipOffsets[i];
	}
	t->tex.m_numMips = numMips;
	t->tex.m_widthLog2 = wLog2;
	t->tex.m_heightLog2 = hLog2;
	t->tex.m_bytesPerPixel = 4;
	c->textures.push_back(t);
	return c->textures.size();
}

// Texture through the reference's own builder (stb_image_resize mips, Texture.cpp:119-199).
SRB_API uint64_t srref_texture_create_rgba8(void* h, const uint8_t* rgba, uint32_t w, uint32_t ht, int calcMips)
{
	RefCtx* c = (RefCtx*)h;
	RefTexture* t = new RefTexture;
	t->tex.CreateFromRGBA8(rgba, w, ht, calcMips != 0);
	c->textures.push_back(t);
	return c->textures.size();
}

SRB_API uint64_t srref_texture_create_file(void* h, const char* path)
{
	RefCtx* c = (RefCtx*)h;
	RefTexture* t = new RefTexture;
	t->tex.CreateFromFile(path);
	if (t->tex.m_texels.Size() == 0)
	{
		delete t;
		return 0;
	}
	c->textures.push_back(t);
	return c->textures.size();
}

SRB_API int srref_texture_get(void* h, uint64_t tex, uint8_t* texelsOut, uint64_t* bytes, uint32_t* mipOffsets,
                              uint32_t* numMips, uint32_t* wLog2, uint32_t* hLog2)
{
	RefCtx* c = (RefCtx*)h;
	if (!tex || tex > c->textures.size())
	{
		return SRB_ERR_INVALID;
	}
	sr::Tex::TextureData& t = c->textures[tex - 1]->tex;
	if (bytes)
	{
		*bytes = t.m_texels.Size();
	}
	if (texelsOut)
	{
		memcpy(texelsOut, t.m_texels.Data(), t.m_texels.Size());
	}
	if (mipOffsets)
	{
		memcpy(mipOffsets, t.m_mipOffsets, sizeof(uint32_t) * sr::Config::c_maxTexDimLog2);
	}
	if (numMips) *numMips = t.m_numMips;
	if (wLog2) *wLog2 = t.m_widthLog2;
	if (hLog2) *hLog2 = t.m_heightLog2;
	return SRB_OK;
}

SRB_API int srref_begin_frame(void* h)
{
	((RefCtx*)h)->ctx->BeginFrame();
	return SRB_OK;
}

SRB_API int srref_clear(void* h, uint32_t color, int clearColour, int clearDepth)
{
	RefCtx* c = (RefCtx*)h;
	c->ctx->ClearFrameBuffer(*c->fb, color, clearColour != 0, clearDepth != 0);
	return SRB_OK;
}

SRB_API int srref_draw_indexed(void* h, const srb_draw_desc* d)
{
	RefCtx* c = (RefCtx*)h;
	sr::DrawCall call;
	FillDraw(c, *d, call);
	if (!call.m_pixelShader)
	{
		return SRB_ERR_UNKNOWN_SHADER;
	}
	c->ctx->DrawIndexed(call);
	return SRB_OK;
}

SRB_API int srref_end_frame(void* h)
{
	((RefCtx*)h)->ctx->EndFrame();
	return SRB_OK;
}

/* Timed loop, all in native code: per frame BeginFrame -> ClearFrameBuffer -> DrawIndexed x n -> EndFrame
 * (the region BASELINE.md §3.3 defines).  `mvps` holds frames*n_draws matrices (or NULL to use the descs').
 * Writes per-frame milliseconds to ms_out[frames]. */
SRB_API int srref_render_frames(void* h, const srb_draw_desc* draws, uint32_t n_draws, const float* mvps,
                                uint32_t frames, uint32_t clear_color, double* ms_out)
{
	RefCtx* c = (RefCtx*)h;
	std::vector<sr::DrawCall> calls(n_draws);
	for (uint32_t i = 0; i < n_draws; ++i)
	{
		FillDraw(c, draws[i], calls[i]);
		if (!calls[i].m_pixelShader)
		{
			return SRB_ERR_UNKNOWN_SHADER;
		}
	}
	for (uint32_t f = 0; f < frames; ++f)
	{
		auto t0 = std::chrono::steady_clock::now();
		c->ctx->BeginFrame();
		c->ctx->ClearFrameBuffer(*c->fb, clear_color);
		for (uint32_t i = 0; i < n_draws; ++i)
		{
			if (mvps)
			{
				memcpy(calls[i].m_mvp.Data(), mvps + (size_t(f) * n_draws + i) * 16, sizeof(float) * 16);
			}
			c->ctx->DrawIndexed(calls[i]);
		}
		c->ctx->EndFrame();
		auto t1 = std::chrono::steady_clock::now();
		if (ms_out)
		{
			ms_out[f] = std::chrono::duration<double, std::milli>(t1 - t0).count();
		}
	}
	return SRB_OK;
}

SRB_API int srref_read_tiles(void* h, void* colourTiles, void* depthTiles, uint64_t depthStride)
{
	RefCtx* c = (RefCtx*)h;
	sr::FrameBufferPlane* p = c->fb->WritePlane();
	uint32_t const n = p->m_tilesX * p->m_tilesY;
	if (colourTiles)
	{
		memcpy(colourTiles, p->m_colourTiles, size_t(n) * sizeof(sr::ColourTile));
	}
	if (depthTiles)
	{
		for (uint32_t i = 0; i < n; ++i)
		{
			memcpy((uint8_t*)depthTiles + i * depthStride, p->m_depthTiles[i].m_depth, sizeof(p->m_depthTiles[i].m_depth));
		}
	}
	return SRB_OK;
}

// RenderContext::Blit (Renderer.cpp:350-372), waited for so the pixels are valid on return.
SRB_API int srref_blit_linear(void* h, uint8_t* linearPixels)
{
	RefCtx* c = (RefCtx*)h;
	uint32_t const idx = c->fb->m_writePlane;
	c->ctx->Blit(*c->fb, linearPixels, nullptr, nullptr);
	c->ctx->m_taskSystem.WaitForCounter(&c->fb->m_jobs[idx].m_counter);
	c->fb->SwapPlanes(); // keep the same write plane for the next frame of this harness
	return SRB_OK;
}

SRB_API int srref_framebuffer_info(void* h, uint32_t* w, uint32_t* ht, uint32_t* tx, uint32_t* ty)
{
	RefCtx* c = (RefCtx*)h;
	sr::FrameBufferPlane* p = c->fb->WritePlane();
	if (w) *w = p->m_width;
	if (ht) *ht = p->m_height;
	if (tx) *tx = p->m_tilesX;
	if (ty) *ty = p->m_tilesY;
	return SRB_OK;
}

/* ---- parity dumps; valid after srref_end_frame and before the next srref_begin_frame ------------------------ */

SRB_API int srref_dump_tile_counts(void* h, uint32_t* counts, uint32_t numTiles)
{
	RefCtx* c = (RefCtx*)h;
	sr::BinContext& b = c->ctx->m_binner;
	if (numTiles != b.m_numBinsX * b.m_numBinsY)
	{
		return SRB_ERR_INVALID;
	}
	for (uint32_t t = 0; t < numTiles; ++t)
	{
		uint32_t n = 0;
		for (uint32_t th = 0; th < b.m_numThreads; ++th)
		{
			sr::ThreadBin& bin = b.LookupThreadBin(th, t % b.m_numBinsX, t / b.m_numBinsX);
			for (uint32_t i = 0; i < bin.m_numChunks; ++i)
			{
				n += bin.m_binChunks[i]->m_numTris;
			}
		}
		counts[t] = n;
	}
	return SRB_OK;
}

SRB_API int srref_dump_tile_tris(void* h, uint32_t tileIdx, srb_tile_tri* out, uint32_t cap, uint32_t* n)
{
	RefCtx* c = (RefCtx*)h;
	std::vector<sr::BinChunk*> chunks;
	SortedChunks(c, tileIdx, chunks);
	uint32_t k = 0;
	for (sr::BinChunk* ch : chunks)
	{
		for (uint32_t t = 0; t < ch->m_numTris; ++t)
		{
			if (k < cap)
			{
				CopyTri(*ch, t, out[k]);
			}
			++k;
		}
	}
	*n = k;
	return k <= cap ? SRB_OK : SRB_ERR_OVERFLOW;
}

/* Pre-depth coverage of every list entry: each triangle is rasterised alone, by the reference's own
 * RasterizeTrisInBin_OutputFragments, against a depth tile cleared to 0.0f. */
SRB_API int srref_dump_tile_coverage(void* h, uint32_t tileIdx, uint64_t* masks, uint32_t capEntries, uint32_t* n)
{
	RefCtx* c = (RefCtx*)h;
	std::vector<sr::BinChunk*> chunks;
	SortedChunks(c, tileIdx, chunks);
	sr::DepthTile* depth = (sr::DepthTile*)aligned_alloc(64, (sizeof(sr::DepthTile) + 63) & ~size_t(63));
	sr::BinChunk* one = (sr::BinChunk*)aligned_alloc(64, (sizeof(sr::BinChunk) + 63) & ~size_t(63));
	sr::DrawCall dummy;
	uint32_t k = 0;
	for (sr::BinChunk* ch : chunks)
	{
		for (uint32_t t = 0; t < ch->m_numTris; ++t, ++k)
		{
			if (k >= capEntries)
			{
				continue;
			}
			one->m_edgeEq[0] = ch->m_edgeEq[t];
			one->m_zOverW[0] = ch->m_zOverW[t];
			one->m_recipW[0] = ch->m_recipW[t];
			one->m_numTris = 1;
			one->m_attribsPerTri = ch->m_attribsPerTri;
			one->m_drawCallIdx = ch->m_drawCallIdx;
			for (uint32_t i = 0; i < 64 * 64; ++i)
			{
				depth->m_depth[i] = sr::Config::c_depthMax;
			}
			c->scratch.Reset();
			sr::FragmentBuffer fb;
			fb.m_fragments = (sr::FragmentBuffer::Frag*)c->scratch.Align(KT_ALIGNOF(sr::FragmentBuffer::Frag));
			fb.m_allocator = &c->scratch;
			sr::RasterizeTrisInBin_OutputFragments(dummy, depth, *one, 0, fb);
			uint64_t* m = masks + size_t(k) * 64;
			memset(m, 0, 64 * sizeof(uint64_t));
			for (uint32_t f = 0; f < fb.m_numFragments; ++f)
			{
				uint32_t const x = fb.m_fragments[f].x, y = fb.m_fragments[f].y;
				m[(y >> 3) * 8 + (x >> 3)] |= 1ull << ((y & 7) * 8 + (x & 7));
			}
		}
	}
	c->scratch.Reset();
	free(depth);
	free(one);
	*n = k;
	return k <= capEntries ? SRB_OK : SRB_ERR_OVERFLOW;
}

/* The ordered fragment stream of one tile (what the reference shades, Rasterizer.cpp:558-575), starting from a
 * depth tile cleared to 0.0f: frag = entryIndex << 12 | y << 6 | x. */
SRB_API int srref_dump_tile_fragments(void* h, uint32_t tileIdx, uint32_t* frags, uint64_t cap, uint64_t* n,
                                      float* depthOut)
{
	RefCtx* c = (RefCtx*)h;
	std::vector<sr::BinChunk*> chunks;
	SortedChunks(c, tileIdx, chunks);
	sr::DepthTile* depth = (sr::DepthTile*)aligned_alloc(64, (sizeof(sr::DepthTile) + 63) & ~size_t(63));
	for (uint32_t i = 0; i < 64 * 64; ++i)
	{
		depth->m_depth[i] = sr::Config::c_depthMax;
	}
	sr::DrawCall dummy;
	uint64_t k = 0;
	uint32_t entryBase = 0;
	for (sr::BinChunk* ch : chunks)
	{
		c->scratch.Reset();
		sr::FragmentBuffer fb;
		fb.m_fragments = (sr::FragmentBuffer::Frag*)c->scratch.Align(KT_ALIGNOF(sr::FragmentBuffer::Frag));
		fb.m_allocator = &c->scratch;
		sr::RasterizeTrisInBin_OutputFragments(dummy, depth, *ch, 0, fb);
		for (uint32_t f = 0; f < fb.m_numFragments; ++f, ++k)
		{
			if (k < cap)
			{
				frags[k] = ((entryBase + fb.m_fragments[f].triIdx) << 12) | (uint32_t(fb.m_fragments[f].y) << 6) |
				           fb.m_fragments[f].x;
			}
		}
		entryBase += ch->m_numTris;
	}
	if (depthOut)
	{
		memcpy(depthOut, depth->m_depth, sizeof(float) * 64 * 64);
	}
	c->scratch.Reset();
	free(depth);
	*n = k;
	return k <= cap ? SRB_OK : SRB_ERR_OVERFLOW;
}

/* The host CPU's RCPPS, the instruction behind _mm256_rcp_ps at Rasterizer.cpp:375-376. */
SRB_API void srref_rcp(const float* in, float* out, uint64_t n)
{
	for (uint64_t i = 0; i < n; ++i)
	{
		out[i] = _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(in[i])));
	}
}

/* Tex::SampleWrap (Texture.cpp:381-452) + RGBA32SoA_To_RGBA8AoS on arbitrary inputs, 8 lanes per call, for sampler
 * unit tests. n must be a multiple of 8. */
SRB_API int srref_sample(void* h, uint64_t tex, const float* u, const float* v, const float* dudx, const float* dudy,
                         const float* dvdx, const float* dvdy, uint32_t* rgba, uint64_t n)
{
	RefCtx* c = (RefCtx*)h;
	if (!tex || tex > c->textures.size() || (n & 7))
	{
		return SRB_ERR_INVALID;
	}
	sr::Tex::TextureData const& t = c->textures[tex - 1]->tex;
	for (uint64_t i = 0; i < n; i += 8)
	{
		__m256 r, g, b, a;
		sr::Tex::SampleWrap(t, _mm256_loadu_ps(u + i), _mm256_loadu_ps(v + i), _mm256_loadu_ps(dudx + i),
		                    _mm256_loadu_ps(dudy + i), _mm256_loadu_ps(dvdx + i), _mm256_loadu_ps(dvdy + i), r, g, b, a,
		                    0xFF);
		KT_ALIGNAS(32) uint32_t px[8];
		sr::simdutil::RGBA32SoA_To_RGBA8AoS(r, g, b, a, px);
		memcpy(rgba + i, px, sizeof(px));
	}
	return SRB_OK;
}

} // extern "C"
