// softrast_b200/Renderer.h — source-compatible C++ shim of the reference's renderer API over the C ABI.
//
// Mirrors, name for name, what scene code uses from the reference:
//   SoftRast/Renderer.h:21-177  ColourTile, DepthTile, FrameBufferPlane, FrameBuffer, PixelShaderFn, GenericDrawBuffer,
//                               DrawCall (+ Set* chain), RenderContext {BeginFrame, ClearFrameBuffer, DrawIndexed,
//                               EndFrame, Blit, Shutdown}
//   SoftRast/Texture.h:21-41    Tex::TextureData {CreateFromRGBA8, Clear, m_texels, m_mipOffsets, ...}
//   Viewer/Shaders.h:71-130     shader::UnlitDiffuseShader / VisualizeNormalsShader / VisualizeUVsShader
// so that code written like Viewer/Scene.cpp:32-65 compiles unchanged and runs on the GPU.  Everything forwards to
// include/softrast_b200.h; nothing here computes pixels.  A pixel shader is selected by the IDENTITY of the function
// pointer stored in the DrawCall (the reference calls it; we look it up); an unknown pointer is a hard error — there
// is no CPU fallback.  Header-only; link with -lsoftrast_b200.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <unordered_map>
#include <vector>

#include "../softrast_b200.h"

namespace sr
{

namespace Config
{
constexpr uint32_t c_binHeightLog2 = SRB_BIN_LOG2;
constexpr uint32_t c_binWidthLog2 = SRB_BIN_LOG2;
constexpr uint32_t c_binHeight = 1u << c_binHeightLog2;
constexpr uint32_t c_binWidth = 1u << c_binWidthLog2;
constexpr int32_t c_subPixelBits = SRB_SUBPIXEL_BITS;
constexpr uint32_t c_maxVaryings = SRB_MAX_VARYINGS;
constexpr uint32_t c_maxTexDimLog2 = SRB_MAX_TEX_DIM_LOG2;
constexpr float c_depthMin = 1.0f; // reverse-Z (SR_USE_REVERSE_Z, Config.h:4,32-38)
constexpr float c_depthMax = 0.0f;
} // namespace Config

inline void SrbCheck(int rc, srb_context* ctx, const char* what)
{
	if (rc != SRB_OK)
	{
		fprintf(stderr, "softrast_b200: %s failed (%d): %s\n", what, rc, ctx ? srb_last_error(ctx) : "");
		abort(); // the reference API returns void everywhere; errors cannot be ignored silently
	}
}

// ---- framebuffer (Renderer.h:21-108) ----------------------------------------------------------------------------
struct ColourTile
{
	static uint32_t const c_bytesPerPixel = 4;
	alignas(32) uint8_t m_colour[Config::c_binHeight * Config::c_binWidth * c_bytesPerPixel];
};

struct DepthTile
{
	alignas(32) float m_depth[Config::c_binWidth * Config::c_binWidth];
	float m_hiZmin;
	float m_hiZmax;
};
static_assert(sizeof(ColourTile) == SRB_COLOUR_TILE_BYTES, "ColourTile layout");
static_assert(sizeof(DepthTile) == SRB_DEPTH_TILE_BYTES, "DepthTile layout");

struct FrameBufferPlane
{
	ColourTile* m_colourTiles = nullptr; // host mirror, valid after RenderContext::EndFrame (if read-back is on)
	DepthTile* m_depthTiles = nullptr;
	uint32_t m_height = 0, m_width = 0, m_tilesY = 0, m_tilesX = 0;
};

struct FrameBuffer
{
	FrameBuffer(FrameBuffer const&) = delete;
	FrameBuffer& operator=(FrameBuffer const&) = delete;
	FrameBuffer(uint32_t _width, uint32_t _height, bool = true, bool = true)
	{
		for (FrameBufferPlane& p : m_bufferedPlanes)
		{
			p.m_width = _width;
			p.m_height = _height;
			p.m_tilesX = (_width + Config::c_binWidth - 1) / Config::c_binWidth;
			p.m_tilesY = (_height + Config::c_binHeight - 1) / Config::c_binHeight;
			size_t const n = size_t(p.m_tilesX) * p.m_tilesY;
			p.m_colourTiles = (ColourTile*)aligned_alloc(64, n * sizeof(ColourTile));
			p.m_depthTiles = (DepthTile*)aligned_alloc(64, ((n * sizeof(DepthTile) + 63) / 64) * 64);
		}
	}
	~FrameBuffer()
	{
		for (FrameBufferPlane& p : m_bufferedPlanes)
		{
			free(p.m_colourTiles);
			free(p.m_depthTiles);
		}
	}
	FrameBufferPlane* WritePlane() { return &m_bufferedPlanes[m_writePlane]; }
	FrameBufferPlane const* ReadPlane() const { return &m_bufferedPlanes[m_writePlane ^ 1]; }
	void SwapPlanes() { m_writePlane ^= 1; }

	FrameBufferPlane m_bufferedPlanes[2];
	uint32_t m_writePlane = 0;
	// device side (created lazily by the RenderContext that first uses this framebuffer)
	srb_context* m_ctx = nullptr;
	srb_handle m_handle = 0;
};

struct Interpolants; // host-side SoA streams do not exist on the GPU path
using PixelShaderFn = void(void const* _uniforms, Interpolants const& _interpolants, uint32_t o_texels[8], uint32_t _execMask);

// ---- textures (Texture.h:21-41) -----------------------------------------------------------------------------------
namespace Tex
{
struct TextureData
{
	TextureData() = default;
	TextureData(TextureData const&) = delete;
	TextureData& operator=(TextureData const&) = delete;
	TextureData(TextureData&&) = default;
	TextureData& operator=(TextureData&&) = default;
	// Texture.cpp:119-199; mips by 2x2 box filter (the reference uses stb_image_resize's default filter)
	// Texture.cpp:122-199: the same tiled / Morton layout and, with _calcMips, the same mip texels as the reference
	// stores (every level filtered from the original image like stbir_resize_uint8 does, SRB_MIPS_STB).
	void CreateFromRGBA8(uint8_t const* _texels, uint32_t _width, uint32_t _height, bool _calcMips = false)
	{
		uint64_t bytes = 0;
		int const mips = _calcMips ? SRB_MIPS_STB : SRB_MIPS_NONE;
		SrbCheck(srb_texture_build_rgba8(nullptr, _width, _height, mips, nullptr, &bytes, m_mipOffsets, &m_numMips),
		         nullptr, "srb_texture_build_rgba8");
		m_texels.resize(bytes);
		SrbCheck(srb_texture_build_rgba8(_texels, _width, _height, mips, m_texels.data(), &bytes, m_mipOffsets, &m_numMips),
		         nullptr, "srb_texture_build_rgba8");
		m_widthLog2 = m_heightLog2 = 0;
		while ((1u << m_widthLog2) < _width) ++m_widthLog2;
		while ((1u << m_heightLog2) < _height) ++m_heightLog2;
		m_bytesPerPixel = 4;
		m_handle = 0;
	}
	// Texture.cpp:103-117: stbi_load(_file, ..., 4) + CreateFromRGBA8(..., true).  PNG and TGA are decoded by the library
	// (same bytes as stb_image); a file that cannot be loaded leaves the texture empty, as in the reference.
	void CreateFromFile(char const* _file)
	{
		Clear();
		uint8_t* rgba = nullptr;
		uint32_t w = 0, h = 0;
		if (srb_image_load_rgba8(_file, &rgba, &w, &h) != SRB_OK)
		{
			fprintf(stderr, "Failed to load texture: %s (%s)\n", _file, srb_model_last_error());
			return;
		}
		CreateFromRGBA8(rgba, w, h, true);
		srb_image_free(rgba);
	}
	void Clear()
	{
		m_texels.clear();
		m_handle = 0;
	}
	std::vector<uint8_t> m_texels;
	uint32_t m_mipOffsets[Config::c_maxTexDimLog2] = {};
	uint32_t m_widthLog2 = 0, m_heightLog2 = 0, m_numMips = 0, m_bytesPerPixel = 0;
	mutable srb_handle m_handle = 0; // device copy, made on first draw
};
} // namespace Tex

// ---- pixel shaders (Viewer/Shaders.h:71-130): distinct symbols whose ADDRESSES select the device shaders -------------
namespace shader
{
inline void HostCallIsAnError(char const* name)
{
	fprintf(stderr, "softrast_b200: %s was called on the host; pixel shaders run on the GPU (no CPU fallback)\n", name);
	abort();
}
inline void UnlitDiffuseShader(void const*, Interpolants const&, uint32_t*, uint32_t) { HostCallIsAnError("UnlitDiffuseShader"); }
inline void VisualizeNormalsShader(void const*, Interpolants const&, uint32_t*, uint32_t) { HostCallIsAnError("VisualizeNormalsShader"); }
inline void VisualizeUVsShader(void const*, Interpolants const&, uint32_t*, uint32_t) { HostCallIsAnError("VisualizeUVsShader"); }
// Viewer/SponzaScene.cpp:13-104 (file-static there; a scene that keeps its own copy registers that pointer instead:
// ctx.RegisterPixelShader(SponzaShader, SRB_SHADER_SPONZA)).  Uniforms = Tex::TextureData*, like UnlitDiffuse; the
// frame constants of SponzaScene.cpp:11 go through RenderContext::SetSponzaConstants.
inline void SponzaShader(void const*, Interpolants const&, uint32_t*, uint32_t) { HostCallIsAnError("SponzaShader"); }
} // namespace shader

// ---- draw call (Renderer.h:112-150) ---------------------------------------------------------------------------------
struct GenericDrawBuffer
{
	void const* m_ptr = nullptr;
	uint32_t m_num = 0;
	uint32_t m_stride = 0;
};

struct Mat4Storage
{
	float m[16]; // column-major like kt::Mat4 (m_cols[c][r] = m[4*c + r]); identity by default
};

struct DrawCall
{
	static const uint32_t UV_OFFSET_INVALID = 0xFFFFFFFF;
	DrawCall()
	{
		memset(&m_mvp, 0, sizeof(m_mvp));
		m_mvp.m[0] = m_mvp.m[5] = m_mvp.m[10] = m_mvp.m[15] = 1.0f;
	}
	DrawCall& SetPixelShader(PixelShaderFn* _fn, void const* _uniforms)
	{
		m_pixelShader = _fn;
		m_pixelUniforms = _uniforms;
		return *this;
	}
	DrawCall& SetIndexBuffer(void const* _buffer, uint32_t const _stride, uint32_t const _num)
	{
		m_indexBuffer = {_buffer, _num, _stride};
		return *this;
	}
	DrawCall& SetPositionBuffer(void const* _buffer, uint32_t const _stride, uint32_t const _num)
	{
		m_positionBuffer = {_buffer, _num, _stride};
		return *this;
	}
	DrawCall& SetAttributeBuffer(void const* _buffer, uint32_t const _stride, uint32_t const _num, uint32_t const _uvOffset = 0)
	{
		m_uvOffset = _uvOffset;
		m_attributeBuffer = {_buffer, _num, _stride};
		return *this;
	}
	DrawCall& SetFrameBuffer(FrameBuffer* _buffer)
	{
		m_frameBufferOwner = _buffer;
		m_frameBuffer = _buffer->WritePlane();
		return *this;
	}
	// accepts kt::Mat4 (16 contiguous column-major floats) or any type of that layout
	template <typename Mat4T>
	DrawCall& SetMVP(Mat4T const& _mvp)
	{
		static_assert(sizeof(Mat4T) == sizeof(float) * 16, "SetMVP expects 16 column-major floats (kt::Mat4)");
		memcpy(m_mvp.m, &_mvp, sizeof(m_mvp.m));
		return *this;
	}

	PixelShaderFn* m_pixelShader = nullptr;
	void const* m_pixelUniforms = nullptr;
	GenericDrawBuffer m_indexBuffer, m_positionBuffer, m_attributeBuffer;
	uint32_t m_uvOffset = 0;
	FrameBufferPlane const* m_frameBuffer = nullptr;
	FrameBuffer* m_frameBufferOwner = nullptr;
	Mat4Storage m_mvp;
	uint32_t m_drawCallIdx = 0;
};

// ---- render context (Renderer.h:153-177) -----------------------------------------------------------------------------
class RenderContext
{
public:
	// Default = the reference's contract: DrawCall buffers are read from the application's arrays EVERY frame (they are
	// re-uploaded at their first use in each frame: SRB_FLAG_UPLOAD_ALWAYS), so editing vertices in place, or freeing an
	// array and allocating another at the same address, behaves as it does with the CPU renderer.  Pass
	// SRB_FLAG_NONE to opt in to CACHED device mirrors (found by pointer, uploaded once): much faster for static scenes,
	// but then every change to a bound array must be announced with Invalidate().
	explicit RenderContext(int _device = 0, uint32_t _flags = SRB_FLAG_UPLOAD_ALWAYS)
	{
		int const rc = srb_create(_device, _flags, &m_ctx);
		SrbCheck(rc, m_ctx, "srb_create");
		RegisterPixelShader(shader::UnlitDiffuseShader, SRB_SHADER_UNLIT_DIFFUSE);
		RegisterPixelShader(shader::VisualizeNormalsShader, SRB_SHADER_VISUALIZE_NORMALS);
		RegisterPixelShader(shader::VisualizeUVsShader, SRB_SHADER_VISUALIZE_UVS);
		RegisterPixelShader(shader::SponzaShader, SRB_SHADER_SPONZA);
	}
	// A context for another frame in flight of the same scene: shares _parent's textures and buffers (srb_create_shared).
	// Both contexts must cache their host buffers (SRB_FLAG_NONE): upload-always contexts cannot share device mirrors.
	RenderContext(RenderContext& _parent, uint32_t _flags)
	{
		int const rc = srb_create_shared(_parent.m_ctx, _flags, &m_ctx);
		SrbCheck(rc, m_ctx, "srb_create_shared");
		m_shaders = _parent.m_shaders;
	}
	~RenderContext() { Shutdown(); }
	RenderContext(RenderContext const&) = delete;
	RenderContext& operator=(RenderContext const&) = delete;

	void Shutdown()
	{
		if (m_ctx)
		{
			srb_destroy(m_ctx);
			m_ctx = nullptr;
		}
	}

	// The function-pointer -> device-shader registry.
	void RegisterPixelShader(PixelShaderFn* _fn, uint32_t _deviceShader) { m_shaders[(void const*)_fn] = _deviceShader; }

	// Replaces writing the file-static g_constants of Viewer/SponzaScene.cpp:11 (done every frame by SponzaScene::Update,
	// :168-187): applies to the frames submitted after the call.
	void SetSponzaConstants(srb_sponza_constants const& _constants)
	{
		SrbCheck(srb_set_sponza_constants(m_ctx, &_constants), m_ctx, "srb_set_sponza_constants");
	}

	// EndFrame copies the finished tiles back into FrameBuffer::WritePlane()'s host arrays like the reference leaves
	// them (default).  Turn it off when only Blit() consumes the frame.
	void SetReadbackOnEndFrame(bool _on) { m_readback = _on; }

	void BeginFrame()
	{
		SrbCheck(srb_begin_frame(m_ctx), m_ctx, "srb_begin_frame");
		m_frameFb = nullptr;
		m_numDraws = 0;
	}

	void ClearFrameBuffer(FrameBuffer& _buffer, uint32_t _color = 0x00000000, bool _clearColour = true, bool _clearDepth = true)
	{
		SrbCheck(srb_clear(m_ctx, Handle(_buffer), _color, _clearColour, _clearDepth), m_ctx, "srb_clear");
		m_frameFb = &_buffer;
	}

	void DrawIndexed(DrawCall const& _call)
	{
		auto it = m_shaders.find((void const*)_call.m_pixelShader);
		if (it == m_shaders.end())
		{
			SrbCheck(SRB_ERR_UNKNOWN_SHADER, nullptr, "DrawIndexed: pixel shader pointer is not registered");
		}
		if (!_call.m_frameBufferOwner)
		{
			SrbCheck(SRB_ERR_INVALID, nullptr, "DrawIndexed: no framebuffer bound (SetFrameBuffer)");
		}
		srb_draw_desc d;
		memset(&d, 0, sizeof(d));
		d.shader = it->second;
		d.uv_offset = _call.m_uvOffset;
		d.framebuffer = Handle(*_call.m_frameBufferOwner);
		if ((d.shader == SRB_SHADER_UNLIT_DIFFUSE || d.shader == SRB_SHADER_SPONZA) && _call.m_pixelUniforms)
		{
			Tex::TextureData const* tex = (Tex::TextureData const*)_call.m_pixelUniforms;
			if (!tex->m_texels.empty())
			{
				if (!tex->m_handle)
				{
					SrbCheck(srb_texture_create(m_ctx, tex->m_texels.data(), tex->m_texels.size(), tex->m_mipOffsets, tex->m_numMips,
					                            tex->m_widthLog2, tex->m_heightLog2, &tex->m_handle),
					         m_ctx, "srb_texture_create");
				}
				d.texture = tex->m_handle;
			}
		}
		d.indices = {0, 0, _call.m_indexBuffer.m_ptr, _call.m_indexBuffer.m_stride, _call.m_indexBuffer.m_num};
		d.positions = {0, 0, _call.m_positionBuffer.m_ptr, _call.m_positionBuffer.m_stride, _call.m_positionBuffer.m_num};
		d.attributes = {0, 0, _call.m_attributeBuffer.m_ptr, _call.m_attributeBuffer.m_stride, _call.m_attributeBuffer.m_num};
		memcpy(d.mvp, _call.m_mvp.m, sizeof(d.mvp));
		SrbCheck(srb_draw_indexed(m_ctx, &d), m_ctx, "srb_draw_indexed");
		m_frameFb = _call.m_frameBufferOwner;
		++m_numDraws;
	}

	void EndFrame()
	{
		SrbCheck(srb_end_frame(m_ctx), m_ctx, "srb_end_frame");
		if (m_readback && m_frameFb)
		{
			FrameBufferPlane* p = m_frameFb->WritePlane();
			SrbCheck(srb_read_tiles(m_ctx, m_frameFb->m_handle, p->m_colourTiles, p->m_depthTiles, sizeof(DepthTile)), m_ctx,
			         "srb_read_tiles");
		}
	}

	void Blit(FrameBuffer& _fb, uint8_t* _linearPixels, void (*_onFinishBlit)(void*) = nullptr, void* _onFinishUser = nullptr)
	{
		SrbCheck(srb_blit_linear(m_ctx, Handle(_fb), _linearPixels, _onFinishBlit, _onFinishUser), m_ctx, "srb_blit_linear");
		_fb.SwapPlanes();
	}

	void Invalidate(void const* _hostBuffer) { SrbCheck(srb_invalidate_host(m_ctx, _hostBuffer), m_ctx, "srb_invalidate_host"); }
	srb_context* Native() { return m_ctx; }

private:
	srb_handle Handle(FrameBuffer& _fb)
	{
		if (!_fb.m_handle || _fb.m_ctx != m_ctx)
		{
			FrameBufferPlane* p = _fb.WritePlane();
			SrbCheck(srb_framebuffer_create(m_ctx, p->m_width, p->m_height, &_fb.m_handle), m_ctx, "srb_framebuffer_create");
			_fb.m_ctx = m_ctx;
		}
		return _fb.m_handle;
	}

	srb_context* m_ctx = nullptr;
	std::unordered_map<void const*, uint32_t> m_shaders;
	FrameBuffer* m_frameFb = nullptr;
	uint32_t m_numDraws = 0;
	bool m_readback = true;
};

} // namespace sr
