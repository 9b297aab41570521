/*
 * softrast_b200.h — C ABI of the B200-native sort-middle frame pipeline.
 *
 * This is the drop-in boundary for the hot path of karltechno/SoftRast:
 * sr::RenderContext::EndFrame() and everything it drives
 *   (reference SoftRast/Renderer.cpp:209-317 -> Binning.cpp:464 -> Rasterizer.cpp:525 ->
 *    Viewer/Shaders.h:71 -> Texture.cpp:381).
 * The reference has no FFI of its own (it is a statically linked C++ class API, SoftRast/Renderer.h:119-177);
 * the entry points below are what a binding of that API needs, one per reference call.  The C++ shim
 * `include/softrast_b200/Renderer.h` keeps the reference's class/method names and forwards here.
 *
 * Plain C: opaque context pointer, integer handles, raw pointers + sizes.  Every call returns an int status
 * (0 = SRB_OK); srb_last_error() gives the text.  There is NO CPU fallback: if no CUDA device is usable
 * srb_create() fails.
 */
#ifndef SOFTRAST_B200_H
#define SOFTRAST_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SRB_API __attribute__((visibility("default")))
#else
#define SRB_API
#endif

typedef struct srb_context srb_context;
typedef uint64_t srb_handle; /* 0 is never a valid handle */

enum srb_status
{
	SRB_OK = 0,
	SRB_ERR_CUDA = 1,            /* a CUDA runtime call failed (text in srb_last_error) */
	SRB_ERR_INVALID = 2,         /* bad argument / bad handle / wrong call order */
	SRB_ERR_OVERFLOW = 3,        /* a device-side capacity was exceeded; the frame is incomplete */
	SRB_ERR_UNKNOWN_SHADER = 4,  /* pixel shader not in the registry (reference would call a host fn ptr) */
	SRB_ERR_NO_DEVICE = 5
};

/* Device-side pixel shaders = the reference's Viewer/Shaders.h functions (PixelShaderFn, Renderer.h:110). */
enum srb_shader
{
	SRB_SHADER_UNLIT_DIFFUSE = 0,     /* Viewer/Shaders.h:71-104  */
	SRB_SHADER_VISUALIZE_NORMALS = 1, /* Viewer/Shaders.h:106-121 */
	SRB_SHADER_VISUALIZE_UVS = 2,     /* Viewer/Shaders.h:123-130 */
	SRB_SHADER_SPONZA = 3,            /* Viewer/SponzaScene.cpp:13-104: sun + 16 point lights + ambient, times the texture */
	SRB_SHADER_COUNT = 4
};

/* Frame constants of SRB_SHADER_SPONZA = SponzaScene::Constants (Viewer/SponzaScene.h:17-47, filled at
 * SponzaScene.cpp:121-187).  sun_dir holds the three broadcast values of m_sunDir[0..2] AS THE SHADER READS THEM (the
 * reference broadcasts sunDir.x into all three, SponzaScene.cpp:135-137); falloff is carried but unused by the shader. */
#define SRB_SPONZA_POINT_LIGHTS 16
typedef struct srb_sponza_light
{
	float pos[3];
	float colour[3];
	float intensity;
	float falloff;
} srb_sponza_light;
typedef struct srb_sponza_constants
{
	float sun_dir[3];
	float ambient[3];
	float pad[2];
	srb_sponza_light lights[SRB_SPONZA_POINT_LIGHTS];
} srb_sponza_constants;

/* Compile-time constants mirrored from SoftRast/Config.h:18-27 (screen size is runtime here). */
#define SRB_BIN_LOG2 6
#define SRB_BIN_DIM 64
#define SRB_SUBPIXEL_BITS 8
#define SRB_MAX_VARYINGS 8
#define SRB_MAX_TEX_DIM_LOG2 14
#define SRB_COLOUR_TILE_BYTES 16384u /* sizeof(sr::ColourTile), Renderer.h:21-28 */
#define SRB_DEPTH_TILE_BYTES 16416u  /* sizeof(sr::DepthTile) incl. hiZ floats + padding, Renderer.h:30-37 */

/* srb_create flags */
#define SRB_FLAG_NONE 0u
#define SRB_FLAG_UPLOAD_ALWAYS 1u /* never cache host-pointer buffers across frames: every array is re-uploaded at its first use in each frame */
/* Set up the attribute planes of ALL varyings of every triangle, as the reference's binner does (Binning.cpp:340-350),
 * so that srb_dump_tile_tris can report them.  Without it only the planes the bound pixel shader reads are set up
 * (UnlitDiffuse: 6, 7 and uv_offset, uv_offset+1; VisualizeNormals: 3..5; VisualizeUVs: 6, 7): same pixels, less work. */
#define SRB_FLAG_FULL_RECORDS 2u

/* One buffer binding = sr::GenericDrawBuffer (Renderer.h:112-117).  Either `buffer` is a resident
 * srb_buffer_create() handle (then `host` is ignored and `offset` is a byte offset into it), or `buffer` is 0 and
 * `host` points to host memory that the library mirrors on the device (cached by pointer+size; with
 * SRB_FLAG_UPLOAD_ALWAYS re-uploaded once per frame; call srb_invalidate_host() after changing the bytes). */
typedef struct srb_buffer_ref
{
	srb_handle buffer;
	uint64_t offset;
	const void* host;
	uint32_t stride; /* bytes */
	uint32_t num;    /* elements (indices / vertices) */
} srb_buffer_ref;

/* POD mirror of sr::DrawCall (Renderer.h:119-150). */
typedef struct srb_draw_desc
{
	uint32_t shader;            /* enum srb_shader  (replaces PixelShaderFn* m_pixelShader) */
	uint32_t uv_offset;         /* m_uvOffset, in floats into the attribute vertex */
	srb_handle texture;         /* m_pixelUniforms for UNLIT_DIFFUSE: Tex::TextureData; 0 = null texture */
	srb_handle framebuffer;     /* m_frameBuffer (the write plane) */
	srb_buffer_ref indices;     /* stride 1, 2 or 4 (Binning.cpp:167-205); num = index count */
	srb_buffer_ref positions;   /* reads 3 floats at offset 0 of each element (Binning.cpp:207-213) */
	srb_buffer_ref attributes;  /* stride = 4*numVaryings <= 32 bytes (Binning.cpp:215-221) */
	float mvp[16];              /* kt::Mat4, column-major: mvp[4*c + r] */
} srb_draw_desc;

typedef struct srb_counters
{
	uint64_t tris_in;        /* input triangles submitted this frame */
	uint64_t tris_setup;     /* triangles surviving clip/cull ("TrisBinned", Binning.cpp:313) */
	uint64_t tris_clipped;   /* input triangles that went through the clipper ("TrisClipped", :516) */
	uint64_t tile_refs;      /* (triangle, tile) pairs */
	uint64_t tiles_nonempty;
	uint64_t max_refs_in_tile;
	uint64_t pixels_covered; /* pixels whose depth was written this frame */
	uint64_t overflow;       /* non-zero if a capacity was exceeded */
} srb_counters;

/* Tile-relative triangle record exactly as the reference stores it per (tile, triangle) in a BinChunk
 * (Binning.h:18-55, filled at Binning.cpp:412-454).  Used only by the parity dumps. */
typedef struct srb_tile_tri
{
	int32_t c[3];
	int32_t dx[3];
	int32_t dy[3];
	uint8_t block_min_x, block_max_x, block_min_y, block_max_y;
	float recip_w[3];  /* c0, dx, dy */
	float z_over_w[3]; /* c0, dx, dy */
	float attr_dx[SRB_MAX_VARYINGS];
	float attr_dy[SRB_MAX_VARYINGS];
	float attr_c[SRB_MAX_VARYINGS];
	uint32_t attribs_per_tri;
	uint32_t draw_idx;
} srb_tile_tri;

/* ---- context ---------------------------------------------------------------------------------------------- */
/* replaces RenderContext::RenderContext (Renderer.cpp:138-150) */
SRB_API int srb_create(int device, uint32_t flags, srb_context** out);
/* A second context on the parent's device that SHARES the parent's textures, buffers and host-buffer mirrors (handles
 * are valid in every context of the family; they live until the last context is destroyed) and has its own
 * framebuffers, frame state and CUDA stream.  This is how several frames of one scene are kept in flight — one context
 * per frame in flight — with ONE copy of the scene in HBM.  Like everything else here: one submitting thread. */
SRB_API int srb_create_shared(srb_context* parent, uint32_t flags, srb_context** out);
/* replaces RenderContext::Shutdown / ~RenderContext (Renderer.cpp:152-159) */
SRB_API void srb_destroy(srb_context* ctx);
SRB_API const char* srb_last_error(srb_context* ctx);
SRB_API const char* srb_version(void);

/* The reference's mip selection uses the CPU's RCPPS approximation (Rasterizer.cpp:375-376).  It is a table on the
 * top `index_bits` mantissa bits; the device replays it.  srb_harvest_rcp_table() reads the table from THIS host's
 * CPU (table must hold 1<<index_bits words; returns the number of index bits needed, 0 on failure).
 * srb_create() harvests automatically; srb_set_rcp_table() overrides (e.g. to replay a golden fixture). */
SRB_API int srb_set_rcp_table(srb_context* ctx, const uint32_t* table, uint32_t index_bits);
SRB_API uint32_t srb_harvest_rcp_table(uint32_t* table, uint32_t max_index_bits);
/* Likewise for RSQRTPS (Viewer/SponzaScene.cpp:66): a table on (exponent parity, top `index_bits` mantissa bits), 2 <<
 * index_bits entries; see srb_host.cpp for the model.  Harvested and installed by srb_create(). */
SRB_API int srb_set_rsqrt_table(srb_context* ctx, const uint32_t* table, uint32_t index_bits);
SRB_API uint32_t srb_harvest_rsqrt_table(uint32_t* table, uint32_t max_index_bits);
/* replaces the file-static g_constants of Viewer/SponzaScene.cpp:11, which SponzaScene::Update rewrites every frame
 * (:168-187): the constants apply to the draws of the frames submitted after the call */
SRB_API int srb_set_sponza_constants(srb_context* ctx, const srb_sponza_constants* constants);

/* The frame constants of the viewer's default scene over time = SponzaScene::Init's light set-up and SponzaScene::Update's
 * animation (Viewer/SponzaScene.cpp:126-160 and :168-187; host code, like there): init seeds 16 point lights from
 * kt::XorShift32's default state, update(dt) moves them (quaternion rotation about per-light axes) and blends their
 * colours, then advances the phase — the same floats, operation for operation, as the reference computes; feed
 * `constants` to srb_set_sponza_constants before the frame's draws. */
typedef struct srb_sponza_scene
{
	float anim_phase; /* m_animPhase */
	struct
	{
		float base_pos[3], rot_offset[3], rot_axis[3], angle, colour_a[3], colour_b[3]; /* PointLightAnim, SponzaScene.h:26-37 */
	} anim[SRB_SPONZA_POINT_LIGHTS];
	srb_sponza_constants constants;
} srb_sponza_scene;
SRB_API void srb_sponza_scene_init(srb_sponza_scene* scene);
SRB_API void srb_sponza_scene_update(srb_sponza_scene* scene, float dt);

/* ---- resources -------------------------------------------------------------------------------------------- */
/* Tex::TextureData (Texture.h:21-41): the tiled/Morton texel blob + mip offsets are uploaded verbatim. */
SRB_API int srb_texture_create(srb_context* ctx, const uint8_t* texels, uint64_t bytes, const uint32_t* mip_offsets,
                               uint32_t num_mips, uint32_t width_log2, uint32_t height_log2, srb_handle* out);
SRB_API int srb_texture_destroy(srb_context* ctx, srb_handle tex);
/* Host-side builder for the reference's layout = TextureData::CreateFromRGBA8 (Texture.cpp:122-199: :73-101 tiling,
 * :159-175 mip placement).  calc_mips: SRB_MIPS_NONE (level 0 only), SRB_MIPS_BOX (a 2x2 box filter from the previous
 * level: cheap, NOT what the reference stores) or SRB_MIPS_STB (the reference's own mips, byte for byte: every level
 * filtered from the original image by a restatement of stb_image_resize's default down-sampling path, Texture.cpp:196).
 * Call with texels_out == NULL to get the required size in *bytes_out. */
#define SRB_MIPS_NONE 0
#define SRB_MIPS_BOX 1
#define SRB_MIPS_STB 2
SRB_API int srb_texture_build_rgba8(const uint8_t* rgba, uint32_t width, uint32_t height, int calc_mips,
                                    uint8_t* texels_out, uint64_t* bytes_out, uint32_t* mip_offsets_out,
                                    uint32_t* num_mips_out);

/* The same builder ON THE DEVICE: uploads the linear RGBA8 image, re-orders level 0 into the tiled /
 * Morton layout with a streaming kernel and filters every further level from the original image with stb_image_resize's
 * down-sampling arithmetic in its order of float additions (calc_mips: SRB_MIPS_NONE or SRB_MIPS_STB) — byte for byte
 * what srb_texture_build_rgba8(SRB_MIPS_STB) and the reference's CreateFromRGBA8 produce, in ~3 ms instead of
 * 0.25-0.5 s per 1024^2 texture on a host core.  `rgba` is borrowed until the call returns.  srb_texture_read copies a
 * texture's blob and description back (texels_out may be NULL to query the size). */
SRB_API int srb_texture_create_rgba8(srb_context* ctx, const uint8_t* rgba, uint32_t width, uint32_t height,
                                     int calc_mips, srb_handle* out);
SRB_API int srb_texture_read(srb_context* ctx, srb_handle tex, uint8_t* texels_out, uint64_t cap, uint64_t* bytes_out,
                             uint32_t* mip_offsets_out, uint32_t* num_mips_out, uint32_t* width_log2_out,
                             uint32_t* height_log2_out);

SRB_API int srb_buffer_create(srb_context* ctx, const void* host, uint64_t bytes, srb_handle* out);
SRB_API int srb_buffer_update(srb_context* ctx, srb_handle buf, uint64_t offset, const void* host, uint64_t bytes);
SRB_API int srb_buffer_destroy(srb_context* ctx, srb_handle buf);
SRB_API int srb_invalidate_host(srb_context* ctx, const void* host);

/* FrameBuffer / FrameBufferPlane::Init (Renderer.cpp:23-57): 64x64 tiles, row-major tiles, two planes. */
SRB_API int srb_framebuffer_create(srb_context* ctx, uint32_t width, uint32_t height, srb_handle* out);
SRB_API int srb_framebuffer_destroy(srb_context* ctx, srb_handle fb);

/* Screen-tile split of one frame across GPUs (BASELINE config 4): the root context exports its framebuffer, every
 * other process imports it (CUDA IPC: the colour tiles are mapped over NVLink/NVSwitch) and all contexts draw the SAME
 * frames, in the same order, after srb_set_tile_ownership(ctx, world, rank): each sets up, bins, rasterises and shades
 * only what touches tiles with tile % world == rank, and its shade kernel stores the finished COLOUR tiles straight into
 * the root's framebuffer — the composite is the store, there is no separate gather.  Depth stays on the GPU that owns
 * the tile (srb_read_tiles on the root returns depth for the root's own tiles only).
 * Completion is signalled on the device, there is no host barrier inside a frame: the last CTA of every rank's shade
 * kernel stamps an arrival flag in the root's memory, the root's shade kernel ends when all stamps of the frame are in
 * (so srb_end_frame / srb_sync on the root returns with the whole frame composited), and the root stamps a release flag
 * when it begins its next frame, which the other ranks' shade kernels wait for before they overwrite the previous frame
 * (they may run at most one frame ahead).  A stamp that does not arrive within 10 s fails the frame (SRB_ERR_CUDA).
 * `handles` is SRB_FB_EXPORT_BYTES bytes; at most 32 ranks. */
#define SRB_FB_EXPORT_BYTES 128u
SRB_API int srb_framebuffer_export(srb_context* ctx, srb_handle fb, void* handles);
SRB_API int srb_framebuffer_import(srb_context* ctx, const void* handles, uint32_t width, uint32_t height,
                                   srb_handle* out);
SRB_API int srb_set_tile_ownership(srb_context* ctx, uint32_t modulus, uint32_t remainder);

/* ---- frame ------------------------------------------------------------------------------------------------ */
/* RenderContext::BeginFrame (Renderer.cpp:201-207) */
SRB_API int srb_begin_frame(srb_context* ctx);
/* RenderContext::ClearFrameBuffer (Renderer.cpp:168-194): depth <- 0.0f (reverse-Z far), colour <- memset(byte) */
SRB_API int srb_clear(srb_context* ctx, srb_handle fb, uint32_t color, int clear_colour, int clear_depth);
/* RenderContext::DrawIndexed (Renderer.cpp:161-166) */
SRB_API int srb_draw_indexed(srb_context* ctx, const srb_draw_desc* draw);
/* RenderContext::EndFrame (Renderer.cpp:209-317): runs the whole pipeline; returns when the frame is complete in
 * the (device-resident) tile buffers.  The _async variant only enqueues; srb_sync() waits. */
SRB_API int srb_end_frame(srb_context* ctx);
SRB_API int srb_end_frame_async(srb_context* ctx);
SRB_API int srb_sync(srb_context* ctx);

/* The viewer's main loop (Viewer/Main.cpp:50-84) run `frames` times in native code — the camera-path batch of
 * BASELINE config 5.  Frame f is rendered by items[f % n_items] (several contexts = several frames in flight on
 * separate CUDA streams; every context must hold the same scene, its `draws` carry that context's handles):
 *   BeginFrame; ClearFrameBuffer(clear_color); DrawIndexed(draws[i]) with mvp = mvps[(f*n_draws + i)*16 ..] (or the
 *   descs' own mvp if mvps == NULL); EndFrame.
 * If colour_out != NULL, frame f's colour tiles (tiles*16384 bytes) are copied to colour_out + f*colour_stride (host
 * memory; pinned memory from srb_host_alloc makes the copy asynchronous).  Returns when every frame is complete. */
typedef struct srb_batch_item
{
	srb_context* ctx;
	const srb_draw_desc* draws;
} srb_batch_item;
SRB_API int srb_render_frames(const srb_batch_item* items, uint32_t n_items, uint32_t n_draws, const float* mvps,
                              uint32_t frames, uint32_t clear_color, void* colour_out, uint64_t colour_stride);
/* Tuning hint: how many frames (contexts of this device) the caller keeps in flight.  >= 4 sizes the grids of the set-up,
 * raster and shade kernels so that the kernels of different frames interleave on the SMs (throughput); fewer favours
 * the latency of one frame.  srb_render_frames sets it from its number of batch items. */
SRB_API int srb_set_frames_in_flight_hint(srb_context* ctx, uint32_t frames_in_flight);
SRB_API void* srb_host_alloc(uint64_t bytes);
/* Pinned host memory with cudaHostAlloc flags: SRB_HOST_WRITE_COMBINED (the CPU only reads what the GPU wrote, rarely:
 * read-back buffers), SRB_HOST_PORTABLE (pinned for every CUDA context of the process). */
#define SRB_HOST_WRITE_COMBINED 1u
#define SRB_HOST_PORTABLE 2u
/* 2 MiB-aligned memory advised to transparent huge pages, touched, then page-locked with cudaHostRegister: fewer I/O
 * translations per DMA when many GPUs of one box copy to the host at once (8 x B200 reading back 8.36 MB frames:
 * 119 -> 165 GB/s over all GPUs, profiles/r02_d2h_ceiling_8gpu.json; one GPU alone: no difference). */
#define SRB_HOST_HUGE_PAGES 4u
SRB_API void* srb_host_alloc_ex(uint64_t bytes, uint32_t flags);
SRB_API void srb_host_free(void* p);
/* Measurement aid: `reps` device-to-host copies of `bytes` bytes each from a device buffer into `host` (any host memory)
 * on the context's stream; *ms = device time of all of them (CUDA events). */
SRB_API int srb_debug_d2h_copies(srb_context* ctx, void* host, uint64_t bytes, uint32_t reps, float* ms);
/* Device-side stopwatch for callers that do not own the library's streams: srb_timer_mark records CUDA event `slot`
 * (0..3) on the context's stream; srb_timer_elapsed waits for (b, slot_b) and returns the milliseconds between
 * (a, slot_a) and (b, slot_b) — a and b may be different contexts on the same device. */
SRB_API int srb_timer_mark(srb_context* ctx, uint32_t slot);
SRB_API int srb_timer_elapsed(srb_context* a, uint32_t slot_a, srb_context* b, uint32_t slot_b, float* ms);
/* Writes `bytes` of scratch device memory on the context's stream (benchmarks use it to evict L2 between steps). */
SRB_API int srb_flush_l2(srb_context* ctx, uint64_t bytes);

/* ---- scene ingestion: sr::Obj::Model (Viewer/Obj.h:13-71, Viewer/Obj.cpp:374-560) ---------------------------- */
/* Host-side, needs no device.  srb_model_load = Obj::Model::Load: if "<path>.bin" exists it is read (the reference's
 * cache in kt::Serialize's byte format, Obj.cpp:15-39,376-397 — files written by either side load in the other),
 * otherwise the OBJ text is parsed with the reference's rules (Obj.cpp:160-312,399-546: vertices de-duplicated per mesh
 * by (pos, uv, normal) index triple in first-use order, quads split 0-1-2 / 0-2-3, negative indices, a mesh per 'g',
 * 16-bit indices up to 65535 vertices, usemtl / mtllib / newmtl / map_Kd) and the cache is written (:548-560).
 * Diffuse maps become Tex::TextureData exactly like TextureData::CreateFromFile (Texture.cpp:103-199): PNG and TGA
 * are decoded here with stb_image's expansion rules, any other format through `decoder` (returns 0 and a malloc'ed
 * width*height*4 RGBA8 image); mips as SRB_MIPS_STB.  Flags = Obj::LoadFlags (Obj.h:52-58) plus two cache switches.
 * Errors and non-fatal notes (a missing MTL or texture leaves the material untextured, like the reference) are in
 * srb_model_last_error() (per thread). */
typedef struct srb_model srb_model;
typedef struct srb_resident_model srb_resident_model;
#define SRB_OBJ_FLIP_WINDING 0x1u
#define SRB_OBJ_GEN_NORMALS 0x2u /* "todo" in the reference: accepted and ignored, as there */
#define SRB_OBJ_FLIP_UVS 0x4u
#define SRB_OBJ_NO_CACHE_READ 0x100u
#define SRB_OBJ_NO_CACHE_WRITE 0x200u
typedef int (*srb_image_decoder)(const char* path, uint8_t** rgba_out, uint32_t* width, uint32_t* height, void* user);
typedef struct srb_mesh_view /* sr::Obj::Mesh, Obj.h:24-43 */
{
	const void* indices;   /* m_indexData */
	uint32_t index_stride; /* 2 (IndexType::u16) or 4 (u32) */
	uint32_t num_indices;  /* m_numIndices */
	const void* vertices;  /* m_vertexData: 32-byte sr::Obj::Vertex {pos[3], norm[3], uv[2]} */
	uint32_t num_vertices;
	uint32_t material;     /* m_matIdx */
} srb_mesh_view;
typedef struct srb_material_view /* sr::Obj::Material, Obj.h:45-53 */
{
	const char* name;
	const uint8_t* texels; /* m_diffuse.m_texels (tiled / Morton / mips); NULL or 0 bytes = no texture */
	uint64_t texel_bytes;
	uint32_t mip_offsets[SRB_MAX_TEX_DIM_LOG2];
	uint32_t num_mips, width_log2, height_log2, bytes_per_pixel;
} srb_material_view;
SRB_API int srb_model_load(const char* path, uint32_t flags, srb_model** out);
SRB_API int srb_model_load_ex(const char* path, uint32_t flags, srb_image_decoder decoder, void* user, srb_model** out);
/* The same with a device at hand (ctx may be NULL = srb_model_load_ex): when the OBJ text is parsed, the materials' textures
 * are tiled and mip-mapped by srb_texture_create_rgba8's kernels instead of on the host — the same bytes in m_texels and
 * in the cache, ~80x sooner for 1024^2 images.  The context is only borrowed for the call. */
SRB_API int srb_model_load_on(srb_context* ctx, const char* path, uint32_t flags, srb_image_decoder decoder, void* user,
                              srb_model** out);
SRB_API void srb_model_free(srb_model* model);
SRB_API const char* srb_model_last_error(void);
SRB_API int srb_model_info(const srb_model* model, uint32_t* num_meshes, uint32_t* num_materials, int* from_cache);
SRB_API int srb_model_mesh(const srb_model* model, uint32_t index, srb_mesh_view* out);
SRB_API int srb_model_material(const srb_model* model, uint32_t index, srb_material_view* out);
SRB_API int srb_model_save_cache(const srb_model* model, const char* bin_path);
/* stbi_load(path, &w, &h, &comp, 4) for the built-in formats (PNG: every colour type / bit depth / interlace; TGA:
 * 24/32-bit and grey, raw or RLE); free with srb_image_free. */
SRB_API int srb_image_load_rgba8(const char* path, uint8_t** rgba_out, uint32_t* width, uint32_t* height);
SRB_API void srb_image_free(uint8_t* rgba);
/* The model made resident on the context's device (one vertex + one index buffer per mesh, one texture per textured
 * material) and the draw list Viewer/Scene.cpp:35-63 issues for it: one draw per mesh, position = attribute buffer =
 * the 32-byte vertices, uv_offset 6, `textured_shader` (SRB_SHADER_UNLIT_DIFFUSE, or SRB_SHADER_SPONZA as
 * Viewer/SponzaScene.cpp:189-215 does) + the material's texture when m_matIdx names a material, else
 * SRB_SHADER_VISUALIZE_NORMALS (UNLIT_DIFFUSE) or the Sponza shader with null uniforms.  draws == NULL returns the
 * count in *n. */
SRB_API int srb_model_make_resident(srb_context* ctx, const srb_model* model, srb_resident_model** out);
SRB_API void srb_resident_model_free(srb_resident_model* resident);
SRB_API int srb_resident_model_draws(const srb_resident_model* resident, srb_handle framebuffer, const float* mvp,
                                     uint32_t textured_shader, srb_draw_desc* draws, uint32_t cap, uint32_t* n);

/* ---- results ---------------------------------------------------------------------------------------------- */
/* Copies the write plane's tiles to host memory in the reference layout: colour tiles are SRB_COLOUR_TILE_BYTES
 * apart, depth tiles `depth_stride` apart (pass SRB_DEPTH_TILE_BYTES to fill a sr::DepthTile array, or 16384 for a
 * packed array).  Either pointer may be NULL. */
SRB_API int srb_read_tiles(srb_context* ctx, srb_handle fb, void* colour_tiles, void* depth_tiles,
                           uint64_t depth_stride);
/* RenderContext::Blit (Renderer.cpp:319-372): de-tile colour into linear RGBA8 (width*height*4 bytes), swap planes,
 * call `on_finish(user)` from a non-submitting thread when the pixels are in host memory. */
SRB_API int srb_blit_linear(srb_context* ctx, srb_handle fb, uint8_t* linear_pixels, void (*on_finish)(void*),
                            void* user);
SRB_API int srb_framebuffer_info(srb_context* ctx, srb_handle fb, uint32_t* width, uint32_t* height,
                                 uint32_t* tiles_x, uint32_t* tiles_y);

/* ---- counters, timing, parity dumps (all valid after srb_end_frame / srb_sync) -------------------------------- */
SRB_API int srb_get_counters(srb_context* ctx, srb_counters* out);
/* Device time of the last frame's kernels in microseconds, in pipeline order; names[i] are static strings. */
SRB_API int srb_get_kernel_times(srb_context* ctx, float* micros, const char** names, uint32_t cap, uint32_t* n);
SRB_API int srb_set_timing(srb_context* ctx, int enabled);
/* Number of kernels this library launched since srb_create (for bench.py's gpu_launches). */
SRB_API uint64_t srb_launch_count(srb_context* ctx);
/* Per-tile triangle list in canonical (draw, triangle, fan) order: number of refs per tile (tiles_x*tiles_y). */
SRB_API int srb_dump_tile_counts(srb_context* ctx, uint32_t* counts, uint32_t num_tiles);
/* The tile-relative records of one tile in list order (what the reference keeps in its sorted BinChunks). */
SRB_API int srb_dump_tile_tris(srb_context* ctx, uint32_t tile_idx, srb_tile_tri* out, uint32_t cap, uint32_t* n);
/* Canonical rank (index into the frame's setup-triangle stream) of each entry of one tile's list. */
SRB_API int srb_dump_tile_ranks(srb_context* ctx, uint32_t tile_idx, uint32_t* out, uint32_t cap, uint32_t* n);
/* Coverage of every list entry of one tile BEFORE the depth-buffer test: for each entry 64 words (block index =
 * (y>>3)*8 + (x>>3)), bit = row*8 + lane, set iff the reference's block loop visits the block
 * (Rasterizer.cpp:221-261) and the sample is inside all edges and z > 0 (Rasterizer.cpp:88-95,134-192). */
SRB_API int srb_dump_tile_coverage(srb_context* ctx, uint32_t tile_idx, uint64_t* masks, uint32_t cap_entries,
                                   uint32_t* n);
/* Re-runs the last frame's tile kernel with a visibility dump: winners[tile*4096 + y*64 + x] = canonical rank of the
 * triangle whose fragment is visible at that pixel, 0xFFFFFFFF where nothing was drawn this frame. */
SRB_API int srb_dump_winners(srb_context* ctx, uint32_t* winners, uint64_t num_pixels);

/* Unit-test entry points: the sampler (Tex::SampleWrap + pack, Texture.cpp:381-452, SIMDUtil.h:87-121) and the RCPPS
 * replay on arbitrary inputs, running the same device code as the tile kernel. */
SRB_API int srb_debug_sample(srb_context* ctx, srb_handle tex, const float* u, const float* v, const float* dudx,
                             const float* dudy, const float* dvdx, const float* dvdy, uint32_t* rgba, uint32_t n);
SRB_API int srb_debug_rcp(srb_context* ctx, const float* in, float* out, uint32_t n);
SRB_API int srb_debug_rsqrt(srb_context* ctx, const float* in, float* out, uint32_t n);

#ifdef __cplusplus
}
#endif
#endif /* SOFTRAST_B200_H */
