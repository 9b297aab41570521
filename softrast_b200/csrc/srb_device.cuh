// srb_device.cuh — device-side data layout and the exact-arithmetic helpers shared by the pipeline kernels.
//
// Everything here is written so that each float operation of the reference (compiled -ffp-contract=off) maps to
// exactly one IEEE-754 round-to-nearest operation on the GPU: explicit __fmul_rn/__fadd_rn/... (never contracted,
// independent of -fmad) and __fmaf_rn only where the reference has an FMA intrinsic.  Integer edge arithmetic is
// 32-bit two's-complement wrap like the reference's int32/AVX2 code.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#define SRB_TILE 64
#define SRB_TILE_LOG2 6
#define SRB_TILE_PIXELS 4096
#define SRB_MAX_VARY 8

namespace srb
{

// ---------------------------------------------------------------------------------------------------------------
// HBM layout
// ---------------------------------------------------------------------------------------------------------------

// Texture descriptor = the fields of sr::Tex::TextureData (reference SoftRast/Texture.h:33-40).
struct TexDev
{
	const uint8_t* texels;
	uint32_t mipOffsets[14];
	uint32_t numMips;
	uint32_t widthLog2;
	uint32_t heightLog2;
	uint32_t bytes;
};

// One recorded draw = sr::DrawCall (reference SoftRast/Renderer.h:119-150) with device pointers.
struct DrawDev
{
	const uint8_t* idx;
	const uint8_t* pos;
	const uint8_t* attr;
	uint32_t idxStride;
	uint32_t posStride;
	uint32_t attrStride;
	uint32_t numTris;
	uint32_t triBase;     // number of input triangles of all earlier draws (canonical order = draw-major)
	uint32_t shader;
	uint32_t uvOffset;
	uint32_t numVaryings; // attrStride / 4
	int32_t texture;      // index into the texture table, -1 = null
	uint32_t planeMask;   // bit j: the attribute plane of varying j is set up (what the shader reads, or all varyings)
	uint32_t chunkBase;   // number of 256-triangle set-up chunks of all earlier draws (a chunk never straddles two draws)
	uint32_t pad;
	float mvp[16];        // column-major
};

// Set-up triangles are identified by a CANONICAL KEY that preserves the reference's single-threaded order
// (draw, triangle, fan): key = g*8 for the single output of an unclipped input triangle g (g = index of the input
// triangle over all draws of the frame, draw-major) and key = g*8 + 1 + f for fan triangle f of a clipped one.
// Records live at a SLOT: slot = g for unclipped triangles (sparse storage, culled triangles leave holes that nobody
// reads), slots >= numInputTris are handed out atomically to clipped fans.  Draw order never depends on where a record
// is stored or on the order atomics ran in: it is carried by the key into the depth resolve (srb_raster.cu).
#define SRB_KEY_UNCLIPPED(g) ((uint32_t)(g) << 3)
#define SRB_KEY_FAN(g, f) (((uint32_t)(g) << 3) + 1u + (uint32_t)(f))
#define SRB_MAX_INPUT_TRIS (1u << 28)

// Raster record: what binning and the rasteriser need, 64 bytes.
// Screen-space edge equations (Binning.cpp:242-259), pixel bounding box (:315-322), z/w plane (:336-338) with the
// vertex-0 value and vertex-0 raster position from which tile-relative constants are derived (:436-444).
struct __align__(16) RasterRec
{
	int32_t c[3];
	int32_t dx[3];
	int32_t dy[3];
	uint16_t xmin, xmax, ymin, ymax;
	float zdx, zdy, z0;
	float r0x, r0y;
};
static_assert(sizeof(RasterRec) == 64, "RasterRec must be 64 bytes");

// Shade record: 1/w plane and the attribute/w planes (Binning.cpp:340-350), 128 bytes.
// `info` packs what the pixel shader needs from the draw so that shading does not chase the draw table:
// shader | min(uvOffset, 255) << 8 | (texture index + 1) << 16 (0 = null texture).  pad[0] = draw index (dumps).
// Planes are stored as (dx, dy, vertex-0 value) triplets, varying j at pl[SRB_PLANE_SLOT(j)]: varyings 6 and 7 (the
// sampler's u, v) come first, so the textured shader touches only the first 64 bytes of the record.
// For a CLIPPED input triangle g, shadeRecs[g] is not a triangle but a redirect: {pad[0] = first fan slot,
// pad[1] = mask of surviving fan indices}; fan f lives at slot pad[0] + popc(pad[1] & ((1 << f) - 1)).
#define SRB_PLANE_SLOT(j) (((j) + 2u) & 7u)
struct __align__(16) ShadeRec
{
	float wdx, wdy, w0;
	uint32_t info;
	float r0x, r0y;
	uint32_t pad[2];
	float pl[SRB_MAX_VARY][3];
};
static_assert(sizeof(ShadeRec) == 128, "ShadeRec must be 128 bytes");

// A (triangle, tile) reference / a surviving triangle: canonical key + record slot.
struct __align__(8) KeySlot
{
	uint32_t key;
	uint32_t slot;
};

// An entry of the survivor list (K1 -> bin fill): the triangle's key and slot and, for a triangle whose bounding box lies in
// ONE tile (four out of five on the hall scene), that tile and its packed block range — the bin fill then appends the
// reference from this entry alone, without fetching the 64-byte record and walking its bin range twice.
// tile = 0x80000000 | tile index, or 0: walk the record's bin range (Binning.cpp:352-410).
struct __align__(16) Survivor
{
	uint32_t key;
	uint32_t slot;
	uint32_t tile;
	uint32_t blocks;
};

// A (triangle, tile) reference as the tile lists hold it: canonical key, record slot, and the 8x8-block range the
// reference's block loops visit inside that tile (Rasterizer.cpp:201-223: begin = blockMin & ~7, end = blockMax
// exclusive), packed bx0 | bx1 << 4 | by0 << 8 | by1 << 12 with bx1/by1 exclusive, so that a rasteriser warp can
// decide from the list alone whether a triangle can touch its part of the tile.
struct __align__(16) TileRef
{
	uint32_t key;
	uint32_t slot;
	uint32_t blocks;
	uint32_t quads; // bit q = qy * 4 + qx: the block range touches the tile's 16x16-pixel quad (qx, qy)
};

__device__ __forceinline__ uint32_t pack_block_range(int32_t minX, int32_t maxX, int32_t minY, int32_t maxY)
{
	uint32_t const bx0 = (uint32_t)minX >> 3, by0 = (uint32_t)minY >> 3;
	uint32_t const bx1 = ((uint32_t)maxX + 7u) >> 3, by1 = ((uint32_t)maxY + 7u) >> 3;
	return bx0 | (bx1 << 4) | (by0 << 8) | (by1 << 12);
}

// Which of the tile's sixteen 16x16 quads (2x2 blocks each) a packed block range touches.
__device__ __forceinline__ uint32_t quad_mask(uint32_t blocks)
{
	uint32_t const bx0 = blocks & 15u, bx1 = (blocks >> 4) & 15u, by0 = (blocks >> 8) & 15u, by1 = (blocks >> 12) & 15u;
	if (bx0 >= bx1 || by0 >= by1)
	{
		return 0u;
	}
	// quads qx0 .. qx1 inclusive, from blocks bx0 .. bx1 - 1
	uint32_t const qx0 = bx0 >> 1, qx1 = (bx1 - 1u) >> 1, qy0 = by0 >> 1, qy1 = (by1 - 1u) >> 1;
	uint32_t const xm = ((2u << qx1) - 1u) & ~((1u << qx0) - 1u); // 4 bits
	uint32_t const rows = ((2u << qy1) - 1u) & ~((1u << qy0) - 1u);
	uint32_t const spread = (rows & 1u) | ((rows & 2u) << 3) | ((rows & 4u) << 6) | ((rows & 8u) << 9); // bit qy -> bit 4*qy
	return spread * xm;
}

// One unit of tile work for the raster kernel: a slice [begin, end) of one tile's reference list.
struct __align__(16) UnitDesc
{
	uint32_t tile;
	uint32_t begin;
	uint32_t end;
	uint32_t unitsInTile;
};

// Frame control block (device), zeroed at the start of every frame.
struct FrameCtl
{
	uint32_t numSurvivors; // triangles surviving clip/cull == entries of the survivor list
	uint32_t numClipQueue; // input triangles queued for the clipper
	uint32_t numFanSlots;  // slots handed out beyond numInputTris
	uint32_t totalRefs;
	uint32_t tilesNonEmpty;
	uint32_t maxRefs;
	uint32_t overflow;     // bit0: fan slots, bit1: tile refs, bit2: units
	uint32_t pixelsCovered;
	uint32_t numUnits;
	uint32_t unitTicket;   // raster: persistent-CTA unit dispenser
	uint32_t unitSize;
	uint32_t ctasDone;     // clip kernel: CTAs that are done (the last one runs the tile scan)
	uint32_t frameDone;    // screen-tile split: shade CTAs that have stored their tiles (the last one stamps the arrival flag)
	uint32_t doneValue;    // screen-tile split: the stamp (set by the host in the head upload, not zero)
	uint32_t pad[2];
};
static_assert(sizeof(FrameCtl) == 64, "FrameCtl is one 64-byte block in front of the draw table");

struct FrameParams
{
	uint32_t width, height;
	uint32_t tilesX, tilesY;
	uint32_t numDraws;
	uint32_t numInputTris;
	uint32_t slotCapacity; // numInputTris + capacity for clipped fans
	uint32_t refCapacity;
	uint32_t unitCapacity;
	uint32_t clearPending; // a clear is folded into this frame: every tile must be written
	uint32_t splitTiles;   // tiles may be split into several units (needs a depth clear)
	uint32_t ownMod;       // screen-tile split across GPUs: this context owns tiles with tile % ownMod == ownRem
	uint32_t ownRem;
	uint32_t minUnit;      // smallest slice of a tile's reference list handed to the rasteriser as one unit
	uint32_t smemHist;     // set-up: the per-CTA tile histogram fits in shared memory (else: global atomics per reference)
	uint32_t smemBase;     // set-up: the draws' triBase / chunkBase table fits in shared memory (else: binary search in global memory)
	uint32_t numChunks;    // set-up: 256-triangle chunks of the frame (sum over draws of ceil(numTris / 256))
};

__device__ __forceinline__ bool tile_owned(const FrameParams& fp, uint32_t tile)
{
	return fp.ownMod <= 1u || (tile % fp.ownMod) == fp.ownRem;
}

// ---------------------------------------------------------------------------------------------------------------
// exact float helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float mulf(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float addf(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float subf(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float divf(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// x86 cvttss2si: truncate; out of range / NaN -> 0x80000000 ("integer indefinite").
__device__ __forceinline__ int32_t cvtt_x86(float f)
{
	return (f >= -2147483648.0f && f < 2147483648.0f) ? __float2int_rz(f) : (int32_t)0x80000000;
}
// x86 cvtps2dq: round to nearest even; out of range / NaN -> 0x80000000.
__device__ __forceinline__ int32_t cvtn_x86(float f)
{
	return (f >= -2147483648.0f && f < 2147483648.0f) ? __float2int_rn(f) : (int32_t)0x80000000;
}
// x86 maxps(a, b): (a > b) ? a : b  -> returns b if either operand is NaN.
__device__ __forceinline__ float max_x86(float a, float b) { return (a > b) ? a : b; }

__device__ __forceinline__ int32_t wrap_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
__device__ __forceinline__ int32_t wrap_sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
__device__ __forceinline__ int32_t wrap_mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }

// The host CPU's RCPPS replayed from its mantissa table (reference Rasterizer.cpp:375-376 uses _mm256_rcp_ps).
// table[i] = bits(RCPPS(1 + i * 2^-bits)); the result for (sign, exponent e, mantissa m) is
// table[m >> (23-bits)] + ((127 - e) << 23), flushed to zero when the exponent underflows.
static __device__ __noinline__ float rcp_x86_special(float x, const uint32_t* __restrict__ table, uint32_t bits)
{
	uint32_t const u = __float_as_uint(x);
	uint32_t const s = u & 0x80000000u;
	uint32_t const e = (u >> 23) & 0xFFu;
	uint32_t const m = u & 0x7FFFFFu;
	if (e == 0u)
	{
		return __uint_as_float(s | 0x7F800000u); // +-0 and denormals (treated as zero) -> +-inf
	}
	if (e == 0xFFu)
	{
		return m ? __uint_as_float(u | 0x00400000u) : __uint_as_float(s); // NaN -> qNaN ; +-inf -> +-0
	}
	int32_t const r = (int32_t)table[m >> (23u - bits)] + ((127 - (int32_t)e) << 23);
	if (r < 0x00800000)
	{
		return __uint_as_float(s); // result would be denormal: flushed to +-0
	}
	return __uint_as_float(s | (uint32_t)r);
}

__device__ __forceinline__ float rcp_x86(float x, const uint32_t* __restrict__ table, uint32_t bits)
{
	uint32_t const u = __float_as_uint(x);
	uint32_t const eb = u & 0x7F800000u;
	if (eb - 0x00800000u < 0x7E000000u)
	{
		// biased exponent 1..252: normal input, normal result (table entries lie in (0.5, 1])
		uint32_t const r = __ldg(&table[(u >> (23u - bits)) & ((1u << bits) - 1u)]) + 0x3F800000u - eb;
		return __uint_as_float((u & 0x80000000u) | r);
	}
	return rcp_x86_special(x, table, bits);
}

// The host CPU's RSQRTPS replayed from its table (reference Viewer/SponzaScene.cpp:66 uses _mm256_rsqrt_ps).
// x = 2^(2k+p) * m: table[(p << bits) | top bits of m's mantissa] with the exponent field lowered by k (srb_host.cpp).
__device__ __forceinline__ float rsqrt_x86(float x, const uint32_t* __restrict__ table, uint32_t bits)
{
	uint32_t const u = __float_as_uint(x);
	uint32_t const e = (u >> 23) & 0xFFu;
	if ((int32_t)u >= 0x00800000 && e != 0xFFu)
	{
		// positive normal
		int32_t const ue = (int32_t)e - 127;
		uint32_t const p = (uint32_t)ue & 1u;
		int32_t const k = (ue - (int32_t)p) >> 1;
		return __uint_as_float(__ldg(&table[(p << bits) | ((u & 0x7FFFFFu) >> (23u - bits))]) - ((uint32_t)k << 23));
	}
	if (e == 0xFFu && (u & 0x7FFFFFu))
	{
		return __uint_as_float(u | 0x00400000u); // NaN -> quiet NaN
	}
	if (e == 0u)
	{
		return __uint_as_float((u & 0x80000000u) | 0x7F800000u); // +-0 and denormals (treated as zero) -> +-inf
	}
	if ((int32_t)u < 0)
	{
		return __uint_as_float(0xFFC00000u); // negative (incl. -inf): the "real indefinite" NaN
	}
	return 0.0f; // +inf
}

// Frame constants of the Sponza pixel shader as the shade kernel reads them (srb_sponza_constants in the C ABI).
struct SponzaLightDev
{
	float pos[3];
	float colour[3];
	float intensity;
	float falloff;
};
struct SponzaDev
{
	float sunDir[3];
	float ambient[3];
	float pad[2];
	SponzaLightDev lights[16];
};

// ---------------------------------------------------------------------------------------------------------------
// bin traversal shared by the count (setup) and fill (bin) kernels — reference Binning.cpp:352-410.
// Calls f(tileIdx) for every bin the reference appends the triangle to, in the reference's loop order.
// ---------------------------------------------------------------------------------------------------------------
struct BinRange
{
	uint32_t bx0, bx1, by0, by1;
	bool check;
};

__device__ __forceinline__ BinRange bin_range(uint32_t xmin, uint32_t xmax, uint32_t ymin, uint32_t ymax)
{
	BinRange r;
	r.by0 = ymin >> SRB_TILE_LOG2;
	r.by1 = ymax >> SRB_TILE_LOG2;
	r.bx0 = xmin >> SRB_TILE_LOG2;
	r.bx1 = xmax >> SRB_TILE_LOG2;
	// Binning.cpp:358-370: numXbins is computed from the Y range too, so the coverage check is skipped whenever
	// the box spans <= 2 bin rows, whatever its width.
	uint32_t const numY = r.by1 - r.by0 + 1;
	r.check = !(numY <= 2);
	return r;
}

// Binning.cpp:381-409: an edge "has a corner inside" if E > 0 at one of the four bin corners (screen-space C).
__device__ __forceinline__ bool bin_overlaps(const int32_t c[3], const int32_t dx[3], const int32_t dy[3], int32_t X0,
                                             int32_t Y0)
{
	int32_t const X1 = X0 + SRB_TILE, Y1 = Y0 + SRB_TILE;
#pragma unroll
	for (int k = 0; k < 3; ++k)
	{
		int32_t const e00 = wrap_add(wrap_add(c[k], wrap_mul(dy[k], X0)), wrap_mul(dx[k], Y0));
		int32_t const e01 = wrap_add(wrap_add(c[k], wrap_mul(dy[k], X0)), wrap_mul(dx[k], Y1));
		int32_t const e10 = wrap_add(wrap_add(c[k], wrap_mul(dy[k], X1)), wrap_mul(dx[k], Y0));
		int32_t const e11 = wrap_add(wrap_add(c[k], wrap_mul(dy[k], X1)), wrap_mul(dx[k], Y1));
		if (!((e00 > 0) | (e01 > 0) | (e10 > 0) | (e11 > 0)))
		{
			return false;
		}
	}
	return true;
}

// Tile-relative quantities of one (triangle, tile) pair — reference Binning.cpp:421-452.
struct TileEdges
{
	int32_t c[3];
	int32_t minX, maxX, minY, maxY; // block bbox, each clamp(v - origin, 0, 64)
};

__device__ __forceinline__ int32_t clampi(int32_t v, int32_t lo, int32_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ TileEdges tile_edges(const RasterRec& r, int32_t X0, int32_t Y0)
{
	TileEdges t;
#pragma unroll
	for (int k = 0; k < 3; ++k)
	{
		t.c[k] = wrap_add(r.c[k], wrap_add(wrap_mul(r.dx[k], Y0), wrap_mul(r.dy[k], X0)));
	}
	t.minX = clampi((int32_t)r.xmin - X0, 0, SRB_TILE);
	t.maxX = clampi((int32_t)r.xmax - X0, 0, SRB_TILE);
	t.minY = clampi((int32_t)r.ymin - Y0, 0, SRB_TILE);
	t.maxY = clampi((int32_t)r.ymax - Y0, 0, SRB_TILE);
	return t;
}

// Plane constant moved to the tile origin: c0 = (dx*sx + dy*sy) + q0 with sx = (float)X0 - r0.x (Binning.cpp:436-452).
__device__ __forceinline__ float plane_c0(float pdx, float pdy, float q0, float sx, float sy)
{
	return addf(addf(mulf(pdx, sx), mulf(pdy, sy)), q0);
}

} // namespace srb
