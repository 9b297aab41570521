// srb_raster.cu — K3 (raster_kernel: per-tile rasterisation + depth resolve) and K4 (shade_kernel: fragment shading),
// plus the parity dump kernels.
//
// Replaces the reference back-end RasterAndShadeBin (SoftRast/Rasterizer.cpp:525-577):
//   RasterizeTrisInBin_OutputFragments (:194-304), ComputeBlockMask8x8[_DepthOnly] (:97-192),
//   ComputeInterpolants (:306-458), ShadeFragmentBuffer (:460-523), the pixel shaders (Viewer/Shaders.h:71-130)
//   and the sampler Tex::SampleWrap (SoftRast/Texture.cpp:381-452, :212-233, :243-379).
//
// Design (not the reference's): a tile's depth and winning-triangle id are resolved as one 64-bit key per pixel,
// key = depthBits << 32 | (0xFFFFFFFE - rank), held in the registers of the lane that owns the pixel.  The reference walks the tile's
// triangles serially with a strict `z > stored` test, so the surviving fragment of a pixel is the one with the largest
// z, and among equal z the FIRST in canonical order: exactly max(key).  Max is order independent, so all 8x8 blocks of
// all triangles are resolved in parallel, 8 lanes per block (one lane per column walking 8 rows, the same evaluation
// order as the reference's 8-wide AVX2 rows, which keeps depth bit-exact).  Only the final
// visible fragment of each pixel is shaded (the reference shades every fragment that passed early-Z when it was
// drawn, then overwrites), one thread per pixel, and colour + depth tiles are written once with coalesced stores in
// the reference's ColourTile/DepthTile layout.
#include "srb_device.cuh"
#include "srb_kernels.h"
#include "../../include/softrast_b200.h"

#include <stdlib.h>

namespace srb
{

namespace
{

// Counters of a statistics build (make STATS=1 -> libsoftrast_b200_stats.so, profiles/stats.py); nothing in the product.
#ifdef SRB_STATS
__device__ unsigned long long g_stats[16];
__device__ int g_nullTaps; // statistics build: every texel tap reads texel 0 of its texture (see profiles/texel_taps.py)
#define SRB_STAT(i, n)                                                      \
	do                                                                      \
	{                                                                       \
		if ((threadIdx.x & 31u) == (uint32_t)(__ffs(__activemask()) - 1))   \
			atomicAdd(&g_stats[i], (unsigned long long)(n));                \
	} while (0)
#define SRB_STAT_LANE(i, n) atomicAdd(&g_stats[i], (unsigned long long)(n))
#else
#define SRB_STAT(i, n) ((void)0)
#define SRB_STAT_LANE(i, n) ((void)0)
#endif

constexpr int kRasterThreads = 128;
constexpr uint32_t kNoWinner = 0xFFFFFFFFu; // low key word of a pixel that has not received a fragment this frame

struct TriTile
{
	int32_t c[3], dx[3], dy[3];
	float zc0, zdx, zdy;
};

__device__ __forceinline__ void load_raster_rec(const RasterRec* __restrict__ recs, uint32_t rank, RasterRec& r)
{
	const uint4* p = reinterpret_cast<const uint4*>(recs + rank);
	uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
	for (int i = 0; i < 4; ++i) d[i] = __ldg(p + i);
}

// Reference coarse test, Rasterizer.cpp:224-261: the "8x8 block" is tested over a 64x64 extent from its origin.
// returns 0 = skip, 1 = full path, 2 = depth-only path (all 12 corners > 0).  e00[k] = edge k at the block origin.
// (All arithmetic is modulo 2^32 like the reference's int32, so the four corners can be formed by adding 64*dx, 64*dy.)
__device__ __forceinline__ int ref_coarse(const TriTile& t, int32_t xB, int32_t yB, int32_t (&e00)[3])
{
	bool all = true, any = true;
#pragma unroll
	for (int k = 0; k < 3; ++k)
	{
		int32_t const a = wrap_add(wrap_add(t.c[k], wrap_mul(t.dy[k], xB)), wrap_mul(t.dx[k], yB));
		int32_t const b = wrap_add(a, (int32_t)((uint32_t)t.dx[k] << 6)); // (xB, yB + 64)
		int32_t const c = wrap_add(a, (int32_t)((uint32_t)t.dy[k] << 6)); // (xB + 64, yB)
		int32_t const d = wrap_add(b, (int32_t)((uint32_t)t.dy[k] << 6)); // (xB + 64, yB + 64)
		e00[k] = a;
		any = any && (max(max(a, b), max(c, d)) > 0);
		all = all && (min(min(a, b), min(c, d)) > 0);
	}
	return any ? (all ? 2 : 1) : 0;
}

// z/w of lane `l`, row 0 of block (xB, yB): Rasterizer.cpp:213 (tileTopLeft = fma(ramp, dx, c0)) and :156-157.
__device__ __forceinline__ float block_z0(const TriTile& t, int32_t xB, int32_t yB, int32_t l)
{
	float const topLeft = fma_((float)l, t.zdx, t.zc0);
	float const z = fma_((float)yB, t.zdy, topLeft);
	return addf(z, mulf((float)xB, t.zdx));
}

// ---------------------------------------------------------------------------------------------------------------
// sampler + shaders
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t part1by1_5(uint32_t x)
{
	x &= 0x1Fu;
	x = (x | (x << 4)) & 0x10Fu;  // ---4 ---- 3210
	x = (x | (x << 2)) & 0x133u;  // ---4 --32 --10
	x = (x | (x << 1)) & 0x155u;  // ---4 -3-2 -1-0
	return x;
}


__device__ __forceinline__ float lerp_fma(float a, float b, float t)
{
	// SIMDUtil.h:141-145: fmadd(t, b, fnmadd(t, a, a))
	return fma_(t, b, fma_(-t, a, a));
}

__device__ __forceinline__ uint32_t pack_channel(float c)
{
	// SIMDUtil.h:87-106: cvtps(fma(c,255,.5)) -> packus_epi32 (sat to u16) -> packus_epi16 (AS SIGNED i16, sat to u8)
	// i < 0 -> 0 (first pack); 32768..65535 and everything saturated to 65535 -> negative as i16 -> 0 (second pack);
	// 256..32767 -> 255.  cvtps2dq's 0x80000000 for NaN / out of range is negative -> 0, which is also what the
	// saturating __float2int_rn gives after this mapping (NaN -> 0, +-overflow -> INT_MAX / INT_MIN -> 0).
	int32_t const i = __float2int_rn(fma_(c, 255.0f, 0.5f));
	return ((uint32_t)i > 32767u) ? 0u : (uint32_t)min(i, 255);
}

__device__ __forceinline__ uint32_t pack_rgba(float r, float g, float b, float a)
{
	return pack_channel(r) | (pack_channel(g) << 8) | (pack_channel(b) << 16) | (pack_channel(a) << 24);
}

// The same pack for channel values that are known to lie in [0, 1 + a rounding error] or to be NaN — what the bilinear
// filter of texels in [0, 1] with weights in [0, 1) produces: there the integer is 0 .. 256 (or the indefinite value for
// NaN), so the two saturating packs reduce to a clamp to 0 .. 255 with NaN -> 0, which is exactly cvt.rni.sat.u8.f32.
__device__ __forceinline__ uint32_t pack_channel_unit(float c)
{
	uint32_t r;
	asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(fma_(c, 255.0f, 0.5f)));
	return r;
}

__device__ __forceinline__ uint32_t pack_rgba_unit(float r, float g, float b, float a)
{
	return pack_channel_unit(r) | (pack_channel_unit(g) << 8) | (pack_channel_unit(b) << 16) | (pack_channel_unit(a) << 24);
}

__device__ __forceinline__ void wrap_coord(float u, uint32_t dim, uint32_t& i0, float& frac)
{
	// Texture.cpp:410-436
	uint32_t const sign = __float_as_uint(u) & 0x80000000u;
	float const au = __uint_as_float(__float_as_uint(u) ^ sign);
	float fr = subf(au, floorf(au));
	if (sign)
	{
		fr = subf(1.0f, fr);
	}
	float const t = mulf((float)dim, fr);
	float const tf = floorf(t);
	frac = subf(t, tf);
	// tf is in [0, dim] or NaN; cvtps2dq(NaN) = 0x80000000 and __float2int_rn(NaN) = 0 agree after the mask
	i0 = (uint32_t)__float2int_rn(tf) & (dim - 1u);
}

// Texel channel -> float: k * (float)byte with k = 1.0f / 255.0f (Texture.cpp:438-456: cvtepi32_ps, then the multiply).
// Computed without the conversion unit (16 conversions per pixel made it the busiest pipe of the shade kernel): a byte
// permute builds the float 2^23 + byte, and ONE fused multiply-add takes the 2^23 out again:
//   fma(2^23 + b, k, -(2^23 * k)) = round((2^23 + b) * k - 2^23 * k) = round(b * k)      (2^23 * k is exact: a power of two)
// which is the reference's single rounding of b * k, bit for bit.
__device__ __forceinline__ void texel_to_float(uint32_t px, float (&o)[4])
{
	float const k = 1.0f / 255.0f;
#ifdef SRB_EXP_I2F_TEXEL
	o[0] = mulf(k, (float)(px & 0xFFu));
	o[1] = mulf(k, (float)((px >> 8) & 0xFFu));
	o[2] = mulf(k, (float)((px >> 16) & 0xFFu));
	o[3] = mulf(k, (float)(px >> 24));
#else
	float const bias = -(8388608.0f * k);
	o[0] = fma_(__uint_as_float(__byte_perm(px, 0x4B000000u, 0x7650)), k, bias);
	o[1] = fma_(__uint_as_float(__byte_perm(px, 0x4B000000u, 0x7651)), k, bias);
	o[2] = fma_(__uint_as_float(__byte_perm(px, 0x4B000000u, 0x7652)), k, bias);
	o[3] = fma_(__uint_as_float(__byte_perm(px, 0x4B000000u, 0x7653)), k, bias);
#endif
}

// The shade kernel's tables in shared memory, read through ONE shared-space base address that the compiler must keep in
// a register (tables_base): left to itself it re-derives the address of every table at every use from the CTA's rank in
// its cluster (S2R SR_CgaCtaId and three more instructions, four times per pixel = 13 of the 345 instructions of a pixel).
struct ShadeTables
{
	TexDev texs[48];      // texture descriptors of the frame (kSmemTexs)
	uint16_t rcp16[2048]; // RCPPS table packed to 16 bits per entry (1 << kFastRcpBits)
	uint16_t spread[32];  // 5-bit Morton spread
};
__shared__ __align__(16) ShadeTables g_sTab;

__device__ __forceinline__ uint32_t tables_base()
{
	uint32_t b = (uint32_t)__cvta_generic_to_shared(&g_sTab);
	asm volatile("" : "+r"(b)); // opaque from here on: not rematerialised
	return b;
}

template <uint32_t kOffset>
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr)
{
	uint16_t v;
	asm("ld.shared.u16 %0, [%1+%2];" : "=h"(v) : "r"(addr), "n"(kOffset));
	return v;
}

template <uint32_t kOffset>
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
	uint32_t v;
	asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(kOffset));
	return v;
}

template <uint32_t kOffset>
__device__ __forceinline__ uint2 lds_u32x2(uint32_t addr)
{
	uint2 v;
	asm("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(addr), "n"(kOffset));
	return v;
}

template <uint32_t kOffset>
__device__ __forceinline__ uint4 lds_u32x4(uint32_t addr)
{
	uint4 v;
	asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr), "n"(kOffset));
	return v;
}

// 5-bit Morton spread table (x -> bits of x at the even positions), filled by the first warp of a CTA.
__device__ __forceinline__ void fill_spread_table(uint16_t* spread)
{
	if (threadIdx.x < 32u)
	{
		spread[threadIdx.x] = (uint16_t)part1by1_5(threadIdx.x);
	}
}

// What the sampler needs of a texture descriptor before it knows the mip level.
struct TexHead
{
	const uint8_t* texels;
	uint32_t numMips, widthLog2, heightLog2, bytes;
};

__device__ __forceinline__ TexHead load_tex_head(const TexDev* t)
{
	// TexDev: texels (8 bytes) ... {numMips, widthLog2} and {heightLog2, bytes} are 8-byte aligned pairs
	TexHead h;
	h.texels = t->texels;
	uint2 const a = *reinterpret_cast<const uint2*>(&t->numMips);
	uint2 const b = *reinterpret_cast<const uint2*>(&t->heightLog2);
	h.numMips = a.x;
	h.widthLog2 = a.y;
	h.heightLog2 = b.x;
	h.bytes = b.y;
	return h;
}

// The same from the shared-memory copy (tab = tables_base()).
__device__ __forceinline__ TexHead load_tex_head_shared(uint32_t tab, uint32_t ti)
{
	static_assert(sizeof(TexDev) == 80 && offsetof(TexDev, numMips) == 64, "TexDev layout");
	uint32_t const at = tab + ti * (uint32_t)sizeof(TexDev);
	uint2 const p = lds_u32x2<0>(at);
	uint4 const q = lds_u32x4<64>(at);
	TexHead h;
	h.texels = reinterpret_cast<const uint8_t*>(((unsigned long long)p.y << 32) | p.x);
	h.numMips = q.x;
	h.widthLog2 = q.y;
	h.heightLog2 = q.z;
	h.bytes = q.w;
	return h;
}

// Tex::SampleWrap + RGBA32SoA_To_RGBA8AoS for one fragment.  `spread` = the 32-entry table above (shared memory);
// `mipOffset(mip)` returns TexDev::mipOffsets[mip] from wherever the caller keeps the descriptor.
// `modulate`: the RGB factors the Sponza shader multiplies the sample by before packing (nullptr = none).
// kQuadTap: test every sample for the 128-bit tap below (off in the frame kernels unless SRB_QUAD_TAPS=1: the vote and its
// predicates cost ten instructions per pixel and the registers of both variants, and 1 in 64 800 warps of the hall qualifies).
template <bool kQuadTap, typename MipOffset>
__device__ __forceinline__ uint32_t sample_wrap(const TexHead& tex, MipOffset&& mipOffset, uint32_t spreadAt, float u,
                                                float v, float dudx, float dudy, float dvdx, float dvdy,
                                                const float* modulate = nullptr)
{
	// CalcMipLevels, Texture.cpp:212-233 (note the mixed axes: dudy*height, dvdx*width)
	uint32_t const numMips = tex.numMips;
	uint2 const logs = make_uint2(tex.widthLog2, tex.heightLog2);
	float const Wt = (float)(1u << logs.x), Ht = (float)(1u << logs.y);
	float const a = mulf(dudx, Wt), b = mulf(dudy, Ht), c = mulf(dvdx, Wt), d = mulf(dvdy, Ht);
	float const du2 = fma_(a, a, mulf(b, b));
	float const dv2 = fma_(c, c, mulf(d, d));
	float const m = max_x86(du2, dv2);
	// The reference takes the exponent of sqrt(m) (ExtractExponent, SIMDUtil.h:24-30).  A correctly rounded square root
	// never crosses a power of two, so for m = 2^e * f its exponent is floor(e / 2); zero/denormal m gives a negative
	// exponent either way (clamped to mip 0) and inf/NaN give >= 64 either way (clamped to the last mip).
	int32_t const e = ((int32_t)((__float_as_uint(m) >> 23) & 0xFFu) - 127) >> 1;
	int32_t const mip = min((int32_t)numMips - 1, max(0, e));
	uint32_t const wl = logs.x - min(logs.x, (uint32_t)mip), hl = logs.y - min(logs.y, (uint32_t)mip);
	uint32_t const w = 1u << wl, h = 1u << hl;
	uint32_t x0, y0;
	float fu, fv;
	wrap_coord(u, w, x0, fu);
	wrap_coord(v, h, y0, fv);
	uint32_t const x1 = (x0 + 1u) & (w - 1u), y1 = (y0 + 1u) & (h - 1u);
	// Texture.cpp:73-101 / :267-291: 32x32 tiles (row-major, max(w,32)/32 per row), Morton inside (x in the even bits,
	// y in the odd bits); 4 bytes per texel
	uint32_t const rowShift = 12u + (wl > 5u ? wl - 5u : 0u);
	uint32_t const ox0 = ((x0 >> 5) << 12) | (lds_u16<0>(spreadAt + ((x0 & 31u) << 1)) << 2);
	uint32_t const ox1 = ((x1 >> 5) << 12) | (lds_u16<0>(spreadAt + ((x1 & 31u) << 1)) << 2);
	uint32_t const oy0 = ((y0 >> 5) << rowShift) | (lds_u16<0>(spreadAt + ((y0 & 31u) << 1)) << 3);
	uint32_t const oy1 = ((y1 >> 5) << rowShift) | (lds_u16<0>(spreadAt + ((y1 & 31u) << 1)) << 3);
	const uint8_t* base = tex.texels + mipOffset((uint32_t)mip);
#ifdef SRB_STATS
	// Differential measurement of what the texel taps cost the memory system: with g_nullTaps set, the four taps of every
	// sample read texel 0 of the texture (always a hit after the first touch); the kernel's L1 / L2 sector and hit counters
	// with and without it differ by exactly the taps' traffic.
	uint32_t const nullMask = g_nullTaps ? 0u : 0xFFFFFFFFu;
	if (g_nullTaps) base = tex.texels;
#else
	uint32_t const nullMask = 0xFFFFFFFFu;
#endif
	uint32_t p00, p10, p11, p01;
	// The 2x2 footprint of a tap with even x0 and y0 is ONE aligned 16-byte group of the Morton order (x in the even
	// bits, y in the odd bits: +1 = x + 1, +2 = y + 1, +3 = both), inside one 32x32 tile: one 128-bit load instead of
	// four 32-bit ones.  Taken when all the lanes that are sampling together agree (a vote, so that the warp does not
	// run both variants): strongly magnified textures.
	bool const quad = kQuadTap && ((x0 | y0) & 1u) == 0u && wl != 0u && hl != 0u;
	if (kQuadTap && __all_sync(__activemask(), quad))
	{
		uint4 const q = __ldg(reinterpret_cast<const uint4*>(base + ((ox0 + oy0) & nullMask)));
		p00 = q.x;
		p10 = q.y;
		p01 = q.z;
		p11 = q.w;
		SRB_STAT(6, 1);
	}
	else
	{
		p00 = __ldg(reinterpret_cast<const uint32_t*>(base + ((ox0 + oy0) & nullMask)));
		p10 = __ldg(reinterpret_cast<const uint32_t*>(base + ((ox1 + oy0) & nullMask)));
		p11 = __ldg(reinterpret_cast<const uint32_t*>(base + ((ox1 + oy1) & nullMask)));
		p01 = __ldg(reinterpret_cast<const uint32_t*>(base + ((ox0 + oy1) & nullMask)));
	}
	SRB_STAT(5, 1);
	float t00[4], t10[4], t11[4], t01[4];
	texel_to_float(p00, t00);
	texel_to_float(p10, t10);
	texel_to_float(p11, t11);
	texel_to_float(p01, t01);
	float out[4];
#pragma unroll
	for (int k = 0; k < 4; ++k)
	{
		float const left = lerp_fma(t00[k], t01[k], fv);
		float const right = lerp_fma(t10[k], t11[k], fv);
		out[k] = lerp_fma(left, right, fu);
	}
	if (modulate)
	{
		// SponzaScene.cpp:99-101: r = radiance[0] * r ...
		out[0] = mulf(modulate[0], out[0]);
		out[1] = mulf(modulate[1], out[1]);
		out[2] = mulf(modulate[2], out[2]);
		return pack_rgba(out[0], out[1], out[2], out[3]);
	}
	return pack_rgba_unit(out[0], out[1], out[2], out[3]);
}

// RCPPS replay from the PACKED shared-memory table (kFastRcpBits index bits, entries (T[i] - 0x3F000000) >> 11: the usual
// table has 12 significant mantissa bits, srb_api.cu checks).  The two look-ups per pixel index the table with the top
// mantissa bits of 1/w at the neighbouring pixels, i.e. with 32 scattered indices per warp: from global memory that is
// up to a dozen cache lines per load, a fifth of the kernel's L1 wavefronts; from shared memory a bank conflict or two.
__device__ __forceinline__ float rcp_x86_packed(float x, uint32_t tab, const uint32_t* __restrict__ table,
                                                uint32_t bits)
{
	uint32_t const u = __float_as_uint(x);
	uint32_t const eb = u & 0x7F800000u;
	if (eb - 0x00800000u < 0x7E000000u)
	{
		uint32_t const entry = (lds_u16<(uint32_t)offsetof(ShadeTables, rcp16)>(tab + ((u >> 11) & 0xFFEu)) << 11) + 0x3F000000u;
		return __uint_as_float((u & 0x80000000u) | (entry + 0x3F800000u - eb));
	}
	return rcp_x86_special(x, table, bits);
}

constexpr uint32_t kSmemTexs = 48; // texture descriptors kept in shared memory by the shade kernel (the rest: global)
static_assert(sizeof(ShadeTables::texs) / sizeof(TexDev) == kSmemTexs && sizeof(ShadeTables::rcp16) == 4096, "ShadeTables");

struct ShadeEnv
{
	const ShadeRec* srecs;
	const TexDev* texs;
	const uint32_t* rcpTable;
	uint32_t rcpBits;
	const uint32_t* rsqrtTable;
	uint32_t rsqrtBits;
	const SponzaDev* sponza; // shared-memory copy of the frame's constants (valid when a draw uses SRB_SHADER_SPONZA)
	bool rcp16;                // g_sTab.rcp16 holds the 11-bit RCPPS table packed to 16 bits per entry (false: not packable)
	const uint32_t* rcpSmem;   // shared-memory copies of the tables for the lit shader (nullptr: unusual table widths)
	const uint32_t* rsqrtSmem;
	uint32_t tab;              // tables_base(): shared-space address of g_sTab
};

// SimdUtil Dot3SoA (SIMDUtil.h:123-126): fmadd(x0, x1, fmadd(y0, y1, z0 * z1))
__device__ __forceinline__ float dot3_soa(float x0, float y0, float z0, float x1, float y1, float z1)
{
	return fma_(x0, x1, fma_(y0, y1, mulf(z0, z1)));
}

// RCPPS / RSQRTPS replay from SHARED-MEMORY copies of the tables, for the common table widths (11 / 10 index bits):
// the lit shader does 48 look-ups per pixel, which is worth 16 KB of shared memory per CTA.  Same results as
// rcp_x86 / rsqrt_x86 (srb_device.cuh), which handle the special inputs and every other table width.
constexpr uint32_t kFastRcpBits = 11, kFastRsqrtBits = 10;

struct SponzaTables
{
	const uint32_t* rcpSmem;   // 1 << kFastRcpBits entries, or nullptr: use the global tables
	const uint32_t* rsqrtSmem; // 2 << kFastRsqrtBits entries
	const uint32_t* rcpTable;
	uint32_t rcpBits;
	const uint32_t* rsqrtTable;
	uint32_t rsqrtBits;
};

__device__ __forceinline__ float rcp_fast(float x, const SponzaTables& t)
{
	uint32_t const u = __float_as_uint(x);
	uint32_t const eb = u & 0x7F800000u;
	if (eb - 0x00800000u < 0x7E000000u)
	{
		uint32_t const r = t.rcpSmem[(u >> (23u - kFastRcpBits)) & ((1u << kFastRcpBits) - 1u)] + 0x3F800000u - eb;
		return __uint_as_float((u & 0x80000000u) | r);
	}
	return rcp_x86_special(x, t.rcpTable, t.rcpBits);
}

__device__ __forceinline__ float rsqrt_fast(float x, const SponzaTables& t)
{
	uint32_t const u = __float_as_uint(x);
	if (u - 0x00800000u < 0x7F000000u) // positive normal
	{
		// (exponent parity, top mantissa bits) = bits 23 .. 23 - kFastRsqrtBits of x with bit 23 flipped (bias 127 is odd)
		uint32_t const idx = ((u >> (23u - kFastRsqrtBits)) & ((2u << kFastRsqrtBits) - 1u)) ^ (1u << kFastRsqrtBits);
		int32_t const ue = (int32_t)(u >> 23) - 127;
		int32_t const k = ue >> 1; // floor((ue - parity) / 2) == ue >> 1
		return __uint_as_float(t.rsqrtSmem[idx] - ((uint32_t)k << 23));
	}
	return rsqrt_x86(x, t.rsqrtTable, t.rsqrtBits);
}

// Lighting of the Sponza pixel shader, Viewer/SponzaScene.cpp:40-93: sun (with the 0.1 "magic bias"), 16 point lights
// (RSQRTPS / RCPPS replayed from the host's tables), ambient.  pn = interpolated position (0..2) and normal (3..5).
template <bool kFast>
__device__ __forceinline__ void sponza_radiance(const SponzaDev* __restrict__ k, const float (&pn)[6], const SponzaTables& t,
                                                float (&radiance)[3])
{
	auto rcp = [&](float x) { return kFast ? rcp_fast(x, t) : rcp_x86(x, t.rcpTable, t.rcpBits); };
	auto rsqrt = [&](float x) { return kFast ? rsqrt_fast(x, t) : rsqrt_x86(x, t.rsqrtTable, t.rsqrtBits); };
	float const px = pn[0], py = pn[1], pz = pn[2], nx = pn[3], ny = pn[4], nz = pn[5];
	// _mm256_max_ps(0.1, nDotL): maxps returns its SECOND operand unless the first is greater
	float const sun = max_x86(0.1f, dot3_soa(nx, ny, nz, k->sunDir[0], k->sunDir[1], k->sunDir[2]));
	float r0 = sun, r1 = sun, r2 = sun;
	// the three table look-ups of a light depend on each other; several lights in flight hide their latency (the sums
	// stay in the reference's order)
#pragma unroll 4
	for (int i = 0; i < 16; ++i)
	{
		const SponzaLightDev& L = k->lights[i];
		float const tx = subf(L.pos[0], px), ty = subf(L.pos[1], py), tz = subf(L.pos[2], pz);
		float const distSq = dot3_soa(tx, ty, tz, tx, ty, tz);
		float const recipDist = rsqrt(distSq);
		float const dist = rcp(recipDist);
		float const lx = mulf(tx, recipDist), ly = mulf(ty, recipDist), lz = mulf(tz, recipDist);
		float const nDotL = max_x86(0.0f, dot3_soa(lx, ly, lz, nx, ny, nz));
		float const atten = rcp(addf(1.0f, fma_(0.1f, dist, mulf(distSq, 0.01f))));
		float const lightRadiance = mulf(nDotL, mulf(L.intensity, atten));
		r0 = addf(r0, mulf(lightRadiance, L.colour[0]));
		r1 = addf(r1, mulf(lightRadiance, L.colour[1]));
		r2 = addf(r2, mulf(lightRadiance, L.colour[2]));
	}
	radiance[0] = addf(r0, k->ambient[0]);
	radiance[1] = addf(r1, k->ambient[1]);
	radiance[2] = addf(r2, k->ambient[2]);
}

// Planes (tile-relative) and values of varyings uvOffset, uvOffset + 1 when they are not 6, 7: out = dx dy c value, twice.
static __device__ __noinline__ void load_deriv_planes(const ShadeRec* __restrict__ rec, uint32_t uo, float sx, float sy,
                                                     float W, float fx, float fy, float (&out)[8])
{
#pragma unroll
	for (uint32_t k = 0; k < 2; ++k)
	{
		const float* q = rec->pl[SRB_PLANE_SLOT(uo + k)];
		float const dx = __ldg(q), dy = __ldg(q + 1);
		float const c = plane_c0(dx, dy, __ldg(q + 2), sx, sy);
		out[4 * k] = dx;
		out[4 * k + 1] = dy;
		out[4 * k + 2] = c;
		out[4 * k + 3] = mulf(W, fma_(dy, fy, fma_(dx, fx, c)));
	}
}

// Interpolants (Rasterizer.cpp:356-400) + pixel shader (Viewer/Shaders.h) for the visible fragment of pixel (x, y) of
// the tile whose origin is (fX0, fY0).  Only the planes the shader reads are fetched.  kTexSmem: every texture
// descriptor of the frame is in shared memory (env.smemTexs).
// kUniform: every draw of the frame is UnlitDiffuse with a non-empty texture and uvOffset 6 (the scene of Viewer/Scene.cpp,
// found by the host when the frame is submitted): no per-pixel dispatch on shader, texture or derivative source.
template <bool kTexSmem, bool kSponza, bool kUniform, bool kQuadTap>
__device__ __forceinline__ uint32_t shade_pixel(const ShadeEnv& env, uint32_t slot, float fX0, float fY0, float fx,
                                                float fy)
{
	const ShadeRec* __restrict__ rec = env.srecs + slot;
	const float4* rv = reinterpret_cast<const float4*>(rec);
	// the first 64 bytes of the record in one go: everything the textured shader needs
	float4 const head = __ldg(rv);      // wdx, wdy, w0, info
	float4 const v1 = __ldg(rv + 1);    // r0x, r0y, (draw / redirect)
	float4 const v2 = __ldg(rv + 2);    // varyings 6, 7: (dx6 dy6 c6 dx7)
	float4 const v3 = __ldg(rv + 3);    //                (dy7 c7) + varying 0
	float const wdx = head.x, wdy = head.y;
	uint32_t const info = __float_as_uint(head.w); // shader | uvOffset << 8 | (texture + 1) << 16
	float const sx = subf(fX0, v1.x), sy = subf(fY0, v1.y);
	float const wc0 = plane_c0(wdx, wdy, head.z, sx, sy);
	float const W = __frcp_rn(fma_(fx, wdx, fma_(fy, wdy, wc0))); // 1.0f / x, correctly rounded

	struct Plane
	{
		float dx, dy, c;
	};
	auto plane = [&](uint32_t j) -> Plane {
		const float* q = rec->pl[SRB_PLANE_SLOT(j)];
		Plane p;
		p.dx = __ldg(q);
		p.dy = __ldg(q + 1);
		p.c = plane_c0(p.dx, p.dy, __ldg(q + 2), sx, sy);
		return p;
	};
	auto eval = [&](const Plane& p) -> float { return mulf(W, fma_(p.dy, fy, fma_(p.dx, fx, p.c))); };

	uint32_t const shader = kUniform ? (uint32_t)SRB_SHADER_UNLIT_DIFFUSE : (info & 0xFFu);
	if (!kUniform && shader == SRB_SHADER_VISUALIZE_NORMALS)
	{
		float const r = fma_(eval(plane(3)), 0.5f, 0.5f), g = fma_(eval(plane(4)), 0.5f, 0.5f),
		            b = fma_(eval(plane(5)), 0.5f, 0.5f);
		return pack_rgba(r, g, b, 1.0f);
	}
	if (!kUniform && shader == SRB_SHADER_VISUALIZE_UVS)
	{
		return pack_rgba(eval(plane(6)), eval(plane(7)), 0.0f, 0.0f);
	}
	// UnlitDiffuseShader (Shaders.h:71-104) and SponzaShader (SponzaScene.cpp:13-104): null / empty texture = white
	if (!kUniform && (info >> 16) == 0u)
	{
		return 0xFFFFFFFFu;
	}
	uint32_t const ti = (info >> 16) - 1u;
	TexHead const tex = kTexSmem ? load_tex_head_shared(env.tab, ti) : load_tex_head(env.texs + ti);
	if (!kUniform && tex.bytes == 0u)
	{
		return 0xFFFFFFFFu;
	}
	Plane pu, pv;
	pu.dx = v2.x;
	pu.dy = v2.y;
	pu.c = plane_c0(v2.x, v2.y, v2.z, sx, sy);
	pv.dx = v2.w;
	pv.dy = v3.x;
	pv.c = plane_c0(v2.w, v3.x, v3.y, sx, sy);
	float const u = eval(pu), v = eval(pv);
	float deriv[4] = {0.0f, 0.0f, 0.0f, 0.0f}; // dudx, dudy, dvdx, dvdy
	uint32_t const uo = kUniform ? 6u : ((info >> 8) & 0xFFu);
	if (uo + 1u < SRB_MAX_VARY)
	{
		float const fx1 = addf(1.0f, fx), fy1 = addf(1.0f, fy);
		float const a10 = fma_(wdx, fx1, fma_(wdy, fy, wc0)), a01 = fma_(wdx, fx, fma_(wdy, fy1, wc0));
		float const W10 = env.rcp16 ? rcp_x86_packed(a10, env.tab, env.rcpTable, env.rcpBits) : rcp_x86(a10, env.rcpTable, env.rcpBits);
		float const W01 = env.rcp16 ? rcp_x86_packed(a01, env.tab, env.rcpTable, env.rcpBits) : rcp_x86(a01, env.rcpTable, env.rcpBits);
		// derivatives come from varyings uvOffset, uvOffset + 1 (Rasterizer.cpp:378-399); the usual case is 6, 7, the
		// planes already in registers
		Plane p0 = pu, p1 = pv;
		float s0 = u, s1 = v;
		if (uo != 6u)
		{
			float t[8];
			load_deriv_planes(rec, uo, sx, sy, W, fx, fy, t);
			p0.dx = t[0], p0.dy = t[1], p0.c = t[2], s0 = t[3];
			p1.dx = t[4], p1.dy = t[5], p1.c = t[6], s1 = t[7];
		}
		deriv[0] = subf(mulf(W10, fma_(p0.dx, fx1, fma_(p0.dy, fy, p0.c))), s0);
		deriv[1] = subf(mulf(W01, fma_(p0.dx, fx, fma_(p0.dy, fy1, p0.c))), s0);
		deriv[2] = subf(mulf(W10, fma_(p1.dx, fx1, fma_(p1.dy, fy, p1.c))), s1);
		deriv[3] = subf(mulf(W01, fma_(p1.dx, fx, fma_(p1.dy, fy1, p1.c))), s1);
	}
	auto mipOffset = [&](uint32_t mip) -> uint32_t {
		return kTexSmem ? lds_u32<(uint32_t)offsetof(TexDev, mipOffsets)>(env.tab + ti * (uint32_t)sizeof(TexDev) + mip * 4u) : __ldg(&env.texs[ti].mipOffsets[mip]);
	};
	if (kSponza && shader == SRB_SHADER_SPONZA)
	{
		float pn[6], radiance[3];
#pragma unroll
		for (uint32_t j = 0; j < 6; ++j)
		{
			pn[j] = eval(plane(j));
		}
		SponzaTables const t = {env.rcpSmem, env.rsqrtSmem, env.rcpTable, env.rcpBits, env.rsqrtTable, env.rsqrtBits};
		if (env.rcpSmem)
		{
			sponza_radiance<true>(env.sponza, pn, t, radiance);
		}
		else
		{
			sponza_radiance<false>(env.sponza, pn, t, radiance);
		}
		return sample_wrap<kQuadTap>(tex, mipOffset, env.tab + (uint32_t)offsetof(ShadeTables, spread), u, v, deriv[0], deriv[1], deriv[2], deriv[3], radiance);
	}
	return sample_wrap<kQuadTap>(tex, mipOffset, env.tab + (uint32_t)offsetof(ShadeTables, spread), u, v, deriv[0], deriv[1], deriv[2], deriv[3]);
}

// ---------------------------------------------------------------------------------------------------------------
// the tile kernel
// ---------------------------------------------------------------------------------------------------------------
// Work decomposition (not the reference's): the unit of work is (a slice of one tile's reference list) x (one 16x16
// QUAD of the tile), and it belongs to ONE WARP from start to finish — warps never wait for each other (no CTA
// barriers; a CTA is only a container of four independent warps).  The warp's 4 lane-groups of 8 lanes each OWN one
// 8x8 block of the quad: lane = one column of the block, its 8 rows' (depth, winner) keys live in 16 registers.
// Nothing but the owner ever touches a pixel, so the depth resolve needs no atomics and no shared-memory key buffer.
//   scan  : 32 list entries per step, one per lane.  The entry carries the triangle's block range inside the tile
//           (TileRef::blocks, written by the bin fill), so entries that cannot touch the quad cost one compare.
//           Entries that can are queued (ballot + popc) in the warp's private shared memory.
//   stage : 32 queued triangles at a time (full warps): fetch the 64-byte record, derive the tile-relative edge
//           constants and z plane (what the reference keeps per BinChunk entry, Binning.cpp:412-454) and append them to
//           the warp's staging area.
//   lists : when the staging area cannot take another step, every lane-group tests the staged triangles against ITS
//           block, 8 triangles at a time (lane = triangle): block loops of Rasterizer.cpp:201-223 and the reference's
//           coarse test (:224-261, incl. its 64x64 extent and the depth-only path); hits are compacted into the
//           group's candidate list with a ballot.
//   raster: every lane-group walks its own list: 3 x LDS.128 fetch the triangle, 8 rows are evaluated in the
//           reference's order (z bit-exact), and (depth, winner) keys are compared as one 64-bit integer and selected
//           in registers.
constexpr int kWarpsPerCta = kRasterThreads / 32;
constexpr int kStageCap = 64; // staged triangles per warp; list entries are 6 bits of index + mode
static_assert(kStageCap == 64, "list entries hold a 6-bit staged index; a scan step adds at most 32");

struct WarpStage
{
	uint4 q0[kStageCap];            // c0 c1 c2 dx0
	uint4 q1[kStageCap];            // dx1 dx2 dy0 dy1
	uint4 q2[kStageCap];            // dy2 zc0 zdx zdy
	uint32_t keyLow[kStageCap];
	uint16_t blocks[kStageCap];     // TileRef::blocks
	uint4 pend[64];                 // scanned entries that can touch the quad, waiting to be staged 32 at a time
	uint8_t list[4][kStageCap];     // per lane-group: staged index | (row mode - 1) << 6: fast, depth-only, general
};

// K3: warps pull work (a unit x one of the tile's 16 quads) from a device-side dispenser and publish the winning keys
// to the per-tile key buffer in HBM/L2 (all-zero between frames: the shade kernel clears what it reads): plain stores
// when the tile is one unit, RED.MAX.64 when the tile's list is split over several units.  Only pixels that received a
// fragment this frame are written.
// The ticket counter is advanced with atom.inc: around an atom.add under `if (lane == 0)` the
// assembler builds its warp-aggregated form (vote, popc, atomic by an elected lane, shuffle), and that shuffle waits for the
// result on the spot — the point of requesting a ticket early is not to wait for it.  (The bound must not be all ones: the
// assembler rewrites that inc as the add.  Counts never get near 2^31: a frame has at most 2^28 triangles.)
__device__ __forceinline__ uint32_t next_ticket(uint32_t* counter)
{
	uint32_t old;
	asm volatile("atom.global.inc.u32 %0, [%1], 0x7fffffff;" : "=r"(old) : "l"(counter) : "memory");
	return old;
}

__global__ void __launch_bounds__(kRasterThreads, 8) raster_kernel(RasterArgs A)
{
	__shared__ WarpStage S_all[kWarpsPerCta];
	uint32_t const lane = threadIdx.x & 31u, grp = lane >> 3;
	WarpStage& S = S_all[threadIdx.x >> 5];
	int32_t const l = (int32_t)(lane & 7u);
	uint32_t const ltMask = (1u << lane) - 1u;
	if (A.ctl->overflow != 0u)
	{
		return; // the host grows the buffers and re-runs the frame
	}
	uint32_t const numJobs = A.ctl->numUnits * 16u;
	uint32_t ticket = 0;
	if (lane == 0)
	{
		ticket = next_ticket(&A.ctl->unitTicket);
	}
	ticket = __shfl_sync(0xFFFFFFFFu, ticket, 0);
	while (ticket < numJobs)
	{
		// the next ticket is requested now and consumed when this job is done: its latency hides behind the job
		uint32_t nextTicket = 0;
		if (lane == 0)
		{
			nextTicket = next_ticket(&A.ctl->unitTicket);
		}
		UnitDesc const d = A.units[ticket >> 4];
		uint32_t const tile = d.tile;
		uint32_t const tileY = A.fp.tilesX == 1u ? tile : __umulhi(tile, A.tilesXMagic);
		int32_t const X0 = (int32_t)(tile - tileY * A.fp.tilesX) * SRB_TILE;
		int32_t const Y0 = (int32_t)tileY * SRB_TILE;
		uint32_t const quad = ticket & 15u;
		uint32_t const qbx = (quad & 3u) * 2u, qby = (quad >> 2) * 2u; // quad origin in blocks
		uint32_t const gbx = qbx + (grp & 1u), gby = qby + (grp >> 1);             // this lane-group's block
		int32_t const xB = (int32_t)(gbx * 8u), yB = (int32_t)(gby * 8u);
		int32_t const xl = xB + l;
		float const fl = (float)l, fxB = (float)xB, fyB = (float)yB;

		// ---- keys of my column: depth bits << 32 | (0xFFFFFFFE - canonical key); low word 0xFFFFFFFF = no fragment yet
		long long key[8];
		if (A.clearDepth)
		{
#pragma unroll
			for (int row = 0; row < 8; ++row) key[row] = (long long)kNoWinner; // depth 0.0f = Config::c_depthMax (reverse Z), Renderer.cpp:168-194
		}
		else
		{
			const uint32_t* depthTile = reinterpret_cast<const uint32_t*>(A.depthTiles + (size_t)tile * 16384u);
#pragma unroll
			for (int row = 0; row < 8; ++row)
			{
				key[row] = (long long)(((unsigned long long)depthTile[(yB + row) * SRB_TILE + xl] << 32) | kNoWinner);
			}
		}

		uint32_t nStaged = 0, nPend = 0;
		uint4 ref = make_uint4(0u, 0u, 0u, 0u);
		if (d.begin + lane < d.end)
		{
			ref = __ldg(reinterpret_cast<const uint4*>(A.refs + d.begin + lane));
		}
		for (uint32_t base = d.begin; base < d.end; base += 32u)
		{
			// ---- scan: entries that can touch my quad go to the pending queue ---------------------------------------
			uint4 const cur = ref;
			bool const valid = base + lane < d.end;
			if (base + 32u + lane < d.end)
			{
				ref = __ldg(reinterpret_cast<const uint4*>(A.refs + base + 32u + lane)); // next step's entry: in flight during this one
			}
			bool const touches = valid && ((cur.w >> quad) & 1u) != 0u; // TileRef::quads
			uint32_t const tm = __ballot_sync(0xFFFFFFFFu, touches);
			if (touches)
			{
				S.pend[nPend + (uint32_t)__popc(tm & ltMask)] = cur;
			}
			nPend += (uint32_t)__popc(tm);
			bool const last = base + 32u >= d.end;
			// ---- stage: 32 pending triangles at a time, so that the record fetch and the set-up run with full warps --
			while (nPend >= 32u || (last && (nPend | nStaged) != 0u)) // at the end of the list: drain both stages
			{
				uint32_t const take = min(32u, nPend);
				__syncwarp();
				if (lane < take)
				{
					uint4 const e = S.pend[nPend - take + lane];
					uint32_t const idx = nStaged + lane;
					RasterRec r;
					load_raster_rec(A.rrecs, e.y, r);
					// edge constants at the tile origin, Binning.cpp:421-427
					int32_t const c0 = wrap_add(r.c[0], wrap_add(wrap_mul(r.dx[0], Y0), wrap_mul(r.dy[0], X0)));
					int32_t const c1 = wrap_add(r.c[1], wrap_add(wrap_mul(r.dx[1], Y0), wrap_mul(r.dy[1], X0)));
					int32_t const c2 = wrap_add(r.c[2], wrap_add(wrap_mul(r.dx[2], Y0), wrap_mul(r.dy[2], X0)));
					float const zc0 = plane_c0(r.zdx, r.zdy, r.z0, subf((float)X0, r.r0x), subf((float)Y0, r.r0y));
					S.q0[idx] = make_uint4((uint32_t)c0, (uint32_t)c1, (uint32_t)c2, (uint32_t)r.dx[0]);
					S.q1[idx] = make_uint4((uint32_t)r.dx[1], (uint32_t)r.dx[2], (uint32_t)r.dy[0], (uint32_t)r.dy[1]);
					S.q2[idx] = make_uint4((uint32_t)r.dy[2], __float_as_uint(zc0), __float_as_uint(r.zdx), __float_as_uint(r.zdy));
					S.keyLow[idx] = 0xFFFFFFFEu - e.x;
					S.blocks[idx] = (uint16_t)e.z;
				}
				nPend -= take;
				nStaged += take;
				if (nStaged <= (uint32_t)kStageCap - 32u && !(last && nPend == 0u))
				{
					continue; // room for another 32: keep filling so that the lists get long
				}
				__syncwarp();

				// ---- candidate list of my block: lane = staged triangle, 8 per step ----------------------------------
				uint32_t n = 0;
				for (uint32_t j0 = 0; j0 < nStaged; j0 += 8u)
				{
					uint32_t const j = j0 + (uint32_t)l;
					int mode = 0;
					if (j < nStaged)
					{
						uint32_t const bl = S.blocks[j];
						if (gbx >= (bl & 15u) && gbx < ((bl >> 4) & 15u) && gby >= ((bl >> 8) & 15u) && gby < ((bl >> 12) & 15u))
						{
							uint4 const q0 = S.q0[j], q1 = S.q1[j], q2 = S.q2[j];
							TriTile tt;
							tt.c[0] = (int32_t)q0.x, tt.c[1] = (int32_t)q0.y, tt.c[2] = (int32_t)q0.z;
							tt.dx[0] = (int32_t)q0.w, tt.dx[1] = (int32_t)q1.x, tt.dx[2] = (int32_t)q1.y;
							tt.dy[0] = (int32_t)q1.z, tt.dy[1] = (int32_t)q1.w, tt.dy[2] = (int32_t)q2.x;
							int32_t e00[3];
							mode = ref_coarse(tt, xB, yB, e00);
							// 1 = fast rows; 2 = the reference's depth-only path; 3 = general rows (a z plane that can reach
							// inf/NaN inside the tile: the fast rows order depth by its bit pattern, which needs finite values)
							float const zlim = 1.0e30f;
							bool const tame = fabsf(__uint_as_float(q2.y)) < zlim && fabsf(__uint_as_float(q2.z)) < zlim &&
							                  fabsf(__uint_as_float(q2.w)) < zlim;
							mode = (mode == 1 && !tame) ? 3 : mode;
							SRB_STAT_LANE(0, 1);
							SRB_STAT_LANE(1, mode != 0 ? 1 : 0);
							if (mode == 1 && A.blockReject)
							{
								// The reference's coarse test looks at a 64x64 extent (quirk C-2), so it lets through every block of
								// the bounding box that the triangle does not touch at all.  The rows of such a block would all fail
								// the edge test: drop it here.  An edge function is linear, so its largest value over the block's
								// 64 pixels is at a corner: e(xB, yB) + 7 * max(dy, 0) + 7 * max(dx, 0); all pixels fail the test
								// e >= 0 when that is negative.  Only for operands small enough that nothing wraps modulo 2^32
								// (the rows themselves compute modulo 2^32 like the reference).
								bool none = false, small = true;
#pragma unroll
								for (int k = 0; k < 3; ++k)
								{
									int32_t const ex = tt.dx[k], ey = tt.dy[k];
									small = small && (uint32_t)e00[k] + (1u << 29) < (1u << 30) && (uint32_t)ex + (1u << 25) < (1u << 26) &&
									        (uint32_t)ey + (1u << 25) < (1u << 26);
									none = none || wrap_add(e00[k], wrap_mul(7, wrap_add(max(ex, 0), max(ey, 0)))) < 0;
								}
								if (small && none)
								{
									mode = 0;
									SRB_STAT_LANE(2, 1);
								}
							}
						}
					}
					uint32_t const hits = (__ballot_sync(0xFFFFFFFFu, mode != 0) >> (grp * 8u)) & 0xFFu;
					if (mode != 0)
					{
						S.list[grp][n + (uint32_t)__popc(hits & ((1u << l) - 1u))] = (uint8_t)(j | ((uint32_t)(mode - 1) << 6));
					}
					n += (uint32_t)__popc(hits);
				}
				__syncwarp();

				// ---- raster: my block's candidates ------------------------------------------------------------------
#ifdef SRB_STATS
				{
					// warp-level walks of the candidate lists, and their length = the longest of the four lists
					uint32_t const longest = max(max(__shfl_sync(0xFFFFFFFFu, n, 0), __shfl_sync(0xFFFFFFFFu, n, 8)),
					                             max(__shfl_sync(0xFFFFFFFFu, n, 16), __shfl_sync(0xFFFFFFFFu, n, 24)));
					if (lane == 0)
					{
						SRB_STAT_LANE(7, 1);
						SRB_STAT_LANE(8, longest);
					}
				}
#endif
				for (uint32_t i = 0; i < n; ++i)
				{
					uint32_t const entry = S.list[grp][i];
					uint32_t const idx = entry & 0x3Fu;
					uint4 const q0 = S.q0[idx], q1 = S.q1[idx], q2 = S.q2[idx];
					uint32_t const kl = S.keyLow[idx];
					int32_t const dx0 = (int32_t)q0.w, dx1 = (int32_t)q1.x, dx2 = (int32_t)q1.y;
					// edge k at (xB + l, yB): all arithmetic modulo 2^32 like the reference's int32 lanes
					int32_t e0 = wrap_add(wrap_add((int32_t)q0.x, wrap_mul((int32_t)q1.z, xl)), wrap_mul(dx0, yB));
					int32_t e1 = wrap_add(wrap_add((int32_t)q0.y, wrap_mul((int32_t)q1.w, xl)), wrap_mul(dx1, yB));
					int32_t e2 = wrap_add(wrap_add((int32_t)q0.z, wrap_mul((int32_t)q2.x, xl)), wrap_mul(dx2, yB));
					float const zdx = __uint_as_float(q2.z), zdy = __uint_as_float(q2.w);
					// z/w of my column, row 0: Rasterizer.cpp:213 (tileTopLeft = fma(ramp, dx, c0)) and :156-157
					float z = addf(fma_(fyB, zdy, fma_(fl, zdx, __uint_as_float(q2.y))), mulf(fxB, zdx));
					// pass <=> inside && z > 0 && z > stored (ordered, strict; Rasterizer.cpp:88-95); equal z: the first in
					// canonical order wins (largest low word).  Keys are compared as SIGNED 64-bit integers: stored depths are
					// >= +0.0f, so their bit patterns are non-negative and ordered like the floats.
#ifdef SRB_STATS
					{
						// statistics build: visits, and visits none of whose 64 pixels is inside the three edges
						int32_t a0 = e0, a1 = e1, a2 = e2;
						uint32_t allOut = 0x80000000u, anyOut = 0u;
						for (int row = 0; row < 8; ++row)
						{
							allOut &= (uint32_t)(a0 | a1 | a2);
							anyOut |= (uint32_t)(a0 | a1 | a2);
							a0 = wrap_add(a0, dx0), a1 = wrap_add(a1, dx1), a2 = wrap_add(a2, dx2);
						}
						uint32_t const am = __activemask();
						uint32_t const outMask = (__ballot_sync(am, (allOut >> 31) != 0u) >> (grp * 8u)) & 0xFFu;
						uint32_t const someOutMask = (__ballot_sync(am, (anyOut >> 31) != 0u) >> (grp * 8u)) & 0xFFu;
						// hierarchical-Z potential: the smallest depth stored in my block right now, and the largest z the candidate
						// can have on it (plane at the corner that maximises it, plus a rounding allowance)
						int32_t lo = (int32_t)(key[0] >> 32);
						for (int row = 1; row < 8; ++row) lo = min(lo, (int32_t)(key[row] >> 32));
						for (int o = 1; o < 8; o <<= 1) lo = min(lo, __shfl_xor_sync(am, lo, o));
						float const zc = z + (zdx > 0.0f ? (7.0f - fl) * zdx : -fl * zdx) + (zdy > 0.0f ? 7.0f * zdy : 0.0f);
						float const zmax = zc + fabsf(zc) * 1e-5f;
						if (l == 0)
						{
							SRB_STAT_LANE(10, lo > 0 ? 1 : 0);
							SRB_STAT_LANE(11, (lo > 0 && (entry & 0xC0u) == 0u && zmax < __int_as_float(lo)) ? 1 : 0);
						}
						if (l == 0)
						{
							SRB_STAT_LANE(3, 1);
							SRB_STAT_LANE(4, (outMask == 0xFFu && (entry & 0xC0u) != 0x40u) ? 1 : 0);
							SRB_STAT_LANE(9, someOutMask == 0u ? 1 : 0); // all 64 pixels inside the three edges
						}
					}
#endif
					if (__builtin_expect((entry & 0xC0u) == 0u, 1))
					{
						// fast rows (finite z): a set sign bit — outside an edge, or z < 0 / -0.0 — makes the candidate negative
						// so that it loses; +0.0 loses against a cleared pixel through the low word (kNoWinner is the maximum)
	#pragma unroll
						for (int row = 0; row < 8; ++row)
						{
							uint32_t const hi = ((uint32_t)(e0 | e1 | e2) & 0x80000000u) | __float_as_uint(z);
							long long const cand = (long long)(((unsigned long long)hi << 32) | kl);
							key[row] = cand > key[row] ? cand : key[row];
							e0 = wrap_add(e0, dx0);
							e1 = wrap_add(e1, dx1);
							e2 = wrap_add(e2, dx2);
							z = addf(z, zdy);
						}
					}
					else
					{
						bool const depthOnly = (entry & 0xC0u) == 0x40u; // the reference's all-corners-inside path skips the edge test
	#pragma unroll
						for (int row = 0; row < 8; ++row)
						{
							bool const inside = depthOnly || ((e0 | e1 | e2) >= 0);
							uint32_t const zb = __float_as_uint(inside ? fmaxf(z, 0.0f) : 0.0f); // NaN -> 0: fails like the ordered compare
							long long const cand = (long long)(((unsigned long long)zb << 32) | kl);
							key[row] = cand > key[row] ? cand : key[row];
							e0 = wrap_add(e0, dx0);
							e1 = wrap_add(e1, dx1);
							e2 = wrap_add(e2, dx2);
							z = addf(z, zdy);
						}
					}
				}
				__syncwarp(); // the staging area is reused
				nStaged = 0;
			}
			__syncwarp(); // the pending queue is written again by the next scan step
		}

		// ---- publish the pixels that received a fragment ---------------------------------------------------------
		unsigned long long* gk = A.tileKeys + (size_t)tile * SRB_TILE_PIXELS + (uint32_t)(yB * SRB_TILE + xl);
		bool const split = d.unitsInTile > 1u;
#pragma unroll
		for (int row = 0; row < 8; ++row)
		{
			if ((uint32_t)key[row] != kNoWinner)
			{
				if (split)
				{
					atomicMax(gk + row * SRB_TILE, (unsigned long long)key[row]);
				}
				else
				{
					gk[row * SRB_TILE] = (unsigned long long)key[row];
				}
			}
		}
		ticket = __shfl_sync(0xFFFFFFFFu, nextTicket, 0);
	}
}

// Canonical key of the visible fragment -> slot of its records (srb_device.cuh).
__device__ __forceinline__ uint32_t slot_of_key(const ShadeRec* __restrict__ srecs, uint32_t key)
{
	uint32_t const g = key >> 3, f = key & 7u;
	if (f == 0u)
	{
		return g;
	}
	uint2 const redirect = __ldg(reinterpret_cast<const uint2*>(&srecs[g].pad[0]));
	return redirect.x + __popc(redirect.y & ((1u << (f - 1u)) - 1u));
}

// K4: one thread per pixel.  A CTA walks chunks of 128 consecutive pixels (two rows) of one tile, so everything that
// depends on the tile is CTA-uniform.  Reads the resolved key, shades the visible fragment (or applies the pending
// clear), writes colour + depth in the reference's ColourTile/DepthTile layout, and zeroes the key.
constexpr int kShadeThreads = 128;

// Spin until *flag (another GPU writes it over NVLink, or this GPU's own earlier kernel) has reached `value`; stamps only
// grow, and the comparison is wrap-safe.  Gives up after ten seconds (a rank that died must not hang the others' GPUs) and
// reports it through the control block (bit 3 of `overflow`).
__device__ __forceinline__ void split_wait(const uint32_t* flag, uint32_t value, FrameCtl* ctl)
{
	unsigned long long t0 = 0;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
	while ((int32_t)(*reinterpret_cast<const volatile uint32_t*>(flag) - value) < 0)
	{
		__nanosleep(200);
		unsigned long long t1;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
		if (t1 - t0 > 10000000000ull)
		{
			atomicOr(&ctl->overflow, 8u);
			return;
		}
	}
}

// kSponza: some draw of the frame uses SRB_SHADER_SPONZA (its lighting loop costs registers the other shaders do not need)
// kFast: the common frame — one GPU (no screen-tile split), no debug output, uniform UnlitDiffuse draws (see shade_pixel)
template <bool kTexSmem, bool kSponza, bool kFast, bool kQuadTap = false>
__global__ void __launch_bounds__(kShadeThreads) shade_kernel(RasterArgs A)
{
	__shared__ __align__(16) SponzaDev s_sponza;
	__shared__ uint32_t s_rcp[kSponza ? (1u << kFastRcpBits) : 1u];
	__shared__ uint32_t s_rsqrt[kSponza ? (2u << kFastRsqrtBits) : 1u];
	bool const fastTables = kSponza && A.rcpBits == kFastRcpBits && A.rsqrtBits == kFastRsqrtBits;
	fill_spread_table(g_sTab.spread);
	if (A.rcp16)
	{
		const uint4* src = reinterpret_cast<const uint4*>(A.rcp16);
		uint4* dst = reinterpret_cast<uint4*>(g_sTab.rcp16);
		for (uint32_t i = threadIdx.x; i < (1u << kFastRcpBits) / 8u; i += kShadeThreads) dst[i] = __ldg(src + i);
	}
	if (fastTables)
	{
		for (uint32_t i = threadIdx.x; i < (1u << kFastRcpBits); i += kShadeThreads) s_rcp[i] = __ldg(A.rcpTable + i);
		for (uint32_t i = threadIdx.x; i < (2u << kFastRsqrtBits); i += kShadeThreads) s_rsqrt[i] = __ldg(A.rsqrtTable + i);
	}
	if (kSponza)
	{
		const uint32_t* src = reinterpret_cast<const uint32_t*>(A.sponza);
		uint32_t* dst = reinterpret_cast<uint32_t*>(&s_sponza);
		for (uint32_t i = threadIdx.x; i < (uint32_t)(sizeof(SponzaDev) / 4u); i += kShadeThreads) dst[i] = __ldg(src + i);
	}
	if (kTexSmem)
	{
		// texture descriptors into shared memory: two dependent global loads per pixel become LDS
		uint32_t const words = min(A.numTexs, kSmemTexs) * (uint32_t)(sizeof(TexDev) / 4u);
		const uint32_t* src = reinterpret_cast<const uint32_t*>(A.texs);
		uint32_t* dst = reinterpret_cast<uint32_t*>(g_sTab.texs);
		for (uint32_t i = threadIdx.x; i < words; i += kShadeThreads) dst[i] = __ldg(src + i);
	}
	__syncthreads();
	if (A.ctl->overflow != 0u)
	{
		// the frame is going to be re-run with larger buffers: the host only needs the counters the binner left
		if (blockIdx.x == 0 && threadIdx.x < (uint32_t)(sizeof(FrameCtl) / 4u))
		{
			A.hostCtl[threadIdx.x] = __ldcg(reinterpret_cast<const uint32_t*>(A.ctl) + threadIdx.x);
		}
		return;
	}
	if (!kFast && A.splitFlags && !A.splitIsRoot)
	{
		// Screen-tile split: this frame's tiles go into the root GPU's framebuffer, which still holds the previous frame
		// until the root says it has been consumed (it stamps the release flag when it begins its next frame).
		if (threadIdx.x == 0)
		{
			split_wait(A.splitFlags + 32, A.ctl->doneValue - 1u, A.ctl);
		}
		__syncthreads();
	}
	ShadeEnv env;
	env.srecs = A.srecs;
	env.texs = A.texs;
	env.rcpTable = A.rcpTable;
	env.rcpBits = A.rcpBits;
	env.rsqrtTable = A.rsqrtTable;
	env.rsqrtBits = A.rsqrtBits;
	env.sponza = &s_sponza;
	env.rcp16 = A.rcp16 != nullptr;
	env.rcpSmem = fastTables ? s_rcp : nullptr;
	env.rsqrtSmem = fastTables ? s_rsqrt : nullptr;
	env.tab = tables_base(); // (after the barrier that publishes the tables)
	// this context's tiles: all of them, or every ownMod-th one in a screen-tile split across GPUs
	uint32_t const numTiles = A.fp.tilesX * A.fp.tilesY;
	uint32_t const mod = kFast ? 1u : max(1u, A.fp.ownMod), rem = (!kFast && A.fp.ownMod > 1u) ? A.fp.ownRem : 0u;
	uint32_t const ownedTiles = numTiles > rem ? (numTiles - rem + mod - 1u) / mod : 0u;
	uint32_t const numChunks = ownedTiles * (SRB_TILE_PIXELS / kShadeThreads);
	uint32_t covered = 0;
	auto key_index = [&](uint32_t chunk) -> uint32_t {
		return ((rem + (chunk >> 5) * mod) << 12) | ((chunk & 31u) << 7) | threadIdx.x;
	};
	// (A second key in flight and an L2 prefetch of the NEXT pixel's record were measured: the shade kernel 38.9 -> 40.7 us
	// on the hall, 119.7 -> 123.5 at 4K, no better with frames in flight: the moves and the prefetch cost more issue
	// slots than the wait they hide.)
	uint32_t nextGp = key_index(blockIdx.x);
	unsigned long long nextKey = blockIdx.x < numChunks ? __ldcg(A.tileKeys + nextGp) : 0ull;
	for (uint32_t chunk = blockIdx.x; chunk < numChunks; chunk += gridDim.x)
	{
		uint32_t const gp = nextGp;
		uint32_t const tile = gp >> 12, p = gp & 4095u;
		unsigned long long const key = nextKey;
		if (chunk + gridDim.x < numChunks)
		{
			nextGp = key_index(chunk + gridDim.x);
			nextKey = __ldcg(A.tileKeys + nextGp); // in flight while this pixel is shaded
		}
		bool const winner = key != 0ull; // the raster kernel publishes winners only
		uint32_t const low = (uint32_t)key;
		if (winner)
		{
			A.tileKeys[gp] = 0ull; // all-zero again for the next frame
		}
		float* depthTile = reinterpret_cast<float*>(A.depthTiles + (size_t)tile * 16384u);
		uint32_t* colourTile = reinterpret_cast<uint32_t*>(A.colourTiles + (size_t)tile * 16384u);
		if (!kFast && A.winnersOut)
		{
			A.winnersOut[gp] = winner ? 0xFFFFFFFEu - low : 0xFFFFFFFFu;
		}
		if (winner)
		{
			uint32_t const ty = A.fp.tilesX == 1u ? tile : __umulhi(tile, A.tilesXMagic);
			uint32_t const tx = tile - ty * A.fp.tilesX;
			uint32_t const slot = slot_of_key(A.srecs, 0xFFFFFFFEu - low);
			colourTile[p] = shade_pixel<kTexSmem, kSponza, kFast, kQuadTap>(env, slot, (float)(tx * SRB_TILE), (float)(ty * SRB_TILE), (float)(p & 63u),
			                            (float)(p >> 6));
			depthTile[p] = __uint_as_float((uint32_t)(key >> 32));
			++covered;
		}
		else
		{
			if (A.clearDepth) depthTile[p] = 0.0f;
			if (A.clearColour) colourTile[p] = A.clearWord;
		}
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) covered += __shfl_xor_sync(0xFFFFFFFFu, covered, o);
	if ((threadIdx.x & 31u) == 0 && covered) atomicAdd(&A.ctl->pixelsCovered, covered);
	// The last CTA to get here publishes the frame: it stores the control block into the host's pinned copy (a 64-byte
	// store over PCIe instead of a copy-engine job per frame), and in a screen-tile split it first stamps this rank's
	// arrival flag in the root's memory (after a system-scope fence by every thread that stored tiles) — the store into the
	// root's framebuffer WAS the composite — and, on the root, waits until every rank's stamp has arrived, so the root's
	// stream — and its srb_sync — completes exactly when the whole frame is in its framebuffer.  No host barrier in a frame.
	__shared__ uint32_t s_last;
	bool const split = !kFast && A.splitFlags;
	if (split)
	{
		__threadfence_system();
	}
	else
	{
		__threadfence();
	}
	__syncthreads();
	if (threadIdx.x == 0)
	{
		s_last = atomicAdd(&A.ctl->frameDone, 1u) == gridDim.x - 1u ? 1u : 0u;
	}
	__syncthreads();
	if (s_last)
	{
		if (split)
		{
			uint32_t const stamp = A.ctl->doneValue;
			if (threadIdx.x == 0)
			{
				__threadfence_system();
				*reinterpret_cast<volatile uint32_t*>(A.splitFlags + A.fp.ownRem) = stamp;
			}
			if (A.splitIsRoot && threadIdx.x < A.fp.ownMod && threadIdx.x < 32u && threadIdx.x != A.fp.ownRem)
			{
				split_wait(A.splitFlags + threadIdx.x, stamp, A.ctl);
			}
			__syncthreads(); // a wait that timed out has set the overflow bit
		}
		if (threadIdx.x < (uint32_t)(sizeof(FrameCtl) / 4u))
		{
			A.hostCtl[threadIdx.x] = __ldcg(reinterpret_cast<const uint32_t*>(A.ctl) + threadIdx.x);
		}
	}
}

// ---------------------------------------------------------------------------------------------------------------
// parity dump kernels (debug only; they reuse the device functions of the production path above)
// ---------------------------------------------------------------------------------------------------------------
__global__ void dump_tile_tris_kernel(RasterArgs A, uint32_t tile, const KeySlot* list, uint32_t count,
                                      srb_tile_tri* out, uint32_t cap)
{
	int32_t const X0 = (int32_t)(tile % A.fp.tilesX) * SRB_TILE;
	int32_t const Y0 = (int32_t)(tile / A.fp.tilesX) * SRB_TILE;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count && i < cap; i += gridDim.x * blockDim.x)
	{
		uint32_t const rank = list[i].slot;
		RasterRec r;
		load_raster_rec(A.rrecs, rank, r);
		ShadeRec const sr = A.srecs[rank];
		TileEdges const te = tile_edges(r, X0, Y0);
		srb_tile_tri o;
		for (int k = 0; k < 3; ++k)
		{
			o.c[k] = te.c[k];
			o.dx[k] = r.dx[k];
			o.dy[k] = r.dy[k];
		}
		o.block_min_x = (uint8_t)te.minX;
		o.block_max_x = (uint8_t)te.maxX;
		o.block_min_y = (uint8_t)te.minY;
		o.block_max_y = (uint8_t)te.maxY;
		float const sx = subf((float)X0, r.r0x), sy = subf((float)Y0, r.r0y);
		o.recip_w[0] = plane_c0(sr.wdx, sr.wdy, sr.w0, sx, sy);
		o.recip_w[1] = sr.wdx;
		o.recip_w[2] = sr.wdy;
		o.z_over_w[0] = plane_c0(r.zdx, r.zdy, r.z0, sx, sy);
		o.z_over_w[1] = r.zdx;
		o.z_over_w[2] = r.zdy;
		uint32_t const nv = A.draws[sr.pad[0]].numVaryings;
		for (uint32_t j = 0; j < SRB_MAX_VARY; ++j)
		{
			bool const on = j < nv;
			const float* q = sr.pl[SRB_PLANE_SLOT(j)];
			o.attr_dx[j] = on ? q[0] : 0.0f;
			o.attr_dy[j] = on ? q[1] : 0.0f;
			o.attr_c[j] = on ? plane_c0(q[0], q[1], q[2], sx, sy) : 0.0f;
		}
		o.attribs_per_tri = nv;
		o.draw_idx = sr.pad[0];
		out[i] = o;
	}
}

// one thread per (entry, 8x8 block): coverage before the depth-buffer test (inside all edges and z > 0), honouring
// the reference's block loop bounds and coarse rejection.
__global__ void dump_tile_coverage_kernel(RasterArgs A, uint32_t tile, const KeySlot* list, uint32_t listCount,
                                          unsigned long long* masks, uint32_t cap)
{
	uint32_t const count = min(listCount, cap);
	int32_t const X0 = (int32_t)(tile % A.fp.tilesX) * SRB_TILE;
	int32_t const Y0 = (int32_t)(tile / A.fp.tilesX) * SRB_TILE;
	for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < count * 64u; w += gridDim.x * blockDim.x)
	{
		uint32_t const i = w >> 6, b = w & 63u;
		int32_t const xB = (int32_t)(b & 7u) * 8, yB = (int32_t)(b >> 3) * 8;
		uint32_t const rank = list[i].slot;
		RasterRec r;
		load_raster_rec(A.rrecs, rank, r);
		TileEdges const te = tile_edges(r, X0, Y0);
		unsigned long long mask = 0ull;
		bool const visited = xB >= (te.minX & ~7) && xB < te.maxX && yB >= (te.minY & ~7) && yB < te.maxY;
		if (visited)
		{
			TriTile tt;
			for (int k = 0; k < 3; ++k)
			{
				tt.c[k] = te.c[k];
				tt.dx[k] = r.dx[k];
				tt.dy[k] = r.dy[k];
			}
			tt.zc0 = plane_c0(r.zdx, r.zdy, r.z0, subf((float)X0, r.r0x), subf((float)Y0, r.r0y));
			tt.zdx = r.zdx;
			tt.zdy = r.zdy;
			int32_t e00[3];
			int const mode = ref_coarse(tt, xB, yB, e00);
			if (mode != 0)
			{
				for (int32_t l = 0; l < 8; ++l)
				{
					int32_t e[3];
					for (int k = 0; k < 3; ++k) e[k] = wrap_add(e00[k], wrap_mul(tt.dy[k], l));
					float z = block_z0(tt, xB, yB, l);
					for (int row = 0; row < 8; ++row)
					{
						bool const inside = (mode == 2) || ((e[0] | e[1] | e[2]) >= 0);
						if (inside && z > 0.0f)
						{
							mask |= 1ull << (row * 8 + l);
						}
						for (int k = 0; k < 3; ++k) e[k] = wrap_add(e[k], tt.dx[k]);
						z = addf(z, tt.zdy);
					}
				}
			}
		}
		masks[w] = mask;
	}
}

// Stand-alone sampler entry for unit tests of the texture path: n fragments, same code as the tile kernel.
__global__ void sample_kernel(const TexDev* texs, uint32_t texIdx, const float* u, const float* v, const float* dudx,
                              const float* dudy, const float* dvdx, const float* dvdy, uint32_t* out, uint32_t n)
{
	fill_spread_table(g_sTab.spread);
	__syncthreads();
	uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
	{
		TexHead const tex = load_tex_head(texs + texIdx);
		out[i] = sample_wrap<true>(tex, [&](uint32_t mip) { return texs[texIdx].mipOffsets[mip]; }, tables_base() + (uint32_t)offsetof(ShadeTables, spread), u[i], v[i], dudx[i],
		                     dudy[i], dvdx[i], dvdy[i]);
	}
}

__global__ void rcp_kernel(const uint32_t* table, uint32_t bits, const float* in, float* out, uint32_t n)
{
	uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
	{
		out[i] = rcp_x86(in[i], table, bits);
	}
}

__global__ void rsqrt_kernel(const uint32_t* table, uint32_t bits, const float* in, float* out, uint32_t n)
{
	uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
	{
		out[i] = rsqrt_x86(in[i], table, bits);
	}
}

// De-tile the colour plane into linear RGBA8 — BlitJobFn, Renderer.cpp:319-347.  The one purely streaming kernel of
// the path: a row of a tile is 256 contiguous bytes on both sides, so a thread moves 16 bytes (4 pixels) and 16
// consecutive threads one tile row; widths that are not a multiple of 4 pixels take the scalar variant.
__global__ void __launch_bounds__(256) detile_kernel_vec(const uint4* __restrict__ colourTiles, uint4* __restrict__ linear,
                                                         uint32_t width4, uint32_t height, uint32_t tilesX)
{
	uint32_t const x4 = blockIdx.x * blockDim.x + threadIdx.x; // in units of 4 pixels
	uint32_t const y = blockIdx.y;
	if (x4 < width4 && y < height)
	{
		uint32_t const tile = (y >> 6) * tilesX + (x4 >> 4);
		linear[(size_t)y * width4 + x4] = __ldcs(colourTiles + ((size_t)tile * 1024u + (y & 63u) * 16u + (x4 & 15u)));
	}
}

// The same with the copy engines of the SM instead of its load/store units (Hopper+ bulk asynchronous copies, UBLKCP in
// SASS): a tile is 16 KiB of contiguous memory, so ONE thread brings it into shared memory with one cp.async.bulk
// (completion on an mbarrier), and each of its 64 rows — 256 contiguous bytes in the linear image — leaves with one bulk
// store.  A CTA issues 65 instructions' worth of copies instead of 1024 load/store pairs, and 14 CTAs per SM keep
// 224 KiB per SM in flight, which is what a copy this short (8.3 MB at 1080p = 2.6 us at the HBM peak) needs.
// Needs rows that start and end on 16-byte boundaries (width % 4 == 0).
constexpr int kDetileBulkThreads = 64;
__global__ void __launch_bounds__(kDetileBulkThreads) detile_bulk_kernel(const uint8_t* __restrict__ colourTiles,
                                                                        uint8_t* __restrict__ linear, uint32_t width,
                                                                        uint32_t height, uint32_t tilesX)
{
	__shared__ __align__(128) uint8_t s_tile[SRB_TILE_PIXELS * 4];
	__shared__ __align__(8) unsigned long long s_mbar;
	uint32_t const tile = blockIdx.x;
	uint32_t const x0 = (tile % tilesX) * SRB_TILE, y0 = (tile / tilesX) * SRB_TILE;
	uint32_t const mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
	uint32_t const sm = (uint32_t)__cvta_generic_to_shared(s_tile);
	if (threadIdx.x == 0)
	{
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(1));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(SRB_TILE_PIXELS * 4) : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sm),
		             "l"(colourTiles + (size_t)tile * (SRB_TILE_PIXELS * 4)), "r"(SRB_TILE_PIXELS * 4), "r"(mbar)
		             : "memory");
	}
	__syncthreads(); // the barrier is initialised before anybody polls it
	uint32_t done = 0;
	while (!done)
	{
		asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
		             : "=r"(done)
		             : "r"(mbar), "r"(0)
		             : "memory");
	}
	uint32_t const row = threadIdx.x, y = y0 + row;
	if (y < height)
	{
		uint32_t const bytes = min((uint32_t)SRB_TILE, width - x0) * 4u;
		asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(linear + ((size_t)y * width + x0) * 4u),
		             "r"(sm + row * (SRB_TILE * 4u)), "r"(bytes)
		             : "memory");
		asm volatile("cp.async.bulk.commit_group;" ::: "memory");
		asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // shared memory must outlive the reads of the copy
	}
}

__global__ void detile_kernel(const uint32_t* __restrict__ colourTiles, uint32_t* __restrict__ linear, uint32_t width,
                              uint32_t height, uint32_t tilesX)
{
	uint32_t const x = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t const y = blockIdx.y;
	if (x < width && y < height)
	{
		uint32_t const tile = (y >> 6) * tilesX + (x >> 6);
		linear[(size_t)y * width + x] = colourTiles[(size_t)tile * 4096u + (y & 63u) * 64u + (x & 63u)];
	}
}

} // namespace

size_t raster_smem_bytes() { return sizeof(WarpStage) * kWarpsPerCta; }

cudaError_t raster_init() { return cudaSuccess; }

void launch_raster(const RasterArgs& A, uint32_t ctas, cudaStream_t stream)
{
	raster_kernel<<<ctas, kRasterThreads, 0, stream>>>(A);
}

void launch_shade(const RasterArgs& A, cudaStream_t stream)
{
	uint32_t const numPixels = A.fp.tilesX * A.fp.tilesY * SRB_TILE_PIXELS;
	uint32_t blocks = (numPixels + kShadeThreads - 1) / kShadeThreads;
	static uint32_t const envCtas = [] {
		const char* e = getenv("SRB_SHADE_CTAS_PER_SM"); // tuning knob for experiments (not part of the ABI)
		return (uint32_t)(e && atoi(e) > 0 ? atoi(e) : 0);
	}();
	uint32_t const maxBlocks = 148u * (envCtas ? envCtas : (A.shadeCtasPerSm ? A.shadeCtasPerSm : 16u));
	if (blocks > maxBlocks) blocks = maxBlocks; // grid-stride: one covered-pixel atomic per warp of a resident CTA
	bool const texSmem = A.numTexs <= kSmemTexs, sponza = A.sponza != nullptr;
	bool const fast = texSmem && !sponza && A.uniformUnlit && A.fp.ownMod <= 1u && !A.winnersOut && !A.splitFlags;
	static bool const quadTaps = getenv("SRB_QUAD_TAPS") != nullptr; // 128-bit texel taps under a warp vote (see sample_wrap)
	if (fast && quadTaps) shade_kernel<true, false, true, true><<<blocks, kShadeThreads, 0, stream>>>(A);
	else if (fast) shade_kernel<true, false, true><<<blocks, kShadeThreads, 0, stream>>>(A);
	else if (texSmem && !sponza) shade_kernel<true, false, false><<<blocks, kShadeThreads, 0, stream>>>(A);
	else if (texSmem) shade_kernel<true, true, false><<<blocks, kShadeThreads, 0, stream>>>(A);
	else if (!sponza) shade_kernel<false, false, false><<<blocks, kShadeThreads, 0, stream>>>(A);
	else shade_kernel<false, true, false><<<blocks, kShadeThreads, 0, stream>>>(A);
}

#ifdef SRB_STATS
void stats_null_taps(int on)
{
	cudaDeviceSynchronize();
	cudaMemcpyToSymbol(g_nullTaps, &on, sizeof(on));
}

void stats_read(unsigned long long* out, bool reset)
{
	cudaDeviceSynchronize();
	cudaMemcpyFromSymbol(out, g_stats, sizeof(g_stats));
	if (reset)
	{
		unsigned long long zero[16] = {0};
		cudaMemcpyToSymbol(g_stats, zero, sizeof(zero));
	}
}
#endif

int raster_ctas_per_sm()
{
	int n = 0;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, raster_kernel, kRasterThreads, 0);
	return n;
}

void launch_dump_tile_tris(const RasterArgs& A, uint32_t tile, const KeySlot* list, uint32_t count, srb_tile_tri* out,
                           uint32_t cap, cudaStream_t stream)
{
	dump_tile_tris_kernel<<<64, 128, 0, stream>>>(A, tile, list, count, out, cap);
}

void launch_dump_tile_coverage(const RasterArgs& A, uint32_t tile, const KeySlot* list, uint32_t count,
                               unsigned long long* masks, uint32_t cap, cudaStream_t stream)
{
	dump_tile_coverage_kernel<<<256, 128, 0, stream>>>(A, tile, list, count, masks, cap);
}

void launch_sample(const TexDev* texs, uint32_t texIdx, const float* u, const float* v, const float* dudx,
                   const float* dudy, const float* dvdx, const float* dvdy, uint32_t* out, uint32_t n,
                   cudaStream_t stream)
{
	sample_kernel<<<(n + 127) / 128, 128, 0, stream>>>(texs, texIdx, u, v, dudx, dudy, dvdx, dvdy, out, n);
}

void launch_rcp(const uint32_t* table, uint32_t bits, const float* in, float* out, uint32_t n, cudaStream_t stream)
{
	rcp_kernel<<<(n + 127) / 128, 128, 0, stream>>>(table, bits, in, out, n);
}

void launch_rsqrt(const uint32_t* table, uint32_t bits, const float* in, float* out, uint32_t n, cudaStream_t stream)
{
	rsqrt_kernel<<<(n + 127) / 128, 128, 0, stream>>>(table, bits, in, out, n);
}

void launch_detile(const uint32_t* colourTiles, uint32_t* linear, uint32_t width, uint32_t height, uint32_t tilesX,
                   cudaStream_t stream)
{
	// Measured (profiles/README.md): the bulk-copy kernel moves the same bytes in the same time (1080p 6.6 - 6.9 us against
	// 6.1 - 6.5 us for the vector kernel, 3840x2160 12.0 - 13.0 against 11.9 - 13.0): a copy this short is bound by its launch
	// and the DRAM round trip, not by load/store instructions.  The vector kernel stays the default; SRB_BULK_DETILE=1 selects
	// the bulk-copy one (one fifth of the warp instructions, which only matters beside a frame that needs the issue slots).
	static bool const bulk = getenv("SRB_BULK_DETILE") != nullptr; // A/B knob (not part of the ABI)
	if ((width & 3u) == 0u && bulk)
	{
		uint32_t const tilesY = (height + SRB_TILE - 1) / SRB_TILE;
		detile_bulk_kernel<<<tilesX * tilesY, kDetileBulkThreads, 0, stream>>>(reinterpret_cast<const uint8_t*>(colourTiles),
		                                                                      reinterpret_cast<uint8_t*>(linear), width, height, tilesX);
		return;
	}
	if ((width & 3u) == 0u)
	{
		uint32_t const width4 = width / 4u;
		dim3 grid((width4 + 255) / 256, height);
		detile_kernel_vec<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(colourTiles), reinterpret_cast<uint4*>(linear),
		                                          width4, height, tilesX);
		return;
	}
	dim3 grid((width + 255) / 256, height);
	detile_kernel<<<grid, 256, 0, stream>>>(colourTiles, linear, width, height, tilesX);
}

} // namespace srb
