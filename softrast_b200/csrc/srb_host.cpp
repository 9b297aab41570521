// srb_host.cpp — host-only helpers of the C ABI that need no CUDA: the texture builder in the reference's storage
// layout and the RCPPS table harvest.  Compiled with the host compiler and linked into libsoftrast_b200.so.
#include "../../include/softrast_b200.h"

#include <immintrin.h>
#include <math.h>
#include <string.h>

#include <vector>

namespace
{

// x bits in even positions, y bits in odd positions (reference SoftRast/Texture.cpp:36-41), 5 bits each.
inline uint32_t Morton5(uint32_t x, uint32_t y)
{
	uint32_t m = 0;
	for (uint32_t b = 0; b < 5; ++b)
	{
		m |= ((x >> b) & 1u) << (2 * b);
		m |= ((y >> b) & 1u) << (2 * b + 1);
	}
	return m;
}

inline uint32_t AlignUp32(uint32_t v) { return (v + 31u) & ~31u; }

// Linear RGBA8 -> 32x32 tiles, Morton order inside a tile (reference Texture.cpp:73-101).
void TileLevel(const uint8_t* src, uint8_t* dst, uint32_t w, uint32_t h)
{
	uint32_t const tilesX = AlignUp32(w) >> 5;
	for (uint32_t y = 0; y < h; ++y)
	{
		for (uint32_t x = 0; x < w; ++x)
		{
			uint32_t const offs = ((y >> 5) * tilesX + (x >> 5)) * 1024u + Morton5(x & 31u, y & 31u);
			memcpy(dst + size_t(offs) * 4, src + (size_t(y) * w + x) * 4, 4);
		}
	}
}

inline uint32_t FloorLog2(uint32_t v)
{
	uint32_t r = 0;
	while (v >>= 1) ++r;
	return r;
}

// ---- the reference's mip filter ------------------------------------------------------------------------------------
// The reference builds every mip level from the ORIGINAL image with stbir_resize_uint8 (Texture.cpp:188-198), i.e.
// stb_image_resize's down-sampling path with its defaults: Mitchell-Netravali kernel, clamped edges, linear colour space,
// no alpha handling (stb_image_resize.h:2464-2472).  This is a restatement of that path for 4 x uint8 channels that
// keeps stb's arithmetic and, above all, its ORDER of float additions, so the bytes come out identical:
//   coefficients  stbir__calculate_filters / _calculate_coefficients_downsample / _normalize_downsample_coefficients
//                 (:1194-1232, :1087-1115, :1117-1190), kernel stbir__filter_mitchell (:824-836)
//   horizontal    stbir__resample_horizontal_downsample (:1526-1655): every input pixel is scattered into the output
//                 pixels it contributes to, input pixels in ascending order
//   vertical      stbir__buffer_loop_downsample / _resample_vertical_downsample (:2166-2205, :1987-2066): every input
//                 row is scattered into the output rows it contributes to, rows in ascending order
//   decode/encode byte / 255.0f (:1275-1283); (uchar)(int)(saturate(f) * 255.0f + 0.5 [double]) (:1733-1758)
// (Compiled without FMA contraction, like the parity build of the reference.)
struct StbAxis
{
	int margin = 0;              // stbir__get_filter_pixel_margin
	std::vector<int> n0, n1;     // per contributor (input pixel incl. margins): first / last output pixel
	std::vector<float> coef;     // 4 per contributor (stbir__get_coefficient_width of a support-2 filter)
	float scale = 1.0f;
};

inline float StbMitchell(float x)
{
	x = (float)fabs(x);
	if (x < 1.0f) return (16 + x * x * (21 * x - 36)) / 18;
	if (x < 2.0f) return (32 + x * (-60 + x * (36 - 7 * x))) / 18;
	return 0.0f;
}

void StbCalculateFilters(StbAxis& a, int inputSize, int outputSize)
{
	float const scale = ((float)outputSize / inputSize) / (1.0f - 0.0f); // stbir__calculate_transform, s0 = 0, s1 = 1
	float const shift = 0.0f * outputSize / (1.0f - 0.0f);
	a.scale = scale;
	int const pixelWidth = (int)ceil(2.0f * 2 / scale); // stbir__get_filter_pixel_width (down-sampling branch)
	a.margin = pixelWidth / 2;
	int const width = 4; // stbir__get_coefficient_width: ceil(support * 2)
	int const num = inputSize + a.margin * 2;
	a.n0.assign(num, 0);
	a.n1.assign(num, -1);
	a.coef.assign(size_t(num) * width + 8, 0.0f); // + slack: stb writes a (zero) fifth coefficient past a group
	float const radius = 2.0f / scale; // support / scale_ratio
	for (int n = 0; n < num; ++n)
	{
		int const nAdj = n - a.margin;
		// stbir__calculate_sample_range_downsample
		float const centre = (float)nAdj + 0.5f;
		float const lo = centre - radius, hi = centre + radius;
		float const outLo = lo * scale - shift, outHi = hi * scale - shift;
		float const outCentre = centre * scale - shift;
		int const first = (int)floor(outLo + 0.5), last = (int)floor(outHi - 0.5);
		// stbir__calculate_coefficients_downsample
		float* g = &a.coef[size_t(n) * width];
		a.n0[n] = first;
		a.n1[n] = last;
		for (int i = 0; i <= last - first; ++i)
		{
			float const outPixelCentre = (float)(i + first) + 0.5f;
			float const x = outPixelCentre - outCentre;
			g[i] = StbMitchell(x) * scale;
		}
		for (int i = last - first; i >= 0; --i)
		{
			if (g[i]) break;
			a.n1[n] = a.n0[n] + i - 1;
		}
	}
	// stbir__normalize_downsample_coefficients
	for (int i = 0; i < outputSize; ++i)
	{
		float total = 0;
		for (int j = 0; j < num; ++j)
		{
			if (i >= a.n0[j] && i <= a.n1[j]) total += a.coef[size_t(j) * width + (i - a.n0[j])];
			else if (i < a.n0[j]) break;
		}
		float const s = 1 / total;
		for (int j = 0; j < num; ++j)
		{
			if (i >= a.n0[j] && i <= a.n1[j]) a.coef[size_t(j) * width + (i - a.n0[j])] *= s;
			else if (i < a.n0[j]) break;
		}
	}
	for (int j = 0; j < num; ++j)
	{
		int skip = 0;
		while (a.coef[size_t(j) * width + skip] == 0) skip++;
		a.n0[j] += skip;
		while (a.n0[j] < 0)
		{
			a.n0[j]++;
			skip++;
		}
		int const range = a.n1[j] - a.n0[j] + 1;
		int const max = width < range ? width : range;
		for (int i = 0; i < max; ++i)
		{
			if (i + skip >= width) break;
			a.coef[size_t(j) * width + i] = a.coef[size_t(j) * width + i + skip];
		}
	}
	for (int j = 0; j < num; ++j)
	{
		a.n1[j] = a.n1[j] < outputSize - 1 ? a.n1[j] : outputSize - 1;
	}
}

inline int ClampIdx(int n, int max) { return n < 0 ? 0 : (n >= max ? max - 1 : n); }

// stbir_resize_uint8(src, iw, ih, 0, dst, ow, oh, 0, 4) for ow <= iw, oh <= ih.
void StbResizeDown4(const uint8_t* src, int iw, int ih, uint8_t* dst, int ow, int oh)
{
	StbAxis H, V;
	StbCalculateFilters(H, iw, ow);
	StbCalculateFilters(V, ih, oh);
	std::vector<float> decode(size_t(iw + 2 * H.margin) * 4), hbuf(size_t(ow) * 4), acc(size_t(ow) * oh * 4, 0.0f);
	float const vRadius = 2.0f / V.scale;
	for (int y = -V.margin; y < ih + V.margin; ++y)
	{
		// stbir__buffer_loop_downsample: rows whose (un-normalised) range misses the output are skipped
		float const centre = (float)y + 0.5f;
		float const outLo = (centre - vRadius) * V.scale - 0.0f, outHi = (centre + vRadius) * V.scale - 0.0f;
		int const first = (int)floor(outLo + 0.5), last = (int)floor(outHi - 0.5);
		if (last < 0 || first >= oh) continue;
		// stbir__decode_scanline (clamped edges)
		const uint8_t* row = src + size_t(ClampIdx(y, ih)) * iw * 4;
		for (int x = -H.margin; x < iw + H.margin; ++x)
		{
			const uint8_t* px = row + size_t(ClampIdx(x, iw)) * 4;
			for (int c = 0; c < 4; ++c) decode[size_t(x + H.margin) * 4 + c] = ((float)px[c]) / 255.0f;
		}
		// stbir__resample_horizontal_downsample
		std::fill(hbuf.begin(), hbuf.end(), 0.0f);
		for (int x = 0; x < iw + 2 * H.margin; ++x)
		{
			const float* in = &decode[size_t(x) * 4];
			for (int k = H.n0[x]; k <= H.n1[x]; ++k)
			{
				float const co = H.coef[size_t(x) * 4 + (k - H.n0[x])];
				float* out = &hbuf[size_t(k) * 4];
				out[0] += in[0] * co;
				out[1] += in[1] * co;
				out[2] += in[2] * co;
				out[3] += in[3] * co;
			}
		}
		// stbir__resample_vertical_downsample
		int const contributor = y + V.margin;
		for (int k = V.n0[contributor]; k <= V.n1[contributor]; ++k)
		{
			float const co = V.coef[size_t(contributor) * 4 + (k - V.n0[contributor])];
			float* out = &acc[size_t(k) * ow * 4];
			for (int i = 0; i < ow * 4; ++i) out[i] += hbuf[i] * co;
		}
	}
	// stbir__encode_scanline
	for (size_t i = 0; i < size_t(ow) * oh * 4; ++i)
	{
		float f = acc[i];
		f = f < 0 ? 0 : (f > 1 ? 1 : f);
		dst[i] = (unsigned char)((int)((f * 255.0f) + 0.5));
	}
}

inline float HostRcp(float x) { return _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(x))); }

inline float HostRsqrt(float x) { return _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x))); }

inline uint32_t Bits(float f)
{
	uint32_t u;
	memcpy(&u, &f, 4);
	return u;
}

inline float FromBits(uint32_t u)
{
	float f;
	memcpy(&f, &u, 4);
	return f;
}

} // namespace

// internal (not exported): the filter tables of one axis for the device-side builder (srb_texbuild.cu / srb_api.cu)
void srb_internal_stb_axis(int inputSize, int outputSize, int* margin, int** n0, int** n1, float** coef, int* numContributors)
{
	StbAxis a;
	StbCalculateFilters(a, inputSize, outputSize);
	int const num = inputSize + 2 * a.margin;
	*margin = a.margin;
	*numContributors = num;
	*n0 = new int[num];
	*n1 = new int[num];
	*coef = new float[size_t(num) * 4];
	memcpy(*n0, a.n0.data(), sizeof(int) * num);
	memcpy(*n1, a.n1.data(), sizeof(int) * num);
	memcpy(*coef, a.coef.data(), sizeof(float) * size_t(num) * 4);
}

void srb_internal_stb_axis_free(int* n0, int* n1, float* coef)
{
	delete[] n0;
	delete[] n1;
	delete[] coef;
}

extern "C"
{

SRB_API int srb_texture_build_rgba8(const uint8_t* rgba, uint32_t width, uint32_t height, int calc_mips,
                                    uint8_t* texels_out, uint64_t* bytes_out, uint32_t* mip_offsets_out,
                                    uint32_t* num_mips_out)
{
	// Same preconditions as TextureData::CreateFromRGBA8 (Texture.cpp:122-130): powers of two, multiples of 32.
	if (!width || !height || (width & (width - 1)) || (height & (height - 1)) || (width & 31u) || (height & 31u) ||
	    FloorLog2(width) >= SRB_MAX_TEX_DIM_LOG2 || FloorLog2(height) >= SRB_MAX_TEX_DIM_LOG2)
	{
		return SRB_ERR_INVALID;
	}
	uint32_t const numMips = calc_mips ? FloorLog2(width > height ? width : height) + 1u : 1u;
	// Mip placement, Texture.cpp:159-175: level k is max(1, dim >> k), each padded to a multiple of 32.
	uint64_t total = 0;
	uint32_t offsets[SRB_MAX_TEX_DIM_LOG2] = {0};
	for (uint32_t m = 0; m < numMips; ++m)
	{
		uint32_t const w = (width >> m) ? (width >> m) : 1u, h = (height >> m) ? (height >> m) : 1u;
		offsets[m] = (uint32_t)total;
		total += uint64_t(AlignUp32(w)) * AlignUp32(h) * 4u;
	}
	if (bytes_out) *bytes_out = total;
	if (num_mips_out) *num_mips_out = numMips;
	if (mip_offsets_out) memcpy(mip_offsets_out, offsets, sizeof(offsets));
	if (!texels_out)
	{
		return SRB_OK;
	}
	if (!rgba)
	{
		return SRB_ERR_INVALID;
	}
	memset(texels_out, 0, total);
	if (calc_mips == SRB_MIPS_STB)
	{
		// the reference's own mips: every level filtered from the original image (Texture.cpp:188-198)
		std::vector<uint8_t> level;
		for (uint32_t m = 0; m < numMips; ++m)
		{
			uint32_t const w = (width >> m) ? (width >> m) : 1u, h = (height >> m) ? (height >> m) : 1u;
			if (m == 0)
			{
				TileLevel(rgba, texels_out + offsets[0], w, h);
				continue;
			}
			level.assign(size_t(w) * h * 4, 0);
			StbResizeDown4(rgba, (int)width, (int)height, level.data(), (int)w, (int)h);
			TileLevel(level.data(), texels_out + offsets[m], w, h);
		}
		return SRB_OK;
	}
	std::vector<uint8_t> cur(rgba, rgba + size_t(width) * height * 4), next;
	uint32_t w = width, h = height;
	for (uint32_t m = 0; m < numMips; ++m)
	{
		if (m > 0)
		{
			// 2x2 box filter, round to nearest; a dimension already at 1 stays 1
			uint32_t const nw = w > 1 ? w / 2 : 1, nh = h > 1 ? h / 2 : 1;
			next.assign(size_t(nw) * nh * 4, 0);
			for (uint32_t y = 0; y < nh; ++y)
			{
				for (uint32_t x = 0; x < nw; ++x)
				{
					uint32_t const x0 = w > 1 ? 2 * x : 0, x1 = w > 1 ? 2 * x + 1 : 0;
					uint32_t const y0 = h > 1 ? 2 * y : 0, y1 = h > 1 ? 2 * y + 1 : 0;
					for (uint32_t c = 0; c < 4; ++c)
					{
						uint32_t const s = cur[(size_t(y0) * w + x0) * 4 + c] + cur[(size_t(y0) * w + x1) * 4 + c] +
						                   cur[(size_t(y1) * w + x0) * 4 + c] + cur[(size_t(y1) * w + x1) * 4 + c];
						next[(size_t(y) * nw + x) * 4 + c] = (uint8_t)((s + 2u) >> 2);
					}
				}
			}
			cur.swap(next);
			w = nw;
			h = nh;
		}
		TileLevel(cur.data(), texels_out + offsets[m], w, h);
	}
	return SRB_OK;
}

/* Reads the RCPPS mantissa table of the CPU this process runs on.  Tries index widths 11..max_index_bits and returns
 * the smallest one that reproduces RCPPS exactly for every one of the 2^23 mantissas at three exponents (and the
 * exponent-shift model across all exponents for a subsample); 0 if none does. */
SRB_API uint32_t srb_harvest_rcp_table(uint32_t* table, uint32_t max_index_bits)
{
	if (!table || max_index_bits < 11)
	{
		return 0;
	}
	if (max_index_bits > 23) max_index_bits = 23;
	for (uint32_t bits = 11; bits <= max_index_bits; ++bits)
	{
		uint32_t const n = 1u << bits;
		for (uint32_t i = 0; i < n; ++i)
		{
			table[i] = Bits(HostRcp(FromBits(0x3F800000u | (i << (23 - bits)))));
		}
		bool ok = true;
		for (uint32_t m = 0; m < (1u << 23) && ok; ++m)
		{
			// exponent 127 (exhaustive), plus two other exponents on a stride
			ok = Bits(HostRcp(FromBits(0x3F800000u | m))) == table[m >> (23 - bits)];
			if (ok && (m & 63u) == 0)
			{
				for (uint32_t e : {1u, 60u, 130u, 200u, 252u})
				{
					int32_t const r = (int32_t)table[m >> (23 - bits)] + ((127 - (int32_t)e) << 23);
					uint32_t const expect = r < 0x00800000 ? 0u : (uint32_t)r;
					if (Bits(HostRcp(FromBits((e << 23) | m))) != expect)
					{
						ok = false;
						break;
					}
				}
			}
		}
		if (ok)
		{
			return bits;
		}
	}
	return 0;
}

/* Reads the RSQRTPS table of the CPU this process runs on (reference Viewer/SponzaScene.cpp:66 uses _mm256_rsqrt_ps).
 * Model: x = 2^(2k+p) * m, p in {0,1}, m in [1,2)  ->  RSQRTPS(x) = 2^-k * T[p][top `bits` bits of m's mantissa], i.e.
 * table[(p << bits) | i] = bits(RSQRTPS(2^p * (1 + i * 2^-bits))) and the result's exponent field is lowered by k.
 * Tries index widths 10..max_index_bits and returns the smallest one that reproduces the instruction exactly for every
 * one of the 2^23 mantissas at both parities (and the exponent model across the exponent range on a subsample); 0 if
 * none does.  `table` needs 2 << max_index_bits entries. */
SRB_API uint32_t srb_harvest_rsqrt_table(uint32_t* table, uint32_t max_index_bits)
{
	if (!table || max_index_bits < 10)
	{
		return 0;
	}
	if (max_index_bits > 23) max_index_bits = 23;
	for (uint32_t bits = 10; bits <= max_index_bits; ++bits)
	{
		uint32_t const n = 1u << bits;
		for (uint32_t p = 0; p < 2; ++p)
		{
			for (uint32_t i = 0; i < n; ++i)
			{
				table[(p << bits) | i] = Bits(HostRsqrt(FromBits(((127u + p) << 23) | (i << (23 - bits)))));
			}
		}
		bool ok = true;
		for (uint32_t m = 0; m < (1u << 23) && ok; ++m)
		{
			for (uint32_t p = 0; p < 2 && ok; ++p)
			{
				ok = Bits(HostRsqrt(FromBits(((127u + p) << 23) | m))) == table[(p << bits) | (m >> (23 - bits))];
			}
			if (ok && (m & 63u) == 0)
			{
				for (uint32_t e : {1u, 2u, 61u, 128u, 131u, 200u, 253u, 254u})
				{
					int32_t const ue = (int32_t)e - 127;          // unbiased exponent = 2k + p
					int32_t const pp = ue & 1, k = (ue - pp) / 2; // floor division for negative exponents
					uint32_t const expect = table[((uint32_t)pp << bits) | (m >> (23 - bits))] - ((uint32_t)k << 23);
					if (Bits(HostRsqrt(FromBits((e << 23) | m))) != expect)
					{
						ok = false;
						break;
					}
				}
			}
		}
		if (ok)
		{
			return bits;
		}
	}
	return 0;
}

// ---- the Sponza scene's frame constants over time (Viewer/SponzaScene.cpp:126-160, :168-187) --------------------------
// Host arithmetic restated operation for operation (this file is compiled without FMA contraction): kt::XorShift32
// (Random.inl:33-41), RandomUnitFloat (:8-21), kt::Lerp = (1 - t) * a + t * b (MathUtil.inl:7-11), Normalize
// (Vec3.inl:113-127), Quat::FromNormalizedAxisAngle (Quat.inl:64-75), Mul(Quat, Vec3) = v + w * t + cross(q, t) with
// t = 2 * cross(q, v) (Quat.inl:116-121), sinf / cosf from libm like kt::Sin / kt::Cos.
namespace
{
struct V3
{
	float x, y, z;
};
inline V3 Cross3(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float LerpF(float a, float b, float t) { return (1.0f - t) * a + t * b; }
inline V3 Normalize3(V3 v)
{
	float const mag = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
	if (mag > 0.0f)
	{
		float const inv = 1.0f / mag;
		v.x *= inv;
		v.y *= inv;
		v.z *= inv;
	}
	return v;
}
struct XorShift32
{
	uint32_t state = 254693456u; // kt::XorShift32's default seed (Random.h:9)
	float Unit()
	{
		uint32_t x = state;
		x ^= x << 13;
		x ^= x >> 17;
		x ^= x << 5;
		state = x;
		return FromBits((0x7Fu << 23) | (x >> 9)) - 1.0f;
	}
	// kt::Vec3(RandomUnitFloat(rng), RandomUnitFloat(rng), RandomUnitFloat(rng)): C++ leaves the evaluation order of the
	// three arguments open; the reference as compiled (GCC here, MSVC originally) evaluates them RIGHT TO LEFT, so the
	// first number drawn is z.  tests/test_oracle.py pins this against the compiled reference.
	V3 Vec()
	{
		V3 v;
		v.z = Unit();
		v.y = Unit();
		v.x = Unit();
		return v;
	}
};
} // namespace

extern "C"
{

SRB_API void srb_sponza_scene_init(srb_sponza_scene* s)
{
	memset(s, 0, sizeof(*s));
	for (int i = 0; i < 3; ++i) s->constants.ambient[i] = 0.1f; // SponzaScene.cpp:126-130
	V3 const sun = Normalize3(V3{0.4f, 0.7f, 0.1f});
	s->constants.sun_dir[0] = s->constants.sun_dir[1] = s->constants.sun_dir[2] = sun.x; // :132-137 broadcasts x three times
	XorShift32 rng;
	for (uint32_t i = 0; i < SRB_SPONZA_POINT_LIGHTS; ++i) // :141-159
	{
		auto& a = s->anim[i];
		srb_sponza_light& l = s->constants.lights[i];
		V3 const ca = rng.Vec(), cb = rng.Vec();
		a.colour_a[0] = ca.x, a.colour_a[1] = ca.y, a.colour_a[2] = ca.z;
		a.colour_b[0] = cb.x, a.colour_b[1] = cb.y, a.colour_b[2] = cb.z;
		a.base_pos[0] = LerpF(-1000.0f, 1000.0f, rng.Unit());
		a.base_pos[1] = LerpF(50.0f, 250.0f, rng.Unit());
		a.base_pos[2] = LerpF(-150.0f, 150.0f, rng.Unit());
		l.falloff = LerpF(500.0f, 2500.0f, rng.Unit());
		l.intensity = LerpF(150.0f, 350.0f, rng.Unit());
		V3 ax = rng.Vec();
		ax = Normalize3(V3{ax.x * 2.0f - 1.0f, ax.y * 2.0f - 1.0f, ax.z * 2.0f - 1.0f});
		a.rot_axis[0] = ax.x, a.rot_axis[1] = ax.y, a.rot_axis[2] = ax.z;
		V3 const ro = rng.Vec();
		a.rot_offset[0] = (ro.x * 2.0f - 1.0f) * 25.0f + 5.0f;
		a.rot_offset[1] = (ro.y * 2.0f - 1.0f) * 25.0f + 5.0f;
		a.rot_offset[2] = (ro.z * 2.0f - 1.0f) * 25.0f + 5.0f;
		a.angle = 0.0f;
	}
}

SRB_API void srb_sponza_scene_update(srb_sponza_scene* s, float dt)
{
	float const sinT = sinf(s->anim_phase) * 0.5f + 0.5f; // :173
	s->anim_phase += dt;
	for (uint32_t i = 0; i < SRB_SPONZA_POINT_LIGHTS; ++i) // :177-191
	{
		auto& a = s->anim[i];
		srb_sponza_light& l = s->constants.lights[i];
		for (int k = 0; k < 3; ++k) l.colour[k] = LerpF(a.colour_a[k], a.colour_b[k], sinT);
		float const half = a.angle * 0.5f;
		float const sn = sinf(half), cs = cosf(half);
		V3 const q{a.rot_axis[0] * sn, a.rot_axis[1] * sn, a.rot_axis[2] * sn};
		V3 const v{a.base_pos[0] - a.rot_offset[0], a.base_pos[1] - a.rot_offset[1], a.base_pos[2] - a.rot_offset[2]};
		V3 const c0 = Cross3(q, v);
		V3 const t{c0.x * 2.0f, c0.y * 2.0f, c0.z * 2.0f};
		V3 const c1 = Cross3(q, t);
		V3 const r{(v.x + t.x * cs) + c1.x, (v.y + t.y * cs) + c1.y, (v.z + t.z * cs) + c1.z};
		l.pos[0] = r.x + a.rot_offset[0];
		l.pos[1] = r.y + a.rot_offset[1];
		l.pos[2] = r.z + a.rot_offset[2];
		a.angle += dt;
	}
}

} // extern "C"

SRB_API const char* srb_version(void) { return "softrast_b200 0.1 (sm_100a)"; }

} // extern "C"
