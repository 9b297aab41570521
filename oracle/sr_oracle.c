/*
 * sr_oracle.c — TEST INFRASTRUCTURE ONLY.  A plain-C, scalar, single-threaded restatement of the reference's
 * sort-middle frame pipeline (karltechno/SoftRast), used as the parity checker ("port") wherever the compiled
 * reference itself (oracle/_ref) is not available, and cross-checked against it bit-for-bit in tests/test_oracle.py.
 *
 * PARITY PIN: this file is pinned against oracle/_ref/libsrref_parity.so (the unmodified reference sources compiled
 * in place, -ffp-contract=off, 1 thread) on seeded scenes, and against the golden fixtures under tests/golden/ that
 * were generated from that library (tests/golden/make_golden.py).  The reference has no tests or golden vectors of
 * its own for this path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this code.  The
 * product (softrast_b200/csrc) never includes, links or calls it.
 *
 * Structure follows the reference, one function per reference function, each citing file:line under /root/reference:
 *   front-end  BinTrisEntry                        SoftRast/Binning.cpp:464-535
 *              ComputeClipMask / ClipPlane         SoftRast/Binning.cpp:56-68, 85-165
 *              BinTransformedAndClippedTri         SoftRast/Binning.cpp:279-456 (SetupEdge :242-259, SetupPlane :261-277)
 *   back-end   RasterAndShadeBin                   SoftRast/Rasterizer.cpp:525-577
 *              RasterizeTrisInBin_OutputFragments  SoftRast/Rasterizer.cpp:194-304
 *              ComputeBlockMask8x8[_DepthOnly]     SoftRast/Rasterizer.cpp:97-192
 *              ComputeInterpolantsDrawCallImpl     SoftRast/Rasterizer.cpp:306-419
 *              ShadeFragmentBuffer                 SoftRast/Rasterizer.cpp:460-523
 *   shaders    Unlit / Normals / UVs               Viewer/Shaders.h:71-130
 *              Sponza (sun + 16 point lights)      Viewer/SponzaScene.cpp:13-104
 *   sampler    SampleWrap, CalcMipLevels, Gather   SoftRast/Texture.cpp:381-452, 212-233, 243-379
 *   pack       RGBA32SoA_To_RGBA8AoS               SoftRast/SIMDUtil.h:87-121
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -mfma -fPIC -shared (no -ffast-math): every float operator below is one
 * IEEE binary32 operation and fmaf() is a hardware FMA, exactly like the reference built with the parity flags.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/softrast_b200.h"

#define BIN 64
#define MAXV 8

typedef struct
{
	const uint8_t* texels;
	uint64_t bytes;
	uint32_t mipOffsets[14];
	uint32_t numMips, wLog2, hLog2;
} OTex;

typedef struct
{
	srb_tile_tri* tris;
	uint32_t n, cap;
} OTile;

typedef struct
{
	uint32_t width, height, tilesX, tilesY;
	uint32_t* colour; /* tiles * 4096 */
	float* depth;     /* tiles * 4096 */
	OTile* tiles;
	OTex* texs;
	uint32_t numTex, capTex;
	srb_draw_desc* draws;
	uint32_t numDraws, capDraws;
	uint32_t rcpTable[1 << 16];
	uint32_t rcpBits;
	uint32_t rsqrtTable[2 << 16];
	uint32_t rsqrtBits;
	srb_sponza_constants sponza; /* g_constants of Viewer/SponzaScene.cpp:11 */
	uint64_t trisSetup, trisClipped;
} OCtx;

/* ------------------------------------------------------------------------------------------------------------ */
static int32_t cvtt(float f) /* cvttss2si */
{
	return (f >= -2147483648.0f && f < 2147483648.0f) ? (int32_t)f : INT32_MIN;
}
static int32_t cvtn(float f) /* cvtps2dq, round to nearest even */
{
	return (f >= -2147483648.0f && f < 2147483648.0f) ? (int32_t)lrintf(f) : INT32_MIN;
}
static uint32_t fbits(float f)
{
	uint32_t u;
	memcpy(&u, &f, 4);
	return u;
}
static float ffrom(uint32_t u)
{
	float f;
	memcpy(&f, &u, 4);
	return f;
}
static int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
static int32_t clampi(int32_t v, int32_t lo, int32_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* RCPPS replay: table on the top mantissa bits + exponent shift. */
static float rcp_x86(const OCtx* c, float x)
{
	uint32_t const u = fbits(x), s = u & 0x80000000u, e = (u >> 23) & 0xFFu, m = u & 0x7FFFFFu;
	if (e == 0) return ffrom(s | 0x7F800000u);
	if (e == 0xFF) return m ? ffrom(u | 0x00400000u) : ffrom(s);
	int32_t const r = (int32_t)c->rcpTable[m >> (23 - c->rcpBits)] + ((127 - (int32_t)e) << 23);
	if (r < 0x00800000) return ffrom(s);
	return ffrom(s | (uint32_t)r);
}

/* RSQRTPS replay: x = 2^(2k+p) * m -> table[(p << bits) | top bits of m] with the exponent lowered by k. */
static float rsqrt_x86(const OCtx* c, float x)
{
	uint32_t const u = fbits(x), e = (u >> 23) & 0xFFu, m = u & 0x7FFFFFu;
	if (e == 0xFF && m) return ffrom(u | 0x00400000u);                 /* NaN -> quiet NaN */
	if (e == 0) return ffrom((u & 0x80000000u) | 0x7F800000u);         /* +-0, denormals -> +-inf */
	if (u & 0x80000000u) return ffrom(0xFFC00000u);                    /* negative (incl. -inf) -> indefinite */
	if (e == 0xFF) return 0.0f;                                        /* +inf -> +0 */
	int32_t const ue = (int32_t)e - 127;
	uint32_t const p = (uint32_t)ue & 1u;
	int32_t const k = (ue - (int32_t)p) / 2;
	return ffrom(c->rsqrtTable[(p << c->rsqrtBits) | (m >> (23 - c->rsqrtBits))] - ((uint32_t)k << 23));
}

/* ---- front-end --------------------------------------------------------------------------------------------- */
typedef struct
{
	float x, y, z, w;
} V4;

/* Binning.cpp:56-68 */
static uint32_t ComputeClipMask(V4 v)
{
	uint32_t m = 0;
	if (v.x + v.w < 0.0f) m |= 1;
	if (v.x - v.w > 0.0f) m |= 2;
	if (v.y + v.w < 0.0f) m |= 4;
	if (v.y - v.w > 0.0f) m |= 8;
	if (v.z < 0.0f) m |= 16;
	if (v.z - v.w > 0.0f) m |= 32;
	return m;
}

/* kt::Lerp, kt/src/kt/inl/MathUtil.inl:7-11 */
static float LerpKt(float a, float b, float t) { return (1.0f - t) * a + t * b; }

typedef struct
{
	V4 verts[2][9];
	float attribs[2][9][MAXV];
	uint32_t numIn, inputIdx;
} ClipBuffer;

/* Binning.cpp:85-165 */
static void ClipPlane(ClipBuffer* b, uint32_t plane)
{
	static const float planes[6][4] = {{1, 0, 0, 1}, {-1, 0, 0, 1}, {0, 1, 0, 1}, {0, -1, 0, 1}, {0, 0, 1, 1}, {0, 0, -1, 1}};
	const float* P = planes[plane];
	V4* in = b->verts[b->inputIdx];
	V4* out = b->verts[b->inputIdx ^ 1];
	float(*ina)[MAXV] = b->attribs[b->inputIdx];
	float(*outa)[MAXV] = b->attribs[b->inputIdx ^ 1];
	uint32_t nOut = 0;
	uint32_t i0 = b->numIn - 1;
	/* kt::Dot(Vec4, Vec4), kt/src/kt/inl/Vec4.inl:162-165 */
	float d0 = P[0] * in[i0].x + P[1] * in[i0].y + P[2] * in[i0].z + P[3] * in[i0].w;
	for (uint32_t i1 = 0; i1 < b->numIn; ++i1)
	{
		float const d1 = P[0] * in[i1].x + P[1] * in[i1].y + P[2] * in[i1].z + P[3] * in[i1].w;
		int const in0 = d0 >= 0.0f, in1 = d1 >= 0.0f;
		if (in0)
		{
			out[nOut] = in[i0];
			memcpy(outa[nOut], ina[i0], sizeof(float) * MAXV);
			++nOut;
		}
		if (in0 ^ in1)
		{
			uint32_t const a = in1 ? i1 : i0, bb = in1 ? i0 : i1;
			float const t = in1 ? d1 / (d1 - d0) : d0 / (d0 - d1);
			out[nOut].x = LerpKt(in[a].x, in[bb].x, t);
			out[nOut].y = LerpKt(in[a].y, in[bb].y, t);
			out[nOut].z = LerpKt(in[a].z, in[bb].z, t);
			out[nOut].w = LerpKt(in[a].w, in[bb].w, t);
			for (uint32_t k = 0; k < MAXV; ++k) outa[nOut][k] = LerpKt(ina[a][k], ina[bb][k], t);
			++nOut;
		}
		d0 = d1;
		i0 = i1;
	}
	b->numIn = nOut;
	b->inputIdx ^= 1;
}

/* Binning.cpp:242-259 */
static void SetupEdge(srb_tile_tri* e, int idx, const int32_t a[2], const int32_t b[2])
{
	int32_t const dy = wsub(b[1], a[1]);
	int32_t const dx = wsub(a[0], b[0]);
	int64_t c = (int64_t)a[1] * (int64_t)wsub(b[0], a[0]) - (int64_t)a[0] * (int64_t)wsub(b[1], a[1]);
	if (dy < 0 || (dy == 0 && dx > 0)) c += 256;
	e->dy[idx] = dy;
	e->dx[idx] = dx;
	e->c[idx] = (int32_t)(uint32_t)(c >> 8);
}

/* Binning.cpp:261-277 */
static void SetupPlane(float K, const float d10[2], const float d20[2], float a10, float a20, float* odx, float* ody)
{
	float const A = d10[1] * a20 - a10 * d20[1];
	float const B = d20[0] * a10 - d10[0] * a20;
	*odx = -A / K;
	*ody = -B / K;
}

static void TileAppend(OTile* t, const srb_tile_tri* r)
{
	if (t->n == t->cap)
	{
		t->cap = t->cap ? t->cap * 2 : 32;
		t->tris = (srb_tile_tri*)realloc(t->tris, sizeof(srb_tile_tri) * t->cap);
	}
	t->tris[t->n++] = *r;
}

static int32_t min3(int32_t a, int32_t b, int32_t c) { a = a < b ? a : b; return a < c ? a : c; }
static int32_t max3(int32_t a, int32_t b, int32_t c) { a = a > b ? a : b; return a > c ? a : c; }

/* Binning.cpp:279-456 */
static void BinTransformedAndClippedTri(OCtx* c, const V4 v[3], const float* attr[3], const srb_draw_desc* d, uint32_t drawIdx)
{
	float const hx = (float)c->width * 0.5f, hy = (float)c->height * 0.5f;
	float iw[3], r[3][2];
	int32_t fp[3][2];
	for (int i = 0; i < 3; ++i)
	{
		iw[i] = 1.0f / v[i].w;
		r[i][0] = iw[i] * v[i].x * hx + hx;
		r[i][1] = iw[i] * v[i].y * -hy + hy;
		fp[i][0] = cvtt(r[i][0] * 256.0f + 0.5f);
		fp[i][1] = cvtt(r[i][1] * 256.0f + 0.5f);
	}
	int64_t area = (int64_t)wsub(fp[2][0], fp[0][0]) * (int64_t)wsub(fp[1][1], fp[0][1]) -
	               (int64_t)wsub(fp[2][1], fp[0][1]) * (int64_t)wsub(fp[1][0], fp[0][0]);
	area >>= 8;
	if (area <= 0) return;
	c->trisSetup++;

	int32_t const xmin = (uint16_t)clampi(wadd(min3(fp[0][0], fp[1][0], fp[2][0]), 255) >> 8, 0, (int32_t)c->width - 1);
	int32_t const ymin = (uint16_t)clampi(wadd(min3(fp[0][1], fp[1][1], fp[2][1]), 255) >> 8, 0, (int32_t)c->height - 1);
	int32_t const xmax = (uint16_t)clampi(wadd(max3(fp[0][0], fp[1][0], fp[2][0]), 255) >> 8, 0, (int32_t)c->width - 1);
	int32_t const ymax = (uint16_t)clampi(wadd(max3(fp[0][1], fp[1][1], fp[2][1]), 255) >> 8, 0, (int32_t)c->height - 1);

	srb_tile_tri edges;
	memset(&edges, 0, sizeof(edges));
	SetupEdge(&edges, 0, fp[0], fp[1]);
	SetupEdge(&edges, 1, fp[1], fp[2]);
	SetupEdge(&edges, 2, fp[2], fp[0]);

	float const d10[2] = {r[1][0] - r[0][0], r[1][1] - r[0][1]};
	float const d20[2] = {r[2][0] - r[0][0], r[2][1] - r[0][1]};
	float const K = d10[0] * d20[1] - d10[1] * d20[0];

	float zdx, zdy, wdx, wdy, adx[MAXV], ady[MAXV];
	SetupPlane(K, d10, d20, v[1].z * iw[1] - v[0].z * iw[0], v[2].z * iw[2] - v[0].z * iw[0], &zdx, &zdy);
	SetupPlane(K, d10, d20, iw[1] - iw[0], iw[2] - iw[0], &wdx, &wdy);
	uint32_t const nAttr = d->attributes.stride / 4;
	for (uint32_t i = 0; i < nAttr; ++i)
	{
		float const a10 = attr[1][i] * iw[1] - attr[0][i] * iw[0];
		float const a20 = attr[2][i] * iw[2] - attr[0][i] * iw[0];
		SetupPlane(K, d10, d20, a10, a20, &adx[i], &ady[i]);
	}

	uint32_t const binYmin = (uint32_t)ymin >> 6, binYmax = (uint32_t)ymax >> 6;
	uint32_t const binXmin = (uint32_t)xmin >> 6, binXmax = (uint32_t)xmax >> 6;
	uint32_t const numYbins = binYmax - binYmin + 1;
	uint32_t const numXbins = binYmax - binYmin + 1; /* sic: Binning.cpp:359 uses the Y range for both */
	int doCheck = 1;
	if (numYbins <= 2 && numXbins <= 2) doCheck = 0;

	for (uint32_t binY = binYmin; binY <= binYmax; ++binY)
	{
		for (uint32_t binX = binXmin; binX <= binXmax; ++binX)
		{
			int32_t const X0 = (int32_t)(binX * BIN), X1 = X0 + BIN, Y0 = (int32_t)(binY * BIN), Y1 = Y0 + BIN;
			if (doCheck)
			{
				int skip = 0;
				for (int k = 0; k < 3; ++k)
				{
					int32_t const e00 = wadd(wadd(edges.c[k], wmul(edges.dy[k], X0)), wmul(edges.dx[k], Y0));
					int32_t const e01 = wadd(wadd(edges.c[k], wmul(edges.dy[k], X0)), wmul(edges.dx[k], Y1));
					int32_t const e10 = wadd(wadd(edges.c[k], wmul(edges.dy[k], X1)), wmul(edges.dx[k], Y0));
					int32_t const e11 = wadd(wadd(edges.c[k], wmul(edges.dy[k], X1)), wmul(edges.dx[k], Y1));
					if (!((e00 > 0) | (e01 > 0) | (e10 > 0) | (e11 > 0))) skip = 1;
				}
				if (skip) continue;
			}
			srb_tile_tri o = edges;
			for (int k = 0; k < 3; ++k)
			{
				o.c[k] = wadd(o.c[k], wadd(wmul(o.dx[k], Y0), wmul(o.dy[k], X0)));
			}
			o.block_min_x = (uint8_t)clampi(xmin - X0, 0, BIN);
			o.block_max_x = (uint8_t)clampi(xmax - X0, 0, BIN);
			o.block_min_y = (uint8_t)clampi(ymin - Y0, 0, BIN);
			o.block_max_y = (uint8_t)clampi(ymax - Y0, 0, BIN);
			float const sx = (float)X0 - r[0][0], sy = (float)Y0 - r[0][1];
			o.recip_w[1] = wdx;
			o.recip_w[2] = wdy;
			o.recip_w[0] = wdx * sx + wdy * sy + iw[0];
			o.z_over_w[1] = zdx;
			o.z_over_w[2] = zdy;
			o.z_over_w[0] = zdx * sx + zdy * sy + v[0].z * iw[0];
			for (uint32_t i = 0; i < nAttr; ++i)
			{
				o.attr_dx[i] = adx[i];
				o.attr_dy[i] = ady[i];
				o.attr_c[i] = adx[i] * sx + ady[i] * sy + attr[0][i] * iw[0];
			}
			o.attribs_per_tri = nAttr;
			o.draw_idx = drawIdx;
			TileAppend(&c->tiles[binY * c->tilesX + binX], &o);
		}
	}
}

static uint32_t FetchIndex(const srb_draw_desc* d, uint32_t i)
{
	/* Binning.cpp:167-205 */
	switch (d->indices.stride)
	{
		case 1: return ((const uint8_t*)d->indices.host)[i];
		case 2: return ((const uint16_t*)d->indices.host)[i];
		default: return ((const uint32_t*)d->indices.host)[i];
	}
}

/* Binning.cpp:464-535 */
static void BinTrisEntry(OCtx* c, const srb_draw_desc* d, uint32_t drawIdx)
{
	uint32_t const numTris = d->indices.num / 3;
	const float* m = d->mvp;
	for (uint32_t t = 0; t < numTris; ++t)
	{
		V4 v[3];
		const float* attr[3];
		for (int i = 0; i < 3; ++i)
		{
			uint32_t const idx = FetchIndex(d, t * 3 + i);
			const float* p = (const float*)((const uint8_t*)d->positions.host + (size_t)idx * d->positions.stride);
			float const x = p[0], y = p[1], z = p[2], w = 1.0f;
			/* kt::Mul(Mat4, Vec4), kt/src/kt/inl/Mat4.inl:285-292 */
			v[i].x = m[0] * x + m[4] * y + m[8] * z + m[12] * w;
			v[i].y = m[1] * x + m[5] * y + m[9] * z + m[13] * w;
			v[i].z = m[2] * x + m[6] * y + m[10] * z + m[14] * w;
			v[i].w = m[3] * x + m[7] * y + m[11] * z + m[15] * w;
			attr[i] = (const float*)((const uint8_t*)d->attributes.host + (size_t)idx * d->attributes.stride);
		}
		uint32_t const c0 = ComputeClipMask(v[0]), c1 = ComputeClipMask(v[1]), c2 = ComputeClipMask(v[2]);
		uint32_t maskOr = c0 | c1 | c2;
		if (maskOr == 0)
		{
			BinTransformedAndClippedTri(c, v, attr, d, drawIdx);
			continue;
		}
		if (c0 & c1 & c2) continue;
		ClipBuffer buf;
		memset(&buf, 0, sizeof(buf));
		buf.numIn = 3;
		for (int i = 0; i < 3; ++i)
		{
			memcpy(buf.attribs[0][i], attr[i], d->attributes.stride);
			buf.verts[0][i] = v[i];
		}
		c->trisClipped++;
		do
		{
			uint32_t const plane = (uint32_t)__builtin_ctz(maskOr);
			maskOr ^= 1u << plane;
			ClipPlane(&buf, plane);
		} while (maskOr && buf.numIn);
		for (uint32_t i = 2; i < buf.numIn; ++i)
		{
			V4 f[3] = {buf.verts[buf.inputIdx][0], buf.verts[buf.inputIdx][i - 1], buf.verts[buf.inputIdx][i]};
			const float* fa[3] = {buf.attribs[buf.inputIdx][0], buf.attribs[buf.inputIdx][i - 1], buf.attribs[buf.inputIdx][i]};
			BinTransformedAndClippedTri(c, f, fa, d, drawIdx);
		}
	}
}

/* ---- back-end: raster -------------------------------------------------------------------------------------- */
/* One 8x8 block of one triangle: Rasterizer.cpp:97-192.  depthOnly selects the _DepthOnly variant. */
static uint64_t ComputeBlockMask8x8(const srb_tile_tri* t, float* depth, int32_t xB, int32_t yB, int depthOnly)
{
	uint64_t mask = 0;
	for (int32_t l = 0; l < 8; ++l)
	{
		/* zOverWPlaneSimd.tileTopLeft = fmadd(ramp, dx, c0), Rasterizer.cpp:213 */
		float const topLeft = fmaf((float)l, t->z_over_w[1], t->z_over_w[0]);
		float z = fmaf((float)yB, t->z_over_w[2], topLeft);
		z = z + ((float)xB * t->z_over_w[1]);
		int32_t e[3];
		for (int k = 0; k < 3; ++k)
		{
			/* tileTopLeftEdge = dy * ramp + c (:219) ; + yTile*dx + xTile*dy (:161) */
			e[k] = wadd(wadd(wmul(t->dy[k], l), t->c[k]), wadd(wmul(yB, t->dx[k]), wmul(xB, t->dy[k])));
		}
		float* dp = depth + xB + l + yB * BIN;
		for (int row = 0; row < 8; ++row)
		{
			int const inside = depthOnly || ((e[0] | e[1] | e[2]) >= 0);
			/* DepthCmpMask, Rasterizer.cpp:88-95: ordered compares, reverse-Z */
			if (inside && z > 0.0f && z > *dp)
			{
				*dp = z;
				mask |= 1ull << (row * 8 + l);
			}
			for (int k = 0; k < 3; ++k) e[k] = wadd(e[k], t->dx[k]);
			z = z + t->z_over_w[2];
			dp += BIN;
		}
	}
	return mask;
}

typedef void (*FragFn)(void* user, uint32_t entry, uint32_t x, uint32_t y);

/* Rasterizer.cpp:194-304 for one triangle (list entry). Returns number of fragments. */
static uint32_t RasterizeTri(const srb_tile_tri* t, float* depth, uint32_t entry, FragFn fn, void* user, uint64_t* masksOut)
{
	uint32_t nfr = 0;
	uint32_t const xBegin = t->block_min_x & ~7u, yBegin = t->block_min_y & ~7u;
	uint32_t const xEnd = t->block_max_x, yEnd = t->block_max_y;
	for (uint32_t yB = yBegin; yB < yEnd; yB += 8)
	{
		for (uint32_t xB = xBegin; xB < xEnd; xB += 8)
		{
			int32_t const X0 = (int32_t)xB, X1 = X0 + BIN, Y0 = (int32_t)yB, Y1 = Y0 + BIN; /* sic: 64-wide extent */
			uint32_t allOut[3];
			for (int k = 0; k < 3; ++k)
			{
				int32_t const e00 = wadd(wadd(t->c[k], wmul(t->dy[k], X0)), wmul(t->dx[k], Y0));
				int32_t const e01 = wadd(wadd(t->c[k], wmul(t->dy[k], X0)), wmul(t->dx[k], Y1));
				int32_t const e10 = wadd(wadd(t->c[k], wmul(t->dy[k], X1)), wmul(t->dx[k], Y0));
				int32_t const e11 = wadd(wadd(t->c[k], wmul(t->dy[k], X1)), wmul(t->dx[k], Y1));
				allOut[k] = (uint32_t)(e00 > 0) | ((uint32_t)(e01 > 0) << 1) | ((uint32_t)(e10 > 0) << 2) | ((uint32_t)(e11 > 0) << 3);
			}
			if (!allOut[0] || !allOut[1] || !allOut[2]) continue;
			int const depthOnly = allOut[0] == 0xF && allOut[1] == 0xF && allOut[2] == 0xF;
			uint64_t mask = ComputeBlockMask8x8(t, depth, (int32_t)xB, (int32_t)yB, depthOnly);
			if (masksOut) masksOut[(yB >> 3) * 8 + (xB >> 3)] = mask;
			while (mask)
			{
				uint32_t const bit = (uint32_t)__builtin_ctzll(mask);
				if (fn) fn(user, entry, (bit & 7) + xB, (bit / 8) + yB);
				++nfr;
				mask ^= 1ull << bit;
			}
		}
	}
	return nfr;
}

/* ---- back-end: sampler + shaders --------------------------------------------------------------------------- */
static uint32_t Morton5(uint32_t x, uint32_t y)
{
	uint32_t m = 0;
	for (uint32_t b = 0; b < 5; ++b)
	{
		m |= ((x >> b) & 1u) << (2 * b);
		m |= ((y >> b) & 1u) << (2 * b + 1);
	}
	return m;
}

static float LerpFma(float a, float b, float t) { return fmaf(t, b, fmaf(-t, a, a)); } /* SIMDUtil.h:141-145 */

static uint32_t PackChannel(float c)
{
	/* SIMDUtil.h:87-106 */
	int32_t i = cvtn(fmaf(c, 255.0f, 0.5f));
	i = i < 0 ? 0 : (i > 65535 ? 65535 : i);            /* _mm_packus_epi32 */
	int32_t const s = (int32_t)(int16_t)(uint16_t)i;   /* reinterpreted as signed 16 */
	return (uint32_t)(s < 0 ? 0 : (s > 255 ? 255 : s)); /* _mm_packus_epi16 */
}

static uint32_t PackRGBA(float r, float g, float b, float a)
{
	return PackChannel(r) | (PackChannel(g) << 8) | (PackChannel(b) << 16) | (PackChannel(a) << 24);
}

static void WrapCoord(float u, uint32_t dim, uint32_t* i0, uint32_t* i1, float* frac)
{
	/* Texture.cpp:410-436 */
	uint32_t const sign = fbits(u) & 0x80000000u;
	float const au = ffrom(fbits(u) ^ sign);
	float fr = au - floorf(au);
	if (sign) fr = 1.0f - fr;
	float const t = (float)dim * fr;
	float const tf = floorf(t);
	*frac = t - tf;
	*i0 = (uint32_t)cvtn(tf) & (dim - 1);
	*i1 = (*i0 + 1) & (dim - 1);
}

static void Texel(const uint8_t* p, float o[4])
{
	float const k = 1.0f / 255.0f;
	for (int i = 0; i < 4; ++i) o[i] = k * (float)p[i];
}

/* Tex::SampleWrap + pack, Texture.cpp:381-452 / :212-233 / :243-379 */
/* modulate: the RGB factors SponzaShader multiplies the sample by before packing (NULL = none) */
static uint32_t SampleWrap(const OTex* tex, float u, float v, float dudx, float dudy, float dvdx, float dvdy,
                           const float* modulate)
{
	float const Wt = (float)(1u << tex->wLog2), Ht = (float)(1u << tex->hLog2);
	float const a = dudx * Wt, b = dudy * Ht, c = dvdx * Wt, d = dvdy * Ht; /* sic: mixed axes, Texture.cpp:217-221 */
	float const du2 = fmaf(a, a, b * b), dv2 = fmaf(c, c, d * d);
	float const mx = du2 > dv2 ? du2 : dv2; /* maxps: second operand on NaN */
	float const m = sqrtf(mx);
	int32_t e = (int32_t)((fbits(m) >> 23) & 0xFFu) - 127;
	if (e < 0) e = 0;
	if (e > (int32_t)tex->numMips - 1) e = (int32_t)tex->numMips - 1;
	uint32_t const mip = (uint32_t)e;
	uint32_t const w = 1u << (tex->wLog2 - (tex->wLog2 < mip ? tex->wLog2 : mip));
	uint32_t const h = 1u << (tex->hLog2 - (tex->hLog2 < mip ? tex->hLog2 : mip));
	uint32_t x0, x1, y0, y1;
	float fu, fv;
	WrapCoord(u, w, &x0, &x1, &fu);
	WrapCoord(v, h, &y0, &y1, &fv);
	uint32_t const mtw = (w > 32 ? w : 32) >> 5;
	const uint8_t* base = tex->texels + tex->mipOffsets[mip];
#define TOFF(x, y) ((((y) >> 5) * mtw + ((x) >> 5)) * 1024u + Morton5((x)&31u, (y)&31u)) * 4u
	float t00[4], t10[4], t11[4], t01[4], out[4];
	Texel(base + TOFF(x0, y0), t00);
	Texel(base + TOFF(x1, y0), t10);
	Texel(base + TOFF(x1, y1), t11);
	Texel(base + TOFF(x0, y1), t01);
#undef TOFF
	for (int k = 0; k < 4; ++k)
	{
		float const left = LerpFma(t00[k], t01[k], fv);
		float const right = LerpFma(t10[k], t11[k], fv);
		out[k] = LerpFma(left, right, fu);
	}
	if (modulate)
	{
		/* SponzaScene.cpp:99-101 */
		out[0] = modulate[0] * out[0];
		out[1] = modulate[1] * out[1];
		out[2] = modulate[2] * out[2];
	}
	return PackRGBA(out[0], out[1], out[2], out[3]);
}

static float MaxPs(float a, float b) { return a > b ? a : b; } /* maxps: the second operand unless a > b */
static float Dot3SoA(float x0, float y0, float z0, float x1, float y1, float z1) /* SIMDUtil.h:123-126 */
{
	return fmaf(x0, x1, fmaf(y0, y1, z0 * z1));
}

/* Lighting of SponzaShader, Viewer/SponzaScene.cpp:40-93. var = interpolated position (0..2), normal (3..5). */
static void SponzaRadiance(const OCtx* c, const float* var, float* radiance)
{
	const srb_sponza_constants* k = &c->sponza;
	float const sun = MaxPs(0.1f, Dot3SoA(var[3], var[4], var[5], k->sun_dir[0], k->sun_dir[1], k->sun_dir[2]));
	radiance[0] = radiance[1] = radiance[2] = sun;
	for (int i = 0; i < SRB_SPONZA_POINT_LIGHTS; ++i)
	{
		const srb_sponza_light* L = &k->lights[i];
		float const tx = L->pos[0] - var[0], ty = L->pos[1] - var[1], tz = L->pos[2] - var[2];
		float const distSq = Dot3SoA(tx, ty, tz, tx, ty, tz);
		float const recipDist = rsqrt_x86(c, distSq);
		float const dist = rcp_x86(c, recipDist);
		float const lx = tx * recipDist, ly = ty * recipDist, lz = tz * recipDist;
		float const nDotL = MaxPs(0.0f, Dot3SoA(lx, ly, lz, var[3], var[4], var[5]));
		float const atten = rcp_x86(c, 1.0f + fmaf(0.1f, dist, distSq * 0.01f));
		float const lightRadiance = nDotL * (L->intensity * atten);
		radiance[0] = radiance[0] + lightRadiance * L->colour[0];
		radiance[1] = radiance[1] + lightRadiance * L->colour[1];
		radiance[2] = radiance[2] + lightRadiance * L->colour[2];
	}
	radiance[0] = radiance[0] + k->ambient[0];
	radiance[1] = radiance[1] + k->ambient[1];
	radiance[2] = radiance[2] + k->ambient[2];
}

/* ComputeInterpolantsDrawCallImpl (Rasterizer.cpp:356-400) + the draw's pixel shader for one fragment. */
static uint32_t ShadeFragment(const OCtx* c, const srb_tile_tri* t, uint32_t x, uint32_t y)
{
	const srb_draw_desc* d = &c->draws[t->draw_idx];
	float const fx = (float)x, fy = (float)y;
	float const W = 1.0f / fmaf(fx, t->recip_w[1], fmaf(fy, t->recip_w[2], t->recip_w[0]));
	float var[MAXV] = {0, 0, 0, 0, 0, 0, 0, 0};
	for (uint32_t i = 0; i < t->attribs_per_tri; ++i)
	{
		var[i] = W * fmaf(t->attr_dy[i], fy, fmaf(t->attr_dx[i], fx, t->attr_c[i]));
	}
	if (d->shader == SRB_SHADER_VISUALIZE_NORMALS)
	{
		return PackRGBA(fmaf(var[3], 0.5f, 0.5f), fmaf(var[4], 0.5f, 0.5f), fmaf(var[5], 0.5f, 0.5f), 1.0f);
	}
	if (d->shader == SRB_SHADER_VISUALIZE_UVS)
	{
		return PackRGBA(var[6], var[7], 0.0f, 0.0f);
	}
	if (!d->texture || c->texs[d->texture - 1].bytes == 0) return 0xFFFFFFFFu; /* Shaders.h:75-79 */
	float deriv[4] = {0, 0, 0, 0};
	uint32_t const uo = d->uv_offset;
	if (uo + 1 < MAXV)
	{
		float const fx1 = 1.0f + fx, fy1 = 1.0f + fy;
		float const W10 = rcp_x86(c, fmaf(t->recip_w[1], fx1, fmaf(t->recip_w[2], fy, t->recip_w[0])));
		float const W01 = rcp_x86(c, fmaf(t->recip_w[1], fx, fmaf(t->recip_w[2], fy1, t->recip_w[0])));
		for (uint32_t k = 0; k < 2; ++k)
		{
			uint32_t const j = uo + k;
			float const s10 = W10 * fmaf(t->attr_dx[j], fx1, fmaf(t->attr_dy[j], fy, t->attr_c[j]));
			float const s01 = W01 * fmaf(t->attr_dx[j], fx, fmaf(t->attr_dy[j], fy1, t->attr_c[j]));
			deriv[2 * k] = s10 - var[j];
			deriv[2 * k + 1] = s01 - var[j];
		}
	}
	if (d->shader == SRB_SHADER_SPONZA)
	{
		float radiance[3];
		SponzaRadiance(c, var, radiance);
		return SampleWrap(&c->texs[d->texture - 1], var[6], var[7], deriv[0], deriv[1], deriv[2], deriv[3], radiance);
	}
	return SampleWrap(&c->texs[d->texture - 1], var[6], var[7], deriv[0], deriv[1], deriv[2], deriv[3], NULL);
}

typedef struct
{
	OCtx* c;
	OTile* tile;
	uint32_t* colour;
} ShadeUser;

static void ShadeCb(void* user, uint32_t entry, uint32_t x, uint32_t y)
{
	ShadeUser* u = (ShadeUser*)user;
	/* ShadeFragmentBuffer scatter, Rasterizer.cpp:514-519: later fragments overwrite earlier ones */
	u->colour[y * BIN + x] = ShadeFragment(u->c, &u->tile->tris[entry], x, y);
}

/* Rasterizer.cpp:525-577: the list is already in (draw, triangle) order, i.e. what the stable sort yields */
static void RasterAndShadeBin(OCtx* c, uint32_t tileIdx)
{
	OTile* t = &c->tiles[tileIdx];
	ShadeUser u = {c, t, c->colour + (size_t)tileIdx * 4096};
	float* depth = c->depth + (size_t)tileIdx * 4096;
	for (uint32_t i = 0; i < t->n; ++i)
	{
		RasterizeTri(&t->tris[i], depth, i, ShadeCb, &u, NULL);
	}
}

/* ---- C ABI ------------------------------------------------------------------------------------------------- */
SRB_API void* sro_create(uint32_t width, uint32_t height)
{
	OCtx* c = (OCtx*)calloc(1, sizeof(OCtx));
	c->width = width;
	c->height = height;
	c->tilesX = (width + BIN - 1) / BIN;
	c->tilesY = (height + BIN - 1) / BIN;
	size_t const n = (size_t)c->tilesX * c->tilesY;
	c->colour = (uint32_t*)calloc(n * 4096, 4);
	c->depth = (float*)calloc(n * 4096, 4);
	c->tiles = (OTile*)calloc(n, sizeof(OTile));
	return c;
}

SRB_API void sro_destroy(void* h)
{
	OCtx* c = (OCtx*)h;
	if (!c) return;
	size_t const n = (size_t)c->tilesX * c->tilesY;
	for (size_t i = 0; i < n; ++i) free(c->tiles[i].tris);
	for (uint32_t i = 0; i < c->numTex; ++i) free((void*)c->texs[i].texels);
	free(c->tiles);
	free(c->texs);
	free(c->draws);
	free(c->colour);
	free(c->depth);
	free(c);
}

SRB_API int sro_set_rcp_table(void* h, const uint32_t* table, uint32_t bits)
{
	OCtx* c = (OCtx*)h;
	if (bits < 1 || bits > 16) return SRB_ERR_INVALID;
	memcpy(c->rcpTable, table, sizeof(uint32_t) << bits);
	c->rcpBits = bits;
	return SRB_OK;
}

SRB_API uint64_t sro_texture_create(void* h, const uint8_t* texels, uint64_t bytes, const uint32_t* mipOffsets,
                                    uint32_t numMips, uint32_t wLog2, uint32_t hLog2)
{
	OCtx* c = (OCtx*)h;
	if (c->numTex == c->capTex)
	{
		c->capTex = c->capTex ? c->capTex * 2 : 8;
		c->texs = (OTex*)realloc(c->texs, sizeof(OTex) * c->capTex);
	}
	OTex* t = &c->texs[c->numTex++];
	memset(t, 0, sizeof(*t));
	uint8_t* copy = (uint8_t*)malloc(bytes ? bytes : 1);
	if (bytes) memcpy(copy, texels, bytes);
	t->texels = copy;
	t->bytes = bytes;
	for (uint32_t i = 0; i < numMips && i < 14; ++i) t->mipOffsets[i] = mipOffsets[i];
	t->numMips = numMips;
	t->wLog2 = wLog2;
	t->hLog2 = hLog2;
	return c->numTex;
}

SRB_API int sro_begin_frame(void* h)
{
	OCtx* c = (OCtx*)h;
	size_t const n = (size_t)c->tilesX * c->tilesY;
	for (size_t i = 0; i < n; ++i) c->tiles[i].n = 0;
	c->numDraws = 0;
	c->trisSetup = c->trisClipped = 0;
	return SRB_OK;
}

/* RenderContext::ClearFrameBuffer, Renderer.cpp:168-194 */
SRB_API int sro_clear(void* h, uint32_t color, int clearColour, int clearDepth)
{
	OCtx* c = (OCtx*)h;
	size_t const n = (size_t)c->tilesX * c->tilesY * 4096;
	if (clearDepth)
	{
		for (size_t i = 0; i < n; ++i) c->depth[i] = 0.0f;
	}
	if (clearColour) memset(c->colour, (int)color, n * 4);
	return SRB_OK;
}

SRB_API int sro_draw_indexed(void* h, const srb_draw_desc* d)
{
	OCtx* c = (OCtx*)h;
	if (d->shader >= SRB_SHADER_COUNT) return SRB_ERR_UNKNOWN_SHADER;
	if (c->numDraws == c->capDraws)
	{
		c->capDraws = c->capDraws ? c->capDraws * 2 : 32;
		c->draws = (srb_draw_desc*)realloc(c->draws, sizeof(srb_draw_desc) * c->capDraws);
	}
	c->draws[c->numDraws++] = *d;
	return SRB_OK;
}

/* RenderContext::EndFrame, Renderer.cpp:209-317, without the task system */
SRB_API int sro_end_frame(void* h)
{
	OCtx* c = (OCtx*)h;
	for (uint32_t i = 0; i < c->numDraws; ++i) BinTrisEntry(c, &c->draws[i], i);
	uint32_t const n = c->tilesX * c->tilesY;
	for (uint32_t t = 0; t < n; ++t)
	{
		if (c->tiles[t].n) RasterAndShadeBin(c, t);
	}
	return SRB_OK;
}

SRB_API int sro_render_frames(void* h, const srb_draw_desc* draws, uint32_t nDraws, const float* mvps, uint32_t frames,
                              uint32_t clearColor, double* msOut)
{
	for (uint32_t f = 0; f < frames; ++f)
	{
		struct timespec t0, t1;
		clock_gettime(CLOCK_MONOTONIC, &t0);
		sro_begin_frame(h);
		sro_clear(h, clearColor, 1, 1);
		for (uint32_t i = 0; i < nDraws; ++i)
		{
			srb_draw_desc d = draws[i];
			if (mvps) memcpy(d.mvp, mvps + ((size_t)f * nDraws + i) * 16, 64);
			sro_draw_indexed(h, &d);
		}
		sro_end_frame(h);
		clock_gettime(CLOCK_MONOTONIC, &t1);
		if (msOut) msOut[f] = (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
	}
	return SRB_OK;
}

SRB_API int sro_read_tiles(void* h, void* colour, void* depth, uint64_t depthStride)
{
	OCtx* c = (OCtx*)h;
	size_t const n = (size_t)c->tilesX * c->tilesY;
	if (colour) memcpy(colour, c->colour, n * 16384);
	if (depth)
	{
		for (size_t i = 0; i < n; ++i) memcpy((uint8_t*)depth + i * depthStride, c->depth + i * 4096, 16384);
	}
	return SRB_OK;
}

SRB_API int sro_get_counters(void* h, srb_counters* out)
{
	OCtx* c = (OCtx*)h;
	memset(out, 0, sizeof(*out));
	out->tris_setup = c->trisSetup;
	out->tris_clipped = c->trisClipped;
	uint32_t const n = c->tilesX * c->tilesY;
	for (uint32_t t = 0; t < n; ++t)
	{
		out->tile_refs += c->tiles[t].n;
		out->tiles_nonempty += c->tiles[t].n != 0;
		if (c->tiles[t].n > out->max_refs_in_tile) out->max_refs_in_tile = c->tiles[t].n;
	}
	for (uint32_t i = 0; i < c->numDraws; ++i) out->tris_in += c->draws[i].indices.num / 3;
	return SRB_OK;
}

SRB_API int sro_dump_tile_counts(void* h, uint32_t* counts, uint32_t numTiles)
{
	OCtx* c = (OCtx*)h;
	if (numTiles != c->tilesX * c->tilesY) return SRB_ERR_INVALID;
	for (uint32_t t = 0; t < numTiles; ++t) counts[t] = c->tiles[t].n;
	return SRB_OK;
}

SRB_API int sro_dump_tile_tris(void* h, uint32_t tile, srb_tile_tri* out, uint32_t cap, uint32_t* n)
{
	OCtx* c = (OCtx*)h;
	OTile* t = &c->tiles[tile];
	*n = t->n;
	memcpy(out, t->tris, sizeof(srb_tile_tri) * (t->n < cap ? t->n : cap));
	return t->n <= cap ? SRB_OK : SRB_ERR_OVERFLOW;
}

SRB_API int sro_dump_tile_coverage(void* h, uint32_t tile, uint64_t* masks, uint32_t capEntries, uint32_t* n)
{
	OCtx* c = (OCtx*)h;
	OTile* t = &c->tiles[tile];
	float* depth = (float*)malloc(4096 * 4);
	*n = t->n;
	for (uint32_t i = 0; i < t->n && i < capEntries; ++i)
	{
		memset(depth, 0, 4096 * 4);
		memset(masks + (size_t)i * 64, 0, 64 * 8);
		RasterizeTri(&t->tris[i], depth, i, NULL, NULL, masks + (size_t)i * 64);
	}
	free(depth);
	return t->n <= capEntries ? SRB_OK : SRB_ERR_OVERFLOW;
}

typedef struct
{
	uint32_t* frags;
	uint64_t cap, n;
} FragUser;

static void FragCb(void* user, uint32_t entry, uint32_t x, uint32_t y)
{
	FragUser* u = (FragUser*)user;
	if (u->n < u->cap) u->frags[u->n] = (entry << 12) | (y << 6) | x;
	u->n++;
}

SRB_API int sro_dump_tile_fragments(void* h, uint32_t tile, uint32_t* frags, uint64_t cap, uint64_t* n, float* depthOut)
{
	OCtx* c = (OCtx*)h;
	OTile* t = &c->tiles[tile];
	float* depth = (float*)calloc(4096, 4);
	FragUser u = {frags, cap, 0};
	for (uint32_t i = 0; i < t->n; ++i) RasterizeTri(&t->tris[i], depth, i, FragCb, &u, NULL);
	if (depthOut) memcpy(depthOut, depth, 4096 * 4);
	free(depth);
	*n = u.n;
	return u.n <= cap ? SRB_OK : SRB_ERR_OVERFLOW;
}

SRB_API int sro_sample(void* h, uint64_t tex, const float* u, const float* v, const float* dudx, const float* dudy,
                       const float* dvdx, const float* dvdy, uint32_t* rgba, uint64_t n)
{
	OCtx* c = (OCtx*)h;
	if (!tex || tex > c->numTex) return SRB_ERR_INVALID;
	for (uint64_t i = 0; i < n; ++i)
	{
		rgba[i] = SampleWrap(&c->texs[tex - 1], u[i], v[i], dudx[i], dudy[i], dvdx[i], dvdy[i], NULL);
	}
	return SRB_OK;
}

SRB_API int sro_set_rsqrt_table(void* h, const uint32_t* table, uint32_t bits)
{
	OCtx* c = (OCtx*)h;
	if (!c || !table || bits < 1 || bits > 16) return SRB_ERR_INVALID;
	memcpy(c->rsqrtTable, table, sizeof(uint32_t) * ((size_t)2 << bits));
	c->rsqrtBits = bits;
	return SRB_OK;
}

SRB_API int sro_set_sponza_constants(void* h, const srb_sponza_constants* k)
{
	if (!h || !k) return SRB_ERR_INVALID;
	((OCtx*)h)->sponza = *k;
	return SRB_OK;
}

SRB_API void sro_rsqrt(void* h, const float* in, float* out, uint64_t n)
{
	for (uint64_t i = 0; i < n; ++i) out[i] = rsqrt_x86((OCtx*)h, in[i]);
}

SRB_API void sro_rcp(void* h, const float* in, float* out, uint64_t n)
{
	for (uint64_t i = 0; i < n; ++i) out[i] = rcp_x86((OCtx*)h, in[i]);
}
