// srb_setup.cu — K1: vertex transform, frustum clipping, triangle set-up, per-tile reference counting and — in its
// tail — the per-tile scan of K2.
//
// Replaces the reference front-end BinTrisEntry + BinTransformedAndClippedTri (SoftRast/Binning.cpp:464-535, :279-456)
// up to, but not including, the per-bin append (that is bin_fill_kernel, srb_bin.cu).
//
//   main loop  : one thread per INPUT triangle over all draws of the frame.  Transform, clip codes, trivial
//                accept/reject.  Unclipped front-facing triangles are set up in place: records are written at
//                slot = input triangle index (no allocation, no ordering dependency between threads).
//                Triangles that cross a frustum plane (a few %, in runs along the frustum's edges) are only QUEUED, so
//                no warp ever serialises behind the clipper.
//   clip_scan_kernel : eight lanes per queued triangle: all eight clip it (Sutherland-Hodgman, same data, same path),
//                then lane i culls and sets up fan triangle i in slots handed out beyond numInputTris.  The LAST CTA to
//                finish — every count is in the global counters by then — runs the tile scan of K2 (srb_scan.cuh) in
//                the kernel's tail: offsets, work units, counters re-zeroed; no launch of its own.
// Draw order is carried by the canonical key (srb_device.cuh), not by where a record is stored.
#include "srb_device.cuh"
#include "srb_kernels.h"
#include "srb_scan.cuh"

#include <algorithm>
#include <stdlib.h>

namespace srb
{

namespace
{

constexpr int kSetupThreads = 256;
#ifndef SRB_DIRECT_THREADS
#define SRB_DIRECT_THREADS 256
#endif
constexpr int kDirectThreads = SRB_DIRECT_THREADS; // the one-triangle-per-thread kernel's CTA
constexpr int kMaxClipVerts = 9; // Binning.cpp:71: 3 + one per frustum plane

__device__ __forceinline__ uint32_t clip_code(float x, float y, float z, float w)
{
	// Binning.cpp:56-68
	uint32_t m = 0;
	if (addf(x, w) < 0.0f) m |= 1u;
	if (subf(x, w) > 0.0f) m |= 2u;
	if (addf(y, w) < 0.0f) m |= 4u;
	if (subf(y, w) > 0.0f) m |= 8u;
	if (z < 0.0f) m |= 16u;
	if (subf(z, w) > 0.0f) m |= 32u;
	return m;
}

// kt::Lerp (kt/src/kt/inl/MathUtil.inl:7-11): (1 - t) * a + t * b
__device__ __forceinline__ float lerp_kt(float a, float b, float t)
{
	return addf(mulf(subf(1.0f, t), a), mulf(t, b));
}

struct Snapped
{
	float rx[3], ry[3], iw[3], zw[3]; // raster x, y, 1/w, z/w
	int32_t fx[3], fy[3];
};

// Viewport transform + 24.8 snap, Binning.cpp:291-303.  1.0f / w is the correctly rounded reciprocal.
__device__ __forceinline__ void snap(const float4 (&v)[3], float hx, float hy, Snapped& s)
{
#pragma unroll
	for (int i = 0; i < 3; ++i)
	{
		s.iw[i] = __frcp_rn(v[i].w);
		s.zw[i] = mulf(v[i].z, s.iw[i]); // Binning.cpp:336: z/w per vertex
		s.rx[i] = addf(mulf(mulf(s.iw[i], v[i].x), hx), hx);
		s.ry[i] = addf(mulf(mulf(s.iw[i], v[i].y), -hy), hy);
		s.fx[i] = cvtt_x86(addf(mulf(s.rx[i], 256.0f), 0.5f));
		s.fy[i] = cvtt_x86(addf(mulf(s.ry[i], 256.0f), 0.5f));
	}
}

// Binning.cpp:305-311: twice the signed area in 24.8, >> 8; <= 0 is culled (back-facing or degenerate).
__device__ __forceinline__ bool front_facing(const Snapped& s)
{
	int64_t a = (int64_t)wrap_sub(s.fx[2], s.fx[0]) * (int64_t)wrap_sub(s.fy[1], s.fy[0]) -
	            (int64_t)wrap_sub(s.fy[2], s.fy[0]) * (int64_t)wrap_sub(s.fx[1], s.fx[0]);
	a >>= 8;
	return a > 0;
}

// SetupEdge, Binning.cpp:242-259.
__device__ __forceinline__ void setup_edge(int32_t ax, int32_t ay, int32_t bx, int32_t by, int32_t& c, int32_t& dx,
                                           int32_t& dy)
{
	dy = wrap_sub(by, ay);
	dx = wrap_sub(ax, bx);
	int64_t cc = (int64_t)ay * (int64_t)wrap_sub(bx, ax) - (int64_t)ax * (int64_t)wrap_sub(by, ay);
	if (dy < 0 || (dy == 0 && dx > 0))
	{
		cc += 256;
	}
	c = (int32_t)(uint32_t)(cc >> 8);
}

// SetupPlane, Binning.cpp:261-277.
__device__ __forceinline__ void setup_plane(float K, float d10x, float d10y, float d20x, float d20y, float a10, float a20,
                                            float& odx, float& ody)
{
	float const A = subf(mulf(d10y, a20), mulf(a10, d20y));
	float const B = subf(mulf(d20x, a10), mulf(d10x, a20));
	odx = divf(-A, K);
	ody = divf(-B, K);
}

__device__ __forceinline__ int32_t min3(int32_t a, int32_t b, int32_t c) { return min(min(a, b), c); }
__device__ __forceinline__ int32_t max3(int32_t a, int32_t b, int32_t c) { return max(max(a, b), c); }

struct SetupArgs
{
	FrameParams fp;
	const DrawDev* draws;
	RasterRec* rasterRecs;
	ShadeRec* shadeRecs;
	Survivor* survivors;
	uint32_t* clipQueue; // [numInputTris] input triangles that cross a frustum plane
	uint32_t* tileCounts;
	uint32_t* offsets;
	uint32_t* cursors;
	UnitDesc* units;
	FrameCtl* ctl;
	uint32_t fuseScan; // the last CTA runs the tile scan (else: tile_scan_kernel is launched after this kernel)
	uint32_t* releaseFlag; // screen-tile split, root GPU only: stamped with (this frame - 1) when the frame begins
};

// Screen-tile split across GPUs: does the bin range hold a tile this context owns (tile % ownMod == ownRem)?
__device__ __forceinline__ bool range_touches_owned(const FrameParams& fp, const BinRange& br)
{
	if (br.bx1 - br.bx0 + 1u >= fp.ownMod)
	{
		return true; // ownMod consecutive tile indices hold every remainder
	}
	for (uint32_t by = br.by0; by <= br.by1; ++by)
	{
		uint32_t const first = (by * fp.tilesX + br.bx0) % fp.ownMod;
		// tiles first .. first + (bx1 - bx0) modulo ownMod: is ownRem among them?
		uint32_t const d = (fp.ownRem + fp.ownMod - first) % fp.ownMod;
		if (d <= br.bx1 - br.bx0)
		{
			return true;
		}
	}
	return false;
}

// Full set-up of one surviving triangle (Binning.cpp:313-350) + tile reference counting (:352-410).
// Returns false when the triangle touches no tile of this context (screen-tile split): nothing is written then.
// s_hist: this CTA's shared-memory tile histogram, or nullptr (counts go straight to the global counters).
__device__ __forceinline__ bool emit_triangle(const Snapped& s, const float* a0, const float* a1,
                                              const float* a2, const DrawDev& draw, uint32_t drawIdx, const FrameParams& fp,
                                              uint32_t slot, RasterRec* __restrict__ rasterRecs,
                                              ShadeRec* __restrict__ shadeRecs, uint32_t* s_hist,
                                              uint32_t* __restrict__ tileCounts, uint2& oneTile)
{
	int32_t const W1 = (int32_t)fp.width - 1, H1 = (int32_t)fp.height - 1;
	uint32_t const xmin = (uint32_t)clampi(wrap_add(min3(s.fx[0], s.fx[1], s.fx[2]), 255) >> 8, 0, W1);
	uint32_t const ymin = (uint32_t)clampi(wrap_add(min3(s.fy[0], s.fy[1], s.fy[2]), 255) >> 8, 0, H1);
	uint32_t const xmax = (uint32_t)clampi(wrap_add(max3(s.fx[0], s.fx[1], s.fx[2]), 255) >> 8, 0, W1);
	uint32_t const ymax = (uint32_t)clampi(wrap_add(max3(s.fy[0], s.fy[1], s.fy[2]), 255) >> 8, 0, H1);
	BinRange const br = bin_range(xmin, xmax, ymin, ymax);
	if (fp.ownMod > 1u && !range_touches_owned(fp, br))
	{
		return false; // another GPU's triangle: skip the expensive half of the set-up
	}
	oneTile = make_uint2(0u, 0u);
	if (br.bx0 == br.bx1 && br.by0 == br.by1)
	{
		// one tile (one bin row: the reference appends without the overlap test, Binning.cpp:358-370): the survivor entry
		// carries the tile and the block range inside it (Binning.cpp:429-433), and the bin fill needs nothing else
		int32_t const X0 = (int32_t)(br.bx0 * SRB_TILE), Y0 = (int32_t)(br.by0 * SRB_TILE);
		oneTile.x = 0x80000000u | (br.by0 * fp.tilesX + br.bx0);
		oneTile.y = pack_block_range(clampi((int32_t)xmin - X0, 0, SRB_TILE), clampi((int32_t)xmax - X0, 0, SRB_TILE),
		                             clampi((int32_t)ymin - Y0, 0, SRB_TILE), clampi((int32_t)ymax - Y0, 0, SRB_TILE));
	}

	int32_t c[3], dx[3], dy[3];
	setup_edge(s.fx[0], s.fy[0], s.fx[1], s.fy[1], c[0], dx[0], dy[0]);
	setup_edge(s.fx[1], s.fy[1], s.fx[2], s.fy[2], c[1], dx[1], dy[1]);
	setup_edge(s.fx[2], s.fy[2], s.fx[0], s.fy[0], c[2], dx[2], dy[2]);

	float const d10x = subf(s.rx[1], s.rx[0]), d10y = subf(s.ry[1], s.ry[0]);
	float const d20x = subf(s.rx[2], s.rx[0]), d20y = subf(s.ry[2], s.ry[0]);
	float const K = subf(mulf(d10x, d20y), mulf(d10y, d20x));

	float const zw0 = s.zw[0];
	float zdx, zdy;
	setup_plane(K, d10x, d10y, d20x, d20y, subf(s.zw[1], zw0), subf(s.zw[2], zw0), zdx, zdy);
	{
		// RasterRec, 64 bytes, assembled in registers (srb_device.cuh)
		uint4* dr = reinterpret_cast<uint4*>(rasterRecs + slot);
		dr[0] = make_uint4((uint32_t)c[0], (uint32_t)c[1], (uint32_t)c[2], (uint32_t)dx[0]);
		dr[1] = make_uint4((uint32_t)dx[1], (uint32_t)dx[2], (uint32_t)dy[0], (uint32_t)dy[1]);
		dr[2] = make_uint4((uint32_t)dy[2], xmin | (xmax << 16), ymin | (ymax << 16), __float_as_uint(zdx));
		dr[3] = make_uint4(__float_as_uint(zdy), __float_as_uint(zw0), __float_as_uint(s.rx[0]), __float_as_uint(s.ry[0]));
	}

	float wdx, wdy;
	setup_plane(K, d10x, d10y, d20x, d20y, subf(s.iw[1], s.iw[0]), subf(s.iw[2], s.iw[0]), wdx, wdy);
	uint32_t const info = (draw.shader & 0xFFu) | (min(draw.uvOffset, 255u) << 8) | ((uint32_t)(draw.texture + 1) << 16);
	{
		// ShadeRec, 128 bytes (srb_device.cuh): head, then the planes in SLOT order — slot s holds varying (s + 6) & 7, i.e.
		// 6, 7, 0 .. 5 (SRB_PLANE_SLOT) — written out 16 bytes at a time as they are completed, so that few values are live.
		// The textured shaders read only the first 64 bytes; the second half is written only when a shader reads it.
		uint4* ds = reinterpret_cast<uint4*>(shadeRecs + slot);
		ds[0] = make_uint4(__float_as_uint(wdx), __float_as_uint(wdy), __float_as_uint(s.iw[0]), info);
		ds[1] = make_uint4(__float_as_uint(s.rx[0]), __float_as_uint(s.ry[0]), drawIdx, 0u);
		bool const secondHalf = (draw.planeMask & 0x100u) != 0u;
		// plane of varying i: (dx, dy, vertex-0 value), zeros when no shader reads it
		auto plane = [&](int i) -> float3 {
			float3 p = make_float3(0.0f, 0.0f, 0.0f);
			if ((draw.planeMask >> i) & 1u)
			{
				float const q0 = mulf(a0[i], s.iw[0]);
				setup_plane(K, d10x, d10y, d20x, d20y, subf(mulf(a1[i], s.iw[1]), q0), subf(mulf(a2[i], s.iw[2]), q0), p.x, p.y);
				p.z = q0;
			}
			return p;
		};
		auto put = [&](int q, float a, float b, float c, float d) {
			ds[q] = make_uint4(__float_as_uint(a), __float_as_uint(b), __float_as_uint(c), __float_as_uint(d));
		};
		float3 const p6 = plane(6), p7 = plane(7);
		put(2, p6.x, p6.y, p6.z, p7.x);
		float3 const p0 = plane(0);
		put(3, p7.y, p7.z, p0.x, p0.y);
		if (secondHalf)
		{
			float3 const p1 = plane(1);
			put(4, p0.z, p1.x, p1.y, p1.z);
			float3 const p2 = plane(2), p3 = plane(3);
			put(5, p2.x, p2.y, p2.z, p3.x);
			float3 const p4 = plane(4);
			put(6, p3.y, p3.z, p4.x, p4.y);
			float3 const p5 = plane(5);
			put(7, p4.z, p5.x, p5.y, p5.z);
		}
	}

	// count the tiles this triangle will be appended to
	for (uint32_t by = br.by0; by <= br.by1; ++by)
	{
		for (uint32_t bx = br.bx0; bx <= br.bx1; ++bx)
		{
			if (br.check && !bin_overlaps(c, dx, dy, (int32_t)(bx * SRB_TILE), (int32_t)(by * SRB_TILE)))
			{
				continue;
			}
			uint32_t const tile = by * fp.tilesX + bx;
			if (s_hist)
			{
				atomicAdd(&s_hist[tile], 1u); // tiles of other GPUs are dropped when the histogram is flushed
			}
			else if (tile_owned(fp, tile))
			{
				atomicAdd(&tileCounts[tile], 1u);
			}
		}
	}
	return true;
}

__device__ __forceinline__ uint32_t fetch_index(const DrawDev& d, uint32_t i)
{
	// Binning.cpp:167-205
	switch (d.idxStride)
	{
		case 1: return d.idx[i];
		case 2: return reinterpret_cast<const uint16_t*>(d.idx)[i];
		default: return reinterpret_cast<const uint32_t*>(d.idx)[i];
	}
}

// kt::Mul(Mat4, Vec4) (kt/src/kt/inl/Mat4.inl:285-292): ((c0*x + c1*y) + c2*z) + c3*w with w = 1
__device__ __forceinline__ float4 transform(const DrawDev& d, const float* p)
{
	float const x = p[0], y = p[1], z = p[2];
	float r[4];
#pragma unroll
	for (int k = 0; k < 4; ++k)
	{
		r[k] = addf(addf(addf(mulf(d.mvp[k], x), mulf(d.mvp[4 + k], y)), mulf(d.mvp[8 + k], z)), mulf(d.mvp[12 + k], 1.0f));
	}
	return make_float4(r[0], r[1], r[2], r[3]);
}

// last d with triBase[d] <= g; the table is in shared memory when it fits, else the draw table itself is searched
__device__ __forceinline__ uint32_t find_draw(const uint32_t* s_triBase, const DrawDev* __restrict__ draws, uint32_t numDraws,
                                              uint32_t g)
{
	uint32_t lo = 0, hi = numDraws;
	while (hi - lo > 1)
	{
		uint32_t const mid = (lo + hi) >> 1;
		uint32_t const b = s_triBase ? s_triBase[mid] : __ldg(&draws[mid].triBase);
		if (b <= g) lo = mid; else hi = mid;
	}
	return lo;
}

constexpr int kClipThreads = 256;

// Clip pass (Binning.cpp:498-533) + the tile scan in the tail.
//
// SIXTEEN lanes share one queued triangle and the polygon lives in shared memory, ONE LANE PER VERTEX: for every frustum
// plane of the triangle's OR-mask (ascending, like the reference) lane i handles the polygon edge (i - 1 -> i) of
// Sutherland-Hodgman — it emits the edge's start vertex if that is inside and the intersection if the edge crosses the
// plane (Binning.cpp:85-165: the same two outputs in the same order) — and the output positions come from a ballot, so a
// plane costs a few dozen instructions instead of a serial walk over up to nine vertices with twelve interpolated
// values each.  Then lane i culls and sets up fan triangle (0, i + 1, i + 2), the <= 7 set-ups side by side.
constexpr int kClipLanes = 16;
constexpr int kClipFloats = 4 + SRB_MAX_VARY; // x y z w + attributes
static_assert(kClipFloats == 12, "a clip vertex is three float4");

__device__ __forceinline__ float plane_dot4(uint32_t plane, float4 v)
{
	// kt::Dot(plane, v) (Vec4.inl:162-165) for the six planes of Binning.cpp:87-97: ((px*x + py*y) + pz*z) + 1*w
	float px = 0.0f, py = 0.0f, pz = 0.0f;
	switch (plane)
	{
		case 0: px = 1.0f; break;
		case 1: px = -1.0f; break;
		case 2: py = 1.0f; break;
		case 3: py = -1.0f; break;
		case 4: pz = 1.0f; break;
		default: pz = -1.0f; break;
	}
	return addf(addf(addf(mulf(px, v.x), mulf(py, v.y)), mulf(pz, v.z)), mulf(1.0f, v.w));
}

__device__ __forceinline__ float4 lerp4(float4 a, float4 b, float t)
{
	return make_float4(lerp_kt(a.x, b.x, t), lerp_kt(a.y, b.y, t), lerp_kt(a.z, b.z, t), lerp_kt(a.w, b.w, t));
}

__global__ void __launch_bounds__(kClipThreads) clip_scan_kernel(const __grid_constant__ SetupArgs A)
{
	extern __shared__ uint32_t s_dyn[]; // [numTiles] counts for the scan (if they fit), then [numDraws] triBase table (if it fits)
	__shared__ uint32_t s_isLast;
	// the polygons of this CTA's 16 lane-groups: two buffers (in / out of a plane) of nine vertices of three float4
	__shared__ float4 s_poly[kClipThreads / kClipLanes][2][kMaxClipVerts][3];
	const FrameParams& fp = A.fp;
	uint32_t const numTiles = fp.tilesX * fp.tilesY;
	uint32_t* const s_counts = fp.smemHist ? s_dyn : nullptr;
	uint32_t* const s_triBase = fp.smemBase ? s_dyn + (fp.smemHist ? numTiles : 0u) : nullptr;
	uint32_t* const s_hist = nullptr; // tile references of clipped triangles are counted straight in the global counters
	uint32_t const n = A.ctl->numClipQueue;
	if (A.releaseFlag && blockIdx.x == 0 && threadIdx.x == 0)
	{
		*reinterpret_cast<volatile uint32_t*>(A.releaseFlag) = A.ctl->doneValue - 1u; // (frames without triangles have no set-up kernel)
	}
	constexpr uint32_t kGroups = kClipThreads / kClipLanes;
	if (s_triBase && blockIdx.x * kGroups < n)
	{
		for (uint32_t i = threadIdx.x; i < fp.numDraws; i += kClipThreads) s_triBase[i] = A.draws[i].triBase;
		__syncthreads();
	}
	float const hx = mulf((float)fp.width, 0.5f), hy = mulf((float)fp.height, 0.5f);
	uint32_t const lane = threadIdx.x & 31u, sub = lane & (kClipLanes - 1u), grpShift = lane & 16u;
	uint32_t const grpMask = 0xFFFFu << grpShift, below = (1u << sub) - 1u;
	float4 (*const poly)[kMaxClipVerts][3] = s_poly[threadIdx.x / kClipLanes];
	uint32_t const groupsPerGrid = gridDim.x * kGroups;
	for (uint32_t base = blockIdx.x * kGroups + (threadIdx.x >> 5) * 2u; base < n; base += groupsPerGrid)
	{
		// (a warp's two groups take two consecutive entries of one iteration, so the whole warp runs the same trip count)
		uint32_t const q = base + (lane >> 4);
		bool const have = q < n;
		uint32_t g = 0, drawIdx = 0, code = 0;
		if (have)
		{
			g = A.clipQueue[q];
			drawIdx = find_draw(s_triBase, A.draws, fp.numDraws, g);
			if (sub < 3u)
			{
				// lane i fetches and transforms vertex i (Binning.cpp:475-485)
				const DrawDev& d = A.draws[drawIdx];
				uint32_t const idx = fetch_index(d, (g - d.triBase) * 3u + sub);
				float4 const v = transform(d, reinterpret_cast<const float*>(d.pos + (size_t)idx * d.posStride));
				const float* ap = reinterpret_cast<const float*>(d.attr + (size_t)idx * d.attrStride);
				float a[SRB_MAX_VARY];
#pragma unroll
				for (int k = 0; k < SRB_MAX_VARY; ++k)
				{
					a[k] = ((uint32_t)k < d.numVaryings) ? ap[k] : 0.0f;
				}
				poly[0][sub][0] = v;
				poly[0][sub][1] = make_float4(a[0], a[1], a[2], a[3]);
				poly[0][sub][2] = make_float4(a[4], a[5], a[6], a[7]);
				code = clip_code(v.x, v.y, v.z, v.w);
			}
		}
		uint32_t maskOr = __reduce_or_sync(0xFFFFFFFFu, code << grpShift) >> grpShift & 0x3Fu; // the group's three codes
		__syncwarp();
		// Binning.cpp:498-523: one plane after the other, ascending bit order, while vertices are left
		uint32_t nVerts = have ? 3u : 0u, src = 0;
		while (__any_sync(0xFFFFFFFFu, maskOr != 0u && nVerts != 0u))
		{
			bool const active = maskOr != 0u && nVerts != 0u;
			uint32_t const plane = active ? (uint32_t)__ffs(maskOr) - 1u : 0u;
			bool const mine = active && sub < nVerts;
			bool inPrev = false, crosses = false;
			float4 c0, c1, c2, x0, x1, x2; // the edge's start vertex; the intersection
			c0 = c1 = c2 = x0 = x1 = x2 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
			if (mine)
			{
				uint32_t const prev = sub == 0u ? nVerts - 1u : sub - 1u;
				float4 const p0 = poly[src][prev][0], q0 = poly[src][sub][0];
				float const dPrev = plane_dot4(plane, p0), dCur = plane_dot4(plane, q0);
				inPrev = dPrev >= 0.0f;
				bool const inCur = dCur >= 0.0f;
				crosses = inPrev != inCur;
				if (inPrev)
				{
					c0 = p0;
					c1 = poly[src][prev][1];
					c2 = poly[src][prev][2];
				}
				if (crosses)
				{
					// the inside vertex is always the first Lerp argument (Binning.cpp:129-155)
					float const t = inCur ? divf(dCur, subf(dCur, dPrev)) : divf(dPrev, subf(dPrev, dCur));
					uint32_t const ia = inCur ? sub : prev, ib = inCur ? prev : sub;
					x0 = lerp4(poly[src][ia][0], poly[src][ib][0], t);
					x1 = lerp4(poly[src][ia][1], poly[src][ib][1], t);
					x2 = lerp4(poly[src][ia][2], poly[src][ib][2], t);
				}
			}
			uint32_t const keepMask = (__ballot_sync(0xFFFFFFFFu, inPrev) & grpMask) >> grpShift;
			uint32_t const crossMask = (__ballot_sync(0xFFFFFFFFu, crosses) & grpMask) >> grpShift;
			__syncwarp(); // every lane has read its inputs
			if (mine)
			{
				uint32_t pos = (uint32_t)__popc(keepMask & below) + (uint32_t)__popc(crossMask & below);
				if (inPrev && pos < (uint32_t)kMaxClipVerts)
				{
					poly[src ^ 1u][pos][0] = c0;
					poly[src ^ 1u][pos][1] = c1;
					poly[src ^ 1u][pos][2] = c2;
					++pos;
				}
				if (crosses && pos < (uint32_t)kMaxClipVerts)
				{
					poly[src ^ 1u][pos][0] = x0;
					poly[src ^ 1u][pos][1] = x1;
					poly[src ^ 1u][pos][2] = x2;
				}
			}
			if (active)
			{
				maskOr ^= 1u << plane;
				nVerts = min((uint32_t)kMaxClipVerts, (uint32_t)__popc(keepMask) + (uint32_t)__popc(crossMask));
				src ^= 1u;
			}
			__syncwarp();
		}
		// fan (0, i-1, i), Binning.cpp:526-533: lane `sub` owns fan triangle i = sub + 2; which ones survive the cull?
		uint32_t const i = sub + 2u;
		bool mine = have && i < nVerts;
		float4 f[3];
		Snapped sn;
		if (mine)
		{
			f[0] = poly[src][0][0];
			f[1] = poly[src][i - 1u][0];
			f[2] = poly[src][i][0];
			snap(f, hx, hy, sn);
			mine = front_facing(sn);
		}
		uint32_t const validMask = (__ballot_sync(0xFFFFFFFFu, mine) & grpMask) >> grpShift;
		uint32_t const nOut = __popc(validMask);
		uint32_t slotBase = 0;
		bool ok = nOut != 0u;
		if (ok && sub == 0u)
		{
			uint32_t const fanBase = atomicAdd(&A.ctl->numFanSlots, nOut);
			slotBase = fp.numInputTris + fanBase;
			if (slotBase + nOut > fp.slotCapacity)
			{
				atomicOr(&A.ctl->overflow, 1u);
				ok = false;
			}
			else
			{
				// redirect record at the (otherwise unused) slot of the input triangle
				*reinterpret_cast<uint2*>(&A.shadeRecs[g].pad[0]) = make_uint2(slotBase, validMask);
			}
		}
		ok = __shfl_sync(0xFFFFFFFFu, (int)ok, (int)grpShift) != 0;
		slotBase = __shfl_sync(0xFFFFFFFFu, slotBase, (int)grpShift);
		bool emitted = false;
		uint2 oneTile = make_uint2(0u, 0u);
		uint32_t const k = __popc(validMask & below);
		if (ok && mine)
		{
			const DrawDev& d = A.draws[drawIdx];
			// (the attribute slots of a clip vertex are the eight floats behind its position)
			emitted = emit_triangle(sn, reinterpret_cast<const float*>(&poly[src][0][1]),
			                        reinterpret_cast<const float*>(&poly[src][i - 1u][1]),
			                        reinterpret_cast<const float*>(&poly[src][i][1]), d, drawIdx, fp, slotBase + k, A.rasterRecs,
			                        A.shadeRecs, s_hist, A.tileCounts, oneTile);
		}
		// survivors: the fan triangles that were set up (in a screen-tile split: those that touch this GPU's tiles)
		uint32_t const em = __ballot_sync(0xFFFFFFFFu, emitted);
		uint32_t sBase = 0;
		if (lane == 0 && em)
		{
			sBase = atomicAdd(&A.ctl->numSurvivors, (uint32_t)__popc(em));
		}
		sBase = __shfl_sync(0xFFFFFFFFu, sBase, 0);
		if (emitted)
		{
			*reinterpret_cast<uint4*>(&A.survivors[sBase + __popc(em & ((1u << lane) - 1u))]) =
				make_uint4(SRB_KEY_FAN(g, i - 2), slotBase + k, oneTile.x, oneTile.y);
		}
		__syncwarp(); // the polygon buffers are reused by the next iteration
	}
	if (!A.fuseScan)
	{
		return;
	}
	// The last CTA to get here runs the tile scan: every count of the frame is in the global counters by then.
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0)
	{
		s_isLast = atomicAdd(&A.ctl->ctasDone, 1u) == gridDim.x - 1u ? 1u : 0u;
	}
	__syncthreads();
	if (s_isLast)
	{
		__threadfence();
		tile_scan_block<kClipThreads>(fp, A.tileCounts, A.offsets, A.cursors, A.units, A.ctl, s_counts);
	}
}

__global__ void __launch_bounds__(kDirectThreads, 1024 / kDirectThreads) setup_direct_kernel(const __grid_constant__ SetupArgs A)
{
	extern __shared__ uint32_t s_dyn[]; // [numTiles] tile histogram (if it fits), then [numDraws] triBase table (if it fits)
	const FrameParams& fp = A.fp;
	uint32_t const numTiles = fp.tilesX * fp.tilesY;
	uint32_t* const s_hist = fp.smemHist ? s_dyn : nullptr;
	uint32_t* const s_triBase = fp.smemBase ? s_dyn + (fp.smemHist ? numTiles : 0u) : nullptr;
	uint32_t const tid = threadIdx.x, lane = tid & 31u;
	if (A.releaseFlag && blockIdx.x == 0 && tid == 0)
	{
		// screen-tile split: the application has consumed the previous frame (it is submitting this one), so the other
		// GPUs may overwrite its tiles in this GPU's framebuffer
		*reinterpret_cast<volatile uint32_t*>(A.releaseFlag) = A.ctl->doneValue - 1u;
	}
	if (s_hist)
	{
		for (uint32_t i = tid; i < numTiles; i += kDirectThreads) s_hist[i] = 0;
	}
	if (s_triBase)
	{
		for (uint32_t i = tid; i < fp.numDraws; i += kDirectThreads) s_triBase[i] = A.draws[i].triBase;
	}
	__syncthreads();
	float const hx = mulf((float)fp.width, 0.5f), hy = mulf((float)fp.height, 0.5f);

	// The grid either covers the input one triangle per thread (one frame in flight: lowest latency) or is a few CTAs per
	// SM striding through it (several frames in flight: the kernel waits on dependent loads most of the time, and a full
	// grid would hold every register of the SMs it runs on, locking the other frames' kernels out).
	for (uint32_t base = blockIdx.x * kDirectThreads; base < fp.numInputTris; base += gridDim.x * kDirectThreads)
	{
		uint32_t const g = base + tid; // global input triangle index, draw-major
		bool survive = false, needsClip = false;
		uint32_t drawIdx = 0;
		uint2 oneTile = make_uint2(0u, 0u);
		if (g < fp.numInputTris)
		{
			drawIdx = find_draw(s_triBase, A.draws, fp.numDraws, g);
			const DrawDev& d = A.draws[drawIdx];
			uint32_t const t = g - d.triBase;
			float4 v[3];
			const float* ap[3];
#pragma unroll
			for (int i = 0; i < 3; ++i)
			{
				uint32_t const idx = fetch_index(d, t * 3 + i);
				v[i] = transform(d, reinterpret_cast<const float*>(d.pos + (size_t)idx * d.posStride));
				ap[i] = reinterpret_cast<const float*>(d.attr + (size_t)idx * d.attrStride);
			}
			uint32_t const c0 = clip_code(v[0].x, v[0].y, v[0].z, v[0].w);
			uint32_t const c1 = clip_code(v[1].x, v[1].y, v[1].z, v[1].w);
			uint32_t const c2 = clip_code(v[2].x, v[2].y, v[2].z, v[2].w);
			if ((c0 | c1 | c2) == 0)
			{
				Snapped s;
				snap(v, hx, hy, s);
				if (front_facing(s))
				{
					survive = emit_triangle(s, ap[0], ap[1], ap[2], d, drawIdx, fp, g, A.rasterRecs, A.shadeRecs, s_hist,
					                        A.tileCounts, oneTile);
				}
			}
			else if ((c0 & c1 & c2) == 0)
			{
				needsClip = true; // Binning.cpp:498-523
			}
		}
		// warp-aggregated appends to the survivor list and the clip queue
		uint32_t const sm = __ballot_sync(0xFFFFFFFFu, survive);
		uint32_t const cm = __ballot_sync(0xFFFFFFFFu, needsClip);
		uint32_t sBase = 0, cBase = 0;
		if (lane == 0)
		{
			if (sm) sBase = atomicAdd(&A.ctl->numSurvivors, (uint32_t)__popc(sm));
			if (cm) cBase = atomicAdd(&A.ctl->numClipQueue, (uint32_t)__popc(cm));
		}
		sBase = __shfl_sync(0xFFFFFFFFu, sBase, 0);
		cBase = __shfl_sync(0xFFFFFFFFu, cBase, 0);
		uint32_t const below = (1u << lane) - 1u;
		if (survive)
		{
			*reinterpret_cast<uint4*>(&A.survivors[sBase + __popc(sm & below)]) = make_uint4(SRB_KEY_UNCLIPPED(g), g, oneTile.x, oneTile.y);
		}
		if (needsClip)
		{
			A.clipQueue[cBase + __popc(cm & below)] = g;
		}
	}
	__syncthreads();
	if (s_hist)
	{
		for (uint32_t i = tid; i < numTiles; i += kDirectThreads)
		{
			uint32_t const n = s_hist[i];
			if (n && tile_owned(fp, i)) atomicAdd(&A.tileCounts[i], n);
		}
	}
}


// ---------------------------------------------------------------------------------------------------------------
// setup_kernel (experiment, not the default): the chunked set-up.  A CTA takes CHUNKS of 256 consecutive triangles of one draw:
//   indices   : the chunk's index bytes arrive in shared memory by ONE bulk asynchronous copy (cp.async.bulk + mbarrier:
//               the stream is contiguous) — or by plain loads when the chunk is not 16-byte aligned;
//   vertices  : a chunk of a mesh references a narrow index range, each vertex 5-6 times (hall scene: 141 k unique
//               vertices for 264 k triangles).  If the range [imin, imax] is at most 768 vertices every vertex of it is
//               transformed ONCE (coalesced reads of consecutive vertices): clip-space position -> clip code, 1/w,
//               raster x / y, their 24.8 snaps, z/w — exactly the per-vertex operations of Binning.cpp:475-303 — into
//               shared memory; otherwise (indices scattered over the mesh) every triangle corner is its own entry;
//   cull      : one thread per triangle: clip codes, trivial accept / reject / queue for the clipper, area cull;
//   compact   : the survivors (40 % on the hall scene) are compacted (ballot + warp counts), so that
//   set-up    : the expensive half — edges, planes, record stores, tile counting (emit_triangle) — runs on FULL warps.
// Same arithmetic, operation for operation, as the per-triangle kernel (setup_direct_kernel, the one the library uses):
// only where and how often it runs differs.  EXPERIMENT, enabled with SRB_SETUP_CHUNKED=1: bit-exact on the whole test
// suite, 10 % fewer warp instructions on the hall scene, but slower (see launch_setup) — the per-triangle kernel's cost
// is spread over fetch, transform, clip codes and control flow, not concentrated in what a vertex cache removes.
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t kChunkTris = kSetupThreads;
constexpr uint32_t kVertCap = 3u * kChunkTris; // 768: one entry per corner in the worst case

struct __align__(16) VertX
{
	float rx, ry, iw, zw; // raster x, y (Binning.cpp:291-299), 1/w, z/w
	int32_t fx, fy;       // 24.8 snaps (:301-303)
	uint32_t code;        // clip code (:56-68)
	uint32_t pad;
};
static_assert(sizeof(VertX) == 32, "VertX is two float4");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one vertex through Binning.cpp:475-485 and :291-303
__device__ __forceinline__ VertX transform_vertex(const DrawDev& d, uint32_t idx, float hx, float hy)
{
	float4 const v = transform(d, reinterpret_cast<const float*>(d.pos + (size_t)idx * d.posStride));
	VertX o;
	o.code = clip_code(v.x, v.y, v.z, v.w);
	o.iw = __frcp_rn(v.w);
	o.zw = mulf(v.z, o.iw);
	o.rx = addf(mulf(mulf(o.iw, v.x), hx), hx);
	o.ry = addf(mulf(mulf(o.iw, v.y), -hy), hy);
	o.fx = cvtt_x86(addf(mulf(o.rx, 256.0f), 0.5f));
	o.fy = cvtt_x86(addf(mulf(o.ry, 256.0f), 0.5f));
	o.pad = 0u;
	return o;
}

__device__ __forceinline__ uint32_t index_from_smem(const uint8_t* raw, uint32_t stride, uint32_t i)
{
	switch (stride)
	{
		case 1: return raw[i];
		case 2: return reinterpret_cast<const uint16_t*>(raw)[i];
		default: return reinterpret_cast<const uint32_t*>(raw)[i];
	}
}

__global__ void __launch_bounds__(kSetupThreads, 4) setup_kernel(const __grid_constant__ SetupArgs A)
{
	extern __shared__ uint32_t s_dyn[]; // [numTiles] tile histogram (if it fits), then [numDraws] chunkBase table (if it fits)
	__shared__ __align__(16) VertX s_vx[kVertCap];                 // 24 KB
	__shared__ __align__(16) uint8_t s_idxRaw[kChunkTris * 3u * 4u]; // the chunk's index bytes, 3 KB
	__shared__ __align__(8) unsigned long long s_mbar;
	__shared__ uint16_t s_surv[kChunkTris];
	__shared__ uint32_t s_warpCount[kSetupThreads / 32];
	__shared__ uint32_t s_range[2]; // min, max vertex index of the chunk
	__shared__ uint32_t s_draw;
	const FrameParams& fp = A.fp;
	uint32_t const numTiles = fp.tilesX * fp.tilesY;
	uint32_t* const s_hist = fp.smemHist ? s_dyn : nullptr;
	uint32_t* const s_chunkBase = fp.smemBase ? s_dyn + (fp.smemHist ? numTiles : 0u) : nullptr;
	uint32_t const tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	if (A.releaseFlag && blockIdx.x == 0 && tid == 0)
	{
		// screen-tile split: the application has consumed the previous frame (it is submitting this one), so the other
		// GPUs may overwrite its tiles in this GPU's framebuffer
		*reinterpret_cast<volatile uint32_t*>(A.releaseFlag) = A.ctl->doneValue - 1u;
	}
	if (s_hist)
	{
		for (uint32_t i = tid; i < numTiles; i += kSetupThreads) s_hist[i] = 0;
	}
	if (s_chunkBase)
	{
		for (uint32_t i = tid; i < fp.numDraws; i += kSetupThreads) s_chunkBase[i] = A.draws[i].chunkBase;
	}
	if (tid == 0)
	{
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_mbar)), "r"(1));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	float const hx = mulf((float)fp.width, 0.5f), hy = mulf((float)fp.height, 0.5f);
	uint32_t phase = 0;

	// The grid either covers the input one chunk per CTA (one frame in flight: lowest latency) or is a few CTAs per SM
	// striding through the chunks (several frames in flight: leaves the SMs' registers to the other frames' kernels).
	for (uint32_t chunk = blockIdx.x; chunk < fp.numChunks; chunk += gridDim.x)
	{
		// ---- which draw, which triangles --------------------------------------------------------------------------
		if (tid == 0)
		{
			uint32_t lo = 0, hi = fp.numDraws; // last d with chunkBase[d] <= chunk
			while (hi - lo > 1)
			{
				uint32_t const mid = (lo + hi) >> 1;
				uint32_t const b = s_chunkBase ? s_chunkBase[mid] : __ldg(&A.draws[mid].chunkBase);
				if (b <= chunk) lo = mid; else hi = mid;
			}
			s_draw = lo;
			s_range[0] = 0xFFFFFFFFu;
			s_range[1] = 0u;
		}
		__syncthreads();
		uint32_t const drawIdx = s_draw;
		const DrawDev& d = A.draws[drawIdx];
		uint32_t const firstTri = (chunk - d.chunkBase) * kChunkTris;
		uint32_t const n = min(kChunkTris, d.numTris - firstTri);
		uint32_t const stride = d.idxStride;

		// ---- indices: one bulk asynchronous copy of the chunk's index bytes (contiguous), plain loads otherwise -------
		const uint8_t* const src = d.idx + (size_t)firstTri * 3u * stride;
		uint32_t const bytes = n * 3u * stride;
		bool const bulk = (((size_t)src | bytes) & 15u) == 0u;
		if (bulk)
		{
			if (tid == 0)
			{
				// (the buffer may last have been written by ordinary stores: order them before the asynchronous proxy's write)
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
				asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_mbar)), "r"(bytes) : "memory");
				asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
				                 smem_u32(s_idxRaw)),
				             "l"(src), "r"(bytes), "r"(smem_u32(&s_mbar))
				             : "memory");
			}
			uint32_t done = 0;
			while (!done)
			{
				asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
				             : "=r"(done)
				             : "r"(smem_u32(&s_mbar)), "r"(phase)
				             : "memory");
			}
			phase ^= 1u;
		}
		else
		{
			for (uint32_t i = tid; i < bytes; i += kSetupThreads) s_idxRaw[i] = __ldg(src + i);
			__syncthreads();
		}
		uint32_t i0 = 0, i1 = 0, i2 = 0;
		if (tid < n)
		{
			i0 = index_from_smem(s_idxRaw, stride, tid * 3u);
			i1 = index_from_smem(s_idxRaw, stride, tid * 3u + 1u);
			i2 = index_from_smem(s_idxRaw, stride, tid * 3u + 2u);
		}
		{
			uint32_t const lo = __reduce_min_sync(0xFFFFFFFFu, tid < n ? min(i0, min(i1, i2)) : 0xFFFFFFFFu);
			uint32_t const hi = __reduce_max_sync(0xFFFFFFFFu, tid < n ? max(i0, max(i1, i2)) : 0u);
			if (lane == 0)
			{
				atomicMin(&s_range[0], lo);
				atomicMax(&s_range[1], hi);
			}
		}
		__syncthreads();
		uint32_t const imin = s_range[0];
		bool const shared = s_range[1] - imin < kVertCap; // every vertex of the range once; else one entry per corner

		// ---- vertices -------------------------------------------------------------------------------------------------
		uint32_t l0, l1, l2; // entries of my triangle's corners in s_vx
		if (shared)
		{
			uint32_t const count = s_range[1] - imin + 1u;
			for (uint32_t v = tid; v < count; v += kSetupThreads)
			{
				s_vx[v] = transform_vertex(d, imin + v, hx, hy);
			}
			l0 = i0 - imin, l1 = i1 - imin, l2 = i2 - imin;
		}
		else
		{
			l0 = tid * 3u, l1 = l0 + 1u, l2 = l0 + 2u;
			if (tid < n)
			{
				s_vx[l0] = transform_vertex(d, i0, hx, hy);
				s_vx[l1] = transform_vertex(d, i1, hx, hy);
				s_vx[l2] = transform_vertex(d, i2, hx, hy);
			}
		}
		__syncthreads();

		// ---- cull: trivial accept / reject / clip queue (Binning.cpp:487-523), area (:305-311) ----------------------------
		bool survive = false, needsClip = false;
		uint32_t const g = d.triBase + firstTri + tid; // global input triangle index, draw-major
		if (tid < n)
		{
			uint32_t const c0 = s_vx[l0].code, c1 = s_vx[l1].code, c2 = s_vx[l2].code;
			if ((c0 | c1 | c2) == 0u)
			{
				int32_t const ax = s_vx[l0].fx, ay = s_vx[l0].fy, bx = s_vx[l1].fx, by = s_vx[l1].fy, cx = s_vx[l2].fx, cy = s_vx[l2].fy;
				int64_t a = (int64_t)wrap_sub(cx, ax) * (int64_t)wrap_sub(by, ay) - (int64_t)wrap_sub(cy, ay) * (int64_t)wrap_sub(bx, ax);
				a >>= 8;
				survive = a > 0;
			}
			else if ((c0 & c1 & c2) == 0u)
			{
				needsClip = true;
			}
		}
		uint32_t const sm = __ballot_sync(0xFFFFFFFFu, survive);
		uint32_t const cm = __ballot_sync(0xFFFFFFFFu, needsClip);
		uint32_t const below = (1u << lane) - 1u;
		if (cm)
		{
			uint32_t cBase = 0;
			if (lane == 0) cBase = atomicAdd(&A.ctl->numClipQueue, (uint32_t)__popc(cm));
			cBase = __shfl_sync(0xFFFFFFFFu, cBase, 0);
			if (needsClip) A.clipQueue[cBase + __popc(cm & below)] = g;
		}
		if (lane == 0) s_warpCount[warp] = (uint32_t)__popc(sm);
		__syncthreads();
		uint32_t before = 0, total = 0;
#pragma unroll
		for (uint32_t w = 0; w < kSetupThreads / 32; ++w)
		{
			uint32_t const c = s_warpCount[w];
			before += w < warp ? c : 0u;
			total += c;
		}
		if (survive) s_surv[before + __popc(sm & below)] = (uint16_t)tid;
		__syncthreads();

		// ---- set-up of the survivors, on full warps ---------------------------------------------------------------------
		for (uint32_t kb = warp * 32u; kb < total; kb += kSetupThreads) // (warp-uniform trip count: full-mask votes below)
		{
			uint32_t const k = kb + lane;
			bool emitted = false;
			uint32_t gs = 0;
			uint2 oneTile = make_uint2(0u, 0u);
			if (k < total)
			{
				uint32_t const t = s_surv[k];
				uint32_t const j0 = index_from_smem(s_idxRaw, stride, t * 3u), j1 = index_from_smem(s_idxRaw, stride, t * 3u + 1u),
				               j2 = index_from_smem(s_idxRaw, stride, t * 3u + 2u);
				uint32_t const e0 = shared ? j0 - imin : t * 3u, e1 = shared ? j1 - imin : t * 3u + 1u,
				               e2 = shared ? j2 - imin : t * 3u + 2u;
				Snapped s;
				{
					const float4* q = reinterpret_cast<const float4*>(s_vx);
					float4 const a0 = q[e0 * 2u], a1 = q[e0 * 2u + 1u], b0 = q[e1 * 2u], b1 = q[e1 * 2u + 1u], c0 = q[e2 * 2u],
					             c1 = q[e2 * 2u + 1u];
					s.rx[0] = a0.x, s.ry[0] = a0.y, s.iw[0] = a0.z, s.zw[0] = a0.w, s.fx[0] = __float_as_int(a1.x), s.fy[0] = __float_as_int(a1.y);
					s.rx[1] = b0.x, s.ry[1] = b0.y, s.iw[1] = b0.z, s.zw[1] = b0.w, s.fx[1] = __float_as_int(b1.x), s.fy[1] = __float_as_int(b1.y);
					s.rx[2] = c0.x, s.ry[2] = c0.y, s.iw[2] = c0.z, s.zw[2] = c0.w, s.fx[2] = __float_as_int(c1.x), s.fy[2] = __float_as_int(c1.y);
				}
				const float* const p0 = reinterpret_cast<const float*>(d.attr + (size_t)j0 * d.attrStride);
				const float* const p1 = reinterpret_cast<const float*>(d.attr + (size_t)j1 * d.attrStride);
				const float* const p2 = reinterpret_cast<const float*>(d.attr + (size_t)j2 * d.attrStride);
				gs = d.triBase + firstTri + t;
				emitted = emit_triangle(s, p0, p1, p2, d, drawIdx, fp, gs, A.rasterRecs, A.shadeRecs, s_hist, A.tileCounts, oneTile);
			}
			// warp-aggregated append to the survivor list (in a screen-tile split: the triangles that touch this GPU's tiles)
			uint32_t const em = __ballot_sync(0xFFFFFFFFu, emitted);
			uint32_t sBase = 0;
			if (lane == 0 && em) sBase = atomicAdd(&A.ctl->numSurvivors, (uint32_t)__popc(em));
			sBase = __shfl_sync(0xFFFFFFFFu, sBase, 0);
			if (emitted)
			{
				*reinterpret_cast<uint4*>(&A.survivors[sBase + __popc(em & below)]) = make_uint4(SRB_KEY_UNCLIPPED(gs), gs, oneTile.x, oneTile.y);
			}
		}
		__syncthreads(); // the chunk's shared memory is reused
	}
	if (s_hist)
	{
		for (uint32_t i = tid; i < numTiles; i += kSetupThreads)
		{
			uint32_t const n = s_hist[i];
			if (n && tile_owned(fp, i)) atomicAdd(&A.tileCounts[i], n);
		}
	}
}

} // namespace

size_t setup_smem_bytes(const FrameParams& fp)
{
	return (size_t(fp.smemHist ? fp.tilesX * fp.tilesY : 0u) + (fp.smemBase ? fp.numDraws : 0u)) * sizeof(uint32_t);
}

void setup_plan_smem(FrameParams& fp)
{
	// what fits into 96 KB of dynamic shared memory: the tile histogram first (it saves an atomic per reference), then the
	// draws' triBase table (it saves global loads in a binary search)
	size_t const budget = 96 * 1024;
	size_t const hist = size_t(fp.tilesX) * fp.tilesY * sizeof(uint32_t);
	fp.smemHist = hist <= 64 * 1024 ? 1u : 0u;
	size_t const left = budget - (fp.smemHist ? hist : 0);
	// (at most 16 KB of it: every CTA loads the table, and a larger one would cost occupancy)
	fp.smemBase = size_t(fp.numDraws) * sizeof(uint32_t) <= std::min<size_t>(left, 16 * 1024) ? 1u : 0u;
}

cudaError_t setup_init()
{
	cudaError_t e = cudaFuncSetAttribute(clip_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
	if (e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(setup_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
	if (e != cudaSuccess) return e;
	return cudaFuncSetAttribute(setup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
}

static SetupArgs make_args(const FrameParams& fp, const DrawDev* draws, RasterRec* rasterRecs, ShadeRec* shadeRecs,
                           Survivor* survivors, uint32_t* clipQueue, uint32_t* tileCounts, uint32_t* offsets, uint32_t* cursors,
                           UnitDesc* units, FrameCtl* ctl, bool fuseScan, uint32_t* releaseFlag)
{
	SetupArgs A;
	A.fp = fp;
	A.draws = draws;
	A.rasterRecs = rasterRecs;
	A.shadeRecs = shadeRecs;
	A.survivors = survivors;
	A.clipQueue = clipQueue;
	A.tileCounts = tileCounts;
	A.offsets = offsets;
	A.cursors = cursors;
	A.units = units;
	A.ctl = ctl;
	A.fuseScan = fuseScan ? 1u : 0u;
	A.releaseFlag = releaseFlag;
	return A;
}

bool launch_setup(const FrameParams& fp, const DrawDev* draws, RasterRec* rasterRecs, ShadeRec* shadeRecs,
                  Survivor* survivors, uint32_t* clipQueue, uint32_t* tileCounts, FrameCtl* ctl, uint32_t ctasPerSm,
                  uint32_t* releaseFlag, cudaStream_t stream)
{
	if (fp.numInputTris == 0)
	{
		return false;
	}
	// The chunked kernel (vertex cache + survivor compaction + bulk index copy) is an EXPERIMENT that did not pay and is
	// kept behind SRB_SETUP_CHUNKED=1 with its measurements (profiles/README.md): 10 % fewer warp instructions on the hall
	// (5.86 M vs 6.55 M), but five barriers per chunk: 12 frames in flight 67.5 vs 65.7 us per frame, 1 M random triangles
	// (no vertex reuse to find) 274 vs 256 us.
	static bool const direct = getenv("SRB_SETUP_CHUNKED") == nullptr;
	uint32_t blocks = direct ? (fp.numInputTris + kDirectThreads - 1) / kDirectThreads : fp.numChunks;
	static uint32_t const envCtas = [] {
		const char* e = getenv("SRB_SETUP_CTAS_PER_SM"); // tuning knob for experiments (not part of the ABI)
		return (uint32_t)(e && atoi(e) > 0 ? atoi(e) : 0);
	}();
	uint32_t const perSm = envCtas ? envCtas : ctasPerSm;
	if (perSm) blocks = std::min(blocks, 148u * (direct ? std::max(1u, perSm * 256u / kDirectThreads) : perSm));
	SetupArgs const A = make_args(fp, draws, rasterRecs, shadeRecs, survivors, clipQueue, tileCounts, nullptr, nullptr, nullptr,
	                              ctl, false, releaseFlag);
	if (direct) setup_direct_kernel<<<blocks, kDirectThreads, setup_smem_bytes(fp), stream>>>(A);
	else setup_kernel<<<blocks, kSetupThreads, setup_smem_bytes(fp), stream>>>(A);
	return true;
}

// The clip pass and — with fuseScan — the tile scan in its tail.  Launched for every frame (a frame without triangles
// still needs its offsets and units).
void launch_clip_scan(const FrameParams& fp, const DrawDev* draws, RasterRec* rasterRecs, ShadeRec* shadeRecs,
                      Survivor* survivors, uint32_t* clipQueue, uint32_t* tileCounts, uint32_t* offsets, uint32_t* cursors,
                      UnitDesc* units, FrameCtl* ctl, bool fuseScan, uint32_t* releaseFlag, cudaStream_t stream)
{
	// one 16-lane group per queued triangle up to ~1.5 % clipped triangles, grid-stride beyond
	uint32_t blocks = (fp.numInputTris / 64 + (kClipThreads / kClipLanes) - 1) / (kClipThreads / kClipLanes);
	blocks = blocks < 1 ? 1 : (blocks > 148u * 4u ? 148u * 4u : blocks);
	SetupArgs const A = make_args(fp, draws, rasterRecs, shadeRecs, survivors, clipQueue, tileCounts, offsets, cursors, units,
	                              ctl, fuseScan, releaseFlag);
	clip_scan_kernel<<<blocks, kClipThreads, setup_smem_bytes(fp), stream>>>(A);
}

} // namespace srb
