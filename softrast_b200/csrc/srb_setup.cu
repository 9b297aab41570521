// srb_setup.cu — K1: vertex transform, frustum clipping, triangle set-up and per-tile reference counting.
//
// Replaces the reference front-end BinTrisEntry + BinTransformedAndClippedTri (SoftRast/Binning.cpp:464-535, :279-456)
// up to, but not including, the per-bin append (that is K2, srb_bin.cu).
//
//   setup_kernel : one thread per INPUT triangle over all draws of the frame.  Transform, clip codes, trivial
//                  accept/reject.  Unclipped front-facing triangles are set up in place: records are written at
//                  slot = input triangle index (no allocation, no ordering dependency between threads).  Triangles
//                  that cross a frustum plane (a few %) are only QUEUED, so no warp ever serialises behind the
//                  clipper.
//   clip_kernel  : one thread per queued triangle: Sutherland-Hodgman, fan, set-up of every surviving fan triangle in
//                  slots handed out beyond numInputTris.
// Draw order is carried by the canonical key (srb_device.cuh), not by where a record is stored.  Both kernels count
// tile references: the main kernel through a shared-memory histogram flushed once per CTA (one global atomic per
// touched tile per CTA), the clip kernel with plain global atomics.
#include "srb_device.cuh"
#include "srb_kernels.h"

#include <algorithm>
#include <stdlib.h>

namespace srb
{

namespace
{

constexpr int kSetupThreads = 256;
constexpr int kMaxClipVerts = 9; // Binning.cpp:71: 3 + one per frustum plane

struct ClipVert
{
	float x, y, z, w;
	float a[SRB_MAX_VARY];
};

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

// kt::Dot(plane, v) (kt/src/kt/inl/Vec4.inl:162-165) for the six planes of Binning.cpp:87-97.
__device__ __forceinline__ float plane_dot(uint32_t plane, const ClipVert& v)
{
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

// Sutherland-Hodgman against one plane, Binning.cpp:85-165.
__device__ uint32_t clip_plane(const ClipVert* in, uint32_t nIn, ClipVert* out, uint32_t plane)
{
	uint32_t nOut = 0;
	uint32_t i0 = nIn - 1;
	float d0 = plane_dot(plane, in[i0]);
	for (uint32_t i1 = 0; i1 < nIn; ++i1)
	{
		float const d1 = plane_dot(plane, in[i1]);
		bool const in0 = d0 >= 0.0f;
		bool const in1 = d1 >= 0.0f;
		if (in0)
		{
			out[nOut++] = in[i0];
		}
		if (in0 != in1)
		{
			// the inside vertex is always the first Lerp argument
			const ClipVert& a = in1 ? in[i1] : in[i0];
			const ClipVert& b = in1 ? in[i0] : in[i1];
			float const t = in1 ? divf(d1, subf(d1, d0)) : divf(d0, subf(d0, d1));
			ClipVert& o = out[nOut++];
			o.x = lerp_kt(a.x, b.x, t);
			o.y = lerp_kt(a.y, b.y, t);
			o.z = lerp_kt(a.z, b.z, t);
			o.w = lerp_kt(a.w, b.w, t);
#pragma unroll
			for (int k = 0; k < SRB_MAX_VARY; ++k)
			{
				o.a[k] = lerp_kt(a.a[k], b.a[k], t);
			}
		}
		d0 = d1;
		i0 = i1;
	}
	return nOut;
}

struct Snapped
{
	float rx[3], ry[3], iw[3];
	int32_t fx[3], fy[3];
};

// Viewport transform + 24.8 snap, Binning.cpp:291-303.
__device__ __forceinline__ void snap(const float4 (&v)[3], float hx, float hy, Snapped& s)
{
#pragma unroll
	for (int i = 0; i < 3; ++i)
	{
		s.iw[i] = divf(1.0f, v[i].w);
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

// Full set-up of one surviving triangle (Binning.cpp:313-350) + tile reference counting (:352-410).
template <bool kSmemHist>
__device__ __forceinline__ void emit_triangle(const float4 (&v)[3], const float* a0, const float* a1, const float* a2,
                                              const DrawDev& draw, uint32_t drawIdx, const FrameParams& fp,
                                              uint32_t slot, RasterRec* __restrict__ rasterRecs,
                                              ShadeRec* __restrict__ shadeRecs, uint32_t* tileCounts)
{
	float const hx = mulf((float)fp.width, 0.5f);
	float const hy = mulf((float)fp.height, 0.5f);
	Snapped s;
	snap(v, hx, hy, s);

	RasterRec rr;
	int32_t const W1 = (int32_t)fp.width - 1, H1 = (int32_t)fp.height - 1;
	rr.xmin = (uint16_t)clampi(wrap_add(min3(s.fx[0], s.fx[1], s.fx[2]), 255) >> 8, 0, W1);
	rr.ymin = (uint16_t)clampi(wrap_add(min3(s.fy[0], s.fy[1], s.fy[2]), 255) >> 8, 0, H1);
	rr.xmax = (uint16_t)clampi(wrap_add(max3(s.fx[0], s.fx[1], s.fx[2]), 255) >> 8, 0, W1);
	rr.ymax = (uint16_t)clampi(wrap_add(max3(s.fy[0], s.fy[1], s.fy[2]), 255) >> 8, 0, H1);
	setup_edge(s.fx[0], s.fy[0], s.fx[1], s.fy[1], rr.c[0], rr.dx[0], rr.dy[0]);
	setup_edge(s.fx[1], s.fy[1], s.fx[2], s.fy[2], rr.c[1], rr.dx[1], rr.dy[1]);
	setup_edge(s.fx[2], s.fy[2], s.fx[0], s.fy[0], rr.c[2], rr.dx[2], rr.dy[2]);

	float const d10x = subf(s.rx[1], s.rx[0]), d10y = subf(s.ry[1], s.ry[0]);
	float const d20x = subf(s.rx[2], s.rx[0]), d20y = subf(s.ry[2], s.ry[0]);
	float const K = subf(mulf(d10x, d20y), mulf(d10y, d20x));

	float const zw0 = mulf(v[0].z, s.iw[0]);
	setup_plane(K, d10x, d10y, d20x, d20y, subf(mulf(v[1].z, s.iw[1]), zw0), subf(mulf(v[2].z, s.iw[2]), zw0), rr.zdx,
	            rr.zdy);
	rr.z0 = zw0;
	rr.r0x = s.rx[0];
	rr.r0y = s.ry[0];

	ShadeRec sr;
	setup_plane(K, d10x, d10y, d20x, d20y, subf(s.iw[1], s.iw[0]), subf(s.iw[2], s.iw[0]), sr.wdx, sr.wdy);
	sr.w0 = s.iw[0];
	sr.info = (draw.shader & 0xFFu) | (min(draw.uvOffset, 255u) << 8) | ((uint32_t)(draw.texture + 1) << 16);
	sr.r0x = s.rx[0];
	sr.r0y = s.ry[0];
	sr.pad[0] = drawIdx;
	sr.pad[1] = 0;
#pragma unroll
	for (int i = 0; i < SRB_MAX_VARY; ++i)
	{
		if ((draw.planeMask >> i) & 1u)
		{
			float const q0 = mulf(a0[i], s.iw[0]);
			setup_plane(K, d10x, d10y, d20x, d20y, subf(mulf(a1[i], s.iw[1]), q0), subf(mulf(a2[i], s.iw[2]), q0),
			            sr.pl[SRB_PLANE_SLOT(i)][0], sr.pl[SRB_PLANE_SLOT(i)][1]);
			sr.pl[SRB_PLANE_SLOT(i)][2] = q0;
		}
		else
		{
			sr.pl[SRB_PLANE_SLOT(i)][0] = sr.pl[SRB_PLANE_SLOT(i)][1] = sr.pl[SRB_PLANE_SLOT(i)][2] = 0.0f;
		}
	}

	{
		uint4* dr = reinterpret_cast<uint4*>(rasterRecs + slot);
		const uint4* srr = reinterpret_cast<const uint4*>(&rr);
#pragma unroll
		for (int i = 0; i < 4; ++i) dr[i] = srr[i];
		uint4* ds = reinterpret_cast<uint4*>(shadeRecs + slot);
		const uint4* ssr = reinterpret_cast<const uint4*>(&sr);
#pragma unroll
		for (int i = 0; i < 4; ++i) ds[i] = ssr[i];
		// varyings 0..5 live in the second half of the record (SRB_PLANE_SLOT): untouched when no shader reads them
		if (draw.planeMask & 0x100u)
		{
#pragma unroll
			for (int i = 4; i < 8; ++i) ds[i] = ssr[i];
		}
	}

	// count the tiles this triangle will be appended to
	BinRange const br = bin_range(rr.xmin, rr.xmax, rr.ymin, rr.ymax);
	for (uint32_t by = br.by0; by <= br.by1; ++by)
	{
		for (uint32_t bx = br.bx0; bx <= br.bx1; ++bx)
		{
			if (br.check && !bin_overlaps(rr.c, rr.dx, rr.dy, (int32_t)(bx * SRB_TILE), (int32_t)(by * SRB_TILE)))
			{
				continue;
			}
			uint32_t const tile = by * fp.tilesX + bx;
			if (kSmemHist || tile_owned(fp, tile))
			{
				atomicAdd(&tileCounts[tile], 1u); // shared-memory histogram in the main kernel
			}
		}
	}
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

__device__ __forceinline__ uint32_t find_draw(const uint32_t* triBase, uint32_t numDraws, uint32_t g)
{
	uint32_t lo = 0, hi = numDraws; // last d with triBase[d] <= g
	while (hi - lo > 1)
	{
		uint32_t const mid = (lo + hi) >> 1;
		if (triBase[mid] <= g) lo = mid; else hi = mid;
	}
	return lo;
}


__global__ void __launch_bounds__(kSetupThreads, 4) setup_kernel(FrameParams fp, const DrawDev* __restrict__ draws,
                                                              RasterRec* __restrict__ rasterRecs,
                                                              ShadeRec* __restrict__ shadeRecs,
                                                              KeySlot* __restrict__ survivors,
                                                              uint32_t* __restrict__ clipQueue,
                                                              uint32_t* __restrict__ tileCounts,
                                                              FrameCtl* __restrict__ ctl)
{
	extern __shared__ uint32_t s_dyn[]; // [numTiles] tile histogram, then [numDraws] triBase table
	uint32_t const numTiles = fp.tilesX * fp.tilesY;
	uint32_t* s_hist = s_dyn;
	uint32_t* s_triBase = s_dyn + numTiles;
	uint32_t const tid = threadIdx.x, lane = tid & 31u;
	for (uint32_t i = tid; i < numTiles; i += kSetupThreads) s_hist[i] = 0;
	for (uint32_t i = tid; i < fp.numDraws; i += kSetupThreads) s_triBase[i] = draws[i].triBase;
	__syncthreads();

	// The grid either covers the input one triangle per thread (one frame in flight: lowest latency) or is a few CTAs per
	// SM striding through it (several frames in flight: the kernel waits on dependent loads most of the time, and a full
	// grid would hold every register of the SMs it runs on, locking the other frames' kernels out).
	for (uint32_t base = blockIdx.x * kSetupThreads; base < fp.numInputTris; base += gridDim.x * kSetupThreads)
	{
	uint32_t const g = base + tid; // global input triangle index, draw-major
	bool survive = false, needsClip = false;
	if (g < fp.numInputTris)
	{
		uint32_t const drawIdx = find_draw(s_triBase, fp.numDraws, g);
		const DrawDev& d = draws[drawIdx];
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
			snap(v, mulf((float)fp.width, 0.5f), mulf((float)fp.height, 0.5f), s);
			if (front_facing(s))
			{
				survive = true;
				emit_triangle<true>(v, ap[0], ap[1], ap[2], d, drawIdx, fp, g, rasterRecs, shadeRecs, s_hist);
			}
		}
		else if ((c0 & c1 & c2) == 0)
		{
			needsClip = true; // Binning.cpp:498-523 runs in clip_kernel
		}
	}
	// warp-aggregated appends to the survivor list and the clip queue
	uint32_t const sm = __ballot_sync(0xFFFFFFFFu, survive);
	uint32_t const cm = __ballot_sync(0xFFFFFFFFu, needsClip);
	uint32_t sBase = 0, cBase = 0;
	if (lane == 0)
	{
		if (sm) sBase = atomicAdd(&ctl->numSurvivors, (uint32_t)__popc(sm));
		if (cm) cBase = atomicAdd(&ctl->numClipQueue, (uint32_t)__popc(cm));
	}
	sBase = __shfl_sync(0xFFFFFFFFu, sBase, 0);
	cBase = __shfl_sync(0xFFFFFFFFu, cBase, 0);
	uint32_t const below = (1u << lane) - 1u;
	if (survive)
	{
		KeySlot ks;
		ks.key = SRB_KEY_UNCLIPPED(g);
		ks.slot = g;
		survivors[sBase + __popc(sm & below)] = ks;
	}
	if (needsClip)
	{
		clipQueue[cBase + __popc(cm & below)] = g;
	}
	}
	__syncthreads();
	for (uint32_t i = tid; i < numTiles; i += kSetupThreads)
	{
		uint32_t const n = s_hist[i];
		if (n && tile_owned(fp, i)) atomicAdd(&tileCounts[i], n);
	}
}

constexpr int kClipThreads = 64;

__global__ void __launch_bounds__(kClipThreads) clip_kernel(FrameParams fp, const DrawDev* __restrict__ draws,
                                                            RasterRec* __restrict__ rasterRecs,
                                                            ShadeRec* __restrict__ shadeRecs,
                                                            KeySlot* __restrict__ survivors,
                                                            const uint32_t* __restrict__ clipQueue,
                                                            uint32_t* __restrict__ tileCounts,
                                                            FrameCtl* __restrict__ ctl)
{
	// Eight lanes share one queued triangle: all of them clip it (same data, same path — no extra time), then lane i
	// culls and sets up fan triangle i, so the <= 7 set-ups of a polygon run side by side instead of one after the other
	// (the kernel is a few hundred triangles of pure latency: 18 -> ~10 us on the hall scene).
	uint32_t const n = ctl->numClipQueue;
	float const hx = mulf((float)fp.width, 0.5f), hy = mulf((float)fp.height, 0.5f);
	uint32_t const lane = threadIdx.x & 31u, sub = lane & 7u, grpShift = lane & 24u;
	uint32_t const groupsPerGrid = gridDim.x * (kClipThreads / 8);
	for (uint32_t base = blockIdx.x * (kClipThreads / 8) + (threadIdx.x >> 5) * 4u; base < n; base += groupsPerGrid)
	{
		// (a warp's four groups take four consecutive entries of one iteration, so the whole warp runs the same trip count)
		uint32_t const q = base + (lane >> 3);
		bool const have = q < n;
		uint32_t g = 0, drawIdx = 0, nVerts = 0, src = 0;
		ClipVert poly[2][kMaxClipVerts];
		if (have)
		{
			g = clipQueue[q];
			// find the draw: last d with triBase <= g
			uint32_t lo = 0, hi = fp.numDraws;
			while (hi - lo > 1)
			{
				uint32_t const mid = (lo + hi) >> 1;
				if (draws[mid].triBase <= g) lo = mid; else hi = mid;
			}
			drawIdx = lo;
			const DrawDev& d = draws[drawIdx];
			uint32_t const t = g - d.triBase;
			uint32_t maskOr = 0;
#pragma unroll
			for (int i = 0; i < 3; ++i)
			{
				uint32_t const idx = fetch_index(d, t * 3 + i);
				float4 const v = transform(d, reinterpret_cast<const float*>(d.pos + (size_t)idx * d.posStride));
				const float* ap = reinterpret_cast<const float*>(d.attr + (size_t)idx * d.attrStride);
				ClipVert& cv = poly[0][i];
				cv.x = v.x; cv.y = v.y; cv.z = v.z; cv.w = v.w;
#pragma unroll
				for (int k = 0; k < SRB_MAX_VARY; ++k)
				{
					cv.a[k] = ((uint32_t)k < d.numVaryings) ? ap[k] : 0.0f;
				}
				maskOr |= clip_code(v.x, v.y, v.z, v.w);
			}
			// Binning.cpp:498-523
			nVerts = 3;
			do
			{
				uint32_t const plane = __ffs(maskOr) - 1;
				maskOr ^= 1u << plane;
				nVerts = clip_plane(poly[src], nVerts, poly[src ^ 1], plane);
				src ^= 1;
			} while (maskOr && nVerts);
		}
		// fan (0, i-1, i), Binning.cpp:526-533: lane `sub` owns fan triangle i = sub + 2; which ones survive the cull?
		uint32_t const i = sub + 2u;
		bool mine = have && i < nVerts;
		float4 f[3];
		if (mine)
		{
			const ClipVert& p0 = poly[src][0];
			const ClipVert& p1 = poly[src][i - 1];
			const ClipVert& p2 = poly[src][i];
			f[0] = make_float4(p0.x, p0.y, p0.z, p0.w);
			f[1] = make_float4(p1.x, p1.y, p1.z, p1.w);
			f[2] = make_float4(p2.x, p2.y, p2.z, p2.w);
			Snapped sn;
			snap(f, hx, hy, sn);
			mine = front_facing(sn);
		}
		uint32_t const validMask = (__ballot_sync(0xFFFFFFFFu, mine) >> grpShift) & 0xFFu;
		uint32_t const nOut = __popc(validMask);
		uint32_t slotBase = 0, survBase = 0;
		bool ok = nOut != 0u;
		if (ok && sub == 0u)
		{
			uint32_t const fanBase = atomicAdd(&ctl->numFanSlots, nOut);
			slotBase = fp.numInputTris + fanBase;
			if (slotBase + nOut > fp.slotCapacity)
			{
				atomicOr(&ctl->overflow, 1u);
				ok = false;
			}
			else
			{
				survBase = atomicAdd(&ctl->numSurvivors, nOut);
				// redirect record at the (otherwise unused) slot of the input triangle
				shadeRecs[g].pad[0] = slotBase;
				shadeRecs[g].pad[1] = validMask;
			}
		}
		ok = __shfl_sync(0xFFFFFFFFu, (int)ok, (int)grpShift) != 0;
		slotBase = __shfl_sync(0xFFFFFFFFu, slotBase, (int)grpShift);
		survBase = __shfl_sync(0xFFFFFFFFu, survBase, (int)grpShift);
		if (ok && mine)
		{
			uint32_t const k = __popc(validMask & ((1u << sub) - 1u));
			const DrawDev& d = draws[drawIdx];
			emit_triangle<false>(f, poly[src][0].a, poly[src][i - 1].a, poly[src][i].a, d, drawIdx, fp, slotBase + k, rasterRecs,
			                     shadeRecs, tileCounts);
			KeySlot ks;
			ks.key = SRB_KEY_FAN(g, i - 2);
			ks.slot = slotBase + k;
			survivors[survBase + k] = ks;
		}
	}
}

} // namespace

size_t setup_smem_bytes(const FrameParams& fp)
{
	return (size_t(fp.tilesX) * fp.tilesY + fp.numDraws) * sizeof(uint32_t);
}

cudaError_t setup_init()
{
	return cudaFuncSetAttribute(setup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
}

bool launch_setup(const FrameParams& fp, const DrawDev* draws, RasterRec* rasterRecs, ShadeRec* shadeRecs,
                  KeySlot* survivors, uint32_t* clipQueue, uint32_t* tileCounts, FrameCtl* ctl, uint32_t ctasPerSm,
                  cudaStream_t stream)
{
	if (fp.numInputTris == 0)
	{
		return false;
	}
	uint32_t blocks = (fp.numInputTris + kSetupThreads - 1) / kSetupThreads;
	static uint32_t const envCtas = [] {
		const char* e = getenv("SRB_SETUP_CTAS_PER_SM"); // tuning knob for experiments (not part of the ABI)
		return (uint32_t)(e && atoi(e) > 0 ? atoi(e) : 0);
	}();
	uint32_t const perSm = envCtas ? envCtas : ctasPerSm;
	if (perSm) blocks = std::min(blocks, 148u * perSm);
	setup_kernel<<<blocks, kSetupThreads, setup_smem_bytes(fp), stream>>>(fp, draws, rasterRecs, shadeRecs, survivors,
	                                                                    clipQueue, tileCounts, ctl);
	return true;
}

bool launch_clip(const FrameParams& fp, const DrawDev* draws, RasterRec* rasterRecs, ShadeRec* shadeRecs,
                 KeySlot* survivors, const uint32_t* clipQueue, uint32_t* tileCounts, FrameCtl* ctl,
                 cudaStream_t stream)
{
	if (fp.numInputTris == 0)
	{
		return false;
	}
	// eight lanes per queued triangle; one group per triangle up to ~1.5 % clipped triangles, grid-stride beyond
	uint32_t blocks = (fp.numInputTris / 64 + (kClipThreads / 8) - 1) / (kClipThreads / 8);
	blocks = blocks < 1 ? 1 : (blocks > 148u * 8u ? 148u * 8u : blocks);
	clip_kernel<<<blocks, kClipThreads, 0, stream>>>(fp, draws, rasterRecs, shadeRecs, survivors, clipQueue, tileCounts,
	                                                ctl);
	return true;
}

} // namespace srb
