// srb_setup.cu — K1: vertex transform, frustum clipping, triangle set-up and per-tile reference counting.
//
// Replaces the reference front-end BinTrisEntry + BinTransformedAndClippedTri (SoftRast/Binning.cpp:464-535, :279-456)
// up to, but not including, the per-bin append (that is K2, srb_bin.cu).
//
// One thread per INPUT triangle over all draws of the frame (draw-major == the reference's canonical order).  Each
// thread produces 0..7 set-up triangles (clipping fans).  Records are written compacted and IN CANONICAL ORDER: a
// block-wide scan plus a single-pass decoupled look-back across blocks gives every output its rank, so
// record index == rank in (draw, triangle, fan) order — the order the single-threaded reference bins in.
#include "srb_device.cuh"
#include "srb_kernels.h"

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
__device__ void emit_triangle(const float4 (&v)[3], const float* a0, const float* a1, const float* a2,
                              const DrawDev& draw, uint32_t drawIdx, const FrameParams& fp, uint32_t rank,
                              RasterRec* __restrict__ rasterRecs, ShadeRec* __restrict__ shadeRecs,
                              uint32_t* __restrict__ tileCounts)
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
	sr.draw = drawIdx;
	sr.r0x = s.rx[0];
	sr.r0y = s.ry[0];
	sr.pad[0] = sr.pad[1] = 0;
#pragma unroll
	for (int i = 0; i < SRB_MAX_VARY; ++i)
	{
		if ((uint32_t)i < draw.numVaryings)
		{
			float const q0 = mulf(a0[i], s.iw[0]);
			setup_plane(K, d10x, d10y, d20x, d20y, subf(mulf(a1[i], s.iw[1]), q0), subf(mulf(a2[i], s.iw[2]), q0),
			            sr.adx[i], sr.ady[i]);
			sr.a0[i] = q0;
		}
		else
		{
			sr.adx[i] = sr.ady[i] = sr.a0[i] = 0.0f;
		}
	}

	if (rank < fp.setupCapacity)
	{
		uint4* dr = reinterpret_cast<uint4*>(rasterRecs + rank);
		const uint4* srr = reinterpret_cast<const uint4*>(&rr);
#pragma unroll
		for (int i = 0; i < 4; ++i) dr[i] = srr[i];
		uint4* ds = reinterpret_cast<uint4*>(shadeRecs + rank);
		const uint4* ssr = reinterpret_cast<const uint4*>(&sr);
#pragma unroll
		for (int i = 0; i < 8; ++i) ds[i] = ssr[i];
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
			atomicAdd(&tileCounts[by * fp.tilesX + bx], 1u);
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

// status in bits 32..33: 0 = not ready, 1 = block aggregate, 2 = inclusive prefix
constexpr unsigned long long kAgg = 1ull << 32;
constexpr unsigned long long kPre = 2ull << 32;

__global__ void __launch_bounds__(kSetupThreads) setup_kernel(FrameParams fp, const DrawDev* __restrict__ draws,
                                                              RasterRec* __restrict__ rasterRecs,
                                                              ShadeRec* __restrict__ shadeRecs,
                                                              uint32_t* __restrict__ tileCounts,
                                                              volatile unsigned long long* lookback,
                                                              FrameCtl* __restrict__ ctl)
{
	__shared__ uint32_t s_vbid;
	__shared__ uint32_t s_warpSum[kSetupThreads / 32];
	__shared__ uint32_t s_blockBase;

	uint32_t const tid = threadIdx.x;
	uint32_t const lane = tid & 31u, warp = tid >> 5;
	if (tid == 0)
	{
		s_vbid = atomicAdd(&ctl->ticket, 1u);
	}
	__syncthreads();
	uint32_t const vbid = s_vbid;
	uint32_t const g = vbid * kSetupThreads + tid; // global input triangle index, draw-major

	// ---- phase 1: transform, classify, clip, cull -> number of output triangles -----------------------------
	ClipVert poly[2][kMaxClipVerts];
	float4 v[3];
	const float* ap[3] = {nullptr, nullptr, nullptr};
	uint32_t drawIdx = 0;
	uint32_t nVerts = 0;   // > 0 only on the clipped path (polygon lives in poly[src])
	uint32_t src = 0;
	uint32_t validMask = 0;
	bool clipped = false;
	float hx = mulf((float)fp.width, 0.5f), hy = mulf((float)fp.height, 0.5f);

	if (g < fp.numInputTris)
	{
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
		uint32_t idx[3];
#pragma unroll
		for (int i = 0; i < 3; ++i)
		{
			idx[i] = fetch_index(d, t * 3 + i);
			const float* p = reinterpret_cast<const float*>(d.pos + (size_t)idx[i] * d.posStride);
			float const x = p[0], y = p[1], z = p[2];
			// kt::Mul(Mat4, Vec4) (kt/src/kt/inl/Mat4.inl:285-292): ((c0*x + c1*y) + c2*z) + c3*w, w = 1
			float r[4];
#pragma unroll
			for (int k = 0; k < 4; ++k)
			{
				r[k] = addf(addf(addf(mulf(d.mvp[k], x), mulf(d.mvp[4 + k], y)), mulf(d.mvp[8 + k], z)),
				            mulf(d.mvp[12 + k], 1.0f));
			}
			v[i] = make_float4(r[0], r[1], r[2], r[3]);
			ap[i] = reinterpret_cast<const float*>(d.attr + (size_t)idx[i] * d.attrStride);
		}
		uint32_t const c0 = clip_code(v[0].x, v[0].y, v[0].z, v[0].w);
		uint32_t const c1 = clip_code(v[1].x, v[1].y, v[1].z, v[1].w);
		uint32_t const c2 = clip_code(v[2].x, v[2].y, v[2].z, v[2].w);
		uint32_t maskOr = c0 | c1 | c2;
		if (maskOr == 0)
		{
			Snapped s;
			snap(v, hx, hy, s);
			validMask = front_facing(s) ? 1u : 0u;
		}
		else if ((c0 & c1 & c2) == 0)
		{
			// Binning.cpp:498-523
			clipped = true;
#pragma unroll
			for (int i = 0; i < 3; ++i)
			{
				ClipVert& cv = poly[0][i];
				cv.x = v[i].x; cv.y = v[i].y; cv.z = v[i].z; cv.w = v[i].w;
#pragma unroll
				for (int k = 0; k < SRB_MAX_VARY; ++k)
				{
					cv.a[k] = ((uint32_t)k < d.numVaryings) ? ap[i][k] : 0.0f;
				}
			}
			nVerts = 3;
			do
			{
				uint32_t const plane = __ffs(maskOr) - 1;
				maskOr ^= 1u << plane;
				nVerts = clip_plane(poly[src], nVerts, poly[src ^ 1], plane);
				src ^= 1;
			} while (maskOr && nVerts);
			// fan (0, i-1, i), Binning.cpp:526-533
			for (uint32_t i = 2; i < nVerts; ++i)
			{
				float4 f[3];
				const ClipVert& p0 = poly[src][0];
				const ClipVert& p1 = poly[src][i - 1];
				const ClipVert& p2 = poly[src][i];
				f[0] = make_float4(p0.x, p0.y, p0.z, p0.w);
				f[1] = make_float4(p1.x, p1.y, p1.z, p1.w);
				f[2] = make_float4(p2.x, p2.y, p2.z, p2.w);
				Snapped s;
				snap(f, hx, hy, s);
				if (front_facing(s))
				{
					validMask |= 1u << (i - 2);
				}
			}
		}
	}
	uint32_t const nOut = __popc(validMask);

	// ---- phase 2: ranks = block scan + decoupled look-back --------------------------------------------------
	uint32_t incl = nOut;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		uint32_t const n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
		if (lane >= (uint32_t)o) incl += n;
	}
	if (lane == 31) s_warpSum[warp] = incl;
	uint32_t const clippedInWarp = __popc(__ballot_sync(0xFFFFFFFFu, clipped));
	if (lane == 0 && clippedInWarp) atomicAdd(&ctl->numClipped, clippedInWarp);
	__syncthreads();
	uint32_t warpBase = 0, blockTotal = 0;
#pragma unroll
	for (int w = 0; w < kSetupThreads / 32; ++w)
	{
		uint32_t const ws = s_warpSum[w];
		if ((uint32_t)w < warp) warpBase += ws;
		blockTotal += ws;
	}
	if (warp == 0)
	{
		uint32_t exclusive = 0;
		if (vbid == 0)
		{
			if (lane == 0)
			{
				__threadfence();
				lookback[0] = kPre | blockTotal;
			}
		}
		else
		{
			if (lane == 0)
			{
				__threadfence();
				lookback[vbid] = kAgg | blockTotal;
			}
			int32_t base = (int32_t)vbid - 1;
			for (;;)
			{
				int32_t const j = base - (int32_t)lane;
				unsigned long long d = kPre; // lanes before block 0 read as "prefix 0"
				if (j >= 0)
				{
					do
					{
						d = lookback[j];
					} while ((d >> 32) == 0ull);
				}
				uint32_t const isPre = __ballot_sync(0xFFFFFFFFu, (d >> 32) == 2ull);
				uint32_t const first = isPre ? (uint32_t)(__ffs(isPre) - 1) : 32u;
				uint32_t val = (lane <= first) ? (uint32_t)d : 0u;
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xFFFFFFFFu, val, o);
				exclusive += val;
				if (isPre) break;
				base -= 32;
			}
			if (lane == 0)
			{
				__threadfence();
				lookback[vbid] = kPre | (unsigned long long)(exclusive + blockTotal);
			}
		}
		if (lane == 0)
		{
			s_blockBase = exclusive;
			if (vbid == gridDim.x - 1)
			{
				uint32_t const total = exclusive + blockTotal;
				ctl->numSetup = total;
				if (total > fp.setupCapacity) atomicOr(&ctl->overflow, 1u);
			}
		}
	}
	__syncthreads();
	uint32_t rank = s_blockBase + warpBase + (incl - nOut);

	// ---- phase 3: full set-up of the survivors ---------------------------------------------------------------
	if (validMask)
	{
		const DrawDev& d = draws[drawIdx];
		if (!clipped)
		{
			emit_triangle(v, ap[0], ap[1], ap[2], d, drawIdx, fp, rank, rasterRecs, shadeRecs, tileCounts);
		}
		else
		{
			for (uint32_t i = 2; i < nVerts; ++i)
			{
				if (validMask & (1u << (i - 2)))
				{
					const ClipVert& p0 = poly[src][0];
					const ClipVert& p1 = poly[src][i - 1];
					const ClipVert& p2 = poly[src][i];
					float4 f[3];
					f[0] = make_float4(p0.x, p0.y, p0.z, p0.w);
					f[1] = make_float4(p1.x, p1.y, p1.z, p1.w);
					f[2] = make_float4(p2.x, p2.y, p2.z, p2.w);
					emit_triangle(f, p0.a, p1.a, p2.a, d, drawIdx, fp, rank, rasterRecs, shadeRecs, tileCounts);
					++rank;
				}
			}
		}
	}
}

} // namespace

void launch_setup(const FrameParams& fp, const DrawDev* draws, RasterRec* rasterRecs, ShadeRec* shadeRecs,
                  uint32_t* tileCounts, unsigned long long* lookback, FrameCtl* ctl, cudaStream_t stream)
{
	if (fp.numInputTris == 0)
	{
		return;
	}
	uint32_t const blocks = (fp.numInputTris + kSetupThreads - 1) / kSetupThreads;
	setup_kernel<<<blocks, kSetupThreads, 0, stream>>>(fp, draws, rasterRecs, shadeRecs, tileCounts, lookback, ctl);
}

uint32_t setup_num_blocks(uint32_t numInputTris)
{
	return (numInputTris + kSetupThreads - 1) / kSetupThreads;
}

} // namespace srb
