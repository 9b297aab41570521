// srb_texbuild.cu — device-side texture builder: Tex::TextureData::CreateFromRGBA8 (reference
// SoftRast/Texture.cpp:119-199) on the GPU.  The linear RGBA8 image is uploaded once; level 0 is re-ordered into the
// reference's 32x32-tiled Morton layout (Texture.cpp:73-101) by a streaming kernel, and every further mip level is
// filtered FROM THE ORIGINAL IMAGE the way stbir_resize_uint8 does it (Texture.cpp:188-198; stb_image_resize.h: Mitchell
// kernel, clamped edges, down-sampling path) and written straight into its tiled place — texel for texel the bytes the
// reference stores.
//
// Exactness: stb's down-sampler SCATTERS — every input pixel (ascending) adds `in * coefficient` into the output pixels it
// contributes to (stbir__resample_horizontal_downsample :1526-1655), and every input row (ascending) adds its filtered row
// into the output rows it contributes to (stbir__resample_vertical_downsample :1987-2066).  For one output value that is
// a sum over its contributors in ascending order, starting from 0.0f, each term rounded as a product and then added.
// The kernels GATHER exactly that sequence (one thread per output value, contributors ascending, __fmul_rn + __fadd_rn),
// with the per-contributor coefficient tables computed on the host by the same code as the host builder
// (srb_host.cpp: srb_internal_stb_axis), so the floats — and after stb's encode ((int)(saturate(f) * 255.0f + 0.5 [double]))
// the bytes — are identical.  tests/test_gpu_texbuild.py compares every byte with the host builder and the reference.
#include <cstdlib>
#include "srb_kernels.h"

namespace srb
{

namespace
{

constexpr int kBatch = 8; // contributors whose loads are in flight together

// x bits in even positions, y bits in odd positions (Texture.cpp:36-41), 5 bits each
__device__ __forceinline__ uint32_t Spread5(uint32_t v) // abcde -> 0a0b0c0d0e
{
	v &= 31u;
	v = (v | (v << 4)) & 0x10Fu;  // a....bcde  -> bit 8 = a, bits 3..0 = bcde
	v = (v | (v << 2)) & 0x133u;  // a..bc..de
	v = (v | (v << 1)) & 0x155u;  // a.b.c.d.e
	return v;
}

__device__ __forceinline__ uint32_t TiledIndex(uint32_t x, uint32_t y, uint32_t tilesX)
{
	return ((y >> 5) * tilesX + (x >> 5)) * 1024u + (Spread5(x) | (Spread5(y) << 1));
}

// Level 0: one thread per 2x2 texel quad.  Morton order keeps a quad contiguous (indices 4q .. 4q+3 = (x,y), (x+1,y),
// (x,y+1), (x+1,y+1)), so a thread reads two 8-byte row pieces and writes one 16-byte vector; a warp covers a 16x8
// texel block (64-byte row segments in, 512 contiguous bytes out).  8 bytes of HBM traffic per texel.
__global__ void __launch_bounds__(256) tex_tile_kernel(const uint2* __restrict__ linear, uint4* __restrict__ dst, uint32_t w, uint32_t h)
{
	uint32_t const quadsPerTile = 256u, tilesX = w >> 5;
	uint32_t const q = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t const numQuads = (w >> 1) * (h >> 1);
	if (q >= numQuads) return;
	uint32_t const tile = q / quadsPerTile, inTile = q % quadsPerTile;
	// inverse of the Morton spread for the quad's (x/2, y/2) inside the tile: 4 bits each
	uint32_t qx = inTile & 0x55u, qy = (inTile >> 1) & 0x55u;
	qx = (qx | (qx >> 1)) & 0x33u;
	qx = (qx | (qx >> 2)) & 0x0Fu;
	qy = (qy | (qy >> 1)) & 0x33u;
	qy = (qy | (qy >> 2)) & 0x0Fu;
	uint32_t const x = (tile % tilesX) * 32u + qx * 2u, y = (tile / tilesX) * 32u + qy * 2u;
	uint2 const top = linear[(size_t(y) * w + x) >> 1];
	uint2 const bottom = linear[(size_t(y + 1) * w + x) >> 1];
	dst[q] = make_uint4(top.x, top.y, bottom.x, bottom.y);
}

// Horizontal pass of one level: hbuf[y * ow + k] = sum over output k's gather list (ascending contributor) of
// decode(pixel(entry.x)) * entry.coefficient.  One thread per (input row, output column), four channels.
// The additions form one dependent chain per channel (that IS the reference's rounding order); everything they consume
// is independent of it, so the entries and pixels of kBatch contributors are loaded before their additions.
__global__ void __launch_bounds__(128) tex_hpass_kernel(const uchar4* __restrict__ linear, float4* __restrict__ hbuf, int iw, int ih,
                                                        int ow, StbAxisDev H)
{
	__shared__ float decode[256];
	for (int i = threadIdx.x; i < 256; i += blockDim.x) decode[i] = __fdiv_rn((float)i, 255.0f); // stbir__decode_scanline
	__syncthreads();
	uint32_t const idx = blockIdx.x * blockDim.x + threadIdx.x; // flat over (row, output column): small levels still fill warps
	if (idx >= uint32_t(ih) * uint32_t(ow)) return;
	int const y = int(idx / uint32_t(ow)), k = int(idx % uint32_t(ow));
	const uchar4* row = linear + size_t(y) * iw;
	float r = 0.0f, g = 0.0f, b = 0.0f, a = 0.0f;
	int const e1 = H.off[k + 1];
	for (int e0 = H.off[k]; e0 < e1; e0 += kBatch)
	{
		int2 ent[kBatch];
		uchar4 px[kBatch];
#pragma unroll
		for (int u = 0; u < kBatch; ++u) ent[u] = H.ent[min(e0 + u, e1 - 1)];
#pragma unroll
		for (int u = 0; u < kBatch; ++u) px[u] = row[ent[u].x];
#pragma unroll
		for (int u = 0; u < kBatch; ++u)
		{
			bool const live = e0 + u < e1;
			float const co = __int_as_float(ent[u].y);
			float const tr = __fadd_rn(r, __fmul_rn(decode[px[u].x], co)), tg = __fadd_rn(g, __fmul_rn(decode[px[u].y], co));
			float const tb = __fadd_rn(b, __fmul_rn(decode[px[u].z], co)), ta = __fadd_rn(a, __fmul_rn(decode[px[u].w], co));
			r = live ? tr : r;
			g = live ? tg : g;
			b = live ? tb : b;
			a = live ? ta : a;
		}
	}
	hbuf[size_t(y) * ow + k] = make_float4(r, g, b, a);
}

__device__ __forceinline__ uint32_t StbEncode(float f) // stbir__encode_scanline, 8-bit linear
{
	f = f < 0.0f ? 0.0f : (f > 1.0f ? 1.0f : f);
	return (uint32_t)(int)__dadd_rn((double)__fmul_rn(f, 255.0f), 0.5) & 255u;
}

// Vertical pass + encode + tiling: out(ky, x) = sum over output row ky's gather list (ascending contributor row, margins =
// clamped rows) of hbuf[entry.x][x] * entry.coefficient; written at the texel's Morton place in the level.
__global__ void __launch_bounds__(128) tex_vpass_kernel(const float4* __restrict__ hbuf, uint32_t* __restrict__ dstLevel, int ih, int ow,
                                                        int oh, StbAxisDev V)
{
	uint32_t const idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= uint32_t(oh) * uint32_t(ow)) return;
	int const ky = int(idx / uint32_t(ow)), x = int(idx % uint32_t(ow));
	float r = 0.0f, g = 0.0f, b = 0.0f, a = 0.0f;
	int const e1 = V.off[ky + 1];
	for (int e0 = V.off[ky]; e0 < e1; e0 += kBatch)
	{
		int2 ent[kBatch];
		float4 p[kBatch];
#pragma unroll
		for (int u = 0; u < kBatch; ++u) ent[u] = V.ent[min(e0 + u, e1 - 1)];
#pragma unroll
		for (int u = 0; u < kBatch; ++u) p[u] = hbuf[size_t(ent[u].x) * ow + x];
#pragma unroll
		for (int u = 0; u < kBatch; ++u)
		{
			bool const live = e0 + u < e1;
			float const co = __int_as_float(ent[u].y);
			float const tr = __fadd_rn(r, __fmul_rn(p[u].x, co)), tg = __fadd_rn(g, __fmul_rn(p[u].y, co));
			float const tb = __fadd_rn(b, __fmul_rn(p[u].z, co)), ta = __fadd_rn(a, __fmul_rn(p[u].w, co));
			r = live ? tr : r;
			g = live ? tg : g;
			b = live ? tb : b;
			a = live ? ta : a;
		}
	}
	uint32_t const tilesX = (uint32_t(ow) + 31u) >> 5;
	dstLevel[TiledIndex(uint32_t(x), uint32_t(ky), tilesX)] = StbEncode(r) | (StbEncode(g) << 8) | (StbEncode(b) << 16) | (StbEncode(a) << 24);
}

// One block per 16 KB chunk of a segment (the block finds its segment by binary search over firstBlock): every array of
// the frame is in flight at once, four 16-byte loads per thread issued before the first store.  Host memory is read
// with ld.volatile-like loads (no stale lines: the application may rewrite its arrays between frames).
__global__ void __launch_bounds__(256) gather_kernel(const GatherSeg* __restrict__ segs, uint32_t n)
{
	uint32_t lo = 0, hi = n;
	while (hi - lo > 1u)
	{
		uint32_t const mid = (lo + hi) >> 1;
		if (segs[mid].firstBlock <= blockIdx.x) lo = mid; else hi = mid;
	}
	GatherSeg const g = segs[lo];
	size_t const base = size_t(blockIdx.x - g.firstBlock) * kGatherChunk;
	if (base >= g.bytes) return;
	size_t const len = min(size_t(kGatherChunk), size_t(g.bytes) - base);
	const uint8_t* srcB = g.src + base;
	uint8_t* dstB = g.dst + base;
	if ((((size_t)srcB | (size_t)dstB | len) & 15u) == 0u)
	{
		const uint4* src = reinterpret_cast<const uint4*>(srcB);
		uint4* dst = reinterpret_cast<uint4*>(dstB);
		uint32_t const n16 = uint32_t(len >> 4);
		uint4 v[4];
#pragma unroll
		for (int k = 0; k < 4; ++k)
		{
			uint32_t const i = threadIdx.x + uint32_t(k) * 256u;
			if (i < n16) v[k] = __ldcv(src + i);
		}
#pragma unroll
		for (int k = 0; k < 4; ++k)
		{
			uint32_t const i = threadIdx.x + uint32_t(k) * 256u;
			if (i < n16) dst[i] = v[k];
		}
	}
	else if ((((size_t)srcB | (size_t)dstB | len) & 3u) == 0u)
	{
		const uint32_t* src = reinterpret_cast<const uint32_t*>(srcB);
		uint32_t* dst = reinterpret_cast<uint32_t*>(dstB);
		for (uint32_t i = threadIdx.x, n4 = uint32_t(len >> 2); i < n4; i += 256u) dst[i] = __ldcv(src + i);
	}
	else
	{
		for (uint32_t i = threadIdx.x; i < uint32_t(len); i += 256u) dstB[i] = __ldcv(srcB + i);
	}
}

// The same gather on bulk asynchronous copies: one warp per 16 KB chunk; lane 0 has the copy engine of the SM fetch the
// chunk from host memory into shared memory (one request stream of large reads instead of 1024 16-byte loads) and store
// it to the mirror.  Chunks that are not 16-byte aligned take the plain loops.
__global__ void __launch_bounds__(32) gather_bulk_kernel(const GatherSeg* __restrict__ segs, uint32_t n)
{
	__shared__ __align__(128) uint8_t s_chunk[kGatherChunk];
	__shared__ __align__(8) unsigned long long s_mbar;
	uint32_t lo = 0, hi = n;
	while (hi - lo > 1u)
	{
		uint32_t const mid = (lo + hi) >> 1;
		if (segs[mid].firstBlock <= blockIdx.x) lo = mid; else hi = mid;
	}
	GatherSeg const g = segs[lo];
	size_t const base = size_t(blockIdx.x - g.firstBlock) * kGatherChunk;
	if (base >= g.bytes) return;
	uint32_t const len = (uint32_t)min(size_t(kGatherChunk), size_t(g.bytes) - base);
	const uint8_t* srcB = g.src + base;
	uint8_t* dstB = g.dst + base;
	if ((((size_t)srcB | (size_t)dstB | len) & 15u) != 0u)
	{
		for (uint32_t i = threadIdx.x; i < len; i += 32u) dstB[i] = __ldcv(srcB + i);
		return;
	}
	uint32_t const mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
	uint32_t const sm = (uint32_t)__cvta_generic_to_shared(s_chunk);
	if (threadIdx.x == 0)
	{
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(1));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(len) : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sm), "l"(srcB),
		             "r"(len), "r"(mbar)
		             : "memory");
		uint32_t done = 0;
		while (!done)
		{
			asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
			             : "=r"(done)
			             : "r"(mbar), "r"(0)
			             : "memory");
		}
		asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dstB), "r"(sm), "r"(len) : "memory");
		asm volatile("cp.async.bulk.commit_group;" ::: "memory");
		asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); // the mirror is complete when the kernel is
	}
}

} // namespace

uint32_t gather_plan(GatherSeg* segs, uint32_t n)
{
	uint32_t blocks = 0;
	for (uint32_t i = 0; i < n; ++i)
	{
		segs[i].firstBlock = blocks;
		segs[i].pad = 0;
		blocks += uint32_t((segs[i].bytes + kGatherChunk - 1) / kGatherChunk);
	}
	return blocks;
}

void launch_gather(const GatherSeg* segs, uint32_t n, uint32_t blocks, cudaStream_t stream)
{
	static bool const bulk = getenv("SRB_GATHER_BULK") != nullptr; // experiment (profiles/README.md)
	if (!(n && blocks)) return;
	if (bulk) gather_bulk_kernel<<<blocks, 32, 0, stream>>>(segs, n);
	else gather_kernel<<<blocks, 256, 0, stream>>>(segs, n);
}

void launch_tex_tile(const uint8_t* linear, uint8_t* dstLevel, uint32_t w, uint32_t h, cudaStream_t stream)
{
	uint32_t const quads = (w >> 1) * (h >> 1);
	tex_tile_kernel<<<(quads + 255u) / 256u, 256, 0, stream>>>(reinterpret_cast<const uint2*>(linear), reinterpret_cast<uint4*>(dstLevel), w, h);
}

void launch_tex_hpass(const uint8_t* linear, float* hbuf, int iw, int ih, int ow, const StbAxisDev& H, cudaStream_t stream)
{
	uint32_t const grid = (uint32_t(ow) * uint32_t(ih) + 127u) / 128u;
	tex_hpass_kernel<<<grid, 128, 0, stream>>>(reinterpret_cast<const uchar4*>(linear), reinterpret_cast<float4*>(hbuf), iw, ih, ow, H);
}

void launch_tex_vpass(const float* hbuf, uint8_t* dstLevel, int ih, int ow, int oh, const StbAxisDev& V, cudaStream_t stream)
{
	uint32_t const grid = (uint32_t(ow) * uint32_t(oh) + 127u) / 128u;
	tex_vpass_kernel<<<grid, 128, 0, stream>>>(reinterpret_cast<const float4*>(hbuf), reinterpret_cast<uint32_t*>(dstLevel), ih, ow, oh, V);
}

} // namespace srb
