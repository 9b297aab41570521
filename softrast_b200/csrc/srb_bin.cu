// srb_bin.cu — K2: binning.  Replaces the reference's per-thread bins (BinContext / ThreadBin / BinChunk,
// SoftRast/Binning.h:15-84, append loop Binning.cpp:372-455) and the per-tile gather + stable radix sort by draw index
// (SoftRast/Rasterizer.cpp:538-553).
//
//   tile scan        : (srb_scan.cuh; runs in the tail of the set-up kernel, or as tile_scan_kernel) exclusive prefix sum
//                      of the per-tile reference counts that K1 produced -> list offsets; cuts every tile's list into
//                      work units of at most `unitSize` references for the raster kernel and re-zeroes the counters.
//   bin_fill_kernel  : CTA-aggregated atomic append of every surviving triangle's (key, slot, block range) to the list of every tile
//                      the reference would bin it to (same overlap decisions, incl. its quirks): shared-memory
//                      histogram, one global atomic per (CTA, touched tile) to reserve a range, shared-memory cursors.
// Lists are SETS: the order inside a list is whatever order the atomics ran in.  The reference's per-tile order
// (stable sort by draw, then binning order) is the order of the canonical keys stored in every entry; the rasteriser
// resolves depth ties by key, so its result is that of the ordered walk, and the parity dumps sort by key.
#include "srb_device.cuh"
#include "srb_kernels.h"
#include "srb_scan.cuh"

namespace srb
{

namespace
{

constexpr int kScanThreads = 1024;

// The tile scan as a kernel of its own (frames without triangles, and SRB_SEPARATE_SCAN=1 for A/B measurements); normally it
// runs in the tail of the set-up kernel (srb_setup.cu).
__global__ void __launch_bounds__(kScanThreads) tile_scan_kernel(FrameParams fp, uint32_t* __restrict__ counts,
                                                                 uint32_t* __restrict__ offsets,
                                                                 uint32_t* __restrict__ cursors,
                                                                 UnitDesc* __restrict__ units, FrameCtl* ctl)
{
	extern __shared__ uint32_t s_counts[]; // [numTiles] if it fits (fp.smemHist)
	tile_scan_block<kScanThreads>(fp, counts, offsets, cursors, units, ctl, fp.smemHist ? s_counts : nullptr);
}

constexpr int kFillThreads = 512;
constexpr uint32_t kFillCtas = 148u * 2u;

// Visits every tile the reference appends this triangle to (Binning.cpp:352-410), in its loop order.
// f(tile, blocks): blocks = the packed 8x8-block range of the triangle inside that tile (TileRef::blocks).
template <typename F>
__device__ __forceinline__ void for_each_bin(const RasterRec* __restrict__ recs, uint32_t slot, const FrameParams& fp, F&& f)
{
	const uint4* p = reinterpret_cast<const uint4*>(recs + slot);
	uint4 const q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2);
	int32_t c[3], dx[3], dy[3];
	c[0] = q0.x; c[1] = q0.y; c[2] = q0.z; dx[0] = q0.w;
	dx[1] = q1.x; dx[2] = q1.y; dy[0] = q1.z; dy[1] = q1.w;
	dy[2] = q2.x;
	uint32_t const xmin = q2.y & 0xFFFFu, xmax = q2.y >> 16, ymin = q2.z & 0xFFFFu, ymax = q2.z >> 16;
	BinRange const br = bin_range(xmin, xmax, ymin, ymax);
	for (uint32_t by = br.by0; by <= br.by1; ++by)
	{
		for (uint32_t bx = br.bx0; bx <= br.bx1; ++bx)
		{
			uint32_t const tile = by * fp.tilesX + bx;
			if (tile_owned(fp, tile) &&
			    (!br.check || bin_overlaps(c, dx, dy, (int32_t)(bx * SRB_TILE), (int32_t)(by * SRB_TILE))))
			{
				// block bbox relative to the tile, Binning.cpp:429-433
				int32_t const X0 = (int32_t)(bx * SRB_TILE), Y0 = (int32_t)(by * SRB_TILE);
				f(tile, pack_block_range(clampi((int32_t)xmin - X0, 0, SRB_TILE), clampi((int32_t)xmax - X0, 0, SRB_TILE),
				                         clampi((int32_t)ymin - Y0, 0, SRB_TILE), clampi((int32_t)ymax - Y0, 0, SRB_TILE)));
			}
		}
	}
}

// Each CTA owns a contiguous chunk of the survivor list and appends it in two passes over shared-memory counters:
// pass A histograms the chunk per tile; then ONE global atomic per touched tile reserves the chunk's range in that
// tile's list (atomics with a return value to one address serialise at L2 latency, so their number per tile is what
// bounds this kernel); pass B hands out the slots inside the reserved ranges with shared-memory atomics.
__global__ void __launch_bounds__(kFillThreads) bin_fill_kernel(FrameParams fp, const RasterRec* __restrict__ recs,
                                                                const Survivor* __restrict__ survivors,
                                                                const uint32_t* __restrict__ offsets,
                                                                uint32_t* __restrict__ cursors,
                                                                TileRef* __restrict__ refs,
                                                                const FrameCtl* __restrict__ ctl)
{
	extern __shared__ uint32_t s_fill[]; // [numTiles] counts / local cursors, [numTiles] reserved bases
	if (ctl->overflow)
	{
		return; // the host grows the buffers and re-runs the frame
	}
	uint32_t const numTiles = fp.tilesX * fp.tilesY;
	uint32_t* s_count = s_fill;
	uint32_t* s_base = s_fill + numTiles;
	uint32_t const numSurvivors = ctl->numSurvivors;
	uint32_t const per = (numSurvivors + gridDim.x - 1) / gridDim.x;
	uint32_t const c0 = min(numSurvivors, blockIdx.x * per), c1 = min(numSurvivors, c0 + per);
	if (c0 == c1)
	{
		return;
	}
	uint32_t const tid = threadIdx.x;
	for (uint32_t i = tid; i < numTiles; i += kFillThreads) s_count[i] = 0;
	__syncthreads();
	for (uint32_t i = c0 + tid; i < c1; i += kFillThreads)
	{
		uint4 const me = __ldg(reinterpret_cast<const uint4*>(survivors + i));
		if (me.z & 0x80000000u)
		{
			atomicAdd(&s_count[me.z & 0x7FFFFFFFu], 1u); // one tile: everything is in the survivor entry
		}
		else
		{
			for_each_bin(recs, me.y, fp, [&](uint32_t tile, uint32_t) { atomicAdd(&s_count[tile], 1u); });
		}
	}
	__syncthreads();
	for (uint32_t i = tid; i < numTiles; i += kFillThreads)
	{
		uint32_t const n = s_count[i];
		if (n)
		{
			s_base[i] = offsets[i] + atomicAdd(&cursors[i], n);
			s_count[i] = 0;
		}
	}
	__syncthreads();
	for (uint32_t i = c0 + tid; i < c1; i += kFillThreads)
	{
		uint4 const me = __ldg(reinterpret_cast<const uint4*>(survivors + i)); // key, slot, tile, blocks
		if (me.z & 0x80000000u)
		{
			uint32_t const tile = me.z & 0x7FFFFFFFu;
			uint32_t const k = atomicAdd(&s_count[tile], 1u);
			*reinterpret_cast<uint4*>(&refs[s_base[tile] + k]) = make_uint4(me.x, me.y, me.w, quad_mask(me.w));
		}
		else
		{
			for_each_bin(recs, me.y, fp, [&](uint32_t tile, uint32_t blocks) {
				uint32_t const k = atomicAdd(&s_count[tile], 1u);
				*reinterpret_cast<uint4*>(&refs[s_base[tile] + k]) = make_uint4(me.x, me.y, blocks, quad_mask(blocks));
			});
		}
	}
}

// Frames with more tiles than the shared-memory counters hold (framebuffers beyond ~8K x 8K): one warp-aggregated global
// atomic per reference instead.  Same lists (they are sets), slower.
__global__ void __launch_bounds__(256) bin_fill_direct_kernel(FrameParams fp, const RasterRec* __restrict__ recs,
                                                              const Survivor* __restrict__ survivors,
                                                              const uint32_t* __restrict__ offsets,
                                                              uint32_t* __restrict__ cursors, TileRef* __restrict__ refs,
                                                              const FrameCtl* __restrict__ ctl)
{
	if (ctl->overflow)
	{
		return;
	}
	uint32_t const numSurvivors = ctl->numSurvivors;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < numSurvivors; i += gridDim.x * blockDim.x)
	{
		Survivor const me = survivors[i];
		for_each_bin(recs, me.slot, fp, [&](uint32_t tile, uint32_t blocks) {
			uint32_t const k = atomicAdd(&cursors[tile], 1u);
			*reinterpret_cast<uint4*>(&refs[offsets[tile] + k]) = make_uint4(me.key, me.slot, blocks, quad_mask(blocks));
		});
	}
}

} // namespace

cudaError_t bin_init()
{
	cudaError_t const e = cudaFuncSetAttribute(tile_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
	if (e != cudaSuccess) return e;
	return cudaFuncSetAttribute(bin_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
}

void launch_tile_scan(const FrameParams& fp, uint32_t* counts, uint32_t* offsets, uint32_t* cursors, UnitDesc* units,
                      FrameCtl* ctl, cudaStream_t stream)
{
	size_t const smem = fp.smemHist ? size_t(fp.tilesX) * fp.tilesY * sizeof(uint32_t) : 0;
	tile_scan_kernel<<<1, kScanThreads, smem, stream>>>(fp, counts, offsets, cursors, units, ctl);
}

bool launch_bin_fill(const FrameParams& fp, const RasterRec* recs, const Survivor* survivors, const uint32_t* offsets,
                     uint32_t* cursors, TileRef* refs, const FrameCtl* ctl, cudaStream_t stream)
{
	if (fp.numInputTris == 0)
	{
		return false;
	}
	uint32_t blocks = (fp.numInputTris + 1023u) / 1024u; // at least ~1k input triangles per CTA
	blocks = blocks < 1 ? 1 : (blocks > kFillCtas ? kFillCtas : blocks);
	size_t const smem = size_t(fp.tilesX) * fp.tilesY * 2 * sizeof(uint32_t);
	if (smem > 96 * 1024)
	{
		bin_fill_direct_kernel<<<kFillCtas * 2u, 256, 0, stream>>>(fp, recs, survivors, offsets, cursors, refs, ctl);
		return true;
	}
	bin_fill_kernel<<<blocks, kFillThreads, smem, stream>>>(fp, recs, survivors, offsets, cursors, refs, ctl);
	return true;
}

} // namespace srb
