// srb_bin.cu — K2: binning.  Replaces the reference's per-thread bins (BinContext / ThreadBin / BinChunk,
// SoftRast/Binning.h:15-84, append loop Binning.cpp:372-455) and the per-tile gather + stable radix sort by draw index
// (SoftRast/Rasterizer.cpp:538-553).
//
//   tile_scan_kernel : exclusive prefix sum of the per-tile reference counts that K1 produced -> list offsets
//   bin_fill_kernel  : one thread per set-up triangle, warp-aggregated atomic append of its RANK to every tile list
//   tile_sort_kernel : per-tile ascending sort of the ranks.  rank == (draw, triangle, fan) order, so the sorted list
//                      is exactly the reference's single-threaded per-tile order, whatever order the atomics ran in.
#include "srb_device.cuh"
#include "srb_kernels.h"

namespace srb
{

namespace
{

constexpr int kScanThreads = 1024;

__global__ void __launch_bounds__(kScanThreads) tile_scan_kernel(uint32_t numTiles, const uint32_t* __restrict__ counts,
                                                                 uint32_t* __restrict__ offsets,
                                                                 uint32_t* __restrict__ cursors, FrameCtl* ctl,
                                                                 uint32_t refCapacity)
{
	__shared__ uint32_t s_warp[32];
	__shared__ uint32_t s_carry;
	__shared__ uint32_t s_max[32];
	__shared__ uint32_t s_nz[32];
	uint32_t const tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	if (tid == 0) s_carry = 0;
	uint32_t localMax = 0, localNz = 0;
	__syncthreads();
	for (uint32_t base = 0; base < numTiles; base += kScanThreads)
	{
		uint32_t const i = base + tid;
		uint32_t const c = i < numTiles ? counts[i] : 0u;
		localMax = max(localMax, c);
		localNz += c ? 1u : 0u;
		uint32_t incl = c;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			uint32_t const n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
			if (lane >= (uint32_t)o) incl += n;
		}
		if (lane == 31) s_warp[warp] = incl;
		__syncthreads();
		if (warp == 0)
		{
			uint32_t w = s_warp[lane];
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				uint32_t const n = __shfl_up_sync(0xFFFFFFFFu, w, o);
				if (lane >= (uint32_t)o) w += n;
			}
			s_warp[lane] = w; // inclusive over warps
		}
		__syncthreads();
		uint32_t const carry = s_carry;
		uint32_t const excl = carry + (warp ? s_warp[warp - 1] : 0u) + incl - c;
		if (i < numTiles)
		{
			offsets[i] = excl;
			cursors[i] = 0;
		}
		__syncthreads();
		if (tid == kScanThreads - 1) s_carry = carry + s_warp[31];
		__syncthreads();
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		localMax = max(localMax, __shfl_xor_sync(0xFFFFFFFFu, localMax, o));
		localNz += __shfl_xor_sync(0xFFFFFFFFu, localNz, o);
	}
	if (lane == 0)
	{
		s_max[warp] = localMax;
		s_nz[warp] = localNz;
	}
	__syncthreads();
	if (tid == 0)
	{
		uint32_t m = 0, nz = 0;
		for (int w = 0; w < 32; ++w)
		{
			m = max(m, s_max[w]);
			nz += s_nz[w];
		}
		offsets[numTiles] = s_carry;
		ctl->totalRefs = s_carry;
		ctl->maxRefs = m;
		ctl->tilesNonEmpty = nz;
		if (s_carry > refCapacity) atomicOr(&ctl->overflow, 2u);
	}
}

constexpr int kFillThreads = 256;

__global__ void __launch_bounds__(kFillThreads) bin_fill_kernel(FrameParams fp, const RasterRec* __restrict__ recs,
                                                                const uint32_t* __restrict__ offsets,
                                                                uint32_t* __restrict__ cursors,
                                                                uint32_t* __restrict__ refs,
                                                                const FrameCtl* __restrict__ ctl)
{
	uint32_t const numSetup = min(ctl->numSetup, fp.setupCapacity);
	if (ctl->totalRefs > fp.refCapacity)
	{
		return; // overflow: the host grows the buffers and re-runs the frame
	}
	uint32_t const lane = threadIdx.x & 31u;
	// whole warps iterate together so that the warp-aggregated append below can use full-mask collectives
	for (uint32_t warpBase = (blockIdx.x * kFillThreads + threadIdx.x) & ~31u; warpBase < numSetup;
	     warpBase += gridDim.x * kFillThreads)
	{
		uint32_t const rank = warpBase + lane;
		bool const live = rank < numSetup;
		int32_t c[3], dx[3], dy[3];
		BinRange br;
		br.bx0 = br.by0 = 1;
		br.bx1 = br.by1 = 0;
		br.check = false;
		if (live)
		{
			const uint4* p = reinterpret_cast<const uint4*>(recs + rank);
			uint4 const q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2);
			c[0] = q0.x; c[1] = q0.y; c[2] = q0.z; dx[0] = q0.w;
			dx[1] = q1.x; dx[2] = q1.y; dy[0] = q1.z; dy[1] = q1.w;
			dy[2] = q2.x;
			uint32_t const xmin = q2.y & 0xFFFFu, xmax = q2.y >> 16, ymin = q2.z & 0xFFFFu, ymax = q2.z >> 16;
			br = bin_range(xmin, xmax, ymin, ymax);
		}
		uint32_t const nbx = live ? br.bx1 - br.bx0 + 1 : 0u;
		uint32_t const nb = live ? nbx * (br.by1 - br.by0 + 1) : 0u;
		uint32_t const maxNb = __reduce_max_sync(0xFFFFFFFFu, nb);
		for (uint32_t k = 0; k < maxNb; ++k)
		{
			uint32_t tile = 0xFFFFFFFFu;
			if (k < nb)
			{
				uint32_t const by = br.by0 + k / nbx, bx = br.bx0 + k % nbx;
				if (!br.check || bin_overlaps(c, dx, dy, (int32_t)(bx * SRB_TILE), (int32_t)(by * SRB_TILE)))
				{
					tile = by * fp.tilesX + bx;
				}
			}
			// warp-aggregated append: lanes that target the same tile share one atomic; lower lanes (= lower
			// ranks) get lower slots
			uint32_t const peers = __match_any_sync(0xFFFFFFFFu, tile);
			if (tile != 0xFFFFFFFFu)
			{
				uint32_t const leader = __ffs(peers) - 1;
				uint32_t base = 0;
				if (lane == leader)
				{
					base = atomicAdd(&cursors[tile], (uint32_t)__popc(peers));
				}
				base = __shfl_sync(peers, base, leader);
				uint32_t const slot = base + __popc(peers & ((1u << lane) - 1u));
				refs[offsets[tile] + slot] = rank;
			}
		}
	}
}

// Ascending sort of one tile's list.  Bitonic network in shared memory for lists up to kSortSmem entries, in global
// memory (same CTA) beyond that.
constexpr int kSortThreads = 256;
constexpr uint32_t kSortSmem = 8192;

__global__ void __launch_bounds__(kSortThreads) tile_sort_kernel(const uint32_t* __restrict__ offsets,
                                                                 uint32_t* __restrict__ refs,
                                                                 const FrameCtl* __restrict__ ctl, uint32_t refCapacity)
{
	__shared__ uint32_t s_keys[kSortSmem];
	__shared__ int s_unsorted;
	if (ctl->totalRefs > refCapacity)
	{
		return;
	}
	uint32_t const tile = blockIdx.x;
	uint32_t const begin = offsets[tile];
	uint32_t const n = offsets[tile + 1] - begin;
	if (n < 2)
	{
		return;
	}
	uint32_t* list = refs + begin;
	uint32_t const tid = threadIdx.x;
	// already sorted? (the warp-aggregated fill keeps most lists in order)
	if (tid == 0) s_unsorted = 0;
	__syncthreads();
	int bad = 0;
	for (uint32_t i = tid + 1; i < n; i += kSortThreads)
	{
		bad |= list[i - 1] > list[i];
	}
	if (bad) s_unsorted = 1;
	__syncthreads();
	if (!s_unsorted)
	{
		return;
	}
	uint32_t p2 = 1;
	while (p2 < n) p2 <<= 1;
	if (p2 <= kSortSmem)
	{
		for (uint32_t i = tid; i < p2; i += kSortThreads)
		{
			s_keys[i] = i < n ? list[i] : 0xFFFFFFFFu;
		}
		__syncthreads();
		for (uint32_t k = 2; k <= p2; k <<= 1)
		{
			for (uint32_t j = k >> 1; j > 0; j >>= 1)
			{
				for (uint32_t i = tid; i < p2; i += kSortThreads)
				{
					uint32_t const ixj = i ^ j;
					if (ixj > i)
					{
						uint32_t const a = s_keys[i], b = s_keys[ixj];
						bool const up = (i & k) == 0;
						if ((a > b) == up)
						{
							s_keys[i] = b;
							s_keys[ixj] = a;
						}
					}
				}
				__syncthreads();
			}
		}
		for (uint32_t i = tid; i < n; i += kSortThreads)
		{
			list[i] = s_keys[i];
		}
	}
	else
	{
		// Large list: bitonic network over a virtual power-of-two array; indices >= n behave as +infinity and are
		// never materialised (a compare-exchange with +infinity in the upper slot of an ascending pair is a no-op;
		// in a descending pair it must move the real key up, which would leave the array — so the network is run
		// in its all-ascending form ("sorting network with flips"), where the partner of i in the first step of
		// each stage is mirrored: i ^ (2k-1) for the top half-cleaner, then plain i ^ j).
		for (uint32_t k = 2; k <= p2; k <<= 1)
		{
			for (uint32_t j = k >> 1; j > 0; j >>= 1)
			{
				bool const first = (j == (k >> 1));
				for (uint32_t i = tid; i < p2; i += kSortThreads)
				{
					uint32_t const partner = first ? (i ^ (k - 1)) : (i ^ j);
					if (partner > i && partner < n)
					{
						uint32_t const a = list[i], b = list[partner];
						if (a > b)
						{
							list[i] = b;
							list[partner] = a;
						}
					}
				}
				__syncthreads();
			}
		}
	}
}

} // namespace

void launch_tile_scan(uint32_t numTiles, const uint32_t* counts, uint32_t* offsets, uint32_t* cursors, FrameCtl* ctl,
                      uint32_t refCapacity, cudaStream_t stream)
{
	tile_scan_kernel<<<1, kScanThreads, 0, stream>>>(numTiles, counts, offsets, cursors, ctl, refCapacity);
}

void launch_bin_fill(const FrameParams& fp, const RasterRec* recs, const uint32_t* offsets, uint32_t* cursors,
                     uint32_t* refs, const FrameCtl* ctl, cudaStream_t stream)
{
	if (fp.numInputTris == 0)
	{
		return;
	}
	uint32_t blocks = (fp.numInputTris + kFillThreads - 1) / kFillThreads;
	if (blocks > 148u * 16u) blocks = 148u * 16u;
	bin_fill_kernel<<<blocks, kFillThreads, 0, stream>>>(fp, recs, offsets, cursors, refs, ctl);
}

void launch_tile_sort(uint32_t numTiles, const uint32_t* offsets, uint32_t* refs, const FrameCtl* ctl,
                      uint32_t refCapacity, cudaStream_t stream)
{
	tile_sort_kernel<<<numTiles, kSortThreads, 0, stream>>>(offsets, refs, ctl, refCapacity);
}

} // namespace srb
