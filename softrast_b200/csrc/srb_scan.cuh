// srb_scan.cuh — the per-tile scan of K2 as a block-level device function, so that it can run either as the TAIL of the
// set-up kernel (the last CTA to finish runs it: no launch, no idle GPU between the two) or as a kernel of its own.
//
// Replaces what the reference gets for free from its per-thread bins (SoftRast/Binning.h:15-84): an exclusive prefix
// sum of the per-tile reference counts K1 produced -> list offsets; every tile's list is cut into work units of at most
// `unitSize` references for the raster kernel (heaviest tiles first) and the counters are re-zeroed for the next frame.
#pragma once
#include "srb_device.cuh"

namespace srb
{

// block-wide inclusive scan: returns the inclusive value, *total = sum over the block.  s_warp: 32 words.
template <int kThreads>
__device__ __forceinline__ uint32_t block_scan_incl(uint32_t v, uint32_t* s_warp, uint32_t* total)
{
	constexpr uint32_t kWarps = kThreads / 32;
	uint32_t const lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	uint32_t incl = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		uint32_t const n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
		if (lane >= (uint32_t)o) incl += n;
	}
	__syncthreads(); // s_warp reuse
	if (lane == 31) s_warp[warp] = incl;
	__syncthreads();
	if (warp == 0)
	{
		uint32_t w = lane < kWarps ? s_warp[lane] : 0u;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			uint32_t const n = __shfl_up_sync(0xFFFFFFFFu, w, o);
			if (lane >= (uint32_t)o) w += n;
		}
		s_warp[lane] = w;
	}
	__syncthreads();
	*total = s_warp[31];
	return incl + (warp ? s_warp[warp - 1] : 0u);
}

// All kThreads threads of ONE block call this.  `sc` = numTiles words of shared memory for the counts (nullptr: the
// counts are re-read from global memory in every pass — frames with more tiles than shared memory holds).
// The counts must be visible to this block (they were produced by atomics of other blocks: the caller fences).
template <int kThreads>
__device__ void tile_scan_block(const FrameParams& fp, uint32_t* __restrict__ counts, uint32_t* __restrict__ offsets,
                                uint32_t* __restrict__ cursors, UnitDesc* __restrict__ units, FrameCtl* ctl, uint32_t* sc)
{
	__shared__ uint32_t s_warp[32];
	__shared__ uint32_t s_max[32];
	__shared__ uint32_t s_nz[32];
	__shared__ uint32_t s_cls[32];
	__shared__ uint32_t s_unitSize;
	constexpr uint32_t kWarps = kThreads / 32;
	uint32_t const numTiles = fp.tilesX * fp.tilesY;
	uint32_t const tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

	// pass 1: offsets; the counters are read once (L2) and re-zeroed for the next frame's K1
	uint32_t carry = 0, localMax = 0, localNz = 0;
	for (uint32_t base = 0; base < numTiles; base += kThreads)
	{
		uint32_t const i = base + tid;
		uint32_t const c = i < numTiles ? __ldcg(counts + i) : 0u;
		localMax = max(localMax, c);
		localNz += c ? 1u : 0u;
		uint32_t total;
		uint32_t const incl = block_scan_incl<kThreads>(c, s_warp, &total);
		if (i < numTiles)
		{
			offsets[i] = carry + incl - c;
			cursors[i] = 0;
			if (sc)
			{
				sc[i] = c;
				counts[i] = 0;
			}
		}
		carry += total;
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
	if (tid < 32) s_cls[tid] = 0;
	__syncthreads();
	uint32_t const totalRefs = carry;
	if (tid == 0)
	{
		uint32_t m = 0, nz = 0;
		for (uint32_t w = 0; w < kWarps; ++w)
		{
			m = max(m, s_max[w]);
			nz += s_nz[w];
		}
		offsets[numTiles] = totalRefs;
		ctl->totalRefs = totalRefs;
		ctl->maxRefs = m;
		ctl->tilesNonEmpty = nz;
		if (totalRefs > fp.refCapacity) atomicOr(&ctl->overflow, 2u);
		// unit size: small enough that the heaviest tile spreads over many warps, large enough that most tiles stay one
		// unit (a split tile pays a merge through global atomics)
		uint32_t u = 0xFFFFFFFFu;
		if (fp.splitTiles)
		{
			u = max(fp.minUnit, (totalRefs / 640u + 31u) & ~31u);
		}
		s_unitSize = u;
		ctl->unitSize = u;
	}
	__syncthreads();
	uint32_t const unitSize = s_unitSize;

	// pass 2: units, HEAVIEST TILES FIRST.  The rasteriser's warps pull units from a dispenser in index order; a unit of
	// a crowded tile takes the longest, so it must not be the one that starts last.  Units are grouped by the size
	// class (log2) of their tile's reference count, classes in descending order, arbitrary order inside a class.
	for (uint32_t base = 0; base < numTiles; base += kThreads)
	{
		uint32_t const i = base + tid;
		if (i < numTiles)
		{
			uint32_t const c = sc ? sc[i] : __ldcg(counts + i);
			if (c)
			{
				atomicAdd(&s_cls[31 - __clz(c)], (c - 1u) / unitSize + 1u); // empty tiles are cleared by the shade kernel
			}
		}
	}
	__syncthreads();
	if (tid == 0)
	{
		uint32_t run = 0;
		for (int k = 31; k >= 0; --k)
		{
			uint32_t const n = s_cls[k];
			s_cls[k] = run;
			run += n;
		}
		s_unitSize = run; // total number of units (unitSize already lives in a register)
	}
	__syncthreads();
	uint32_t const numUnits = s_unitSize;
	for (uint32_t base = 0; base < numTiles; base += kThreads)
	{
		uint32_t const i = base + tid;
		if (i < numTiles)
		{
			uint32_t const c = sc ? sc[i] : __ldcg(counts + i);
			if (!sc) counts[i] = 0; // ready for the next frame's K1
			if (c)
			{
				uint32_t const nu = (c - 1u) / unitSize + 1u;
				uint32_t const first = atomicAdd(&s_cls[31 - __clz(c)], nu);
				uint32_t const begin = offsets[i]; // written by this very thread in pass 1
				for (uint32_t k = 0; k < nu; ++k)
				{
					if (first + k < fp.unitCapacity)
					{
						UnitDesc d;
						d.tile = i;
						d.begin = begin + k * unitSize;
						// unitSize is 0xFFFFFFFF when tiles must not be split (no depth clear): no 32-bit overflow here
						d.end = begin + (uint32_t)min((unsigned long long)c, (unsigned long long)(k + 1u) * unitSize);
						d.unitsInTile = nu;
						*reinterpret_cast<uint4*>(units + first + k) = make_uint4(d.tile, d.begin, d.end, d.unitsInTile);
					}
				}
			}
		}
	}
	if (tid == 0)
	{
		ctl->numUnits = min(numUnits, fp.unitCapacity);
		if (numUnits > fp.unitCapacity) atomicOr(&ctl->overflow, 4u);
	}
}

} // namespace srb
