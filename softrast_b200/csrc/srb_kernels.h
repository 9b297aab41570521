// srb_kernels.h — internal launcher interface between the host C-ABI layer (srb_api.cu) and the kernels.
#pragma once
#include "srb_device.cuh"

struct srb_tile_tri;

namespace srb
{

struct RasterArgs
{
	FrameParams fp;
	const uint32_t* offsets; // numTiles + 1
	const uint32_t* refs;
	const RasterRec* rrecs;
	const ShadeRec* srecs;
	const DrawDev* draws;
	const TexDev* texs;
	const uint32_t* rcpTable;
	uint32_t rcpBits;
	uint8_t* colourTiles; // 16384 bytes per tile
	uint8_t* depthTiles;  // 16384 bytes per tile (packed; the reference's 16416-byte stride is applied on read-back)
	uint32_t clearWord;
	int clearColour;
	int clearDepth;
	FrameCtl* ctl;
	uint32_t* winnersOut; // debug only: canonical rank of the visible fragment per pixel (nullptr in production)
};

// K1
void launch_setup(const FrameParams& fp, const DrawDev* draws, RasterRec* rasterRecs, ShadeRec* shadeRecs,
                  uint32_t* tileCounts, unsigned long long* lookback, FrameCtl* ctl, cudaStream_t stream);
uint32_t setup_num_blocks(uint32_t numInputTris);
// K2
void launch_tile_scan(uint32_t numTiles, const uint32_t* counts, uint32_t* offsets, uint32_t* cursors, FrameCtl* ctl,
                      uint32_t refCapacity, cudaStream_t stream);
void launch_bin_fill(const FrameParams& fp, const RasterRec* recs, const uint32_t* offsets, uint32_t* cursors,
                     uint32_t* refs, const FrameCtl* ctl, cudaStream_t stream);
void launch_tile_sort(uint32_t numTiles, const uint32_t* offsets, uint32_t* refs, const FrameCtl* ctl,
                      uint32_t refCapacity, cudaStream_t stream);
// K3 + K4
cudaError_t raster_init();
size_t raster_smem_bytes();
void launch_raster_shade(const RasterArgs& A, cudaStream_t stream);
// blit
void launch_detile(const uint32_t* colourTiles, uint32_t* linear, uint32_t width, uint32_t height, uint32_t tilesX,
                   cudaStream_t stream);
// parity / unit-test entry points
void launch_dump_tile_tris(const RasterArgs& A, uint32_t tile, srb_tile_tri* out, uint32_t cap, cudaStream_t stream);
void launch_dump_tile_coverage(const RasterArgs& A, uint32_t tile, unsigned long long* masks, uint32_t cap,
                               cudaStream_t stream);
void launch_sample(const TexDev* texs, uint32_t texIdx, const float* u, const float* v, const float* dudx,
                   const float* dudy, const float* dvdx, const float* dvdy, uint32_t* out, uint32_t n,
                   cudaStream_t stream);
void launch_rcp(const uint32_t* table, uint32_t bits, const float* in, float* out, uint32_t n, cudaStream_t stream);

} // namespace srb
