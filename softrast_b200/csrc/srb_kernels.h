// srb_kernels.h — internal launcher interface between the host C-ABI layer (srb_api.cu) and the kernels.
#pragma once
#include "srb_device.cuh"

struct srb_tile_tri;

namespace srb
{

struct RasterArgs
{
	FrameParams fp;
	uint32_t tilesXMagic;    // ceil(2^32 / tilesX): tile / tilesX == __umulhi(tile, magic) for tilesX >= 2
	const uint32_t* offsets; // numTiles + 1
	const TileRef* refs;
	const UnitDesc* units;
	unsigned long long* tileKeys;  // numTiles * 4096 resolved (depth, winner) keys; all zero between frames
	const RasterRec* rrecs;
	const ShadeRec* srecs;
	const DrawDev* draws;
	const TexDev* texs;
	uint32_t numTexs;
	const uint32_t* rcpTable;
	uint32_t rcpBits;
	const uint16_t* rcp16; // the 11-bit table packed to 16 bits per entry (4 KB, copied to shared memory by the shade kernel), or nullptr
	const uint32_t* rsqrtTable;
	uint32_t rsqrtBits;
	const SponzaDev* sponza; // constants of SRB_SHADER_SPONZA for this frame (nullptr if no draw uses it)
	uint8_t* colourTiles; // 16384 bytes per tile
	uint8_t* depthTiles;  // 16384 bytes per tile (packed; the reference's 16416-byte stride is applied on read-back)
	uint32_t clearWord;
	int clearColour;
	int clearDepth;
	FrameCtl* ctl;
	uint32_t shadeCtasPerSm; // resident CTAs per SM the shade grid is sized for (0 = default 16)
	uint32_t blockReject;    // rasteriser: drop (triangle, 8x8 block) pairs whose 64 pixels all fail the edge test before the rows
	uint32_t uniformUnlit;   // every draw of the frame: UnlitDiffuse, non-empty texture, uvOffset 6 (a specialised shade kernel)
	// Screen-tile split of one frame across GPUs (nullptr / 0 otherwise).  splitFlags lives in the ROOT GPU's memory (for
	// the other ranks: peer memory over NVLink): [r] = arrival stamp of rank r ("my tiles of frame n are in the root's
	// framebuffer"), [32] = release stamp of the root ("frame n has been consumed: its tiles may be overwritten").
	uint32_t* splitFlags;
	uint32_t splitIsRoot;
	// The frame's control block in pinned host memory (device address): the last shade CTA stores the final block there
	// itself, so a frame needs no 64-byte copy through the copy engine behind its colour read-back.
	uint32_t* hostCtl;
	uint32_t* winnersOut; // debug only: canonical rank of the visible fragment per pixel (nullptr in production)
};

// K1: set-up; then the clip pass with (fuseScan) the tile scan of K2 in its tail
cudaError_t setup_init();
void setup_plan_smem(FrameParams& fp); // decides fp.smemHist / fp.smemBase from the tile and draw counts
size_t setup_smem_bytes(const FrameParams& fp);
bool launch_setup(const FrameParams& fp, const DrawDev* draws, RasterRec* rasterRecs, ShadeRec* shadeRecs,
                  Survivor* survivors, uint32_t* clipQueue, uint32_t* tileCounts, FrameCtl* ctl, uint32_t ctasPerSm,
                  uint32_t* releaseFlag, cudaStream_t stream); // ctasPerSm: 0 = one triangle per thread, else a grid of that many CTAs per SM
void launch_clip_scan(const FrameParams& fp, const DrawDev* draws, RasterRec* rasterRecs, ShadeRec* shadeRecs,
                      Survivor* survivors, uint32_t* clipQueue, uint32_t* tileCounts, uint32_t* offsets, uint32_t* cursors,
                      UnitDesc* units, FrameCtl* ctl, bool fuseScan, uint32_t* releaseFlag, cudaStream_t stream);
// K2
cudaError_t bin_init();
void launch_tile_scan(const FrameParams& fp, uint32_t* counts, uint32_t* offsets, uint32_t* cursors, UnitDesc* units,
                      FrameCtl* ctl, cudaStream_t stream);
bool launch_bin_fill(const FrameParams& fp, const RasterRec* recs, const Survivor* survivors, const uint32_t* offsets,
                     uint32_t* cursors, TileRef* refs, const FrameCtl* ctl, cudaStream_t stream);
// K3 + K4
cudaError_t raster_init();
size_t raster_smem_bytes();
int raster_ctas_per_sm();
#ifdef SRB_STATS
void stats_read(unsigned long long* out, bool reset); // counters of the statistics build
void stats_null_taps(int on);
#endif
void launch_raster(const RasterArgs& A, uint32_t ctas, cudaStream_t stream);
void launch_shade(const RasterArgs& A, cudaStream_t stream);
// blit
void launch_detile(const uint32_t* colourTiles, uint32_t* linear, uint32_t width, uint32_t height, uint32_t tilesX,
                   cudaStream_t stream);
// texture builder (srb_texbuild.cu).  One axis of stb_image_resize's down-sampling filter for one mip level, on the device,
// as gather lists: output k sums entries off[k] .. off[k+1]-1 in that order (= ascending contributor, stb's order), each
// entry = (index of the input pixel / row, already clamped to the image; coefficient).
struct StbAxisDev
{
	const int2* ent; // .x = source index, .y = coefficient bits
	const int* off;  // outputs + 1
};
void launch_tex_tile(const uint8_t* linear, uint8_t* dstLevel, uint32_t w, uint32_t h, cudaStream_t stream);
void launch_tex_hpass(const uint8_t* linear, float* hbuf, int iw, int ih, int ow, const StbAxisDev& H, cudaStream_t stream);
void launch_tex_vpass(const float* hbuf, uint8_t* dstLevel, int ih, int ow, int oh, const StbAxisDev& V, cudaStream_t stream);
// SRB_FLAG_UPLOAD_ALWAYS with pinned application arrays: ONE kernel per frame pulls every array of the frame's draws out
// of host memory (mapped, read over PCIe by the SMs) into its device mirror, instead of one DMA per array.
struct GatherSeg
{
	const uint8_t* src; // device-visible alias of the pinned host array
	uint8_t* dst;
	unsigned long long bytes;
	uint32_t firstBlock; // the segment is copied by blocks firstBlock .. firstBlock + ceil(bytes / kGatherChunk) - 1
	uint32_t pad;
};
constexpr uint32_t kGatherChunk = 16384; // bytes per block: 256 threads x 4 x 16 bytes, all loads in flight together
// fills firstBlock of every segment and returns the number of blocks
uint32_t gather_plan(GatherSeg* segs, uint32_t n);
void launch_gather(const GatherSeg* segs, uint32_t n, uint32_t blocks, cudaStream_t stream);
// parity / unit-test entry points
void launch_dump_tile_tris(const RasterArgs& A, uint32_t tile, const KeySlot* list, uint32_t count, srb_tile_tri* out,
                           uint32_t cap, cudaStream_t stream);
void launch_dump_tile_coverage(const RasterArgs& A, uint32_t tile, const KeySlot* list, uint32_t count,
                               unsigned long long* masks, uint32_t cap, cudaStream_t stream);
void launch_sample(const TexDev* texs, uint32_t texIdx, const float* u, const float* v, const float* dudx,
                   const float* dudy, const float* dvdx, const float* dvdy, uint32_t* out, uint32_t n,
                   cudaStream_t stream);
void launch_rcp(const uint32_t* table, uint32_t bits, const float* in, float* out, uint32_t n, cudaStream_t stream);
void launch_rsqrt(const uint32_t* table, uint32_t bits, const float* in, float* out, uint32_t n, cudaStream_t stream);

} // namespace srb

// srb_host.cpp: stbir__calculate_filters for one axis of a down-sampling resize (the tables both texture builders use).
// n0 / n1 / coef are indexed by contributor (inputSize + 2 * margin of them), coef has 4 entries per contributor.
void srb_internal_stb_axis(int inputSize, int outputSize, int* margin, int** n0, int** n1, float** coef, int* numContributors);
void srb_internal_stb_axis_free(int* n0, int* n1, float* coef);
