// srb_api.cu — the C ABI (include/softrast_b200.h): context, resources, frame orchestration.
//
// Host-side counterpart of sr::RenderContext (reference SoftRast/Renderer.cpp:138-372).  Where the reference's EndFrame
// pushes front-end tasks, waits, pushes one task per tile and waits again (Renderer.cpp:209-317), this layer enqueues
// six kernels on one CUDA stream with no host synchronisation in between:
//   setup (+ tile counting) -> clip -> tile scan -> bin fill -> raster -> shade.
// There is no CPU fallback: without a CUDA device srb_create fails.
#include "../../include/softrast_b200.h"
#include "srb_kernels.h"

#include <cuda_runtime.h>
#include <stdarg.h>
#include <sys/mman.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <unordered_map>
#include <map>
#include <mutex>
#include <vector>

using namespace srb;

namespace
{

struct Texture
{
	uint8_t* dev = nullptr;
	TexDev desc{};
	bool alive = false;
};

struct Buffer
{
	uint8_t* dev = nullptr;
	uint64_t bytes = 0;
	bool alive = false;
};

struct HostMirror
{
	uint64_t bytes = 0;
	srb_handle buffer = 0;
	uint64_t uploadedInFrame = 0; // SRB_FLAG_UPLOAD_ALWAYS: frame serial of the last upload (one upload per array per frame)
	uint64_t uploadedBytes = 0;
	const uint8_t* hostDev = nullptr; // device-visible alias of the host array if it is pinned / registered, else nullptr
};

struct FrameBufferDev
{
	uint32_t width = 0, height = 0, tilesX = 0, tilesY = 0;
	uint8_t* colour[2] = {nullptr, nullptr}; // 16384 bytes per tile
	uint8_t* depth[2] = {nullptr, nullptr};  // 16384 bytes per tile
	uint32_t writePlane = 0;
	uint32_t* linear = nullptr; // blit staging (device)
	cudaEvent_t planeRead[2] = {nullptr, nullptr}; // recorded on the blit stream when a blit has finished READING that plane
	bool planeBusy[2] = {false, false};
	bool pendingClearColour = false, pendingClearDepth = false;
	uint32_t clearWord = 0;
	bool alive = false;
	bool imported = false; // colour[0] is another process's memory (cudaIpcOpenMemHandle); depth stays local
	uint8_t* rootDepth = nullptr; // imported: the root's depth allocation (only the split flags behind its tiles are used)
	int exportedPlane = -1; // the plane srb_framebuffer_export handed out (its depth allocation carries the split flags)
};

// Every depth plane is allocated with this many extra bytes behind its tiles: the flags of a screen-tile split
// (RasterArgs::splitFlags) live there, so that they travel with the framebuffer's IPC handle.
constexpr size_t kSplitFlagBytes = 256;

struct BlitCallback
{
	void (*fn)(void*);
	void* user;
};

constexpr int kMaxTimers = 10;

// One instantiated CUDA graph of a frame: head upload -> set-up (+ clip, tile scan) -> bin fill -> raster -> shade ->
// control block read-back.  Everything that varies from frame to frame (the control block, the draw table with its MVPs)
// travels through the pinned head buffer, so a camera path re-launches the same graph; `key` holds every value the nodes
// were captured with (kernel arguments, grids), and a frame whose values differ captures a new graph.
struct FrameGraph
{
	std::vector<uint8_t> key;
	cudaGraphExec_t exec = nullptr;
	uint64_t lastUse = 0;
	uint32_t kernels = 0;
};

// Scene resources (textures, buffers, mirrors of borrowed host buffers): owned by one context or shared by several
// contexts of one device (srb_create_shared), e.g. the frames in flight of a camera-path batch — ONE copy of the scene
// in HBM/L2 whatever the number of frames in flight.
struct Resources
{
	int refs = 1;
	std::vector<Texture> textures;
	std::vector<Buffer> buffers;
	std::unordered_map<const void*, HostMirror> mirrors;
	TexDev* dTexs = nullptr;
	uint32_t dTexsCap = 0;
	uint64_t texGeneration = 1; // bumped whenever the texture table changes; contexts re-validate lazily
	uint64_t uploadedGeneration = 0;
	uint64_t frameSerial = 0; // counts srb_begin_frame calls of all contexts of the family
};

} // namespace

struct srb_context
{
	int device = 0;
	uint32_t flags = 0;
	cudaStream_t stream = nullptr;
	cudaStream_t blitStream = nullptr; // de-tile + read-back + callback of Blit run beside the next frame (the planes are double-buffered)
	std::string error;

	Resources* res = nullptr; // textures, buffers, host mirrors (possibly shared with sibling contexts)
	std::vector<FrameBufferDev> fbs;

	uint32_t* dRcp = nullptr;
	uint32_t rcpBits = 0;
	uint16_t* dRcp16 = nullptr; // packed copy of an 11-bit table with 12 significant mantissa bits (the usual one), else nullptr
	uint32_t* dRsqrt = nullptr;
	uint32_t rsqrtBits = 0;
	srb_sponza_constants sponza{}; // SRB_SHADER_SPONZA frame constants (host copy)
	SponzaDev* dSponza = nullptr;
	bool sponzaDirty = true;
	bool recUsesSponza = false, frameUsesSponza = false;

	// frame recording
	bool inFrame = false;
	std::vector<DrawDev> recDraws;   // being recorded by srb_draw_indexed
	uint32_t recInputTris = 0;
	srb_handle recFb = 0;
	std::vector<DrawDev> draws;      // of the submitted frame (kept for the overflow re-run)
	uint32_t numInputTris = 0;
	srb_handle frameFb = 0;

	// frame device state
	// The frame's control block and draw table live in ONE device allocation, [FrameCtl][DrawDev x draws], so that one
	// upload per frame resets the former (64 zero bytes in front) and fills the latter: dCtl and dDraws point into dHead.
	uint8_t* dHead = nullptr;
	uint32_t dHeadCap = 0; // bytes
	uint8_t* hHead = nullptr; // pinned staging copy of the head: the upload is a plain DMA (and a memcpy node of the frame graph)
	uint32_t hHeadCap = 0;
	DrawDev* dDraws = nullptr;
	RasterRec* dRaster = nullptr;   // [slotCap]
	ShadeRec* dShade = nullptr;     // [slotCap]
	Survivor* dSurvivors = nullptr; // [slotCap]
	uint32_t slotCap = 0;           // numInputTris + fanCap
	uint32_t fanCap = 0;            // slots available to clipped fans
	uint32_t* dClipQueue = nullptr; // [clipQueueCap] input triangles that cross a frustum plane
	uint32_t clipQueueCap = 0;
	TileRef* dRefs = nullptr;
	uint32_t refCap = 0;
	UnitDesc* dUnits = nullptr;
	uint32_t unitCap = 0;
	uint32_t* dTileCounts = nullptr;
	uint32_t* dTileOffsets = nullptr;
	uint32_t* dTileCursors = nullptr;
	unsigned long long* dTileKeys = nullptr;
	uint32_t tilesCap = 0;
	uint32_t rasterCtas = 0;
	uint64_t frameSerial = 0; // this frame's serial number (see Resources::frameSerial)
	std::vector<GatherSeg> gather; // SRB_FLAG_UPLOAD_ALWAYS: pinned arrays to pull into their mirrors before this frame's kernels
	GatherSeg* dGather = nullptr;
	uint32_t dGatherCap = 0;
	uint32_t rasterCtasLatency = 0, rasterCtasThroughput = 0; // persistent grid sizes for one / several frames in flight
	uint32_t shadeCtasPerSm = 0;                              // 0 = default
	uint32_t setupCtasPerSm = 0;                              // 0 = one triangle per thread
	FrameCtl* dCtl = nullptr;
	FrameCtl* hCtl = nullptr; // pinned, mapped: written by the last shade CTA of a frame
	uint32_t* hCtlDev = nullptr; // its device address
	bool fuseScan = true;     // the tile scan runs in the tail of the set-up kernel (SRB_SEPARATE_SCAN=1: as its own launch)
	bool useGraphs = true;    // a frame is one CUDA graph launch (SRB_NO_GRAPH=1: plain stream launches)
	std::vector<FrameGraph> graphs; // instantiated frame graphs, keyed by everything their nodes were captured with
	uint64_t graphClock = 0;
	// read-back that belongs to the submitted frame (srb_render_frames): re-issued if the frame has to be re-run
	void* readbackDst = nullptr;
	size_t readbackBytes = 0;
	void* nextReadbackDst = nullptr; // set before srb_end_frame_async, becomes readbackDst of the frame it submits
	size_t nextReadbackBytes = 0;
	std::vector<srb_handle> recFbs; // framebuffers touched by this frame's clears and draws, in order of first use
	std::vector<srb_handle> recDrawFb; // framebuffer of every recorded draw

	// last submitted frame (for overflow re-run, dumps, counters)
	bool framePending = false;
	bool frameValid = false;
	RasterArgs lastArgs{};
	bool lastClearColour = false, lastClearDepth = false; // the clear folded into the submitted frame
	uint32_t lastClearWord = 0;
	srb_counters counters{};

	uint32_t ownMod = 1, ownRem = 0;
	uint32_t* releaseFlag = nullptr; // this frame's (root of a screen-tile split only)
	uint32_t splitSerial = 0; // frames submitted in a screen-tile split: the arrival / release stamp (same on every rank)
	uint32_t minUnit = 128; // tuning knob (SRB_MIN_UNIT): smallest raster work unit in tile references (256 -> 128: the 720p cube grid's rasteriser 41 -> 35 us, others unchanged)
	cudaEvent_t marks[4] = {};
	uint8_t* dFlush = nullptr;
	uint64_t flushBytes = 0;
	uint64_t flushCount = 0;

	bool timing = false;
	cudaEvent_t ev[kMaxTimers] = {};
	float kernelMicros[kMaxTimers] = {};
	cudaEvent_t evBlit[2] = {}; // around the de-tile kernel of the last Blit (timing mode)
	bool blitTimed = false;
	uint64_t launches = 0;
};

namespace
{

int Fail(srb_context* c, int code, const char* fmt, ...)
{
	if (c)
	{
		char buf[512];
		va_list ap;
		va_start(ap, fmt);
		vsnprintf(buf, sizeof(buf), fmt, ap);
		va_end(ap);
		c->error = buf;
	}
	return code;
}

#define SRB_CUDA(c, call)                                                                                  \
	do                                                                                                     \
	{                                                                                                      \
		cudaError_t e_ = (call);                                                                           \
		if (e_ != cudaSuccess)                                                                             \
		{                                                                                                  \
			return Fail((c), SRB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
			            __LINE__);                                                                         \
		}                                                                                                  \
	} while (0)

// Before a resource of the family (texture, buffer, host mirror) is freed or overwritten: contexts created with
// srb_create_shared render from the same resources on their own streams, so waiting for this context's stream is not
// enough — the whole device is drained when the family has more than one member.
void QuiesceResources(srb_context* c)
{
	if (c->res && c->res->refs > 1)
	{
		cudaDeviceSynchronize();
	}
	else
	{
		cudaStreamSynchronize(c->stream);
	}
}

int Bind(srb_context* c)
{
	SRB_CUDA(c, cudaSetDevice(c->device));
	return SRB_OK;
}

FrameBufferDev* GetFb(srb_context* c, srb_handle h)
{
	if (!h || h > c->fbs.size() || !c->fbs[h - 1].alive)
	{
		return nullptr;
	}
	return &c->fbs[h - 1];
}

void DropGraphs(srb_context* c)
{
	for (FrameGraph& g : c->graphs)
	{
		if (g.exec) cudaGraphExecDestroy(g.exec);
	}
	c->graphs.clear();
}

template <typename T>
int Grow(srb_context* c, T*& ptr, uint32_t& cap, uint64_t need, uint64_t extra = 0)
{
	if (need <= cap && ptr)
	{
		return SRB_OK;
	}
	if (need > 0xFFFFFFF0ull)
	{
		return Fail(c, SRB_ERR_OVERFLOW, "capacity request too large");
	}
	if (ptr)
	{
		SRB_CUDA(c, cudaStreamSynchronize(c->stream));
		DropGraphs(c); // their nodes were captured with the old address
		SRB_CUDA(c, cudaFree(ptr));
		ptr = nullptr;
	}
	SRB_CUDA(c, cudaMalloc((void**)&ptr, (need + extra) * sizeof(T)));
	cap = (uint32_t)need;
	return SRB_OK;
}

// Device address of a draw buffer binding; host pointers are mirrored (and cached) on the device.
int Resolve(srb_context* c, const srb_buffer_ref& ref, uint64_t bytes, uint32_t align, const uint8_t** out)
{
	*out = nullptr;
	// the kernels read indices as uint16 / uint32 and positions / attributes as floats: a misaligned offset into a device
	// buffer would fault on the device and poison the CUDA context (host arrays are mirrored into aligned device memory,
	// so their own alignment does not matter — the reference reads them at any alignment)
	if (ref.buffer && align > 1u && (ref.offset % align) != 0)
	{
		return Fail(c, SRB_ERR_INVALID, "draw buffer offset %llu is not a multiple of %u bytes", (unsigned long long)ref.offset, align);
	}
	if (ref.buffer)
	{
		if (ref.buffer > c->res->buffers.size() || !c->res->buffers[ref.buffer - 1].alive)
		{
			return Fail(c, SRB_ERR_INVALID, "bad buffer handle");
		}
		Buffer& b = c->res->buffers[ref.buffer - 1];
		if (ref.offset > b.bytes || bytes > b.bytes - ref.offset)
		{
			return Fail(c, SRB_ERR_INVALID, "buffer binding out of range (%llu + %llu > %llu)",
			            (unsigned long long)ref.offset, (unsigned long long)bytes, (unsigned long long)b.bytes);
		}
		*out = b.dev + ref.offset;
		return SRB_OK;
	}
	if (!ref.host)
	{
		return bytes == 0 ? SRB_OK : Fail(c, SRB_ERR_INVALID, "draw buffer has neither a handle nor a host pointer");
	}
	auto it = c->res->mirrors.find(ref.host);
	bool const always = (c->flags & SRB_FLAG_UPLOAD_ALWAYS) != 0;
	if (it != c->res->mirrors.end() && it->second.bytes >= bytes && !always)
	{
		*out = c->res->buffers[it->second.buffer - 1].dev;
		return SRB_OK;
	}
	if (it != c->res->mirrors.end() && it->second.bytes >= bytes)
	{
		// upload-always: once per array and frame (a draw binds the same vertex array as positions AND attributes, and
		// the draws of a frame may share arrays — the reference reads them in place, so one copy per frame is its view too)
		Buffer& b = c->res->buffers[it->second.buffer - 1];
		HostMirror& m = it->second;
		if (m.uploadedInFrame != c->frameSerial || m.uploadedBytes < bytes)
		{
			if (m.hostDev)
			{
				// pinned array: joins the frame's gather list — ONE batched copy pulls all of them (Submit), instead of a copy
				// call each (25 draws = 50 small copies cost 500 us per frame)
				c->gather.push_back(GatherSeg{m.hostDev, b.dev, bytes, 0u, 0u});
			}
			else
			{
				SRB_CUDA(c, cudaMemcpyAsync(b.dev, ref.host, bytes, cudaMemcpyHostToDevice, c->stream));
			}
			m.uploadedInFrame = c->frameSerial;
			m.uploadedBytes = bytes;
		}
		*out = b.dev;
		return SRB_OK;
	}
	if (it != c->res->mirrors.end())
	{
		srb_buffer_destroy(c, it->second.buffer);
		c->res->mirrors.erase(it);
	}
	srb_handle h = 0;
	int const rc = srb_buffer_create(c, ref.host, bytes, &h);
	if (rc != SRB_OK)
	{
		return rc;
	}
	HostMirror fresh;
	fresh.bytes = bytes;
	fresh.buffer = h;
	fresh.uploadedInFrame = c->frameSerial; // srb_buffer_create has just copied the bytes
	fresh.uploadedBytes = bytes;
	if (always)
	{
		cudaPointerAttributes attr;
		if (cudaPointerGetAttributes(&attr, ref.host) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
		{
			fresh.hostDev = static_cast<const uint8_t*>(attr.devicePointer);
		}
		cudaGetLastError(); // (an unregistered pointer is not an error here)
	}
	c->res->mirrors[ref.host] = fresh;
	*out = c->res->buffers[h - 1].dev;
	return SRB_OK;
}

int UploadTexTable(srb_context* c)
{
	Resources* r = c->res;
	if (r->uploadedGeneration == r->texGeneration)
	{
		return SRB_OK;
	}
	if (r->refs > 1)
	{
		// sibling contexts may have frames in flight that read the table: let them finish before it is replaced
		SRB_CUDA(c, cudaDeviceSynchronize());
	}
	uint32_t const n = (uint32_t)std::max<size_t>(1, c->res->textures.size());
	int rc = Grow(c, c->res->dTexs, c->res->dTexsCap, n);
	if (rc != SRB_OK)
	{
		return rc;
	}
	std::vector<TexDev> tmp(n);
	for (size_t i = 0; i < c->res->textures.size(); ++i)
	{
		tmp[i] = c->res->textures[i].desc;
	}
	SRB_CUDA(c, cudaMemcpyAsync(c->res->dTexs, tmp.data(), n * sizeof(TexDev), cudaMemcpyHostToDevice, c->stream));
	SRB_CUDA(c, cudaStreamSynchronize(c->stream)); // tmp goes out of scope
	r->uploadedGeneration = r->texGeneration;
	return SRB_OK;
}

// The kernels of one frame on stream s, in order.  `timed`: a CUDA event after every step (srb_set_timing).
// Called directly or under stream capture (FrameGraph).  Returns the number of kernels it launched.
int EnqueueFrame(srb_context* c, const FrameParams& fp, const RasterArgs& A, size_t headBytes, bool timed, cudaStream_t s,
                 uint32_t* kernelsOut)
{
	int t = 1;
	uint32_t kernels = 0;
	// one upload: a zeroed control block followed by the draw table
	SRB_CUDA(c, cudaMemcpyAsync(c->dHead, c->hHead, headBytes, cudaMemcpyHostToDevice, s));
	if (timed) SRB_CUDA(c, cudaEventRecord(c->ev[t++], s));
	if (launch_setup(fp, c->dDraws, c->dRaster, c->dShade, c->dSurvivors, c->dClipQueue, c->dTileCounts, c->dCtl,
	                 c->setupCtasPerSm, c->releaseFlag, s))
	{
		kernels++;
	}
	if (timed) SRB_CUDA(c, cudaEventRecord(c->ev[t++], s));
	launch_clip_scan(fp, c->dDraws, c->dRaster, c->dShade, c->dSurvivors, c->dClipQueue, c->dTileCounts, c->dTileOffsets,
	                 c->dTileCursors, c->dUnits, c->dCtl, c->fuseScan, c->releaseFlag, s);
	kernels++;
	if (timed) SRB_CUDA(c, cudaEventRecord(c->ev[t++], s));
	if (!c->fuseScan)
	{
		launch_tile_scan(fp, c->dTileCounts, c->dTileOffsets, c->dTileCursors, c->dUnits, c->dCtl, s);
		kernels++;
	}
	if (timed) SRB_CUDA(c, cudaEventRecord(c->ev[t++], s));
	if (launch_bin_fill(fp, c->dRaster, c->dSurvivors, c->dTileOffsets, c->dTileCursors, c->dRefs, c->dCtl, s))
	{
		kernels++;
	}
	if (timed) SRB_CUDA(c, cudaEventRecord(c->ev[t++], s));
	launch_raster(A, c->rasterCtas, s);
	kernels++;
	if (timed) SRB_CUDA(c, cudaEventRecord(c->ev[t++], s));
	launch_shade(A, s);
	kernels++;
	if (timed) SRB_CUDA(c, cudaEventRecord(c->ev[t++], s));
	// (no copy of the control block: the last shade CTA stores it into c->hCtl itself)
	*kernelsOut = kernels;
	return SRB_OK;
}

// Enqueue the pipeline for the recorded frame.
int Submit(srb_context* c)
{
	FrameBufferDev* fb = GetFb(c, c->frameFb);
	if (!fb)
	{
		return Fail(c, SRB_ERR_INVALID, "no framebuffer bound for this frame (srb_clear or srb_draw_indexed first)");
	}
	int rc = UploadTexTable(c);
	if (rc != SRB_OK)
	{
		return rc;
	}
	uint32_t const numTiles = fb->tilesX * fb->tilesY;
	uint32_t const numDraws = (uint32_t)c->draws.size();

	// capacities (grown on demand; an overflow detected on the device re-runs the frame with larger buffers)
	if (c->numInputTris > SRB_MAX_INPUT_TRIS)
	{
		return Fail(c, SRB_ERR_OVERFLOW, "more than %u triangles in one frame", SRB_MAX_INPUT_TRIS);
	}
	c->fanCap = std::max<uint32_t>(c->fanCap, std::max<uint32_t>(4096u, c->numInputTris / 8));
	uint32_t const wantSlots = c->numInputTris + c->fanCap;
	if (wantSlots > c->slotCap)
	{
		uint32_t const want = wantSlots + wantSlots / 8;
		uint32_t cap = c->slotCap;
		rc = Grow(c, c->dRaster, cap, want);
		if (rc != SRB_OK) return rc;
		cap = c->slotCap;
		rc = Grow(c, c->dShade, cap, want);
		if (rc != SRB_OK) return rc;
		cap = c->slotCap;
		rc = Grow(c, c->dSurvivors, cap, want);
		if (rc != SRB_OK) return rc;
		c->slotCap = want;
	}
	if (c->numInputTris > c->clipQueueCap)
	{
		rc = Grow(c, c->dClipQueue, c->clipQueueCap, c->numInputTris + c->numInputTris / 8);
		if (rc != SRB_OK) return rc;
	}
	uint32_t const wantRefs = std::max<uint32_t>(1u << 20, 2 * c->numInputTris);
	if (wantRefs > c->refCap)
	{
		rc = Grow(c, c->dRefs, c->refCap, wantRefs);
		if (rc != SRB_OK) return rc;
	}
	uint32_t const wantUnits = numTiles + c->refCap / c->minUnit + 64u;
	if (wantUnits > c->unitCap)
	{
		rc = Grow(c, c->dUnits, c->unitCap, wantUnits);
		if (rc != SRB_OK) return rc;
	}
	if (numTiles + 1 > c->tilesCap)
	{
		uint32_t cap = c->tilesCap;
		rc = Grow(c, c->dTileCounts, cap, numTiles + 1);
		if (rc != SRB_OK) return rc;
		cap = c->tilesCap;
		rc = Grow(c, c->dTileOffsets, cap, numTiles + 1);
		if (rc != SRB_OK) return rc;
		cap = c->tilesCap;
		rc = Grow(c, c->dTileCursors, cap, numTiles + 1);
		if (rc != SRB_OK) return rc;
		cap = c->tilesCap;
		rc = Grow(c, c->dTileKeys, cap, uint64_t(numTiles + 1) * 4096u);
		if (rc != SRB_OK) return rc;
		c->tilesCap = numTiles + 1;
		// counters and merge buffers are kept zero BETWEEN frames by the kernels themselves
		SRB_CUDA(c, cudaMemsetAsync(c->dTileCounts, 0, (numTiles + 1) * sizeof(uint32_t), c->stream));
		SRB_CUDA(c, cudaMemsetAsync(c->dTileKeys, 0, size_t(numTiles + 1) * 4096u * sizeof(unsigned long long), c->stream));
	}
	size_t const headBytes = sizeof(FrameCtl) + size_t(numDraws) * sizeof(DrawDev);
	rc = Grow(c, c->dHead, c->dHeadCap, sizeof(FrameCtl) + size_t(std::max<uint32_t>(64u, numDraws)) * sizeof(DrawDev));
	if (rc != SRB_OK) return rc;
	if (c->hHeadCap < c->dHeadCap)
	{
		SRB_CUDA(c, cudaStreamSynchronize(c->stream));
		if (c->hHead) cudaFreeHost(c->hHead);
		c->hHead = nullptr;
		SRB_CUDA(c, cudaHostAlloc((void**)&c->hHead, c->dHeadCap, cudaHostAllocDefault));
		c->hHeadCap = c->dHeadCap;
	}
	c->dCtl = reinterpret_cast<FrameCtl*>(c->dHead);
	c->dDraws = reinterpret_cast<DrawDev*>(c->dHead + sizeof(FrameCtl));

	FrameParams fp;
	memset(&fp, 0, sizeof(fp));
	fp.width = fb->width;
	fp.height = fb->height;
	fp.tilesX = fb->tilesX;
	fp.tilesY = fb->tilesY;
	fp.numDraws = numDraws;
	fp.numInputTris = c->numInputTris;
	fp.slotCapacity = c->numInputTris + c->fanCap;
	fp.refCapacity = c->refCap;
	fp.unitCapacity = c->unitCap;
	fp.clearPending = (c->lastClearColour || c->lastClearDepth) ? 1u : 0u;
	fp.splitTiles = c->lastClearDepth ? 1u : 0u;
	fp.ownMod = c->ownMod;
	fp.ownRem = c->ownRem;
	fp.minUnit = c->minUnit;
	{
		// set-up chunks: 256 consecutive triangles of ONE draw
		uint32_t chunks = 0;
		for (DrawDev& d : c->draws)
		{
			d.chunkBase = chunks;
			chunks += (d.numTris + 255u) / 256u;
		}
		fp.numChunks = chunks;
	}
	setup_plan_smem(fp); // what does not fit into shared memory (huge tile or draw counts) stays in global memory

	RasterArgs A;
	memset(&A, 0, sizeof(A));
	A.fp = fp;
	A.tilesXMagic = fp.tilesX >= 2 ? (uint32_t)((0x100000000ull + fp.tilesX - 1) / fp.tilesX) : 0u;
	A.offsets = c->dTileOffsets;
	A.refs = c->dRefs;
	A.units = c->dUnits;
	A.tileKeys = c->dTileKeys;
	A.rrecs = c->dRaster;
	A.srecs = c->dShade;
	A.draws = c->dDraws;
	A.texs = c->res->dTexs;
	A.numTexs = (uint32_t)c->res->textures.size();
	A.rcpTable = c->dRcp;
	A.rcpBits = c->rcpBits;
	A.rcp16 = c->dRcp16;
	A.rsqrtTable = c->dRsqrt;
	A.rsqrtBits = c->rsqrtBits;
	A.sponza = c->frameUsesSponza ? c->dSponza : nullptr;
	A.colourTiles = fb->colour[fb->writePlane];
	A.depthTiles = fb->depth[fb->writePlane];
	A.clearWord = c->lastClearWord;
	A.clearColour = c->lastClearColour ? 1 : 0;
	A.clearDepth = c->lastClearDepth ? 1 : 0;
	A.ctl = c->dCtl;
	A.hostCtl = c->hCtlDev;
	A.winnersOut = nullptr;
	A.shadeCtasPerSm = c->shadeCtasPerSm;
	{
		static bool const noReject = getenv("SRB_NO_BLOCK_REJECT") != nullptr; // A/B knob (not part of the ABI)
		A.blockReject = noReject ? 0u : 1u;
	}
	uint32_t* releaseFlag = nullptr;
	if (c->ownMod > 1u && (fb->imported || fb->exportedPlane >= 0))
	{
		// a screen-tile split across processes: the flags sit behind the depth tiles of the root's exported plane
		if (!fb->imported && fb->exportedPlane != (int)fb->writePlane)
		{
			return Fail(c, SRB_ERR_INVALID, "the exported framebuffer's planes were swapped (Blit) during a screen-tile split");
		}
		if (c->ownMod > 32u)
		{
			return Fail(c, SRB_ERR_INVALID, "a screen-tile split has at most 32 ranks");
		}
		A.splitFlags = reinterpret_cast<uint32_t*>((fb->imported ? fb->rootDepth : fb->depth[fb->writePlane]) + size_t(numTiles) * 16384u);
		A.splitIsRoot = fb->imported ? 0u : 1u;
		releaseFlag = A.splitIsRoot ? A.splitFlags + 32 : nullptr;
	}
	c->releaseFlag = releaseFlag;
	{
		// the scene of Viewer/Scene.cpp:35-63: every draw UnlitDiffuse with a non-empty texture and uvOffset 6
		static bool const noUniform = getenv("SRB_NO_UNIFORM_SHADE") != nullptr; // A/B knob (not part of the ABI)
		bool uniform = numDraws > 0 && !noUniform;
		for (const DrawDev& d : c->draws)
		{
			uniform = uniform && d.shader == SRB_SHADER_UNLIT_DIFFUSE && d.uvOffset == 6u && d.texture >= 0 &&
			          c->res->textures[(size_t)d.texture].desc.bytes != 0u;
		}
		A.uniformUnlit = uniform ? 1u : 0u;
	}

	cudaStream_t s = c->stream;
	if (fb->planeBusy[fb->writePlane])
	{
		// a Blit may still be de-tiling the plane this frame is about to overwrite (two frames ago): wait on the device
		SRB_CUDA(c, cudaStreamWaitEvent(s, fb->planeRead[fb->writePlane], 0));
		fb->planeBusy[fb->writePlane] = false;
	}
	if (c->timing) SRB_CUDA(c, cudaEventRecord(c->ev[0], s));
	if (!c->gather.empty())
	{
		uint32_t const n = (uint32_t)c->gather.size();
		// One batched copy per frame (cudaMemcpyBatchAsync: the copy engine reads the arrays with large requests and
		// shares the link with the colour read-back better than loads from the SMs do): 51.7 GB/s against 39.3 GB/s for
		// the gather kernel, frames/s with geometry up and colour back every frame 3 740 -> 4 700 (profiles/README.md).
		// SRB_GATHER_KERNEL=1, or a driver without the batched copy: the gather kernel (SRB_GATHER_BULK=1: its variant on
		// bulk asynchronous copies).
		static bool useKernel = getenv("SRB_GATHER_KERNEL") != nullptr || getenv("SRB_GATHER_BULK") != nullptr;
		if (!useKernel)
		{
			std::vector<void*> dsts, srcs;
			std::vector<size_t> sizes;
			for (const GatherSeg& g : c->gather)
			{
				if (g.bytes != 0) // (a draw without indices or vertices binds empty arrays; the batch rejects empty copies)
				{
					dsts.push_back(g.dst);
					srcs.push_back(const_cast<uint8_t*>(g.src));
					sizes.push_back((size_t)g.bytes);
				}
			}
			cudaMemcpyAttributes attr = {};
			attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream; // the arrays are the application's: read them in stream order
			attr.flags = cudaMemcpyFlagPreferOverlapWithCompute;
			size_t attrIdx = 0, failIdx = 0;
			size_t const copies = sizes.size();
			cudaError_t const e = copies == 0 ? cudaSuccess
			                                  : cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), copies, &attr, &attrIdx, 1, &failIdx, s);
			if (e == cudaErrorNotSupported || e == cudaErrorCallRequiresNewerDriver)
			{
				cudaGetLastError();
				useKernel = true;
			}
			else if (e != cudaSuccess)
			{
				cudaGetLastError();
				return Fail(c, SRB_ERR_CUDA, "cudaMemcpyBatchAsync: %s (copy %zu of %zu: dst %p src %p bytes %zu)", cudaGetErrorString(e),
				            failIdx, copies, failIdx < copies ? dsts[failIdx] : nullptr, failIdx < copies ? srcs[failIdx] : nullptr,
				            failIdx < copies ? sizes[failIdx] : (size_t)0);
			}
		}
		if (useKernel)
		{
			rc = Grow(c, c->dGather, c->dGatherCap, n);
			if (rc != SRB_OK) return rc;
			uint32_t const blocks = gather_plan(c->gather.data(), n);
			SRB_CUDA(c, cudaMemcpyAsync(c->dGather, c->gather.data(), n * sizeof(GatherSeg), cudaMemcpyHostToDevice, s)); // pageable: staged
			launch_gather(c->dGather, n, blocks, s);
			c->launches++;
		}
		c->gather.clear();
	}
	if (c->frameUsesSponza && c->sponzaDirty)
	{
		// pageable source: staged by the runtime before the call returns
		SRB_CUDA(c, cudaMemcpyAsync(c->dSponza, &c->sponza, sizeof(SponzaDev), cudaMemcpyHostToDevice, s));
		c->sponzaDirty = false;
	}
	// the head: a zeroed control block followed by the draw table, staged in pinned memory (the previous frame of this
	// context has completed — Finish() — so the buffer is free)
	memset(c->hHead, 0, sizeof(FrameCtl));
	reinterpret_cast<FrameCtl*>(c->hHead)->doneValue = c->splitSerial; // screen-tile split: this frame's arrival stamp
	if (numDraws) memcpy(c->hHead + sizeof(FrameCtl), c->draws.data(), size_t(numDraws) * sizeof(DrawDev));

	uint32_t kernels = 0;
	if (c->useGraphs && !c->timing)
	{
		// everything the nodes of the graph are captured with; the per-frame values travel through the head
		RasterArgs const& keyA = A;
		std::vector<uint8_t> key(sizeof(FrameParams) + sizeof(RasterArgs) + 8 * sizeof(void*) + 4 * sizeof(uint32_t));
		uint8_t* k = key.data();
		memcpy(k, &fp, sizeof(fp)); k += sizeof(fp);
		memcpy(k, &keyA, sizeof(keyA)); k += sizeof(keyA);
		const void* ptrs[8] = {c->dHead, c->hHead, c->dSurvivors, c->dTileCounts, c->dTileCursors, c->dClipQueue, c->hCtl, c->dRefs};
		memcpy(k, ptrs, sizeof(ptrs)); k += sizeof(ptrs);
		uint32_t const vals[4] = {c->setupCtasPerSm, c->rasterCtas, (uint32_t)headBytes, c->fuseScan ? 1u : 0u};
		memcpy(k, vals, sizeof(vals));
		FrameGraph* hit = nullptr;
		for (FrameGraph& g : c->graphs)
		{
			if (g.key == key) hit = &g;
		}
		if (!hit)
		{
			if (c->graphs.size() >= 8)
			{
				size_t oldest = 0;
				for (size_t i = 1; i < c->graphs.size(); ++i)
				{
					if (c->graphs[i].lastUse < c->graphs[oldest].lastUse) oldest = i;
				}
				cudaGraphExecDestroy(c->graphs[oldest].exec);
				c->graphs.erase(c->graphs.begin() + oldest);
			}
			cudaGraph_t graph = nullptr;
			SRB_CUDA(c, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
			uint32_t nk = 0;
			rc = EnqueueFrame(c, fp, A, headBytes, false, s, &nk);
			cudaError_t const e = cudaStreamEndCapture(s, &graph);
			if (rc != SRB_OK)
			{
				if (graph) cudaGraphDestroy(graph);
				return rc;
			}
			SRB_CUDA(c, e);
			FrameGraph g;
			g.key = key;
			g.kernels = nk;
			cudaError_t const ei = cudaGraphInstantiate(&g.exec, graph, 0);
			cudaGraphDestroy(graph);
			SRB_CUDA(c, ei);
			c->graphs.push_back(std::move(g));
			hit = &c->graphs.back();
		}
		hit->lastUse = ++c->graphClock;
		SRB_CUDA(c, cudaGraphLaunch(hit->exec, s));
		kernels = hit->kernels;
	}
	else
	{
		rc = EnqueueFrame(c, fp, A, headBytes, c->timing, s, &kernels);
		if (rc != SRB_OK) return rc;
	}
	c->launches += kernels;
	if (c->readbackDst)
	{
		// the frame's colour tiles go back to the caller's (pinned) memory as part of the frame: a re-run re-issues it
		SRB_CUDA(c, cudaMemcpyAsync(c->readbackDst, fb->colour[fb->writePlane], c->readbackBytes, cudaMemcpyDeviceToHost, s));
	}
	SRB_CUDA(c, cudaGetLastError());

	c->lastArgs = A;
	c->framePending = true;
	c->frameValid = false;
	return SRB_OK;
}

// Wait for the submitted frame; grow + re-run on device-side overflow.
int Finish(srb_context* c)
{
	if (!c->framePending)
	{
		SRB_CUDA(c, cudaStreamSynchronize(c->stream));
		return SRB_OK;
	}
	for (int attempt = 0; attempt < 4; ++attempt)
	{
		SRB_CUDA(c, cudaStreamSynchronize(c->stream));
		FrameCtl const& h = *c->hCtl;
		if (h.overflow == 0)
		{
			break;
		}
		if (h.overflow & 8u)
		{
			c->framePending = false;
			return Fail(c, SRB_ERR_CUDA, "screen-tile split: another rank's stamp did not arrive within 10 s");
		}
		if (attempt == 3)
		{
			c->framePending = false;
			return Fail(c, SRB_ERR_OVERFLOW, "frame still overflows after growing (fan slots %u refs %u units %u)",
			            h.numFanSlots, h.totalRefs, h.numUnits);
		}
		if (h.overflow & 1u)
		{
			c->fanCap = std::max<uint32_t>(2 * c->fanCap, h.numFanSlots + h.numFanSlots / 8 + 1024);
		}
		if (h.overflow & 2u)
		{
			uint32_t const want = h.totalRefs + h.totalRefs / 8 + 1024;
			int rc = Grow(c, c->dRefs, c->refCap, want);
			if (rc != SRB_OK) return rc;
		}
		if (h.overflow & 4u)
		{
			int rc = Grow(c, c->dUnits, c->unitCap, 2 * c->unitCap + 1024);
			if (rc != SRB_OK) return rc;
		}
		// the tile counters / merge buffers may be mid-frame dirty after an aborted frame: reset them
		{
			uint32_t const nt = c->lastArgs.fp.tilesX * c->lastArgs.fp.tilesY;
			SRB_CUDA(c, cudaMemsetAsync(c->dTileCounts, 0, (nt + 1) * sizeof(uint32_t), c->stream));
			SRB_CUDA(c, cudaMemsetAsync(c->dTileKeys, 0, size_t(nt + 1) * 4096u * sizeof(unsigned long long), c->stream));
		}
		int const rc = Submit(c);
		if (rc != SRB_OK)
		{
			return rc;
		}
	}
	FrameCtl const& h = *c->hCtl;
	c->counters.tris_in = c->numInputTris;
	c->counters.tris_setup = h.numSurvivors;
	c->counters.tris_clipped = h.numClipQueue;
	c->counters.tile_refs = h.totalRefs;
	c->counters.tiles_nonempty = h.tilesNonEmpty;
	c->counters.max_refs_in_tile = h.maxRefs;
	c->counters.pixels_covered = h.pixelsCovered;
	c->counters.overflow = h.overflow;
	if (c->timing)
	{
		for (int i = 0; i + 1 < 8; ++i) // 8 events -> 7 intervals
		{
			float ms = 0.0f;
			cudaEventElapsedTime(&ms, c->ev[i], c->ev[i + 1]);
			c->kernelMicros[i] = ms * 1000.0f;
		}
		if (c->fuseScan)
		{
			// the tile scan ran in the clip kernel's tail: what lies between the two events is the event itself
			c->kernelMicros[2] += c->kernelMicros[3];
			c->kernelMicros[3] = 0.0f;
		}
	}
	c->framePending = false;
	c->frameValid = true;
	return SRB_OK;
}

// A frame's clears and draws may address several framebuffers (DrawCall::SetFrameBuffer is per draw, Renderer.h:129).
void NoteFrameBuffer(srb_context* c, srb_handle h)
{
	if (!c->recFb) c->recFb = h;
	if (std::find(c->recFbs.begin(), c->recFbs.end(), h) == c->recFbs.end()) c->recFbs.push_back(h);
}

void ConsumeClear(srb_context* c)
{
	// the pending clear belongs to THIS frame: consume it now (an overflow re-run re-uses the saved copy)
	FrameBufferDev* fb = GetFb(c, c->frameFb);
	if (fb)
	{
		c->lastClearColour = fb->pendingClearColour;
		c->lastClearDepth = fb->pendingClearDepth;
		c->lastClearWord = fb->clearWord;
		fb->pendingClearColour = false;
		fb->pendingClearDepth = false;
	}
}

void CUDART_CB BlitDone(void* p)
{
	BlitCallback* cb = (BlitCallback*)p;
	if (cb->fn)
	{
		cb->fn(cb->user);
	}
	delete cb;
}

} // namespace

extern "C"
{

static int CreateContext(int device, uint32_t flags, Resources* shared, srb_context** out)
{
	if (!out)
	{
		return SRB_ERR_INVALID;
	}
	*out = nullptr;
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n)
	{
		return SRB_ERR_NO_DEVICE; // no CPU fallback by design
	}
	srb_context* c = new srb_context;
	c->device = device;
	c->flags = flags;
	if (shared)
	{
		c->res = shared;
		shared->refs++;
	}
	else
	{
		c->res = new Resources;
	}
	*out = c; // returned even on failure below so that srb_last_error() works; caller must srb_destroy it
	SRB_CUDA(c, cudaSetDevice(device));
	SRB_CUDA(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	if (getenv("SRB_BLIT_SAME_STREAM")) // tuning knob for experiments: blits queue on the frame stream (no overlap)
	{
		c->blitStream = c->stream;
	}
	else
	{
		SRB_CUDA(c, cudaStreamCreateWithFlags(&c->blitStream, cudaStreamNonBlocking));
	}
	c->dHeadCap = (uint32_t)(sizeof(FrameCtl) + 64u * sizeof(DrawDev));
	SRB_CUDA(c, cudaMalloc((void**)&c->dHead, c->dHeadCap));
	SRB_CUDA(c, cudaMemset(c->dHead, 0, c->dHeadCap));
	SRB_CUDA(c, cudaHostAlloc((void**)&c->hHead, c->dHeadCap, cudaHostAllocDefault));
	c->hHeadCap = c->dHeadCap;
	// tuning knobs for A/B measurements (not part of the ABI)
	c->fuseScan = getenv("SRB_SEPARATE_SCAN") == nullptr;
	c->useGraphs = getenv("SRB_NO_GRAPH") == nullptr;
	c->dCtl = reinterpret_cast<FrameCtl*>(c->dHead);
	c->dDraws = reinterpret_cast<DrawDev*>(c->dHead + sizeof(FrameCtl));
	SRB_CUDA(c, cudaHostAlloc((void**)&c->hCtl, sizeof(FrameCtl), cudaHostAllocMapped));
	memset(c->hCtl, 0, sizeof(FrameCtl));
	SRB_CUDA(c, cudaHostGetDevicePointer((void**)&c->hCtlDev, c->hCtl, 0));
	for (int i = 0; i < kMaxTimers; ++i)
	{
		SRB_CUDA(c, cudaEventCreate(&c->ev[i]));
	}
	for (int i = 0; i < 4; ++i)
	{
		SRB_CUDA(c, cudaEventCreate(&c->marks[i]));
	}
	for (int i = 0; i < 2; ++i)
	{
		SRB_CUDA(c, cudaEventCreate(&c->evBlit[i]));
	}
	SRB_CUDA(c, raster_init());
	SRB_CUDA(c, setup_init());
	SRB_CUDA(c, bin_init());
	{
		cudaDeviceProp prop;
		SRB_CUDA(c, cudaGetDeviceProperties(&prop, device));
		int const perSm = std::max(1, raster_ctas_per_sm());
		// persistent warps pulling work from a dispenser.  5 CTAs (20 warps) per SM instead of the 8 that would fit: the
		// rasteriser is issue-bound, and leaving registers free lets another frame's kernels share the SM (measured:
		// best frames/s with frames in flight, profiles/README.md)
		c->rasterCtasLatency = (uint32_t)(prop.multiProcessorCount * std::min(perSm, 5));
		c->rasterCtasThroughput = (uint32_t)(prop.multiProcessorCount * std::min(perSm, 3));
		c->rasterCtas = c->rasterCtasLatency;
		// tuning knobs for experiments (not part of the ABI)
		if (const char* e = getenv("SRB_RASTER_CTAS_PER_SM"))
		{
			c->rasterCtas = (uint32_t)(prop.multiProcessorCount * std::max(1, std::min(perSm, atoi(e))));
			c->rasterCtasLatency = c->rasterCtasThroughput = c->rasterCtas;
		}
		if (const char* e = getenv("SRB_MIN_UNIT"))
		{
			c->minUnit = (uint32_t)std::max(32, atoi(e)) & ~31u;
		}
	}
	// RCPPS table of this host's CPU (reference Rasterizer.cpp:375-376)
	std::vector<uint32_t> table(1u << 16);
	uint32_t bits = srb_harvest_rcp_table(table.data(), 16);
	if (bits == 0)
	{
		table.resize(1u << 23);
		bits = srb_harvest_rcp_table(table.data(), 23);
	}
	if (bits == 0)
	{
		return Fail(c, SRB_ERR_INVALID, "could not model this CPU's RCPPS with a mantissa table");
	}
	int rc = srb_set_rcp_table(c, table.data(), bits);
	if (rc != SRB_OK) return rc;
	// RSQRTPS table of this host's CPU (reference Viewer/SponzaScene.cpp:66)
	table.assign(2u << 16, 0u);
	bits = srb_harvest_rsqrt_table(table.data(), 16);
	if (bits == 0)
	{
		return Fail(c, SRB_ERR_INVALID, "could not model this CPU's RSQRTPS with an (exponent parity, mantissa) table");
	}
	SRB_CUDA(c, cudaMalloc((void**)&c->dSponza, sizeof(SponzaDev)));
	memset(&c->sponza, 0, sizeof(c->sponza));
	return srb_set_rsqrt_table(c, table.data(), bits);
}

SRB_API int srb_create(int device, uint32_t flags, srb_context** out) { return CreateContext(device, flags, nullptr, out); }

SRB_API int srb_create_shared(srb_context* parent, uint32_t flags, srb_context** out)
{
	if (!parent || !parent->res)
	{
		return SRB_ERR_INVALID;
	}
	if ((flags | parent->flags) & SRB_FLAG_UPLOAD_ALWAYS)
	{
		// upload-always re-writes the ONE device mirror of a host array every frame, on the uploading context's stream,
		// while a sibling's frame in flight may still be reading it: give every such context its own resources instead
		if (out) *out = nullptr;
		return Fail(parent, SRB_ERR_INVALID, "SRB_FLAG_UPLOAD_ALWAYS contexts cannot share resources (srb_create_shared)");
	}
	return CreateContext(parent->device, flags, parent->res, out);
}

SRB_API void srb_destroy(srb_context* c)
{
	if (!c)
	{
		return;
	}
	cudaSetDevice(c->device);
	if (c->stream) cudaStreamSynchronize(c->stream);
	if (c->blitStream) cudaStreamSynchronize(c->blitStream);
	if (c->res && --c->res->refs == 0)
	{
		for (Texture& t : c->res->textures) cudaFree(t.dev);
		for (Buffer& b : c->res->buffers) cudaFree(b.dev);
		cudaFree(c->res->dTexs);
		delete c->res;
	}
	c->res = nullptr;
	for (FrameBufferDev& f : c->fbs)
	{
		if (f.imported)
		{
			if (f.colour[0]) cudaIpcCloseMemHandle(f.colour[0]);
			if (f.rootDepth) cudaIpcCloseMemHandle(f.rootDepth);
			cudaFree(f.depth[0]);
			continue;
		}
		for (int p = 0; p < 2; ++p)
		{
			cudaFree(f.colour[p]);
			cudaFree(f.depth[p]);
		}
		cudaFree(f.linear);
	}
	cudaFree(c->dRcp);
	cudaFree(c->dRcp16);
	cudaFree(c->dRsqrt);
	cudaFree(c->dSponza);
	DropGraphs(c);
	cudaFree(c->dHead); // control block + draw table
	if (c->hHead) cudaFreeHost(c->hHead);
	cudaFree(c->dGather);
	cudaFree(c->dClipQueue);
	cudaFree(c->dRaster);
	cudaFree(c->dShade);
	cudaFree(c->dSurvivors);
	cudaFree(c->dRefs);
	cudaFree(c->dUnits);
	cudaFree(c->dTileCounts);
	cudaFree(c->dTileOffsets);
	cudaFree(c->dTileCursors);
	cudaFree(c->dTileKeys);
	cudaFree(c->dFlush);
	if (c->hCtl) cudaFreeHost(c->hCtl);
	for (int i = 0; i < kMaxTimers; ++i)
	{
		if (c->ev[i]) cudaEventDestroy(c->ev[i]);
	}
	for (int i = 0; i < 2; ++i)
	{
		if (c->evBlit[i]) cudaEventDestroy(c->evBlit[i]);
	}
	if (c->stream) cudaStreamDestroy(c->stream);
	if (c->blitStream && c->blitStream != c->stream) cudaStreamDestroy(c->blitStream);
	delete c;
}

SRB_API const char* srb_last_error(srb_context* c) { return c ? c->error.c_str() : "null context"; }

SRB_API int srb_set_rcp_table(srb_context* c, const uint32_t* table, uint32_t index_bits)
{
	if (!c || !table || index_bits < 1 || index_bits > 23)
	{
		return Fail(c, SRB_ERR_INVALID, "bad rcp table");
	}
	int rc = Bind(c);
	if (rc != SRB_OK) return rc;
	SRB_CUDA(c, cudaStreamSynchronize(c->stream));
	if (c->dRcp) cudaFree(c->dRcp);
	c->dRcp = nullptr;
	size_t const bytes = (size_t(1) << index_bits) * sizeof(uint32_t);
	SRB_CUDA(c, cudaMalloc((void**)&c->dRcp, bytes));
	SRB_CUDA(c, cudaMemcpy(c->dRcp, table, bytes, cudaMemcpyHostToDevice));
	c->rcpBits = index_bits;
	// the shade kernel keeps a 16-bit copy of the usual table in shared memory: entries 0x3F000000 + (e << 11), e < 2^16
	if (c->dRcp16) cudaFree(c->dRcp16);
	c->dRcp16 = nullptr;
	DropGraphs(c);
	static bool const noPacked = getenv("SRB_NO_PACKED_RCP") != nullptr; // A/B knob (not part of the ABI)
	if (index_bits == 11 && !noPacked)
	{
		std::vector<uint16_t> packed(1u << 11);
		bool ok = true;
		for (uint32_t i = 0; i < (1u << 11) && ok; ++i)
		{
			uint32_t const d = table[i] - 0x3F000000u;
			ok = table[i] >= 0x3F000000u && (d & 0x7FFu) == 0u && (d >> 11) <= 0xFFFFu;
			packed[i] = (uint16_t)(d >> 11);
		}
		if (ok)
		{
			SRB_CUDA(c, cudaMalloc((void**)&c->dRcp16, packed.size() * sizeof(uint16_t)));
			SRB_CUDA(c, cudaMemcpy(c->dRcp16, packed.data(), packed.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
		}
	}
	return SRB_OK;
}

SRB_API int srb_set_rsqrt_table(srb_context* c, const uint32_t* table, uint32_t index_bits)
{
	if (!c || !table || index_bits < 1 || index_bits > 22)
	{
		return Fail(c, SRB_ERR_INVALID, "bad rsqrt table");
	}
	int rc = Bind(c);
	if (rc != SRB_OK) return rc;
	SRB_CUDA(c, cudaStreamSynchronize(c->stream));
	if (c->dRsqrt) cudaFree(c->dRsqrt);
	c->dRsqrt = nullptr;
	size_t const bytes = (size_t(2) << index_bits) * sizeof(uint32_t);
	SRB_CUDA(c, cudaMalloc((void**)&c->dRsqrt, bytes));
	SRB_CUDA(c, cudaMemcpy(c->dRsqrt, table, bytes, cudaMemcpyHostToDevice));
	c->rsqrtBits = index_bits;
	return SRB_OK;
}

SRB_API int srb_set_sponza_constants(srb_context* c, const srb_sponza_constants* k)
{
	if (!c || !k)
	{
		return Fail(c, SRB_ERR_INVALID, "null sponza constants");
	}
	static_assert(sizeof(srb_sponza_constants) == sizeof(SponzaDev), "ABI and device layouts of the Sponza constants differ");
	c->sponza = *k;
	c->sponzaDirty = true;
	return SRB_OK;
}

SRB_API int srb_texture_create(srb_context* c, const uint8_t* texels, uint64_t bytes, const uint32_t* mip_offsets,
                               uint32_t num_mips, uint32_t width_log2, uint32_t height_log2, srb_handle* out)
{
	if (!c || !out || num_mips > SRB_MAX_TEX_DIM_LOG2 || width_log2 >= SRB_MAX_TEX_DIM_LOG2 ||
	    height_log2 >= SRB_MAX_TEX_DIM_LOG2 || (bytes && (!texels || !mip_offsets || !num_mips)))
	{
		return Fail(c, SRB_ERR_INVALID, "bad texture description");
	}
	int rc = Bind(c);
	if (rc != SRB_OK) return rc;
	Texture t;
	t.alive = true;
	memset(&t.desc, 0, sizeof(t.desc));
	if (bytes)
	{
		// every mip must fit: the sampler may address any texel of the padded level (Texture.cpp:159-175)
		for (uint32_t m = 0; m < num_mips; ++m)
		{
			uint32_t const w = std::max(1u, (1u << width_log2) >> m), h = std::max(1u, (1u << height_log2) >> m);
			uint64_t const need = uint64_t((w + 31u) & ~31u) * ((h + 31u) & ~31u) * 4u;
			if (mip_offsets[m] + need > bytes)
			{
				return Fail(c, SRB_ERR_INVALID, "mip %u (offset %u, %llu bytes) exceeds the texel blob (%llu bytes)", m,
				            mip_offsets[m], (unsigned long long)need, (unsigned long long)bytes);
			}
			t.desc.mipOffsets[m] = mip_offsets[m];
		}
		SRB_CUDA(c, cudaMalloc((void**)&t.dev, bytes));
		SRB_CUDA(c, cudaMemcpy(t.dev, texels, bytes, cudaMemcpyHostToDevice));
	}
	t.desc.texels = t.dev;
	t.desc.numMips = num_mips;
	t.desc.widthLog2 = width_log2;
	t.desc.heightLog2 = height_log2;
	t.desc.bytes = (uint32_t)bytes;
	c->res->textures.push_back(t);
	c->res->texGeneration++;
	*out = c->res->textures.size();
	return SRB_OK;
}

namespace
{
// One axis of one mip level's filter, packed for the device as gather lists (StbAxisDev): for every output the
// contributors that add to it, in ascending order — stb's order — each as (clamped source index, coefficient), from the
// per-contributor tables of srb_internal_stb_axis.  The lists depend only on (input size, output size); they are
// computed once per process and size pair (a scene's textures share a handful of sizes, and a square texture uses the
// same lists for both axes).
struct AxisPack
{
	std::vector<uint32_t> words; // entries (2 words each), then out + 1 offsets, padded to an even number of words
	uint32_t numEntries = 0;
	int out = 0;
};

const AxisPack& GetAxisPack(int inputSize, int outputSize)
{
	static std::mutex mtx;
	static std::map<std::pair<int, int>, AxisPack> cache;
	std::lock_guard<std::mutex> lock(mtx);
	auto it = cache.find({inputSize, outputSize});
	if (it != cache.end()) return it->second;
	AxisPack& p = cache[{inputSize, outputSize}];
	int *n0 = nullptr, *n1 = nullptr, margin = 0, num = 0;
	float* coef = nullptr;
	srb_internal_stb_axis(inputSize, outputSize, &margin, &n0, &n1, &coef, &num);
	p.out = outputSize;
	std::vector<uint32_t> count(outputSize + 1, 0);
	for (int j = 0; j < num; ++j)
	{
		for (int k = std::max(n0[j], 0); k <= n1[j] && k < outputSize; ++k) count[k + 1]++;
	}
	for (int k = 0; k < outputSize; ++k) count[k + 1] += count[k];
	p.numEntries = count[outputSize];
	p.words.assign(size_t(p.numEntries) * 2 + ((size_t(outputSize) + 2) & ~size_t(1)), 0u);
	std::vector<uint32_t> cursor(count.begin(), count.end() - 1);
	for (int j = 0; j < num; ++j) // ascending j: every list ends up in ascending contributor order
	{
		int const src = std::min(std::max(j - margin, 0), inputSize - 1); // clamped edge (stbir__edge_wrap, STBIR_EDGE_CLAMP)
		for (int k = std::max(n0[j], 0); k <= n1[j] && k < outputSize; ++k)
		{
			uint32_t const e = cursor[k]++;
			p.words[size_t(e) * 2] = uint32_t(src);
			memcpy(&p.words[size_t(e) * 2 + 1], &coef[size_t(j) * 4 + (k - n0[j])], 4);
		}
	}
	memcpy(&p.words[size_t(p.numEntries) * 2], count.data(), sizeof(uint32_t) * (outputSize + 1));
	srb_internal_stb_axis_free(n0, n1, coef);
	return p;
}

StbAxisDev AxisAt(const uint32_t* d, const AxisPack& p)
{
	StbAxisDev a;
	a.ent = reinterpret_cast<const int2*>(d);
	a.off = reinterpret_cast<const int*>(d + size_t(p.numEntries) * 2);
	return a;
}
} // namespace

SRB_API int srb_texture_create_rgba8(srb_context* c, const uint8_t* rgba, uint32_t width, uint32_t height, int calc_mips,
                                     srb_handle* out)
{
	if (!c || !out || !rgba)
	{
		return Fail(c, SRB_ERR_INVALID, "srb_texture_create_rgba8: null argument");
	}
	if (calc_mips != SRB_MIPS_NONE && calc_mips != SRB_MIPS_STB)
	{
		return Fail(c, SRB_ERR_INVALID, "srb_texture_create_rgba8 builds SRB_MIPS_NONE or SRB_MIPS_STB (box mips: srb_texture_build_rgba8)");
	}
	uint64_t bytes = 0;
	uint32_t offsets[SRB_MAX_TEX_DIM_LOG2] = {0}, numMips = 0;
	if (srb_texture_build_rgba8(nullptr, width, height, calc_mips, nullptr, &bytes, offsets, &numMips) != SRB_OK)
	{
		return Fail(c, SRB_ERR_INVALID, "texture size must be a power of two >= 32 and < 16384 (Texture.cpp:122-129)");
	}
	int rc = Bind(c);
	if (rc != SRB_OK) return rc;
	Texture t;
	t.alive = true;
	memset(&t.desc, 0, sizeof(t.desc));
	uint8_t* linear = nullptr;
	float* hbuf = nullptr;
	uint32_t* tables = nullptr;
	// The scratch buffers (linear image, horizontally filtered image, filter tables) come from the device's stream-ordered
	// pool, which keeps up to 1 GiB for the next build: cudaFree of buffers this size costs more than the kernels
	// (measured: 120 ms to free the 200 MB scratch of a 4096^2 build).
	{
		cudaMemPool_t pool = nullptr;
		SRB_CUDA(c, cudaDeviceGetDefaultMemPool(&pool, c->device));
		uint64_t keep = 1ull << 30;
		SRB_CUDA(c, cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
	}
	auto cleanup = [&](bool keepTexture) {
		if (linear) cudaFreeAsync(linear, c->stream);
		if (hbuf) cudaFreeAsync(hbuf, c->stream);
		if (tables) cudaFreeAsync(tables, c->stream);
		cudaStreamSynchronize(c->stream);
		if (!keepTexture) cudaFree(t.dev);
	};
	// SRB_TEXBUILD_TRACE=1: host-side time stamps of the phases on stderr (tuning aid, not part of the ABI)
	static bool const trace = getenv("SRB_TEXBUILD_TRACE") != nullptr;
	auto const tStart = std::chrono::steady_clock::now();
	auto stamp = [&](const char* what) {
		if (!trace) return;
		cudaStreamSynchronize(c->stream);
		fprintf(stderr, "texbuild %ux%u %-10s %8.3f ms\n", width, height, what,
		        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tStart).count());
	};
	auto build = [&]() -> int {
		size_t const imageBytes = size_t(width) * height * 4;
		SRB_CUDA(c, cudaMalloc((void**)&t.dev, bytes));
		SRB_CUDA(c, cudaMemsetAsync(t.dev, 0, bytes, c->stream)); // padding of levels smaller than a 32x32 tile
		SRB_CUDA(c, cudaMallocAsync((void**)&linear, imageBytes, c->stream));
		SRB_CUDA(c, cudaMemcpyAsync(linear, rgba, imageBytes, cudaMemcpyHostToDevice, c->stream));
		stamp("upload");
		launch_tex_tile(linear, t.dev + offsets[0], width, height, c->stream); // Texture.cpp:73-101,180
		c->launches++;
		stamp("tile");
		if (numMips > 1)
		{
			// filter tables of every level (both axes) in one upload
			std::vector<const AxisPack*> packs;
			std::vector<size_t> at;
			size_t words = 0;
			for (uint32_t m = 1; m < numMips; ++m)
			{
				for (int axis = 0; axis < 2; ++axis)
				{
					uint32_t const in = axis ? height : width;
					packs.push_back(&GetAxisPack(int(in), int(std::max(1u, in >> m))));
					at.push_back(words);
					words += packs.back()->words.size();
				}
			}
			std::vector<uint32_t> all(words);
			for (size_t i = 0; i < packs.size(); ++i) memcpy(&all[at[i]], packs[i]->words.data(), packs[i]->words.size() * 4);
			SRB_CUDA(c, cudaMallocAsync((void**)&tables, words * 4, c->stream));
			SRB_CUDA(c, cudaMemcpyAsync(tables, all.data(), words * 4, cudaMemcpyHostToDevice, c->stream)); // pageable: staged before it returns
			// the widest horizontally filtered image is level 1's: height rows of width/2 float4
			SRB_CUDA(c, cudaMallocAsync((void**)&hbuf, size_t(height) * std::max(1u, width >> 1) * 16, c->stream));
			stamp("tables");
			for (uint32_t m = 1; m < numMips; ++m) // Texture.cpp:188-198: every level from the ORIGINAL image
			{
				int const ow = int(std::max(1u, width >> m)), oh = int(std::max(1u, height >> m));
				StbAxisDev const H = AxisAt(tables + at[(m - 1) * 2], *packs[(m - 1) * 2]);
				StbAxisDev const V = AxisAt(tables + at[(m - 1) * 2 + 1], *packs[(m - 1) * 2 + 1]);
				launch_tex_hpass(linear, hbuf, int(width), int(height), ow, H, c->stream);
				launch_tex_vpass(hbuf, t.dev + offsets[m], int(height), ow, oh, V, c->stream);
				c->launches += 2;
				stamp("level");
			}
		}
		SRB_CUDA(c, cudaStreamSynchronize(c->stream)); // `rgba` is borrowed only for the duration of the call
		SRB_CUDA(c, cudaGetLastError());
		return SRB_OK;
	};
	rc = build();
	cleanup(rc == SRB_OK);
	stamp("freed");
	if (rc != SRB_OK) return rc;
	for (uint32_t m = 0; m < numMips; ++m) t.desc.mipOffsets[m] = offsets[m];
	t.desc.texels = t.dev;
	t.desc.numMips = numMips;
	t.desc.widthLog2 = 31u - (uint32_t)__builtin_clz(width);
	t.desc.heightLog2 = 31u - (uint32_t)__builtin_clz(height);
	t.desc.bytes = (uint32_t)bytes;
	c->res->textures.push_back(t);
	c->res->texGeneration++;
	*out = c->res->textures.size();
	return SRB_OK;
}

SRB_API int srb_texture_read(srb_context* c, srb_handle tex, uint8_t* texels_out, uint64_t cap, uint64_t* bytes_out,
                             uint32_t* mip_offsets_out, uint32_t* num_mips_out, uint32_t* width_log2_out,
                             uint32_t* height_log2_out)
{
	if (!c || !tex || tex > c->res->textures.size() || !c->res->textures[tex - 1].alive)
	{
		return Fail(c, SRB_ERR_INVALID, "bad texture handle");
	}
	Texture const& t = c->res->textures[tex - 1];
	if (bytes_out) *bytes_out = t.desc.bytes;
	if (mip_offsets_out) memcpy(mip_offsets_out, t.desc.mipOffsets, sizeof(t.desc.mipOffsets));
	if (num_mips_out) *num_mips_out = t.desc.numMips;
	if (width_log2_out) *width_log2_out = t.desc.widthLog2;
	if (height_log2_out) *height_log2_out = t.desc.heightLog2;
	if (texels_out)
	{
		if (cap < t.desc.bytes) return Fail(c, SRB_ERR_INVALID, "srb_texture_read: buffer too small");
		int const rc = Bind(c);
		if (rc != SRB_OK) return rc;
		SRB_CUDA(c, cudaStreamSynchronize(c->stream));
		SRB_CUDA(c, cudaMemcpy(texels_out, t.dev, t.desc.bytes, cudaMemcpyDeviceToHost));
	}
	return SRB_OK;
}

SRB_API int srb_texture_destroy(srb_context* c, srb_handle tex)
{
	if (!c || !tex || tex > c->res->textures.size() || !c->res->textures[tex - 1].alive)
	{
		return Fail(c, SRB_ERR_INVALID, "bad texture handle");
	}
	Bind(c);
	QuiesceResources(c);
	Texture& t = c->res->textures[tex - 1];
	cudaFree(t.dev);
	t.dev = nullptr;
	t.alive = false;
	memset(&t.desc, 0, sizeof(t.desc));
	c->res->texGeneration++;
	return SRB_OK;
}

SRB_API int srb_buffer_create(srb_context* c, const void* host, uint64_t bytes, srb_handle* out)
{
	if (!c || !out)
	{
		return Fail(c, SRB_ERR_INVALID, "bad arguments");
	}
	int rc = Bind(c);
	if (rc != SRB_OK) return rc;
	Buffer b;
	b.alive = true;
	b.bytes = bytes;
	// 16 bytes of slack: position fetches read a Vec3 at the start of the last element
	SRB_CUDA(c, cudaMalloc((void**)&b.dev, bytes + 16));
	if (host && bytes)
	{
		SRB_CUDA(c, cudaMemcpyAsync(b.dev, host, bytes, cudaMemcpyHostToDevice, c->stream));
		SRB_CUDA(c, cudaStreamSynchronize(c->stream));
	}
	c->res->buffers.push_back(b);
	*out = c->res->buffers.size();
	return SRB_OK;
}

SRB_API int srb_buffer_update(srb_context* c, srb_handle buf, uint64_t offset, const void* host, uint64_t bytes)
{
	if (!c || !buf || buf > c->res->buffers.size() || !c->res->buffers[buf - 1].alive || !host ||
	    offset > c->res->buffers[buf - 1].bytes || bytes > c->res->buffers[buf - 1].bytes - offset)
	{
		return Fail(c, SRB_ERR_INVALID, "bad buffer update");
	}
	int rc = Bind(c);
	if (rc != SRB_OK) return rc;
	QuiesceResources(c); // frames in flight (of any context of the family) may be reading the buffer
	SRB_CUDA(c, cudaMemcpyAsync(c->res->buffers[buf - 1].dev + offset, host, bytes, cudaMemcpyHostToDevice, c->stream));
	SRB_CUDA(c, cudaStreamSynchronize(c->stream));
	return SRB_OK;
}

SRB_API int srb_buffer_destroy(srb_context* c, srb_handle buf)
{
	if (!c || !buf || buf > c->res->buffers.size() || !c->res->buffers[buf - 1].alive)
	{
		return Fail(c, SRB_ERR_INVALID, "bad buffer handle");
	}
	Bind(c);
	QuiesceResources(c);
	cudaFree(c->res->buffers[buf - 1].dev);
	c->res->buffers[buf - 1] = Buffer{};
	return SRB_OK;
}

SRB_API int srb_invalidate_host(srb_context* c, const void* host)
{
	if (!c)
	{
		return SRB_ERR_INVALID;
	}
	auto it = c->res->mirrors.find(host);
	if (it != c->res->mirrors.end())
	{
		srb_buffer_destroy(c, it->second.buffer);
		c->res->mirrors.erase(it);
	}
	return SRB_OK;
}

SRB_API int srb_framebuffer_create(srb_context* c, uint32_t width, uint32_t height, srb_handle* out)
{
	if (!c || !out || !width || !height || width > 65535 || height > 65535)
	{
		return Fail(c, SRB_ERR_INVALID, "bad framebuffer size");
	}
	int rc = Bind(c);
	if (rc != SRB_OK) return rc;
	FrameBufferDev f;
	f.alive = true;
	f.width = width;
	f.height = height;
	f.tilesX = (width + SRB_BIN_DIM - 1) / SRB_BIN_DIM; // Renderer.cpp:45-46
	f.tilesY = (height + SRB_BIN_DIM - 1) / SRB_BIN_DIM;
	size_t const bytes = size_t(f.tilesX) * f.tilesY * 16384u;
	for (int p = 0; p < 2; ++p)
	{
		SRB_CUDA(c, cudaMalloc((void**)&f.colour[p], bytes));
		SRB_CUDA(c, cudaMalloc((void**)&f.depth[p], bytes + kSplitFlagBytes));
		SRB_CUDA(c, cudaMemset(f.colour[p], 0, bytes));
		SRB_CUDA(c, cudaMemset(f.depth[p], 0, bytes + kSplitFlagBytes));
	}
	SRB_CUDA(c, cudaMalloc((void**)&f.linear, size_t(width) * height * 4));
	for (int p = 0; p < 2; ++p)
	{
		SRB_CUDA(c, cudaEventCreateWithFlags(&f.planeRead[p], cudaEventDisableTiming));
	}
	c->fbs.push_back(f);
	*out = c->fbs.size();
	return SRB_OK;
}

SRB_API int srb_framebuffer_export(srb_context* c, srb_handle h, void* handles)
{
	FrameBufferDev* f = c ? GetFb(c, h) : nullptr;
	if (!f || !handles || f->imported)
	{
		return Fail(c, SRB_ERR_INVALID, "bad framebuffer export");
	}
	int rc = Bind(c);
	if (rc != SRB_OK) return rc;
	static_assert(2 * sizeof(cudaIpcMemHandle_t) <= SRB_FB_EXPORT_BYTES, "export blob too small");
	cudaIpcMemHandle_t hs[2];
	SRB_CUDA(c, cudaIpcGetMemHandle(&hs[0], f->colour[f->writePlane]));
	SRB_CUDA(c, cudaIpcGetMemHandle(&hs[1], f->depth[f->writePlane]));
	memcpy(handles, hs, sizeof(hs));
	f->exportedPlane = (int)f->writePlane;
	return SRB_OK;
}

SRB_API int srb_framebuffer_import(srb_context* c, const void* handles, uint32_t width, uint32_t height, srb_handle* out)
{
	if (!c || !handles || !out || !width || !height)
	{
		return Fail(c, SRB_ERR_INVALID, "bad framebuffer import");
	}
	int rc = Bind(c);
	if (rc != SRB_OK) return rc;
	cudaIpcMemHandle_t hs[2];
	memcpy(hs, handles, sizeof(hs));
	FrameBufferDev f;
	f.alive = true;
	f.imported = true;
	f.width = width;
	f.height = height;
	f.tilesX = (width + SRB_BIN_DIM - 1) / SRB_BIN_DIM;
	f.tilesY = (height + SRB_BIN_DIM - 1) / SRB_BIN_DIM;
	SRB_CUDA(c, cudaIpcOpenMemHandle((void**)&f.colour[0], hs[0], cudaIpcMemLazyEnablePeerAccess));
	SRB_CUDA(c, cudaIpcOpenMemHandle((void**)&f.rootDepth, hs[1], cudaIpcMemLazyEnablePeerAccess));
	// Only COLOUR is composited into the root's framebuffer.  Depth is per-GPU state of the tiles a GPU owns (frames
	// without a depth clear read it back), so it stays in this GPU's memory: half the bytes over the root's NVLink ingress.
	size_t const bytes = size_t(f.tilesX) * f.tilesY * 16384u;
	SRB_CUDA(c, cudaMalloc((void**)&f.depth[0], bytes));
	SRB_CUDA(c, cudaMemset(f.depth[0], 0, bytes));
	f.colour[1] = f.colour[0];
	f.depth[1] = f.depth[0];
	c->fbs.push_back(f);
	*out = c->fbs.size();
	return SRB_OK;
}

SRB_API int srb_set_tile_ownership(srb_context* c, uint32_t modulus, uint32_t remainder)
{
	if (!c || modulus == 0 || remainder >= modulus)
	{
		return Fail(c, SRB_ERR_INVALID, "bad tile ownership");
	}
	c->ownMod = modulus;
	c->ownRem = remainder;
	return SRB_OK;
}

SRB_API int srb_framebuffer_destroy(srb_context* c, srb_handle h)
{
	FrameBufferDev* f = c ? GetFb(c, h) : nullptr;
	if (!f)
	{
		return Fail(c, SRB_ERR_INVALID, "bad framebuffer handle");
	}
	Bind(c);
	cudaStreamSynchronize(c->stream);
	cudaStreamSynchronize(c->blitStream);
	for (int p = 0; p < 2; ++p)
	{
		if (f->planeRead[p]) cudaEventDestroy(f->planeRead[p]);
	}
	if (f->imported)
	{
		cudaIpcCloseMemHandle(f->colour[0]);
		cudaIpcCloseMemHandle(f->rootDepth);
		cudaFree(f->depth[0]);
	}
	else
	{
		for (int p = 0; p < 2; ++p)
		{
			cudaFree(f->colour[p]);
			cudaFree(f->depth[p]);
		}
		cudaFree(f->linear);
	}
	*f = FrameBufferDev{};
	return SRB_OK;
}

SRB_API int srb_framebuffer_info(srb_context* c, srb_handle h, uint32_t* width, uint32_t* height, uint32_t* tiles_x,
                                 uint32_t* tiles_y)
{
	FrameBufferDev* f = c ? GetFb(c, h) : nullptr;
	if (!f)
	{
		return Fail(c, SRB_ERR_INVALID, "bad framebuffer handle");
	}
	if (width) *width = f->width;
	if (height) *height = f->height;
	if (tiles_x) *tiles_x = f->tilesX;
	if (tiles_y) *tiles_y = f->tilesY;
	return SRB_OK;
}

SRB_API int srb_begin_frame(srb_context* c)
{
	if (!c)
	{
		return SRB_ERR_INVALID;
	}
	c->frameSerial = ++c->res->frameSerial; // unique across the contexts of a family (they share the mirrors)
	c->recDraws.clear();
	c->recInputTris = 0;
	c->recUsesSponza = false;
	c->recFb = 0;
	c->recFbs.clear();
	c->recDrawFb.clear();
	c->inFrame = true;
	return SRB_OK;
}

SRB_API int srb_clear(srb_context* c, srb_handle h, uint32_t color, int clear_colour, int clear_depth)
{
	FrameBufferDev* f = c ? GetFb(c, h) : nullptr;
	if (!f)
	{
		return Fail(c, SRB_ERR_INVALID, "bad framebuffer handle");
	}
	NoteFrameBuffer(c, h);
	// The clear is folded into the tile kernel: every tile of the frame is written exactly once.
	if (clear_colour)
	{
		uint32_t const b = color & 0xFFu; // memset semantics, Renderer.cpp:191
		f->clearWord = b | (b << 8) | (b << 16) | (b << 24);
		f->pendingClearColour = true;
	}
	if (clear_depth)
	{
		f->pendingClearDepth = true;
	}
	return SRB_OK;
}

SRB_API int srb_draw_indexed(srb_context* c, const srb_draw_desc* d)
{
	if (!c || !d)
	{
		return SRB_ERR_INVALID;
	}
	if (!c->inFrame)
	{
		return Fail(c, SRB_ERR_INVALID, "srb_draw_indexed outside srb_begin_frame/srb_end_frame");
	}
	if (d->shader >= SRB_SHADER_COUNT)
	{
		return Fail(c, SRB_ERR_UNKNOWN_SHADER, "pixel shader %u is not in the device registry", d->shader);
	}
	if (!GetFb(c, d->framebuffer))
	{
		return Fail(c, SRB_ERR_INVALID, "bad framebuffer handle in draw");
	}
	if (d->indices.stride != 1 && d->indices.stride != 2 && d->indices.stride != 4)
	{
		return Fail(c, SRB_ERR_INVALID, "index stride must be 1, 2 or 4 (Binning.cpp:167-205)");
	}
	if (d->attributes.stride > 4 * SRB_MAX_VARYINGS || (d->attributes.stride & 3u))
	{
		return Fail(c, SRB_ERR_INVALID, "attribute stride must be a multiple of 4 and <= 32 bytes");
	}
	if (d->positions.stride < 12 || (d->positions.stride & 3u))
	{
		return Fail(c, SRB_ERR_INVALID, "position stride must be a multiple of 4 and >= 12 bytes");
	}
	if (d->texture && (d->texture > c->res->textures.size() || !c->res->textures[d->texture - 1].alive || d->texture > 0xFFFEu))
	{
		return Fail(c, SRB_ERR_INVALID, "bad texture handle in draw (at most 65534 textures per context)");
	}
	int rc = Bind(c);
	if (rc != SRB_OK) return rc;
	NoteFrameBuffer(c, d->framebuffer);
	DrawDev dd;
	memset(&dd, 0, sizeof(dd));
	uint32_t const numTris = d->indices.num / 3; // Renderer.cpp:247
	rc = Resolve(c, d->indices, uint64_t(d->indices.num) * d->indices.stride, d->indices.stride, &dd.idx);
	if (rc != SRB_OK) return rc;
	rc = Resolve(c, d->positions, uint64_t(d->positions.num) * d->positions.stride, 4, &dd.pos);
	if (rc != SRB_OK) return rc;
	rc = Resolve(c, d->attributes, uint64_t(d->attributes.num) * d->attributes.stride, 4, &dd.attr);
	if (rc != SRB_OK) return rc;
	if (numTris && (!dd.idx || !dd.pos || (d->attributes.stride && !dd.attr)))
	{
		return Fail(c, SRB_ERR_INVALID, "draw is missing a buffer");
	}
	dd.idxStride = d->indices.stride;
	dd.posStride = d->positions.stride;
	dd.attrStride = d->attributes.stride;
	dd.numTris = numTris;
	dd.triBase = c->recInputTris;
	dd.shader = d->shader;
	dd.uvOffset = d->uv_offset;
	dd.numVaryings = d->attributes.stride / 4;
	{
		// planes the pixel shader reads (Viewer/Shaders.h:71-130; derivatives come from uvOffset, uvOffset + 1,
		// Rasterizer.cpp:378-399)
		uint32_t const all = dd.numVaryings >= 32u ? 0xFFFFFFFFu : ((1u << dd.numVaryings) - 1u);
		uint32_t need = all;
		if (!(c->flags & SRB_FLAG_FULL_RECORDS))
		{
			switch (d->shader)
			{
				case SRB_SHADER_VISUALIZE_NORMALS: need = 0x38u; break;
				case SRB_SHADER_VISUALIZE_UVS: need = 0xC0u; break;
				case SRB_SHADER_SPONZA: // position, normal, uv (SponzaScene.cpp:25-32) + the derivative pair
					need = d->texture ? (0xFFu | (d->uv_offset < 32u ? 1u << d->uv_offset : 0u) |
					                     (d->uv_offset + 1u < 32u ? 1u << (d->uv_offset + 1u) : 0u))
					                  : 0u;
					break;
				case SRB_SHADER_UNLIT_DIFFUSE:
					need = d->texture ? (0xC0u | (d->uv_offset < 32u ? 1u << d->uv_offset : 0u) |
					                     (d->uv_offset + 1u < 32u ? 1u << (d->uv_offset + 1u) : 0u))
					                  : 0u; // a null texture shades constant white
					break;
				default: break;
			}
		}
		// bit 8: the shader may read the record's second half (varyings 0..5), so it must be written even if only with zeros
		dd.planeMask = (need & all & 0xFFu) | ((need & 0x3Fu) ? 0x100u : 0u);
	}
	dd.texture = d->texture ? (int32_t)(d->texture - 1) : -1;
	memcpy(dd.mvp, d->mvp, sizeof(dd.mvp));
	if (uint64_t(c->recInputTris) + numTris > 0x7FFFFFFFull)
	{
		return Fail(c, SRB_ERR_OVERFLOW, "too many triangles in one frame");
	}
	c->recInputTris += numTris;
	c->recUsesSponza = c->recUsesSponza || d->shader == SRB_SHADER_SPONZA;
	c->recDraws.push_back(dd);
	c->recDrawFb.push_back(d->framebuffer);
	return SRB_OK;
}

SRB_API int srb_end_frame_async(srb_context* c)
{
	if (!c || !c->inFrame)
	{
		return Fail(c, SRB_ERR_INVALID, "srb_end_frame without srb_begin_frame");
	}
	int rc = Bind(c);
	if (rc != SRB_OK) return rc;
	if (c->framePending)
	{
		rc = Finish(c); // the previous frame's overflow check must precede re-use of its buffers
		if (rc != SRB_OK) return rc;
	}
	c->inFrame = false;
	if (!c->recFb)
	{
		return SRB_OK; // nothing recorded
	}
	c->readbackDst = c->nextReadbackDst;
	c->readbackBytes = c->nextReadbackBytes;
	c->nextReadbackDst = nullptr;
	c->frameUsesSponza = c->recUsesSponza;
	if (c->ownMod > 1u)
	{
		c->splitSerial++; // one stamp per frame, the same on every rank (a frame that is re-run keeps its stamp)
	}
	if (c->recFbs.size() <= 1)
	{
		c->draws.swap(c->recDraws);
		c->numInputTris = c->recInputTris;
		c->frameFb = c->recFb;
		ConsumeClear(c);
		return Submit(c);
	}
	// Several framebuffers in one frame: the reference bins all draws together and its tile tasks write through each draw's
	// own framebuffer pointer (Rasterizer.cpp:525-577), i.e. the framebuffers are independent — so the frame is the sequence
	// of one pipeline pass per framebuffer over its draws (submission order kept inside a pass).  Passes but the last are
	// completed before the next one re-uses the context's frame state; counters add up.
	srb_counters sum{};
	for (size_t k = 0; k < c->recFbs.size(); ++k)
	{
		srb_handle const h = c->recFbs[k];
		c->draws.clear();
		uint32_t tris = 0;
		for (size_t i = 0; i < c->recDraws.size(); ++i)
		{
			if (c->recDrawFb[i] != h) continue;
			DrawDev dd = c->recDraws[i];
			dd.triBase = tris;
			tris += dd.numTris;
			c->draws.push_back(dd);
		}
		c->numInputTris = tris;
		c->frameFb = h;
		ConsumeClear(c);
		void* const rb = c->readbackDst;
		if (h != c->recFb) c->readbackDst = nullptr; // (srb_render_frames reads back the first framebuffer)
		rc = Submit(c);
		c->readbackDst = rb;
		if (rc != SRB_OK) return rc;
		rc = Finish(c);
		if (rc != SRB_OK) return rc;
		sum.tris_in += c->counters.tris_in;
		sum.tris_setup += c->counters.tris_setup;
		sum.tris_clipped += c->counters.tris_clipped;
		sum.tile_refs += c->counters.tile_refs;
		sum.tiles_nonempty += c->counters.tiles_nonempty;
		sum.max_refs_in_tile = std::max(sum.max_refs_in_tile, c->counters.max_refs_in_tile);
		sum.pixels_covered += c->counters.pixels_covered;
		sum.overflow |= c->counters.overflow;
	}
	c->counters = sum;
	return SRB_OK;
}

SRB_API int srb_sync(srb_context* c)
{
	if (!c)
	{
		return SRB_ERR_INVALID;
	}
	int rc = Bind(c);
	if (rc != SRB_OK) return rc;
	rc = Finish(c);
	if (rc != SRB_OK) return rc;
	SRB_CUDA(c, cudaStreamSynchronize(c->blitStream)); // blits in flight (their callbacks have run when this returns)
	return SRB_OK;
}

SRB_API int srb_end_frame(srb_context* c)
{
	int rc = srb_end_frame_async(c);
	if (rc != SRB_OK)
	{
		return rc;
	}
	return srb_sync(c);
}

SRB_API int srb_read_tiles(srb_context* c, srb_handle h, void* colour_tiles, void* depth_tiles, uint64_t depth_stride)
{
	FrameBufferDev* f = c ? GetFb(c, h) : nullptr;
	if (!f)
	{
		return Fail(c, SRB_ERR_INVALID, "bad framebuffer handle");
	}
	int rc = srb_sync(c);
	if (rc != SRB_OK) return rc;
	size_t const n = size_t(f->tilesX) * f->tilesY;
	if (colour_tiles)
	{
		SRB_CUDA(c, cudaMemcpyAsync(colour_tiles, f->colour[f->writePlane], n * 16384u, cudaMemcpyDeviceToHost, c->stream));
	}
	if (depth_tiles)
	{
		if (depth_stride < 16384u)
		{
			return Fail(c, SRB_ERR_INVALID, "depth stride < 16384");
		}
		SRB_CUDA(c, cudaMemcpy2DAsync(depth_tiles, depth_stride, f->depth[f->writePlane], 16384u, 16384u, n,
		                              cudaMemcpyDeviceToHost, c->stream));
	}
	SRB_CUDA(c, cudaStreamSynchronize(c->stream));
	return SRB_OK;
}

SRB_API int srb_blit_linear(srb_context* c, srb_handle h, uint8_t* linear_pixels, void (*on_finish)(void*), void* user)
{
	FrameBufferDev* f = c ? GetFb(c, h) : nullptr;
	if (!f || !linear_pixels || f->imported)
	{
		return Fail(c, SRB_ERR_INVALID, "bad blit arguments (blit from the process that owns the framebuffer)");
	}
	int rc = Bind(c);
	if (rc != SRB_OK) return rc;
	if (c->framePending)
	{
		rc = Finish(c);
		if (rc != SRB_OK) return rc;
	}
	// The frame is complete (the frame stream was synchronised above, or is idle).  Like the reference's blit job
	// (Renderer.cpp:319-372) the de-tile, the read-back and the callback run BESIDE the next frame, which renders into
	// the other plane: they go to the blit stream, and a later frame that wants this plane back waits for planeRead.
	SRB_CUDA(c, cudaStreamSynchronize(c->stream));
	if (c->timing) SRB_CUDA(c, cudaEventRecord(c->evBlit[0], c->blitStream));
	launch_detile(reinterpret_cast<const uint32_t*>(f->colour[f->writePlane]), f->linear, f->width, f->height, f->tilesX,
	              c->blitStream);
	c->launches++;
	if (c->timing) SRB_CUDA(c, cudaEventRecord(c->evBlit[1], c->blitStream));
	c->blitTimed = c->timing;
	SRB_CUDA(c, cudaEventRecord(f->planeRead[f->writePlane], c->blitStream));
	f->planeBusy[f->writePlane] = true;
	SRB_CUDA(c, cudaMemcpyAsync(linear_pixels, f->linear, size_t(f->width) * f->height * 4, cudaMemcpyDeviceToHost,
	                            c->blitStream));
	BlitCallback* cb = new BlitCallback{on_finish, user};
	SRB_CUDA(c, cudaLaunchHostFunc(c->blitStream, BlitDone, cb));
	f->writePlane ^= 1u; // FrameBuffer::SwapPlanes, Renderer.cpp:369
	return SRB_OK;
}

SRB_API int srb_get_counters(srb_context* c, srb_counters* out)
{
	if (!c || !out)
	{
		return SRB_ERR_INVALID;
	}
	int rc = srb_sync(c);
	if (rc != SRB_OK) return rc;
	*out = c->counters;
	return SRB_OK;
}

SRB_API int srb_set_timing(srb_context* c, int enabled)
{
	if (!c) return SRB_ERR_INVALID;
	c->timing = enabled != 0;
	return SRB_OK;
}

SRB_API int srb_get_kernel_times(srb_context* c, float* micros, const char** names, uint32_t cap, uint32_t* n)
{
	// "clip" includes the tile scan unless it was launched separately (SRB_SEPARATE_SCAN), in which case "tile_scan" is non-zero
	static const char* kNames[8] = {"upload+reset", "setup", "clip", "tile_scan", "bin_fill", "raster", "shade", "detile"};
	if (!c || !n)
	{
		return SRB_ERR_INVALID;
	}
	if (c->blitTimed)
	{
		// the de-tile kernel of the last Blit issued in timing mode (0 if none)
		float ms = 0.0f;
		if (cudaEventSynchronize(c->evBlit[1]) == cudaSuccess && cudaEventElapsedTime(&ms, c->evBlit[0], c->evBlit[1]) == cudaSuccess)
		{
			c->kernelMicros[7] = ms * 1000.0f;
		}
	}
	*n = 8;
	for (uint32_t i = 0; i < 8 && i < cap; ++i)
	{
		if (micros) micros[i] = c->kernelMicros[i];
		if (names) names[i] = kNames[i];
	}
	return SRB_OK;
}

SRB_API uint64_t srb_launch_count(srb_context* c) { return c ? c->launches : 0; }

SRB_API int srb_dump_tile_counts(srb_context* c, uint32_t* counts, uint32_t num_tiles)
{
	if (!c || !counts)
	{
		return SRB_ERR_INVALID;
	}
	int rc = srb_sync(c);
	if (rc != SRB_OK) return rc;
	if (!c->frameValid || num_tiles != c->lastArgs.fp.tilesX * c->lastArgs.fp.tilesY)
	{
		return Fail(c, SRB_ERR_INVALID, "no completed frame / wrong tile count");
	}
	std::vector<uint32_t> offs(num_tiles + 1);
	SRB_CUDA(c, cudaMemcpy(offs.data(), c->dTileOffsets, (num_tiles + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	for (uint32_t i = 0; i < num_tiles; ++i)
	{
		counts[i] = offs[i + 1] - offs[i];
	}
	return SRB_OK;
}

static int TileRange(srb_context* c, uint32_t tile, uint32_t* begin, uint32_t* count)
{
	int rc = srb_sync(c);
	if (rc != SRB_OK) return rc;
	if (!c->frameValid || tile >= c->lastArgs.fp.tilesX * c->lastArgs.fp.tilesY)
	{
		return Fail(c, SRB_ERR_INVALID, "no completed frame / bad tile index");
	}
	uint32_t o[2];
	SRB_CUDA(c, cudaMemcpy(o, c->dTileOffsets + tile, sizeof(o), cudaMemcpyDeviceToHost));
	*begin = o[0];
	*count = o[1] - o[0];
	return SRB_OK;
}

// The list of one tile in CANONICAL order: its (key, slot) entries sorted by key (the lists themselves are sets).
static int SortedTileList(srb_context* c, uint32_t tile, std::vector<KeySlot>& list)
{
	uint32_t begin, count;
	int rc = TileRange(c, tile, &begin, &count);
	if (rc != SRB_OK) return rc;
	list.resize(count);
	if (count)
	{
		std::vector<TileRef> raw(count);
		SRB_CUDA(c, cudaMemcpy(raw.data(), c->dRefs + begin, count * sizeof(TileRef), cudaMemcpyDeviceToHost));
		for (uint32_t i = 0; i < count; ++i)
		{
			list[i].key = raw[i].key;
			list[i].slot = raw[i].slot;
		}
		std::sort(list.begin(), list.end(), [](KeySlot const& a, KeySlot const& b) { return a.key < b.key; });
	}
	return SRB_OK;
}

SRB_API int srb_dump_tile_ranks(srb_context* c, uint32_t tile, uint32_t* out, uint32_t cap, uint32_t* n)
{
	if (!c || !n)
	{
		return SRB_ERR_INVALID;
	}
	std::vector<KeySlot> list;
	int rc = SortedTileList(c, tile, list);
	if (rc != SRB_OK) return rc;
	*n = (uint32_t)list.size();
	for (uint32_t i = 0; out && i < list.size() && i < cap; ++i)
	{
		out[i] = list[i].key;
	}
	return list.size() <= cap ? SRB_OK : SRB_ERR_OVERFLOW;
}

SRB_API int srb_dump_tile_tris(srb_context* c, uint32_t tile, srb_tile_tri* out, uint32_t cap, uint32_t* n)
{
	if (!c || !n)
	{
		return SRB_ERR_INVALID;
	}
	if (out && !(c->flags & SRB_FLAG_FULL_RECORDS))
	{
		return Fail(c, SRB_ERR_INVALID, "srb_dump_tile_tris reports every varying's plane: create the context with SRB_FLAG_FULL_RECORDS");
	}
	std::vector<KeySlot> list;
	int rc = SortedTileList(c, tile, list);
	if (rc != SRB_OK) return rc;
	uint32_t const count = (uint32_t)list.size();
	*n = count;
	uint32_t const m = std::min(cap, count);
	if (out && m)
	{
		srb_tile_tri* d = nullptr;
		KeySlot* dl = nullptr;
		SRB_CUDA(c, cudaMalloc((void**)&d, m * sizeof(srb_tile_tri)));
		SRB_CUDA(c, cudaMalloc((void**)&dl, count * sizeof(KeySlot)));
		SRB_CUDA(c, cudaMemcpy(dl, list.data(), count * sizeof(KeySlot), cudaMemcpyHostToDevice));
		launch_dump_tile_tris(c->lastArgs, tile, dl, count, d, m, c->stream);
		c->launches++;
		cudaError_t e = cudaMemcpyAsync(out, d, m * sizeof(srb_tile_tri), cudaMemcpyDeviceToHost, c->stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
		cudaFree(d);
		cudaFree(dl);
		SRB_CUDA(c, e);
	}
	return count <= cap ? SRB_OK : SRB_ERR_OVERFLOW;
}

SRB_API int srb_dump_tile_coverage(srb_context* c, uint32_t tile, uint64_t* masks, uint32_t cap_entries, uint32_t* n)
{
	if (!c || !n)
	{
		return SRB_ERR_INVALID;
	}
	std::vector<KeySlot> list;
	int rc = SortedTileList(c, tile, list);
	if (rc != SRB_OK) return rc;
	uint32_t const count = (uint32_t)list.size();
	*n = count;
	uint32_t const m = std::min(cap_entries, count);
	if (masks && m)
	{
		unsigned long long* d = nullptr;
		KeySlot* dl = nullptr;
		SRB_CUDA(c, cudaMalloc((void**)&d, size_t(m) * 64 * sizeof(uint64_t)));
		SRB_CUDA(c, cudaMalloc((void**)&dl, count * sizeof(KeySlot)));
		SRB_CUDA(c, cudaMemcpy(dl, list.data(), count * sizeof(KeySlot), cudaMemcpyHostToDevice));
		launch_dump_tile_coverage(c->lastArgs, tile, dl, count, d, m, c->stream);
		c->launches++;
		cudaError_t e = cudaMemcpyAsync(masks, d, size_t(m) * 64 * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
		cudaFree(d);
		cudaFree(dl);
		SRB_CUDA(c, e);
	}
	return count <= cap_entries ? SRB_OK : SRB_ERR_OVERFLOW;
}

SRB_API int srb_dump_winners(srb_context* c, uint32_t* winners, uint64_t num_pixels)
{
	if (!c || !winners)
	{
		return SRB_ERR_INVALID;
	}
	int rc = srb_sync(c);
	if (rc != SRB_OK) return rc;
	uint64_t const n = uint64_t(c->lastArgs.fp.tilesX) * c->lastArgs.fp.tilesY * 4096u;
	if (!c->frameValid || num_pixels != n)
	{
		return Fail(c, SRB_ERR_INVALID, "no completed frame / wrong pixel count");
	}
	// Re-run the tile kernel of the last frame into scratch tiles (the framebuffer is left untouched).
	RasterArgs A = c->lastArgs;
	uint8_t* scratch = nullptr;
	SRB_CUDA(c, cudaMalloc((void**)&scratch, n * 12));
	A.colourTiles = scratch;
	A.depthTiles = scratch + n * 4;
	A.winnersOut = reinterpret_cast<uint32_t*>(scratch + n * 8);
	A.clearColour = 1;
	A.clearDepth = 1;
	if (!c->lastClearDepth)
	{
		cudaFree(scratch);
		return Fail(c, SRB_ERR_INVALID, "srb_dump_winners needs a frame that began with a depth clear");
	}
	// the unit table and lists of the last frame are still in place; rewind the unit dispenser
	SRB_CUDA(c, cudaMemsetAsync(&c->dCtl->unitTicket, 0, sizeof(uint32_t), c->stream));
	launch_raster(A, c->rasterCtas, c->stream);
	launch_shade(A, c->stream);
	c->launches += 2;
	cudaError_t e = cudaMemcpyAsync(winners, A.winnersOut, n * 4, cudaMemcpyDeviceToHost, c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	cudaFree(scratch);
	SRB_CUDA(c, e);
	return SRB_OK;
}

SRB_API void* srb_host_alloc(uint64_t bytes)
{
	void* p = nullptr;
	if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess)
	{
		return nullptr;
	}
	return p;
}

namespace
{
std::mutex g_hugeMutex;
std::unordered_map<void*, size_t> g_hugeAllocs; // srb_host_alloc_ex(SRB_HOST_HUGE_PAGES) blocks: registered, not cudaHostAlloc'ed
} // namespace

SRB_API void* srb_host_alloc_ex(uint64_t bytes, uint32_t flags)
{
	void* p = nullptr;
	if (flags & SRB_HOST_HUGE_PAGES)
	{
		size_t const huge = size_t(2) << 20;
		size_t const n = (size_t(bytes) + huge - 1) & ~(huge - 1);
		if (posix_memalign(&p, huge, n) != 0)
		{
			return nullptr;
		}
		madvise(p, n, MADV_HUGEPAGE);
		memset(p, 0, n); // touch: the pages exist (as huge pages where the kernel grants them) before they are locked
		if (cudaHostRegister(p, n, (flags & SRB_HOST_PORTABLE) ? cudaHostRegisterPortable : cudaHostRegisterDefault) != cudaSuccess)
		{
			cudaGetLastError();
			free(p);
			return nullptr;
		}
		std::lock_guard<std::mutex> lock(g_hugeMutex);
		g_hugeAllocs[p] = n;
		return p;
	}
	unsigned int f = cudaHostAllocDefault;
	if (flags & SRB_HOST_WRITE_COMBINED) f |= cudaHostAllocWriteCombined;
	if (flags & SRB_HOST_PORTABLE) f |= cudaHostAllocPortable;
	if (cudaHostAlloc(&p, bytes, f) != cudaSuccess)
	{
		return nullptr;
	}
	return p;
}

SRB_API int srb_debug_d2h_copies(srb_context* c, void* host, uint64_t bytes, uint32_t reps, float* ms)
{
	if (!c || !host || !bytes || !ms)
	{
		return SRB_ERR_INVALID;
	}
	int rc = Bind(c);
	if (rc != SRB_OK) return rc;
	if (bytes > c->flushBytes)
	{
		if (c->dFlush) cudaFree(c->dFlush);
		c->dFlush = nullptr;
		SRB_CUDA(c, cudaMalloc((void**)&c->dFlush, bytes));
		c->flushBytes = bytes;
	}
	SRB_CUDA(c, cudaEventRecord(c->marks[2], c->stream));
	for (uint32_t i = 0; i < reps; ++i)
	{
		SRB_CUDA(c, cudaMemcpyAsync(host, c->dFlush, bytes, cudaMemcpyDeviceToHost, c->stream));
	}
	SRB_CUDA(c, cudaEventRecord(c->marks[3], c->stream));
	SRB_CUDA(c, cudaEventSynchronize(c->marks[3]));
	SRB_CUDA(c, cudaEventElapsedTime(ms, c->marks[2], c->marks[3]));
	return SRB_OK;
}

SRB_API void srb_host_free(void* p)
{
	if (!p)
	{
		return;
	}
	{
		std::lock_guard<std::mutex> lock(g_hugeMutex);
		auto it = g_hugeAllocs.find(p);
		if (it != g_hugeAllocs.end())
		{
			g_hugeAllocs.erase(it);
			cudaHostUnregister(p);
			free(p);
			return;
		}
	}
	cudaFreeHost(p);
}

SRB_API int srb_timer_mark(srb_context* c, uint32_t slot)
{
	if (!c || slot >= 4)
	{
		return SRB_ERR_INVALID;
	}
	int rc = Bind(c);
	if (rc != SRB_OK) return rc;
	SRB_CUDA(c, cudaEventRecord(c->marks[slot], c->stream));
	return SRB_OK;
}

SRB_API int srb_timer_elapsed(srb_context* a, uint32_t slot_a, srb_context* b, uint32_t slot_b, float* ms)
{
	if (!a || !b || slot_a >= 4 || slot_b >= 4 || !ms || a->device != b->device)
	{
		return SRB_ERR_INVALID;
	}
	int rc = Bind(b);
	if (rc != SRB_OK) return rc;
	SRB_CUDA(b, cudaEventSynchronize(b->marks[slot_b]));
	SRB_CUDA(b, cudaEventElapsedTime(ms, a->marks[slot_a], b->marks[slot_b]));
	return SRB_OK;
}

SRB_API int srb_flush_l2(srb_context* c, uint64_t bytes)
{
	if (!c)
	{
		return SRB_ERR_INVALID;
	}
	int rc = Bind(c);
	if (rc != SRB_OK) return rc;
	if (bytes > c->flushBytes)
	{
		if (c->dFlush) cudaFree(c->dFlush);
		c->dFlush = nullptr;
		SRB_CUDA(c, cudaMalloc((void**)&c->dFlush, bytes));
		c->flushBytes = bytes;
	}
	SRB_CUDA(c, cudaMemsetAsync(c->dFlush, (int)(c->flushCount++ & 0xFF), bytes, c->stream));
	return SRB_OK;
}

SRB_API int srb_set_frames_in_flight_hint(srb_context* c, uint32_t frames_in_flight)
{
	if (!c)
	{
		return SRB_ERR_INVALID;
	}
	// With several frames in flight (several contexts of one device) the kernels of different frames share the SMs:
	// smaller resident footprints let them interleave — an SM that holds CTAs of different kernels (integer-heavy raster,
	// load-bound set-up, FP- and fetch-heavy shade) issues more than one that holds many CTAs of one kernel (measured:
	// +10 % frames/s at 8 in flight over the full-size grids, profiles/README.md); a single frame in flight wants the
	// larger grids (lower latency).
	bool const many = frames_in_flight >= 4u;
	if (!getenv("SRB_RASTER_CTAS_PER_SM"))
	{
		c->rasterCtas = many ? c->rasterCtasThroughput : c->rasterCtasLatency;
	}
	c->shadeCtasPerSm = many ? 6u : 0u;
	c->setupCtasPerSm = many ? 1u : 0u; // the set-up kernel waits on dependent loads: one CTA per SM leaves the registers to the others
	return SRB_OK;
}

SRB_API int srb_render_frames(const srb_batch_item* items, uint32_t n_items, uint32_t n_draws, const float* mvps,
                              uint32_t frames, uint32_t clear_color, void* colour_out, uint64_t colour_stride)
{
	if (!items || !n_items)
	{
		return SRB_ERR_INVALID;
	}
	for (uint32_t i = 0; i < n_items; ++i)
	{
		srb_set_frames_in_flight_hint(items[i].ctx, n_items);
	}
	std::vector<srb_draw_desc> d(n_draws);
	for (uint32_t f = 0; f < frames; ++f)
	{
		srb_batch_item const& it = items[f % n_items];
		srb_context* c = it.ctx;
		int rc = srb_begin_frame(c);
		if (rc != SRB_OK) return rc;
		srb_handle fbh = n_draws ? it.draws[0].framebuffer : 0;
		if (fbh)
		{
			rc = srb_clear(c, fbh, clear_color, 1, 1);
			if (rc != SRB_OK) return rc;
		}
		for (uint32_t i = 0; i < n_draws; ++i)
		{
			d[i] = it.draws[i];
			if (mvps) memcpy(d[i].mvp, mvps + (size_t(f) * n_draws + i) * 16, sizeof(float) * 16);
			rc = srb_draw_indexed(c, &d[i]);
			if (rc != SRB_OK) return rc;
		}
		if (colour_out && fbh)
		{
			// the read-back is part of the frame: Submit() issues it behind the shade kernel, and a frame that has to be re-run
			// after a capacity overflow (detected at this context's next end_frame or sync) issues it again
			FrameBufferDev* fb = GetFb(c, fbh);
			c->nextReadbackDst = (uint8_t*)colour_out + size_t(f) * colour_stride;
			c->nextReadbackBytes = size_t(fb->tilesX) * fb->tilesY * 16384u;
		}
		rc = srb_end_frame_async(c);
		if (rc != SRB_OK) return rc;
	}
	for (uint32_t i = 0; i < n_items && i < frames; ++i)
	{
		int rc = srb_sync(items[i].ctx);
		if (rc != SRB_OK) return rc;
	}
	return SRB_OK;
}

#ifdef SRB_STATS
/* statistics build only (make STATS=1): the device counters of srb_raster.cu */
SRB_API void srb_debug_stats(uint64_t* out16, int reset)
{
	static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "");
	stats_read(reinterpret_cast<unsigned long long*>(out16), reset != 0);
}
SRB_API void srb_debug_null_taps(int on) { stats_null_taps(on); }
#endif

/* Unit-test entry points for the sampler and the RCPPS replay (same device code as the tile kernel). */
SRB_API int srb_debug_sample(srb_context* c, srb_handle tex, const float* u, const float* v, const float* dudx,
                             const float* dudy, const float* dvdx, const float* dvdy, uint32_t* rgba, uint32_t n)
{
	if (!c || !tex || tex > c->res->textures.size() || !c->res->textures[tex - 1].alive || !n)
	{
		return Fail(c, SRB_ERR_INVALID, "bad sample arguments");
	}
	int rc = Bind(c);
	if (rc != SRB_OK) return rc;
	rc = UploadTexTable(c);
	if (rc != SRB_OK) return rc;
	float* d = nullptr;
	uint32_t* o = nullptr;
	SRB_CUDA(c, cudaMalloc((void**)&d, size_t(n) * 6 * sizeof(float)));
	SRB_CUDA(c, cudaMalloc((void**)&o, size_t(n) * sizeof(uint32_t)));
	const float* src[6] = {u, v, dudx, dudy, dvdx, dvdy};
	for (int i = 0; i < 6; ++i)
	{
		SRB_CUDA(c, cudaMemcpyAsync(d + size_t(i) * n, src[i], size_t(n) * sizeof(float), cudaMemcpyHostToDevice, c->stream));
	}
	launch_sample(c->res->dTexs, (uint32_t)tex - 1, d, d + n, d + 2 * size_t(n), d + 3 * size_t(n), d + 4 * size_t(n),
	              d + 5 * size_t(n), o, n, c->stream);
	c->launches++;
	cudaError_t e = cudaMemcpyAsync(rgba, o, size_t(n) * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	cudaFree(d);
	cudaFree(o);
	SRB_CUDA(c, e);
	return SRB_OK;
}

SRB_API int srb_debug_rcp(srb_context* c, const float* in, float* out, uint32_t n)
{
	if (!c || !in || !out || !n)
	{
		return Fail(c, SRB_ERR_INVALID, "bad rcp arguments");
	}
	int rc = Bind(c);
	if (rc != SRB_OK) return rc;
	float* d = nullptr;
	SRB_CUDA(c, cudaMalloc((void**)&d, size_t(n) * 2 * sizeof(float)));
	SRB_CUDA(c, cudaMemcpyAsync(d, in, size_t(n) * sizeof(float), cudaMemcpyHostToDevice, c->stream));
	launch_rcp(c->dRcp, c->rcpBits, d, d + n, n, c->stream);
	c->launches++;
	cudaError_t e = cudaMemcpyAsync(out, d + n, size_t(n) * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	cudaFree(d);
	SRB_CUDA(c, e);
	return SRB_OK;
}

SRB_API int srb_debug_rsqrt(srb_context* c, const float* in, float* out, uint32_t n)
{
	if (!c || !in || !out || !n)
	{
		return Fail(c, SRB_ERR_INVALID, "bad rsqrt arguments");
	}
	int rc = Bind(c);
	if (rc != SRB_OK) return rc;
	float* d = nullptr;
	SRB_CUDA(c, cudaMalloc((void**)&d, size_t(n) * 2 * sizeof(float)));
	SRB_CUDA(c, cudaMemcpyAsync(d, in, size_t(n) * sizeof(float), cudaMemcpyHostToDevice, c->stream));
	launch_rsqrt(c->dRsqrt, c->rsqrtBits, d, d + n, n, c->stream);
	c->launches++;
	cudaError_t e = cudaMemcpyAsync(out, d + n, size_t(n) * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	cudaFree(d);
	SRB_CUDA(c, e);
	return SRB_OK;
}

} // extern "C"
