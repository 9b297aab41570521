// srb_model.cpp — scene ingestion: the reference's sr::Obj::Model (Viewer/Obj.h:13-71) as host code of
// the C ABI: Wavefront OBJ + MTL parsing with the reference's exact mesh-building rules (Viewer/Obj.cpp:374-560), its
// `.bin` cache in kt::Serialize's byte format (Obj.cpp:15-39, kt/src/kt/inl/Serialization.inl:8-33,102-109,
// SoftRast/Texture.cpp:18-26), diffuse textures built like Tex::TextureData::CreateFromFile (Texture.cpp:103-199; PNG and
// TGA decoded here, other formats through a caller-supplied decoder), and the step the reference leaves to the scene
// code (Viewer/Scene.cpp:35-63): making the model resident on the device and filling one draw per mesh.
// Host compiler only, no CUDA; the resident part goes through the C ABI's own srb_buffer_create / srb_texture_create.
#include "../../include/softrast_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <zlib.h>

#include <exception>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace
{

thread_local std::string g_modelError;

int Fail(int rc, const char* fmt, const char* a = "", const char* b = "")
{
	char buf[1400];
	snprintf(buf, sizeof(buf), fmt, a, b);
	g_modelError = buf;
	return rc;
}

struct ObjVertex // sr::Obj::Vertex, Obj.h:16-21
{
	float pos[3], norm[3], uv[2];
};
static_assert(sizeof(ObjVertex) == 32, "Obj::Vertex is 32 bytes");

struct Mesh // sr::Obj::Mesh, Obj.h:24-43
{
	std::vector<uint8_t> indexData;
	uint32_t indexType = 0; // sr::IndexType (SoftRastTypes.h:15-19): 0 = u16, 1 = u32
	uint32_t numIndices = 0;
	std::vector<ObjVertex> vertexData;
	uint32_t matIdx = 0;
};

struct Texture // sr::Tex::TextureData, Texture.h:21-41
{
	std::vector<uint8_t> texels;
	uint32_t mipOffsets[SRB_MAX_TEX_DIM_LOG2] = {};
	uint32_t widthLog2 = 0, heightLog2 = 0, numMips = 0, bytesPerPixel = 0;
};

struct Material // sr::Obj::Material, Obj.h:45-53
{
	std::string name; // kt::String128: at most 127 characters
	Texture diffuse;
};

} // namespace

struct srb_model
{
	std::vector<Mesh> meshes;
	std::vector<Material> materials;
	bool fromCache = false;
};

struct srb_resident_model
{
	srb_context* ctx = nullptr;
	std::vector<srb_handle> vertexBufs, indexBufs, textures; // per mesh, per mesh, per material (0 = no texels)
	std::vector<uint32_t> numVerts, numIndices, indexStride, matIdx;
	uint32_t numMaterials = 0;
};

namespace
{

// ---- image decoding --------------------------------------------------------------------------------------------
// The reference decodes with stbi_load(file, &x, &y, &comp, 4) (Texture.cpp:107).  PNG and TGA are lossless, so a
// conforming decoder + stb's rules for expanding to 4 x 8 bits gives the same bytes: grey -> (g,g,g,255), grey+alpha ->
// (g,g,g,a), RGB -> (r,g,b,255), palette -> palette entry + tRNS alpha, 16 bits -> the high byte, tRNS colour key ->
// alpha 0, 1/2/4-bit grey scaled by 255/85/17 (stb_image.h: stbi__parse_png_file, stbi__convert_format,
// stbi__convert_16_to_8, stbi__compute_transparency, stbi__depth_scale_table).

bool ReadFile(const char* path, std::vector<uint8_t>& out)
{
	FILE* f = fopen(path, "rb");
	if (!f) return false;
	fseek(f, 0, SEEK_END);
	long const n = ftell(f);
	fseek(f, 0, SEEK_SET);
	out.resize(n > 0 ? size_t(n) : 0);
	size_t const got = out.empty() ? 0 : fread(out.data(), 1, out.size(), f);
	fclose(f);
	return got == out.size();
}

// no texture can be larger than 2^14 texels a side (Config::c_maxTexDimLog2); a header that claims more is rejected before
// anything is allocated for it
constexpr uint32_t kMaxImageDim = 1u << SRB_MAX_TEX_DIM_LOG2;

inline uint32_t Be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

inline int Paeth(int a, int b, int c)
{
	int const p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
	if (pa <= pb && pa <= pc) return a;
	return pb <= pc ? b : c;
}

// Undoes the per-scanline filters of one (sub-)image in place; `raw` holds h * (1 + rowBytes) bytes.
bool PngUnfilter(uint8_t* raw, uint32_t h, size_t rowBytes, uint32_t bpp /* bytes per complete pixel, >= 1 */)
{
	std::vector<uint8_t> zero(rowBytes, 0);
	const uint8_t* prev = zero.data();
	for (uint32_t y = 0; y < h; ++y)
	{
		uint8_t* line = raw + size_t(y) * (rowBytes + 1);
		uint8_t const filter = line[0];
		uint8_t* cur = line + 1;
		if (filter > 4) return false;
		for (size_t i = 0; i < rowBytes; ++i)
		{
			int const a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
			int add = 0;
			switch (filter)
			{
				case 1: add = a; break;
				case 2: add = b; break;
				case 3: add = (a + b) >> 1; break;
				case 4: add = Paeth(a, b, c); break;
				default: break;
			}
			cur[i] = uint8_t(cur[i] + add);
		}
		prev = cur;
	}
	return true;
}

struct PngInfo
{
	uint32_t w = 0, h = 0, depth = 0, colourType = 0, interlace = 0;
	uint8_t palette[256][4];
	uint32_t paletteLen = 0;
	bool hasKey = false;
	uint16_t key[3] = {0, 0, 0};
};

inline uint32_t PngChannels(uint32_t colourType)
{
	switch (colourType)
	{
		case 0: return 1;
		case 2: return 3;
		case 3: return 1;
		case 4: return 2;
		default: return 4;
	}
}

// One unfiltered scanline of `w` pixels -> RGBA8 written at dst + (x * xStep) * 4.
void PngExpandRow(const PngInfo& P, const uint8_t* row, uint32_t w, uint8_t* dst, uint32_t xStep)
{
	uint32_t const ch = PngChannels(P.colourType);
	for (uint32_t x = 0; x < w; ++x)
	{
		uint16_t s[4] = {0, 0, 0, 0};
		if (P.depth == 16)
		{
			for (uint32_t k = 0; k < ch; ++k) s[k] = uint16_t((row[(x * ch + k) * 2] << 8) | row[(x * ch + k) * 2 + 1]);
		}
		else if (P.depth == 8)
		{
			for (uint32_t k = 0; k < ch; ++k) s[k] = row[x * ch + k];
		}
		else // 1, 2, 4 bits: grey or palette, most significant bits first
		{
			uint32_t const bit = x * P.depth;
			s[0] = uint16_t((row[bit >> 3] >> (8 - P.depth - (bit & 7))) & ((1u << P.depth) - 1));
		}
		uint8_t* o = dst + size_t(x) * xStep * 4;
		if (P.colourType == 3)
		{
			uint32_t const i = s[0];
			if (i < P.paletteLen)
			{
				memcpy(o, P.palette[i], 4);
			}
			else // stb reads past the declared palette into its zero-initialised table
			{
				o[0] = o[1] = o[2] = 0;
				o[3] = 255;
			}
			continue;
		}
		bool keyed = false;
		if (P.hasKey && (P.colourType == 0 || P.colourType == 2))
		{
			keyed = true;
			for (uint32_t k = 0; k < ch; ++k) keyed = keyed && s[k] == P.key[k];
		}
		uint8_t v[4];
		for (uint32_t k = 0; k < ch; ++k)
		{
			if (P.depth == 16) v[k] = uint8_t(s[k] >> 8);
			else if (P.depth == 8) v[k] = uint8_t(s[k]);
			else v[k] = uint8_t(s[k] * (P.depth == 1 ? 255 : P.depth == 2 ? 85 : 17));
		}
		switch (P.colourType)
		{
			case 0: o[0] = o[1] = o[2] = v[0]; o[3] = keyed ? 0 : 255; break;
			case 2: o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = keyed ? 0 : 255; break;
			case 4: o[0] = o[1] = o[2] = v[0]; o[3] = v[1]; break;
			default: memcpy(o, v, 4); break;
		}
	}
}

bool DecodePng(const std::vector<uint8_t>& file, std::vector<uint8_t>& rgba, uint32_t& w, uint32_t& h)
{
	static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
	if (file.size() < 8 + 25 || memcmp(file.data(), sig, 8) != 0) return false;
	PngInfo P;
	memset(P.palette, 0, sizeof(P.palette));
	std::vector<uint8_t> idat;
	size_t pos = 8;
	bool haveHeader = false, done = false;
	while (!done && pos + 8 <= file.size())
	{
		uint32_t const len = Be32(&file[pos]);
		const uint8_t* type = &file[pos + 4];
		const uint8_t* data = &file[pos + 8];
		if (pos + 12 + size_t(len) > file.size() + 4) return false; // (the CRC of the last chunk may be cut off)
		if (pos + 8 + size_t(len) > file.size()) return false;
		if (!memcmp(type, "IHDR", 4))
		{
			if (len != 13) return false;
			P.w = Be32(data);
			P.h = Be32(data + 4);
			P.depth = data[8];
			P.colourType = data[9];
			P.interlace = data[12];
			if (!P.w || !P.h || P.w > kMaxImageDim || P.h > kMaxImageDim) return false;
			if (data[10] || data[11] || P.interlace > 1) return false;
			bool const depthOk = P.depth == 1 || P.depth == 2 || P.depth == 4 || P.depth == 8 || P.depth == 16;
			bool const typeOk = P.colourType == 0 || P.colourType == 2 || P.colourType == 3 || P.colourType == 4 || P.colourType == 6;
			if (!depthOk || !typeOk) return false;
			if (P.colourType == 3 && P.depth == 16) return false;
			if ((P.colourType == 2 || P.colourType == 4 || P.colourType == 6) && P.depth < 8) return false;
			haveHeader = true;
		}
		else if (!memcmp(type, "PLTE", 4))
		{
			if (len > 256 * 3 || len % 3) return false;
			P.paletteLen = len / 3;
			for (uint32_t i = 0; i < P.paletteLen; ++i)
			{
				P.palette[i][0] = data[i * 3];
				P.palette[i][1] = data[i * 3 + 1];
				P.palette[i][2] = data[i * 3 + 2];
				P.palette[i][3] = 255;
			}
		}
		else if (!memcmp(type, "tRNS", 4))
		{
			if (!haveHeader) return false;
			if (P.colourType == 3)
			{
				if (len > P.paletteLen) return false;
				for (uint32_t i = 0; i < len; ++i) P.palette[i][3] = data[i];
			}
			else if (P.colourType == 0 || P.colourType == 2)
			{
				uint32_t const ch = PngChannels(P.colourType);
				if (len != ch * 2) return false;
				P.hasKey = true;
				for (uint32_t k = 0; k < ch; ++k)
				{
					uint16_t const v = uint16_t((data[k * 2] << 8) | data[k * 2 + 1]);
					P.key[k] = P.depth == 16 ? v : uint16_t(v & 255);
				}
			}
			else
			{
				return false;
			}
		}
		else if (!memcmp(type, "IDAT", 4))
		{
			idat.insert(idat.end(), data, data + len);
		}
		else if (!memcmp(type, "IEND", 4))
		{
			done = true;
		}
		pos += 12 + size_t(len);
	}
	if (!haveHeader || idat.empty()) return false;
	if (P.colourType == 3 && !P.paletteLen) return false;

	uint32_t const bitsPerPixel = PngChannels(P.colourType) * P.depth;
	uint32_t const bpp = bitsPerPixel >= 8 ? bitsPerPixel / 8 : 1;
	auto rowBytesOf = [&](uint32_t pw) { return (size_t(pw) * bitsPerPixel + 7) / 8; };

	// sub-images: the whole picture, or the seven Adam7 passes
	static const uint32_t xOrig[7] = {0, 4, 0, 2, 0, 1, 0}, yOrig[7] = {0, 0, 4, 0, 2, 0, 1};
	static const uint32_t xSpc[7] = {8, 8, 4, 4, 2, 2, 1}, ySpc[7] = {8, 8, 8, 4, 4, 2, 2};
	uint32_t const passes = P.interlace ? 7 : 1;
	size_t total = 0;
	uint32_t pw[7], ph[7];
	for (uint32_t p = 0; p < passes; ++p)
	{
		pw[p] = P.interlace ? (P.w - xOrig[p] + xSpc[p] - 1) / xSpc[p] : P.w;
		ph[p] = P.interlace ? (P.h - yOrig[p] + ySpc[p] - 1) / ySpc[p] : P.h;
		if (pw[p] && ph[p]) total += (rowBytesOf(pw[p]) + 1) * ph[p];
	}
	std::vector<uint8_t> raw(total);
	z_stream zs;
	memset(&zs, 0, sizeof(zs));
	if (inflateInit(&zs) != Z_OK) return false;
	zs.next_in = idat.data();
	zs.avail_in = uInt(idat.size());
	zs.next_out = raw.data();
	zs.avail_out = uInt(raw.size());
	int const zr = inflate(&zs, Z_FINISH);
	size_t const produced = raw.size() - zs.avail_out;
	inflateEnd(&zs);
	if ((zr != Z_STREAM_END && zr != Z_OK && zr != Z_BUF_ERROR) || produced != raw.size()) return false;

	w = P.w;
	h = P.h;
	rgba.assign(size_t(w) * h * 4, 0);
	size_t offs = 0;
	for (uint32_t p = 0; p < passes; ++p)
	{
		if (!pw[p] || !ph[p]) continue;
		size_t const rb = rowBytesOf(pw[p]);
		if (!PngUnfilter(&raw[offs], ph[p], rb, bpp)) return false;
		for (uint32_t y = 0; y < ph[p]; ++y)
		{
			const uint8_t* row = &raw[offs + size_t(y) * (rb + 1) + 1];
			if (P.interlace)
			{
				uint32_t const oy = yOrig[p] + y * ySpc[p];
				PngExpandRow(P, row, pw[p], &rgba[(size_t(oy) * w + xOrig[p]) * 4], xSpc[p]);
			}
			else
			{
				PngExpandRow(P, row, pw[p], &rgba[size_t(y) * w * 4], 1);
			}
		}
		offs += (rb + 1) * ph[p];
	}
	return true;
}

// Truevision TGA: true-colour 24/32 bits and 8-bit grey, raw or run-length encoded, either vertical origin
// (stb_image.h: stbi__tga_load).  Colour-mapped and 15/16-bit files are not handled (decoder hook).
bool DecodeTga(const std::vector<uint8_t>& file, std::vector<uint8_t>& rgba, uint32_t& w, uint32_t& h)
{
	if (file.size() < 18) return false;
	const uint8_t* p = file.data();
	uint32_t const idLen = p[0], cmapType = p[1], imgType = p[2];
	uint32_t const cmapLen = p[5] | (p[6] << 8), cmapBits = p[7];
	w = p[12] | (p[13] << 8);
	h = p[14] | (p[15] << 8);
	uint32_t const bits = p[16], desc = p[17];
	bool const rle = imgType >= 8;
	uint32_t const kind = imgType & 7; // 2 = true colour, 3 = grey
	if (cmapType != 0 || (kind != 2 && kind != 3) || !w || !h || w > kMaxImageDim || h > kMaxImageDim) return false;
	if (!((kind == 2 && (bits == 24 || bits == 32)) || (kind == 3 && bits == 8))) return false;
	uint32_t const bytes = bits / 8;
	size_t pos = 18 + size_t(idLen) + size_t(cmapLen) * ((cmapBits + 7) / 8);
	rgba.assign(size_t(w) * h * 4, 0);
	bool const topDown = (desc >> 5) & 1; // stb: inverted = 1 - bit 5; inverted rows are flipped
	uint8_t px[4] = {0, 0, 0, 255};
	uint32_t run = 0;
	bool runIsRle = false, havePx = false;
	for (size_t i = 0; i < size_t(w) * h; ++i)
	{
		bool read = true;
		if (rle)
		{
			if (run == 0)
			{
				if (pos >= file.size()) return false;
				uint8_t const c = file[pos++];
				run = 1 + (c & 127);
				runIsRle = (c >> 7) != 0;
				havePx = false;
			}
			read = !runIsRle || !havePx;
			--run;
		}
		if (read)
		{
			if (pos + bytes > file.size()) return false;
			if (kind == 3)
			{
				px[0] = px[1] = px[2] = file[pos];
				px[3] = 255;
			}
			else
			{
				px[0] = file[pos + 2]; // BGR(A) on disk
				px[1] = file[pos + 1];
				px[2] = file[pos];
				px[3] = bytes == 4 ? file[pos + 3] : 255;
			}
			pos += bytes;
			havePx = true;
		}
		size_t const y = i / w, x = i % w;
		size_t const oy = topDown ? y : (h - 1 - y);
		memcpy(&rgba[(oy * w + x) * 4], px, 4);
	}
	return true;
}

bool EndsWithNoCase(const char* s, const char* suffix)
{
	size_t const n = strlen(s), m = strlen(suffix);
	return n >= m && strcasecmp(s + n - m, suffix) == 0;
}

int LoadImageRgba8(const char* path, srb_image_decoder decoder, void* user, std::vector<uint8_t>& rgba, uint32_t& w, uint32_t& h)
{
	std::vector<uint8_t> file;
	if (!ReadFile(path, file)) return Fail(SRB_ERR_INVALID, "cannot read image file %s", path);
	if (DecodePng(file, rgba, w, h)) return SRB_OK;
	if (EndsWithNoCase(path, ".tga") && DecodeTga(file, rgba, w, h)) return SRB_OK;
	if (decoder)
	{
		uint8_t* px = nullptr;
		if (decoder(path, &px, &w, &h, user) == 0 && px)
		{
			rgba.assign(px, px + size_t(w) * h * 4);
			free(px);
			return SRB_OK;
		}
	}
	return Fail(SRB_ERR_INVALID, "unsupported image format (built in: PNG, TGA 24/32-bit/grey; pass a decoder for others): %s", path);
}

inline bool IsPow2(uint32_t v) { return v && !(v & (v - 1)); }

// Tex::TextureData::CreateFromFile (Texture.cpp:103-117) + CreateFromRGBA8(..., _calcMips = true) (:119-199).  The
// reference asserts power-of-two sizes that are multiples of its 32-texel tile; such an image is an error here.
int BuildDiffuse(const char* path, srb_image_decoder decoder, void* user, srb_context* ctx, Texture& t)
{
	std::vector<uint8_t> rgba;
	uint32_t w = 0, h = 0;
	int rc = LoadImageRgba8(path, decoder, user, rgba, w, h);
	if (rc != SRB_OK) return rc;
	if (!IsPow2(w) || !IsPow2(h) || (w % 32) || (h % 32) || w >= (1u << SRB_MAX_TEX_DIM_LOG2) || h >= (1u << SRB_MAX_TEX_DIM_LOG2))
	{
		return Fail(SRB_ERR_INVALID, "texture size must be a power of two >= 32 (Texture.cpp:122-129): %s", path);
	}
	uint64_t bytes = 0;
	if (ctx)
	{
		// a device is at hand: tiling and the stb mip chain run as CUDA kernels (srb_texture_create_rgba8: the same bytes,
		// ~3 ms instead of ~0.25-0.5 s per 1024^2 image) and the blob comes back for m_texels / the .bin cache
		srb_handle tex = 0;
		rc = srb_texture_create_rgba8(ctx, rgba.data(), w, h, SRB_MIPS_STB, &tex);
		if (rc != SRB_OK) return Fail(rc, "srb_texture_create_rgba8 failed for %s: %s", path, srb_last_error(ctx));
		rc = srb_texture_read(ctx, tex, nullptr, 0, &bytes, t.mipOffsets, &t.numMips, &t.widthLog2, &t.heightLog2);
		if (rc == SRB_OK)
		{
			t.texels.resize(bytes);
			rc = srb_texture_read(ctx, tex, t.texels.data(), bytes, nullptr, nullptr, nullptr, nullptr, nullptr);
		}
		srb_texture_destroy(ctx, tex);
		if (rc != SRB_OK) return Fail(rc, "srb_texture_read failed for %s: %s", path, srb_last_error(ctx));
		t.bytesPerPixel = 4;
		return SRB_OK;
	}
	rc = srb_texture_build_rgba8(nullptr, w, h, SRB_MIPS_STB, nullptr, &bytes, t.mipOffsets, &t.numMips);
	if (rc != SRB_OK) return Fail(rc, "srb_texture_build_rgba8 failed for %s", path);
	t.texels.resize(bytes);
	rc = srb_texture_build_rgba8(rgba.data(), w, h, SRB_MIPS_STB, t.texels.data(), &bytes, t.mipOffsets, &t.numMips);
	if (rc != SRB_OK) return Fail(rc, "srb_texture_build_rgba8 failed for %s", path);
	t.widthLog2 = t.heightLog2 = 0;
	while ((1u << t.widthLog2) < w) ++t.widthLog2;
	while ((1u << t.heightLog2) < h) ++t.heightLog2;
	t.bytesPerPixel = 4;
	return SRB_OK;
}

// ---- OBJ / MTL text ---------------------------------------------------------------------------------------------

// Obj.cpp:132-151: leading blanks/tabs skipped; trailing blanks, tabs, CR, LF cut — but never the first character.
char* StripLine(char* buff)
{
	char* ret = buff;
	while (*ret == ' ' || *ret == '\t') ++ret;
	size_t const len = strlen(ret);
	if (len)
	{
		char* t = ret + (len - 1);
		while (t != ret && (*t == ' ' || *t == '\t' || *t == '\r' || *t == '\n')) *t-- = '\0';
	}
	return ret;
}

// Directory part of a path including its last separator (kt::FilePath::GetPath, FilePath.cpp:176-186).  The reference
// then joins with kt::FilePath::Append, which for a path without a directory yields "/name" (a rooted path); here a bare
// file name resolves next to the OBJ, i.e. in the working directory.
std::string DirOf(const char* path)
{
	std::string s(path);
	size_t const k = s.find_last_of("/\\");
	return k == std::string::npos ? std::string() : s.substr(0, k + 1);
}

std::string JoinPath(const std::string& dir, const char* name)
{
	std::string s = dir + name;
	for (char& c : s)
	{
		if (c == '\\') c = '/';
	}
	return s;
}

struct FaceKey // Obj.cpp:49-59 TempFace
{
	uint32_t pos, uv, norm;
	bool operator==(const FaceKey& o) const { return pos == o.pos && uv == o.uv && norm == o.norm; }
};
struct FaceKeyHash
{
	size_t operator()(const FaceKey& k) const
	{
		uint64_t h = 1469598103934665603ull;
		for (uint32_t v : {k.pos, k.uv, k.norm}) h = (h ^ v) * 1099511628211ull;
		return size_t(h ^ (h >> 29));
	}
};

struct Parser // Obj.cpp:71-130 MeshParserState
{
	std::unordered_map<FaceKey, uint32_t, FaceKeyHash> faceMap;
	std::vector<ObjVertex> verts;
	std::vector<uint32_t> indices;
	std::vector<float> pos, uv, norm; // 3 / 2 / 3 floats per entry; NOT reset between meshes (indices are file-global)

	void Finalize(Mesh& m, uint32_t matIdx) // Obj.cpp:90-127
	{
		m.matIdx = matIdx;
		if (verts.empty()) return;
		m.indexType = verts.size() > 0xFFFFu ? 1u : 0u;
		if (m.indexType)
		{
			m.indexData.resize(indices.size() * 4);
			memcpy(m.indexData.data(), indices.data(), m.indexData.size());
		}
		else
		{
			m.indexData.resize(indices.size() * 2);
			uint16_t* d = reinterpret_cast<uint16_t*>(m.indexData.data());
			for (uint32_t i : indices) *d++ = uint16_t(i);
		}
		m.numIndices = uint32_t(indices.size());
		m.vertexData = verts;
		faceMap.clear();
		verts.clear();
		indices.clear();
	}
};

inline int32_t FixupIndex(int32_t idx, int32_t total) // Obj.cpp:153-158
{
	if (!idx) return 0;
	return idx < 0 ? total + idx : idx - 1;
}

// ---- "f" lines ------------------------------------------------------------------------------------------------------
// What the reference's face parser accepts (Viewer/Obj.cpp:160-312) decides the vertex ORDER of a mesh, so it is
// reproduced rule for rule — as a small tokenizer, a corner reader and an emitter, not as the reference writes it:
//   * at most four corners per face; a quad becomes the triangles (0, 1, 2) and (0, 2, 3), longer polygons are cut off;
//   * a corner is "p", "p/t", "p//n" or "p/t/n"; indices are 1-based or negative (relative to the end); a missing index
//     reads as entry 0; a corner whose position index is 0 (or absent) ends the face;
//   * a face with one or two corners still appends that many indices;
//   * vertices are de-duplicated per mesh by their (p, t, n) triple, in first-use order.
struct ObjCorner
{
	int32_t p = 0, t = 0, n = 0; // as written in the file (0 = absent)
};

// An optionally signed decimal number after blanks; no digits reads as 0.  The magnitude wraps modulo 2^32 like the
// reference's int32 accumulation does on this ABI.
int32_t ReadObjInt(const char*& at)
{
	while (*at == ' ' || *at == '\t') ++at;
	bool const negative = *at == '-';
	if (negative) ++at;
	uint32_t magnitude = 0;
	for (; *at >= '0' && *at <= '9'; ++at) magnitude = magnitude * 10u + uint32_t(*at - '0');
	return negative ? -int32_t(magnitude) : int32_t(magnitude);
}

// Reads the corners of one face; returns how many there are (0 .. 4).
int ReadObjCorners(const char* at, ObjCorner (&out)[4])
{
	int count = 0;
	while (*at != '\0' && count < 4)
	{
		ObjCorner c;
		c.p = ReadObjInt(at);
		if (c.p == 0) break;
		if (*at == '/')
		{
			++at;
			if (*at == '/') // "p//n"
			{
				++at;
				c.n = ReadObjInt(at);
			}
			else // "p/t" or "p/t/n"
			{
				c.t = ReadObjInt(at);
				if (*at == '/')
				{
					++at;
					c.n = ReadObjInt(at);
				}
			}
		}
		out[count++] = c;
	}
	return count;
}

// Appends the index of the corner's vertex, creating the vertex at its first use.  false = an index out of range.
bool EmitObjCorner(Parser& S, const ObjCorner& c)
{
	uint32_t const nPos = uint32_t(S.pos.size() / 3), nUv = uint32_t(S.uv.size() / 2), nNorm = uint32_t(S.norm.size() / 3);
	FaceKey k;
	k.pos = uint32_t(FixupIndex(c.p, int32_t(nPos)));
	k.uv = uint32_t(FixupIndex(c.t, int32_t(nUv)));
	k.norm = uint32_t(FixupIndex(c.n, int32_t(nNorm)));
	auto const found = S.faceMap.find(k);
	if (found != S.faceMap.end())
	{
		S.indices.push_back(found->second);
		return true;
	}
	uint32_t const fresh = uint32_t(S.verts.size());
	S.indices.push_back(fresh);
	S.faceMap.emplace(k, fresh);
	if (k.pos >= nPos || (nNorm && k.norm >= nNorm) || (nUv && k.uv >= nUv))
	{
		return false;
	}
	ObjVertex v;
	memset(&v, 0, sizeof(v));
	memcpy(v.pos, &S.pos[size_t(k.pos) * 3], sizeof(v.pos));
	if (nNorm) memcpy(v.norm, &S.norm[size_t(k.norm) * 3], sizeof(v.norm));
	if (nUv) memcpy(v.uv, &S.uv[size_t(k.uv) * 2], sizeof(v.uv));
	S.verts.push_back(v);
	return true;
}

bool ParseFace(Parser& S, const char* line)
{
	ObjCorner corner[4];
	int const n = ReadObjCorners(line + 2, corner);
	static const int kQuadOrder[6] = {0, 1, 2, 0, 2, 3};
	int const emit = n == 4 ? 6 : n;
	for (int i = 0; i < emit; ++i)
	{
		if (!EmitObjCorner(S, corner[n == 4 ? kQuadOrder[i] : i]))
		{
			return false;
		}
	}
	return true;
}

// Obj.cpp:314-356.  "newmtl <name>" opens a material, "map_Kd <file>" (prefix match, like the reference) loads its
// diffuse texture relative to the OBJ's directory.  A texture that cannot be built leaves the material without texels,
// as Texture.cpp:108-112 does (the draw then shades white, Shaders.h:77-82); the reason is kept in the error text.
void ParseMaterials(FILE* f, srb_model& M, const std::string& root, srb_image_decoder decoder, void* user, srb_context* ctx,
                    std::string& warnings)
{
	char buff[2048];
	Material* cur = nullptr;
	size_t curIdx = 0;
	while (fgets(buff, sizeof(buff), f))
	{
		char* line = StripLine(buff);
		if (*line == 'm' && strncmp(line, "map_Kd", 6) == 0)
		{
			if (!cur) continue; // "No newmtl directive, can't parse mtl!"
			char* name = StripLine(line + 6);
			std::string const path = JoinPath(root, name);
			M.materials[curIdx].diffuse = Texture(); // CreateFromFile starts with Clear()
			if (BuildDiffuse(path.c_str(), decoder, user, ctx, M.materials[curIdx].diffuse) != SRB_OK)
			{
				M.materials[curIdx].diffuse = Texture();
				warnings += g_modelError + "\n";
			}
		}
		else if (*line == 'n' && strncmp(line, "newmtl", 6) == 0)
		{
			M.materials.emplace_back();
			curIdx = M.materials.size() - 1;
			cur = &M.materials[curIdx];
			cur->name = StripLine(line + 6);
			if (cur->name.size() > 127) cur->name.resize(127); // kt::String128
		}
	}
}

// ---- the .bin cache (kt::Serialize) -------------------------------------------------------------------------------
// Model  = Array<Mesh>, Array<Material>                                   (Obj.cpp:15-20)
// Mesh   = u32 indexType, Array<u8> indexData, u32 numIndices, Array<Vertex 32 B>, u32 matIdx   (Obj.cpp:22-30)
// Material = TextureData, String128 name (u32 length + characters)        (Obj.cpp:32-37, Serialization.inl:102-109)
// TextureData = Array<u8> texels, u32 widthLog2, u32 heightLog2, u32 bytesPerPixel, u32 mipOffsets[14], u32 numMips
//                                                                         (Texture.cpp:18-26)
// Array<T> = u32 count followed by the elements                           (Serialization.inl:8-40)
struct Writer
{
	std::vector<uint8_t> out;
	void Bytes(const void* p, size_t n)
	{
		const uint8_t* b = static_cast<const uint8_t*>(p);
		out.insert(out.end(), b, b + n);
	}
	void U32(uint32_t v) { Bytes(&v, 4); }
};

struct Reader
{
	const uint8_t* p;
	size_t left;
	bool ok = true;
	bool Bytes(void* dst, size_t n)
	{
		if (n > left)
		{
			ok = false;
			return false;
		}
		if (n) memcpy(dst, p, n);
		p += n;
		left -= n;
		return true;
	}
	uint32_t U32()
	{
		uint32_t v = 0;
		Bytes(&v, 4);
		return v;
	}
};

void SerializeModel(const srb_model& M, Writer& W)
{
	W.U32(uint32_t(M.meshes.size()));
	for (const Mesh& m : M.meshes)
	{
		W.U32(m.indexType);
		W.U32(uint32_t(m.indexData.size()));
		W.Bytes(m.indexData.data(), m.indexData.size());
		W.U32(m.numIndices);
		W.U32(uint32_t(m.vertexData.size()));
		W.Bytes(m.vertexData.data(), m.vertexData.size() * sizeof(ObjVertex));
		W.U32(m.matIdx);
	}
	W.U32(uint32_t(M.materials.size()));
	for (const Material& mat : M.materials)
	{
		const Texture& t = mat.diffuse;
		W.U32(uint32_t(t.texels.size()));
		W.Bytes(t.texels.data(), t.texels.size());
		W.U32(t.widthLog2);
		W.U32(t.heightLog2);
		W.U32(t.bytesPerPixel);
		W.Bytes(t.mipOffsets, sizeof(t.mipOffsets));
		W.U32(t.numMips);
		W.U32(uint32_t(mat.name.size()));
		W.Bytes(mat.name.data(), mat.name.size());
	}
}

bool DeserializeModel(srb_model& M, Reader& R)
{
	uint32_t const numMeshes = R.U32();
	if (!R.ok || size_t(numMeshes) * 20 > R.left) return false;
	M.meshes.resize(numMeshes);
	for (Mesh& m : M.meshes)
	{
		m.indexType = R.U32();
		uint32_t const idxBytes = R.U32();
		if (!R.ok || idxBytes > R.left) return false;
		m.indexData.resize(idxBytes);
		R.Bytes(m.indexData.data(), idxBytes);
		m.numIndices = R.U32();
		uint32_t const nv = R.U32();
		if (!R.ok || size_t(nv) * sizeof(ObjVertex) > R.left) return false;
		m.vertexData.resize(nv);
		R.Bytes(m.vertexData.data(), size_t(nv) * sizeof(ObjVertex));
		m.matIdx = R.U32();
		if (!R.ok || m.indexType > 1u || size_t(m.numIndices) * (m.indexType ? 4 : 2) > m.indexData.size()) return false;
		// every index must address a vertex of this mesh: a corrupt or foreign cache would otherwise make a later draw of
		// the resident model read out of bounds on the device (the caller falls back to re-parsing the OBJ)
		for (uint32_t i = 0; i < m.numIndices; ++i)
		{
			uint32_t const idx = m.indexType ? reinterpret_cast<const uint32_t*>(m.indexData.data())[i]
			                                 : reinterpret_cast<const uint16_t*>(m.indexData.data())[i];
			if (idx >= nv) return false;
		}
	}
	uint32_t const numMats = R.U32();
	if (!R.ok || size_t(numMats) * 80 > R.left) return false;
	M.materials.resize(numMats);
	for (Material& mat : M.materials)
	{
		Texture& t = mat.diffuse;
		uint32_t const tb = R.U32();
		if (!R.ok || tb > R.left) return false;
		t.texels.resize(tb);
		R.Bytes(t.texels.data(), tb);
		t.widthLog2 = R.U32();
		t.heightLog2 = R.U32();
		t.bytesPerPixel = R.U32();
		R.Bytes(t.mipOffsets, sizeof(t.mipOffsets));
		t.numMips = R.U32();
		uint32_t const len = R.U32();
		if (!R.ok || len > R.left || len > 127) return false;
		mat.name.resize(len);
		R.Bytes(&mat.name[0], len);
		if (!R.ok) return false;
	}
	return R.ok;
}

bool FileExists(const char* path)
{
	struct stat st;
	return stat(path, &st) == 0 && S_ISREG(st.st_mode);
}

// Model::Load's text path, Obj.cpp:399-560.
int ParseObj(const char* path, uint32_t flags, srb_image_decoder decoder, void* user, srb_context* ctx, srb_model& M,
             std::string& warnings)
{
	FILE* f = fopen(path, "r");
	if (!f) return Fail(SRB_ERR_INVALID, "Failed to open obj file: %s", path);
	char buff[2048];
	Parser S;
	std::string const root = DirOf(path);
	uint32_t curMat = 0;
	int rc = SRB_OK;
	while (rc == SRB_OK && fgets(buff, sizeof(buff), f))
	{
		char* line = StripLine(buff);
		switch (line[0])
		{
			case 'v':
			{
				float v[3];
				if (line[1] == ' ')
				{
					if (sscanf(line + 2, "%f %f %f", v, v + 1, v + 2) != 3) rc = Fail(SRB_ERR_INVALID, "Failed to parse obj, Bad vertex pos! (%s)", path);
					else S.pos.insert(S.pos.end(), v, v + 3);
				}
				else if (line[1] == 't')
				{
					if (sscanf(line + 3, "%f %f", v, v + 1) != 2)
					{
						rc = Fail(SRB_ERR_INVALID, "Failed to parse obj, Bad uv coord! (%s)", path);
						break;
					}
					if (flags & SRB_OBJ_FLIP_UVS) v[1] = 1.0f - v[1];
					S.uv.insert(S.uv.end(), v, v + 2);
				}
				else if (line[1] == 'n')
				{
					if (sscanf(line + 3, "%f %f %f", v, v + 1, v + 2) != 3) rc = Fail(SRB_ERR_INVALID, "Failed to parse obj, Bad vertex normal! (%s)", path);
					else S.norm.insert(S.norm.end(), v, v + 3);
				}
			}
			break;
			case 'f':
			{
				if (!ParseFace(S, line)) rc = Fail(SRB_ERR_INVALID, "Failed to parse obj, Bad vertex face! (%s)", path);
			}
			break;
			case 'g':
			{
				if (!S.verts.empty())
				{
					M.meshes.emplace_back();
					S.Finalize(M.meshes.back(), curMat);
				}
			}
			break;
			case 'u':
			{
				if (strncmp(line, "usemtl", 6) == 0)
				{
					char* mtl = StripLine(line + 6);
					for (uint32_t i = 0; i < M.materials.size(); ++i)
					{
						if (M.materials[i].name == mtl)
						{
							curMat = i;
							break;
						}
					}
				}
			}
			break;
			case 'm':
			{
				if (strncmp("mtllib", line, 6) == 0)
				{
					char* name = line + 6;
					while (*name == ' ' || *name == '\t') ++name;
					if (!*name) break; // "Invalid material name in obj file"
					std::string const mtlPath = JoinPath(root, name);
					FILE* mf = fopen(mtlPath.c_str(), "r");
					if (!mf)
					{
						// the reference logs this and then reads from the null FILE* (Obj.cpp:518-524); here the
						// model simply has no materials from that library
						warnings += "Failed to open material file: " + mtlPath + "\n";
						break;
					}
					ParseMaterials(mf, M, root, decoder, user, ctx, warnings);
					fclose(mf);
				}
			}
			break;
			default: break;
		}
	}
	fclose(f);
	if (rc != SRB_OK) return rc;
	if (!S.verts.empty())
	{
		M.meshes.emplace_back();
		S.Finalize(M.meshes.back(), curMat);
	}
	if (flags & SRB_OBJ_FLIP_WINDING) // Obj.cpp:61-69, 533-546
	{
		for (Mesh& m : M.meshes)
		{
			if (m.indexType)
			{
				uint32_t* b = reinterpret_cast<uint32_t*>(m.indexData.data());
				for (uint32_t i = 0; i + 2 < m.numIndices; i += 3) std::swap(b[i + 1], b[i + 2]);
			}
			else
			{
				uint16_t* b = reinterpret_cast<uint16_t*>(m.indexData.data());
				for (uint32_t i = 0; i + 2 < m.numIndices; i += 3) std::swap(b[i + 1], b[i + 2]);
			}
		}
	}
	return SRB_OK;
}

} // namespace

// =====================================================================================================================

SRB_API const char* srb_model_last_error(void) { return g_modelError.c_str(); }

SRB_API int srb_image_load_rgba8(const char* path, uint8_t** rgba_out, uint32_t* width, uint32_t* height)
{
	if (!path || !rgba_out || !width || !height) return Fail(SRB_ERR_INVALID, "srb_image_load_rgba8: null argument");
	try
	{
		std::vector<uint8_t> px;
		int const rc = LoadImageRgba8(path, nullptr, nullptr, px, *width, *height);
		if (rc != SRB_OK) return rc;
		*rgba_out = static_cast<uint8_t*>(malloc(px.size()));
		if (!*rgba_out) return Fail(SRB_ERR_INVALID, "out of memory");
		memcpy(*rgba_out, px.data(), px.size());
		return SRB_OK;
	}
	catch (const std::exception& e) // no C++ exception crosses the C ABI
	{
		return Fail(SRB_ERR_INVALID, "srb_image_load_rgba8(%s): %s", path, e.what());
	}
}

SRB_API void srb_image_free(uint8_t* rgba) { free(rgba); }

static int ModelLoadImpl(const char* path, uint32_t flags, srb_image_decoder decoder, void* user, srb_context* ctx, srb_model** out);

SRB_API int srb_model_load_ex(const char* path, uint32_t flags, srb_image_decoder decoder, void* user, srb_model** out)
{
	return srb_model_load_on(nullptr, path, flags, decoder, user, out);
}

SRB_API int srb_model_load_on(srb_context* ctx, const char* path, uint32_t flags, srb_image_decoder decoder, void* user,
                              srb_model** out)
{
	if (!path || !out) return Fail(SRB_ERR_INVALID, "srb_model_load: null argument");
	*out = nullptr;
	try
	{
		return ModelLoadImpl(path, flags, decoder, user, ctx, out);
	}
	catch (const std::exception& e) // no C++ exception crosses the C ABI
	{
		return Fail(SRB_ERR_INVALID, "srb_model_load(%s): %s", path, e.what());
	}
}

static int ModelLoadImpl(const char* path, uint32_t flags, srb_image_decoder decoder, void* user, srb_context* ctx, srb_model** out)
{
	g_modelError.clear();
	std::string const binPath = std::string(path) + ".bin";
	std::unique_ptr<srb_model> M(new srb_model()); // (freed if anything below throws)
	if (!(flags & SRB_OBJ_NO_CACHE_READ) && FileExists(binPath.c_str())) // Obj.cpp:376-397
	{
		std::vector<uint8_t> file;
		if (ReadFile(binPath.c_str(), file))
		{
			Reader R{file.data(), file.size()};
			if (DeserializeModel(*M, R))
			{
				M->fromCache = true;
				*out = M.release();
				return SRB_OK;
			}
			// the reference has no error checking here ("todo", Obj.cpp:387); a truncated cache is re-parsed instead
			*M = srb_model();
		}
	}
	std::string warnings;
	int const rc = ParseObj(path, flags, decoder, user, ctx, *M, warnings);
	if (rc != SRB_OK)
	{
		return rc;
	}
	if (!(flags & SRB_OBJ_NO_CACHE_WRITE)) // Obj.cpp:548-560
	{
		Writer W;
		SerializeModel(*M, W);
		FILE* cf = fopen(binPath.c_str(), "wb");
		if (cf)
		{
			fwrite(W.out.data(), 1, W.out.size(), cf);
			fclose(cf);
		}
		else
		{
			warnings += "Failed to write obj cache file " + binPath + "\n";
		}
	}
	g_modelError = warnings; // non-fatal notes (missing MTL / texture); empty when everything loaded
	*out = M.release();
	return SRB_OK;
}

SRB_API int srb_model_load(const char* path, uint32_t flags, srb_model** out)
{
	return srb_model_load_ex(path, flags, nullptr, nullptr, out);
}

SRB_API void srb_model_free(srb_model* model) { delete model; }

SRB_API int srb_model_info(const srb_model* model, uint32_t* num_meshes, uint32_t* num_materials, int* from_cache)
{
	if (!model) return Fail(SRB_ERR_INVALID, "srb_model_info: null model");
	if (num_meshes) *num_meshes = uint32_t(model->meshes.size());
	if (num_materials) *num_materials = uint32_t(model->materials.size());
	if (from_cache) *from_cache = model->fromCache ? 1 : 0;
	return SRB_OK;
}

SRB_API int srb_model_mesh(const srb_model* model, uint32_t index, srb_mesh_view* out)
{
	if (!model || !out || index >= model->meshes.size()) return Fail(SRB_ERR_INVALID, "srb_model_mesh: bad argument");
	const Mesh& m = model->meshes[index];
	out->indices = m.indexData.data();
	out->index_stride = m.indexType ? 4u : 2u;
	out->num_indices = m.numIndices;
	out->vertices = m.vertexData.data();
	out->num_vertices = uint32_t(m.vertexData.size());
	out->material = m.matIdx;
	return SRB_OK;
}

SRB_API int srb_model_material(const srb_model* model, uint32_t index, srb_material_view* out)
{
	if (!model || !out || index >= model->materials.size()) return Fail(SRB_ERR_INVALID, "srb_model_material: bad argument");
	const Material& mat = model->materials[index];
	out->name = mat.name.c_str();
	out->texels = mat.diffuse.texels.data();
	out->texel_bytes = mat.diffuse.texels.size();
	memcpy(out->mip_offsets, mat.diffuse.mipOffsets, sizeof(out->mip_offsets));
	out->num_mips = mat.diffuse.numMips;
	out->width_log2 = mat.diffuse.widthLog2;
	out->height_log2 = mat.diffuse.heightLog2;
	out->bytes_per_pixel = mat.diffuse.bytesPerPixel;
	return SRB_OK;
}

SRB_API int srb_model_save_cache(const srb_model* model, const char* bin_path)
{
	if (!model || !bin_path) return Fail(SRB_ERR_INVALID, "srb_model_save_cache: null argument");
	Writer W;
	SerializeModel(*model, W);
	FILE* cf = fopen(bin_path, "wb");
	if (!cf) return Fail(SRB_ERR_INVALID, "Failed to write obj cache file %s.", bin_path);
	size_t const n = fwrite(W.out.data(), 1, W.out.size(), cf);
	fclose(cf);
	return n == W.out.size() ? SRB_OK : Fail(SRB_ERR_INVALID, "short write to %s", bin_path);
}

// ---- resident model: what Viewer/Scene.cpp:35-63 does per frame with host pointers, done once with device buffers ------

SRB_API void srb_resident_model_free(srb_resident_model* rm)
{
	if (!rm) return;
	for (srb_handle h : rm->vertexBufs) if (h) srb_buffer_destroy(rm->ctx, h);
	for (srb_handle h : rm->indexBufs) if (h) srb_buffer_destroy(rm->ctx, h);
	for (srb_handle h : rm->textures) if (h) srb_texture_destroy(rm->ctx, h);
	delete rm;
}

SRB_API int srb_model_make_resident(srb_context* ctx, const srb_model* model, srb_resident_model** out)
{
	if (!ctx || !model || !out) return Fail(SRB_ERR_INVALID, "srb_model_make_resident: null argument");
	*out = nullptr;
	srb_resident_model* rm = new srb_resident_model();
	rm->ctx = ctx;
	rm->numMaterials = uint32_t(model->materials.size());
	int rc = SRB_OK;
	for (const Material& mat : model->materials)
	{
		srb_handle h = 0;
		const Texture& t = mat.diffuse;
		if (!t.texels.empty() && t.numMips) // no texels: null texture, the shader returns white (Shaders.h:77-82)
		{
			rc = srb_texture_create(ctx, t.texels.data(), t.texels.size(), t.mipOffsets, t.numMips, t.widthLog2, t.heightLog2, &h);
			if (rc != SRB_OK) break;
		}
		rm->textures.push_back(h);
	}
	for (size_t i = 0; rc == SRB_OK && i < model->meshes.size(); ++i)
	{
		const Mesh& m = model->meshes[i];
		srb_handle vb = 0, ib = 0;
		if (!m.vertexData.empty() && m.numIndices)
		{
			rc = srb_buffer_create(ctx, m.vertexData.data(), m.vertexData.size() * sizeof(ObjVertex), &vb);
			if (rc == SRB_OK) rc = srb_buffer_create(ctx, m.indexData.data(), m.indexData.size(), &ib);
		}
		rm->vertexBufs.push_back(vb);
		rm->indexBufs.push_back(ib);
		rm->numVerts.push_back(uint32_t(m.vertexData.size()));
		rm->numIndices.push_back(m.numIndices);
		rm->indexStride.push_back(m.indexType ? 4u : 2u);
		rm->matIdx.push_back(m.matIdx);
	}
	if (rc != SRB_OK)
	{
		Fail(rc, "srb_model_make_resident: %s", srb_last_error(ctx));
		srb_resident_model_free(rm);
		return rc;
	}
	*out = rm;
	return SRB_OK;
}

SRB_API int srb_resident_model_draws(const srb_resident_model* rm, srb_handle framebuffer, const float* mvp,
                                     uint32_t textured_shader, srb_draw_desc* draws, uint32_t cap, uint32_t* n)
{
	if (!rm || !n) return Fail(SRB_ERR_INVALID, "srb_resident_model_draws: null argument");
	uint32_t count = 0;
	for (size_t i = 0; i < rm->vertexBufs.size(); ++i)
	{
		if (!rm->vertexBufs[i]) continue; // a mesh without indices draws nothing
		if (draws && count < cap)
		{
			srb_draw_desc& d = draws[count];
			memset(&d, 0, sizeof(d));
			d.framebuffer = framebuffer;
			d.uv_offset = 6; // offsetof(Obj::Vertex, uv) / sizeof(float), Scene.cpp:41
			d.attributes.buffer = rm->vertexBufs[i];
			d.attributes.stride = sizeof(ObjVertex);
			d.attributes.num = rm->numVerts[i];
			d.positions = d.attributes;
			d.indices.buffer = rm->indexBufs[i];
			d.indices.stride = rm->indexStride[i];
			d.indices.num = rm->numIndices[i];
			bool const hasMat = rm->matIdx[i] < rm->numMaterials;
			d.texture = hasMat ? rm->textures[rm->matIdx[i]] : 0;
			// Scene.cpp:51-59: no material -> VisualizeNormals; SponzaScene.cpp:204-213: SponzaShader with null uniforms
			d.shader = (hasMat || textured_shader == SRB_SHADER_SPONZA) ? textured_shader : uint32_t(SRB_SHADER_VISUALIZE_NORMALS);
			if (mvp) memcpy(d.mvp, mvp, sizeof(d.mvp));
		}
		++count;
	}
	*n = count;
	if (draws && count > cap) return Fail(SRB_ERR_OVERFLOW, "srb_resident_model_draws: more draws than capacity");
	return SRB_OK;
}
