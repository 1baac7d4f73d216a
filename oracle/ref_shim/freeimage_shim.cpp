/* Test-infrastructure shim: a minimal stand-in for the FreeImage 3.17 entry points the
 * reference calls from Others/image_loader.cpp:31-95 (FreeImage ships only as a Windows
 * .lib/.dll in the reference tree, so it cannot be linked here).  It decodes
 *   - "<file>.rgba8" side-cars written by the harness (u32 width, u32 height, then
 *     width*height*4 bytes RGBA, row 0 = top) for formats we do not decode (JPG/TGA/PNG);
 *   - uncompressed 24/32-bit BMP natively.
 * The bitmap it hands back is what FreeImage would hand back after
 * FreeImage_ConvertTo24Bits: bottom-up rows, BGR order, pitch padded to 4 bytes.
 * Not product code. */
#include "lib\free_image\FreeImage.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

struct shim_bitmap { unsigned w, h, pitch; std::vector<BYTE> bits; };

static std::string norm_path(const char* f)
{
	std::string p(f);
	for (auto& c : p) if (c == '\\') c = '/';
	return p;
}

static bool read_file(const std::string& p, std::vector<BYTE>& out)
{
	FILE* fp = fopen(p.c_str(), "rb");
	if (!fp) return false;
	fseek(fp, 0, SEEK_END); long n = ftell(fp); fseek(fp, 0, SEEK_SET);
	out.resize(n > 0 ? n : 0);
	size_t got = n > 0 ? fread(out.data(), 1, n, fp) : 0;
	fclose(fp);
	return got == (size_t)n;
}

static shim_bitmap* from_rgba_topdown(const BYTE* rgba, unsigned w, unsigned h)
{
	shim_bitmap* b = new shim_bitmap();
	b->w = w; b->h = h; b->pitch = (w * 3 + 3) & ~3u;
	b->bits.assign((size_t)b->pitch * h, 0);
	for (unsigned y = 0; y < h; y++)
	{
		BYTE* dst = b->bits.data() + (size_t)(h - 1 - y) * b->pitch;
		const BYTE* src = rgba + (size_t)y * w * 4;
		for (unsigned x = 0; x < w; x++)
		{
			dst[x * 3 + FI_RGBA_RED] = src[x * 4 + 0];
			dst[x * 3 + FI_RGBA_GREEN] = src[x * 4 + 1];
			dst[x * 3 + FI_RGBA_BLUE] = src[x * 4 + 2];
		}
	}
	return b;
}

static unsigned rd32(const BYTE* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((unsigned)p[3] << 24); }
static unsigned rd16(const BYTE* p) { return p[0] | (p[1] << 8); }

static shim_bitmap* load_any(const char* filename)
{
	std::string p = norm_path(filename);
	std::vector<BYTE> buf;
	if (read_file(p + ".rgba8", buf) && buf.size() >= 8)
	{
		unsigned w = rd32(buf.data()), h = rd32(buf.data() + 4);
		if (buf.size() == 8 + (size_t)w * h * 4) return from_rgba_topdown(buf.data() + 8, w, h);
	}
	if (!read_file(p, buf) || buf.size() < 54 || buf[0] != 'B' || buf[1] != 'M') return nullptr;
	unsigned off = rd32(&buf[10]);
	int w = (int)rd32(&buf[18]), h = (int)rd32(&buf[22]);
	unsigned bpp = rd16(&buf[28]), comp = rd32(&buf[30]);
	if ((bpp != 24 && bpp != 32) || (comp != 0 && comp != 3) || w <= 0 || h == 0) return nullptr;
	bool bottom_up = h > 0; unsigned ah = h > 0 ? h : -h;
	unsigned bytes = bpp / 8, pitch = (w * bytes + 3) & ~3u;
	if (buf.size() < off + (size_t)pitch * ah) return nullptr;
	std::vector<BYTE> rgba((size_t)w * ah * 4);
	for (unsigned y = 0; y < ah; y++)
	{
		const BYTE* src = &buf[off + (size_t)(bottom_up ? ah - 1 - y : y) * pitch];
		for (int x = 0; x < w; x++)
		{
			rgba[((size_t)y * w + x) * 4 + 0] = src[x * bytes + 2];
			rgba[((size_t)y * w + x) * 4 + 1] = src[x * bytes + 1];
			rgba[((size_t)y * w + x) * 4 + 2] = src[x * bytes + 0];
			rgba[((size_t)y * w + x) * 4 + 3] = 255;
		}
	}
	return from_rgba_topdown(rgba.data(), w, ah);
}

extern "C" {
void DLL_CALLCONV FreeImage_Initialise(BOOL) {}
void DLL_CALLCONV FreeImage_DeInitialise() {}
FREE_IMAGE_FORMAT DLL_CALLCONV FreeImage_GetFileType(const char* filename, int)
{
	std::string p = norm_path(filename);
	FILE* a = fopen((p + ".rgba8").c_str(), "rb");
	if (a) { fclose(a); return FIF_RAW; }
	FILE* b = fopen(p.c_str(), "rb");
	if (!b) return FIF_UNKNOWN;
	unsigned char m[2] = { 0, 0 };
	size_t n = fread(m, 1, 2, b); fclose(b);
	return (n == 2 && m[0] == 'B' && m[1] == 'M') ? FIF_BMP : FIF_UNKNOWN;
}
FREE_IMAGE_FORMAT DLL_CALLCONV FreeImage_GetFIFFromFilename(const char*) { return FIF_UNKNOWN; }
BOOL DLL_CALLCONV FreeImage_FIFSupportsReading(FREE_IMAGE_FORMAT fif) { return fif == FIF_BMP || fif == FIF_RAW; }
FIBITMAP* DLL_CALLCONV FreeImage_Load(FREE_IMAGE_FORMAT, const char* filename, int)
{
	shim_bitmap* b = load_any(filename);
	if (!b) return nullptr;
	FIBITMAP* f = new FIBITMAP(); f->data = b; return f;
}
void DLL_CALLCONV FreeImage_Unload(FIBITMAP* dib)
{
	/* the reference unloads the source after ConvertTo24Bits, which returns the same
	 * pixels here; ref-count by making ConvertTo24Bits deep-copy. */
	if (!dib) return;
	delete (shim_bitmap*)dib->data; delete dib;
}
FIBITMAP* DLL_CALLCONV FreeImage_ConvertTo24Bits(FIBITMAP* dib)
{
	FIBITMAP* f = new FIBITMAP(); f->data = new shim_bitmap(*(shim_bitmap*)dib->data); return f;
}
unsigned DLL_CALLCONV FreeImage_GetWidth(FIBITMAP* dib) { return ((shim_bitmap*)dib->data)->w; }
unsigned DLL_CALLCONV FreeImage_GetHeight(FIBITMAP* dib) { return ((shim_bitmap*)dib->data)->h; }
unsigned DLL_CALLCONV FreeImage_GetPitch(FIBITMAP* dib) { return ((shim_bitmap*)dib->data)->pitch; }
BYTE* DLL_CALLCONV FreeImage_GetBits(FIBITMAP* dib) { return ((shim_bitmap*)dib->data)->bits.data(); }
FREE_IMAGE_TYPE DLL_CALLCONV FreeImage_GetImageType(FIBITMAP*) { return FIT_BITMAP; }
unsigned DLL_CALLCONV FreeImage_GetBPP(FIBITMAP*) { return 24; }
}
