// Output side of the render path: what Main/window.cpp:712-740 `screenshot()` does with
// lodepng (RGBA8 PNG of the displayed image), plus a float dump (PFM) and a checkpoint of the accumulation
// (Core/image.h:10-23: pixels + pass_counter) so a render can be resumed or merged.  Dependency-free: the PNG
// is written with stored (uncompressed) deflate blocks.
#include "image_out.h"

#include <cstdio>
#include <cstring>
#include <vector>

namespace ptb
{

namespace
{

uint32_t crc_table[256];
bool crc_ready = false;

void crc_init()
{
	for (uint32_t n = 0; n < 256; n++)
	{
		uint32_t c = n;
		for (int k = 0; k < 8; k++) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
		crc_table[n] = c;
	}
	crc_ready = true;
}

uint32_t crc32(uint32_t crc, const uint8_t* p, size_t n)
{
	if (!crc_ready) crc_init();
	crc = ~crc;
	for (size_t i = 0; i < n; i++) crc = crc_table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
	return ~crc;
}

void put_be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }

bool write_chunk(FILE* f, const char type[4], const std::vector<uint8_t>& data)
{
	std::vector<uint8_t> head;
	put_be32(head, (uint32_t)data.size());
	head.insert(head.end(), type, type + 4);
	uint32_t crc = crc32(0, (const uint8_t*)type, 4);
	if (!data.empty()) crc = crc32(crc, data.data(), data.size());
	std::vector<uint8_t> tail;
	put_be32(tail, crc);
	return fwrite(head.data(), 1, head.size(), f) == head.size() && (data.empty() || fwrite(data.data(), 1, data.size(), f) == data.size()) &&
		fwrite(tail.data(), 1, 4, f) == 4;
}

} // namespace

bool write_png_rgb8(const std::string& path, const uint8_t* rgb, int width, int height, std::string& err)
{
	if (!rgb || width <= 0 || height <= 0) { err = "[Error]write_png: empty image"; return false; }
	FILE* f = fopen(path.c_str(), "wb");
	if (!f) { err = "[Error]write_png: cannot open " + path; return false; }
	static const uint8_t sig[8] = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
	bool ok = fwrite(sig, 1, 8, f) == 8;
	std::vector<uint8_t> ihdr;
	put_be32(ihdr, (uint32_t)width); put_be32(ihdr, (uint32_t)height);
	ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);   // 8-bit RGBA like the reference's screenshot
	ok = ok && write_chunk(f, "IHDR", ihdr);

	// raw scanlines: filter byte 0 + RGBA
	const size_t stride = (size_t)width * 4 + 1;
	std::vector<uint8_t> raw(stride * height);
	for (int y = 0; y < height; y++)
	{
		uint8_t* row = &raw[stride * y];
		row[0] = 0;
		const uint8_t* src = rgb + (size_t)y * width * 3;
		for (int x = 0; x < width; x++) { row[1 + x * 4] = src[x * 3]; row[2 + x * 4] = src[x * 3 + 1]; row[3 + x * 4] = src[x * 3 + 2]; row[4 + x * 4] = 255; }
	}
	// zlib stream of stored blocks
	std::vector<uint8_t> z;
	z.reserve(raw.size() + raw.size() / 65535 * 5 + 16);
	z.push_back(0x78); z.push_back(0x01);
	uint32_t a = 1, b = 0;
	for (size_t i = 0; i < raw.size(); i++) { a = (a + raw[i]) % 65521u; b = (b + a) % 65521u; }
	size_t pos = 0;
	while (pos < raw.size())
	{
		size_t n = raw.size() - pos < 65535 ? raw.size() - pos : 65535;
		z.push_back(pos + n == raw.size() ? 1 : 0);
		z.push_back(n & 0xff); z.push_back(n >> 8); z.push_back(~n & 0xff); z.push_back((~n >> 8) & 0xff);
		z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
		pos += n;
	}
	put_be32(z, (b << 16) | a);
	ok = ok && write_chunk(f, "IDAT", z);
	ok = ok && write_chunk(f, "IEND", std::vector<uint8_t>());
	ok = (fclose(f) == 0) && ok;
	if (!ok) err = "[Error]write_png: short write to " + path;
	return ok;
}

bool write_pfm_rgb(const std::string& path, const float* rgb, int width, int height, float scale, std::string& err)
{
	if (!rgb || width <= 0 || height <= 0) { err = "[Error]write_pfm: empty image"; return false; }
	FILE* f = fopen(path.c_str(), "wb");
	if (!f) { err = "[Error]write_pfm: cannot open " + path; return false; }
	fprintf(f, "PF\n%d %d\n-1.0\n", width, height);     // negative scale: little-endian; rows bottom to top
	std::vector<float> row((size_t)width * 3);
	bool ok = true;
	for (int y = height - 1; y >= 0 && ok; y--)
	{
		const float* src = rgb + (size_t)y * width * 3;
		for (int i = 0; i < width * 3; i++) row[i] = src[i] * scale;
		ok = fwrite(row.data(), sizeof(float), row.size(), f) == row.size();
	}
	ok = (fclose(f) == 0) && ok;
	if (!ok) err = "[Error]write_pfm: short write to " + path;
	return ok;
}

static const char kMagic[8] = { 'P', 'T', 'B', '2', '0', '0', 'C', '2' };

bool write_checkpoint(const std::string& path, const CheckpointHeader& h, const float* sum_rgb, std::string& err)
{
	FILE* f = fopen(path.c_str(), "wb");
	if (!f) { err = "[Error]checkpoint: cannot open " + path; return false; }
	const size_t n = (size_t)h.width * h.height * 3;
	// the checksum covers the header too: a flipped pass counter or camera would resume a different render
	uint32_t crc = crc32(0, (const uint8_t*)&h, sizeof(h));
	crc = crc32(crc, (const uint8_t*)sum_rgb, n * sizeof(float));
	bool ok = fwrite(kMagic, 1, 8, f) == 8 && fwrite(&h, sizeof(h), 1, f) == 1 && fwrite(&crc, 4, 1, f) == 1 && fwrite(sum_rgb, sizeof(float), n, f) == n;
	ok = (fclose(f) == 0) && ok;
	if (!ok) err = "[Error]checkpoint: short write to " + path;
	return ok;
}

bool read_checkpoint(const std::string& path, CheckpointHeader& h, std::vector<float>& sum_rgb, std::string& err)
{
	FILE* f = fopen(path.c_str(), "rb");
	if (!f) { err = "[Error]checkpoint: cannot open " + path; return false; }
	char magic[8];
	uint32_t crc = 0;
	bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, kMagic, 8) == 0 && fread(&h, sizeof(h), 1, f) == 1 && fread(&crc, 4, 1, f) == 1;
	if (ok && (h.width <= 0 || h.height <= 0 || h.width > 65536 || h.height > 65536 || h.pass_counter < 0)) ok = false;
	if (ok)
	{
		// the file must hold exactly the announced pixels: checked BEFORE allocating them
		const size_t n = (size_t)h.width * h.height * 3;
		const long here = ftell(f);
		ok = here >= 0 && fseek(f, 0, SEEK_END) == 0 && ftell(f) >= here && (size_t)(ftell(f) - here) == n * sizeof(float) && fseek(f, here, SEEK_SET) == 0;
		if (ok)
		{
			sum_rgb.resize(n);
			uint32_t want = crc32(0, (const uint8_t*)&h, sizeof(h));
			ok = fread(sum_rgb.data(), sizeof(float), n, f) == n && crc32(want, (const uint8_t*)sum_rgb.data(), n * sizeof(float)) == crc;
		}
	}
	fclose(f);
	if (!ok) err = "[Error]checkpoint: " + path + " is not a valid ptb200 checkpoint (bad magic, size or checksum)";
	return ok;
}

} // namespace ptb
