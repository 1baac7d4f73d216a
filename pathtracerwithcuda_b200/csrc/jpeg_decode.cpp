// JPEG decoder (baseline and progressive Huffman) for the scene front-end.
//
// The reference decodes images through FreeImage (Others/image_loader.cpp:31-95: FreeImage_Load + ConvertTo24Bits, a
// binary-only library built on IJG libjpeg); its own assets are two cube maps of baseline 4:2:0 JPEGs
// (res/texture/lycksele, maskonaive).  It passes load flags 0 (image_loader.cpp:45), which FreeImage documents as
// JPEG_DEFAULT = JPEG_FAST, "load the file as fast as possible, sacrificing some quality"
// (lib/free_image/FreeImage.h:693-695): libjpeg then runs with dct_method = JDCT_IFAST and do_fancy_upsampling = FALSE.
// Three modes (set_jpeg_mode / ptb_set_jpeg_decode), restating the published IJG algorithms:
//   * kJpegReference (default) — what the reference gets: sequential Huffman decoding (ITU T.81 Annex F); the AAN "ifast"
//     integer inverse DCT (jidctfst.c: 8-bit constants, multipliers pre-scaled by the AAN factors in jddctmgr.c, NO rounding in
//     the descales, 10-bit wrapped range limit); chroma replicated (jdsample.c h2v1/h2v2_upsample == jdmerge.c); fixed-point
//     YCbCr -> RGB with the constants of IJG libjpeg 9a (what FreeImage 3.17 bundles): G = Y - 0.344136286 Cb - 0.714136286 Cr;
//   * kJpegFast — the same with libjpeg 6b's rounded constants 0.34414 / 0.71414 (one table entry differs by 1/65536):
//     bit-identical to libjpeg-turbo run with dct_method=JDCT_IFAST, do_fancy_upsampling=FALSE (pinned through
//     oracle/jpeg_lib_shim.c on the reference's own cube maps, committed golden vectors and live synthetic files,
//     tests/test_image_out.py);
//   * kJpegAccurate — "islow" IDCT (jidctint.c), triangle-filter chroma upsampling, 6b constants: bit-identical to
//     libjpeg-turbo's default decode, i.e. PIL.
// The IFAST + replication part of the reference mode is therefore pinned; its colour constant (22553 vs 22554 in one table)
// follows the published libjpeg 9a source and is not pinned by a binary here.
// Every scan goes into per-component coefficient arrays (non-interleaved scans walk the component's own block grid), so
// multi-scan baseline files and progressive files (Annex G: DC/AC first + refinement scans, EOB runs) share one path; the
// quantisation table is latched at a component's first scan as libjpeg does.  Complete progressive files need none of
// libjpeg's block smoothing (all coefficients exact); a file cut short is rejected rather than approximated.
// Arithmetic-coded / 12-bit / CMYK files are rejected (side-car needed).
#include "scene.h"

#include <cstring>

namespace ptb
{

namespace
{

const int kZigzag[64] = { 0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
	35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };

struct HuffTable
{
	bool present = false;
	uint8_t bits[17] = { 0 };
	uint8_t vals[256] = { 0 };
	int mincode[18], maxcode[18], valptr[18];
	bool build()
	{
		int code = 0, k = 0;
		for (int l = 1; l <= 16; l++)
		{
			valptr[l] = k;
			mincode[l] = code;
			code += bits[l];
			k += bits[l];
			maxcode[l] = bits[l] ? code - 1 : -1;
			code <<= 1;
			if (k > 256) return false;
		}
		maxcode[17] = 0x7fffffff;
		return true;
	}
};

struct Component
{
	int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
	int width = 0, height = 0;          // real sample dimensions (ceil(image * h / hmax))
	int bw = 0, bh = 0;                 // blocks covering the real samples: the grid of a non-interleaved scan
	int pbw = 0, pbh = 0;               // blocks padded to whole MCUs: the grid of interleaved scans and of `coefs`
	int stride = 0, rows = 0;           // plane, padded to whole MCUs
	std::vector<uint8_t> plane;
	std::vector<int16_t> coefs;         // pbw * pbh blocks of 64, natural order
	int dc_pred = 0;
	bool seen = false;                  // appeared in a scan: quantisation table latched (as libjpeg does at a component's first scan)
	uint16_t quant[64] = { 0 };
	int ifast_mult[64] = { 0 };
	int8_t known_al[64];                // progressive: successive-approximation bit still missing per coefficient (-1 = never coded)
};

struct BitSource
{
	const uint8_t* p; size_t n, pos;
	uint32_t acc = 0; int cnt = 0;
	bool hit_marker = false;
	BitSource(const uint8_t* d, size_t len, size_t start) : p(d), n(len), pos(start) {}
	void fill()
	{
		while (cnt <= 24)
		{
			int byte = 0;
			if (!hit_marker && pos < n)
			{
				byte = p[pos];
				if (byte == 0xff)
				{
					if (pos + 1 < n && p[pos + 1] == 0x00) pos += 2;
					else { hit_marker = true; byte = 0; }     // a marker: feed zeros, like libjpeg's "insert_fake_data"
				}
				else pos++;
			}
			acc |= (uint32_t)byte << (24 - cnt);
			cnt += 8;
		}
	}
	int get(int k)
	{
		if (k == 0) return 0;
		if (cnt < k) fill();
		int v = (int)(acc >> (32 - k));
		acc <<= k; cnt -= k;
		return v;
	}
	void reset() { acc = 0; cnt = 0; hit_marker = false; }
};

inline int huff_decode(BitSource& bs, const HuffTable& t)
{
	int code = 0;
	for (int l = 1; l <= 16; l++)
	{
		code = (code << 1) | bs.get(1);
		if (t.maxcode[l] >= 0 && code <= t.maxcode[l] && code >= t.mincode[l]) return t.vals[t.valptr[l] + code - t.mincode[l]];
	}
	return -1;
}

inline int huff_extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

// jidctint.c (islow)
const int CONST_BITS = 13, PASS1_BITS = 2;
const int FIX_0_298631336 = 2446, FIX_0_390180644 = 3196, FIX_0_541196100 = 4433, FIX_0_765366865 = 6270, FIX_0_899976223 = 7373,
	FIX_1_175875602 = 9633, FIX_1_501321110 = 12299, FIX_1_847759065 = 15137, FIX_1_961570560 = 16069, FIX_2_053119869 = 16819,
	FIX_2_562915447 = 20995, FIX_3_072711026 = 25172;

inline int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
inline uint8_t range_limit(int x) { x += 128; return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }

void idct_islow(const int16_t* coef, const uint16_t* quant, uint8_t* out, int stride)
{
	int ws[64];
	for (int c = 0; c < 8; c++)
	{
		const int16_t* in = coef + c;
		const uint16_t* q = quant + c;
		int* w = ws + c;
		if (in[8] == 0 && in[16] == 0 && in[24] == 0 && in[32] == 0 && in[40] == 0 && in[48] == 0 && in[56] == 0)
		{
			int dc = (in[0] * q[0]) * (1 << PASS1_BITS);
			for (int r = 0; r < 8; r++) w[r * 8] = dc;
			continue;
		}
		int z2 = in[16] * q[16], z3 = in[48] * q[48];
		int z1 = (z2 + z3) * FIX_0_541196100;
		int tmp2 = z1 + z3 * (-FIX_1_847759065);
		int tmp3 = z1 + z2 * FIX_0_765366865;
		z2 = in[0] * q[0]; z3 = in[32] * q[32];
		int tmp0 = (z2 + z3) * (1 << CONST_BITS);
		int tmp1 = (z2 - z3) * (1 << CONST_BITS);
		int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
		tmp0 = in[56] * q[56]; tmp1 = in[40] * q[40]; tmp2 = in[24] * q[24]; tmp3 = in[8] * q[8];
		z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
		int z4 = tmp1 + tmp3;
		int z5 = (z3 + z4) * FIX_1_175875602;
		tmp0 *= FIX_0_298631336; tmp1 *= FIX_2_053119869; tmp2 *= FIX_3_072711026; tmp3 *= FIX_1_501321110;
		z1 *= -FIX_0_899976223; z2 *= -FIX_2_562915447; z3 *= -FIX_1_961570560; z4 *= -FIX_0_390180644;
		z3 += z5; z4 += z5;
		tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
		w[0] = descale(tmp10 + tmp3, CONST_BITS - PASS1_BITS); w[56] = descale(tmp10 - tmp3, CONST_BITS - PASS1_BITS);
		w[8] = descale(tmp11 + tmp2, CONST_BITS - PASS1_BITS); w[48] = descale(tmp11 - tmp2, CONST_BITS - PASS1_BITS);
		w[16] = descale(tmp12 + tmp1, CONST_BITS - PASS1_BITS); w[40] = descale(tmp12 - tmp1, CONST_BITS - PASS1_BITS);
		w[24] = descale(tmp13 + tmp0, CONST_BITS - PASS1_BITS); w[32] = descale(tmp13 - tmp0, CONST_BITS - PASS1_BITS);
	}
	for (int r = 0; r < 8; r++)
	{
		const int* w = ws + r * 8;
		uint8_t* o = out + (size_t)r * stride;
		int z2 = w[2], z3 = w[6];
		int z1 = (z2 + z3) * FIX_0_541196100;
		int tmp2 = z1 + z3 * (-FIX_1_847759065);
		int tmp3 = z1 + z2 * FIX_0_765366865;
		int tmp0 = (w[0] + w[4]) * (1 << CONST_BITS);
		int tmp1 = (w[0] - w[4]) * (1 << CONST_BITS);
		int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
		tmp0 = w[7]; tmp1 = w[5]; tmp2 = w[3]; tmp3 = w[1];
		z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
		int z4 = tmp1 + tmp3;
		int z5 = (z3 + z4) * FIX_1_175875602;
		tmp0 *= FIX_0_298631336; tmp1 *= FIX_2_053119869; tmp2 *= FIX_3_072711026; tmp3 *= FIX_1_501321110;
		z1 *= -FIX_0_899976223; z2 *= -FIX_2_562915447; z3 *= -FIX_1_961570560; z4 *= -FIX_0_390180644;
		z3 += z5; z4 += z5;
		tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
		const int sh = CONST_BITS + PASS1_BITS + 3;
		o[0] = range_limit(descale(tmp10 + tmp3, sh)); o[7] = range_limit(descale(tmp10 - tmp3, sh));
		o[1] = range_limit(descale(tmp11 + tmp2, sh)); o[6] = range_limit(descale(tmp11 - tmp2, sh));
		o[2] = range_limit(descale(tmp12 + tmp1, sh)); o[5] = range_limit(descale(tmp12 - tmp1, sh));
		o[3] = range_limit(descale(tmp13 + tmp0, sh)); o[4] = range_limit(descale(tmp13 - tmp0, sh));
	}
}

// jidctfst.c ("ifast", AAN): multipliers = quant * AAN scale factors (jddctmgr.c: 14-bit table, DESCALE by 14 - IFAST_SCALE_BITS),
// 8-bit constants, descales WITHOUT rounding, output through the 10-bit wrapped range-limit table.
// kAanScales[8 i + j] = round(2^14 s_i s_j), s_0 = 1, s_k = sqrt(2) cos(k pi / 16).
const int16_t kAanScales[64] = {
	16384, 22725, 21407, 19266, 16384, 12873, 8867, 4520, 22725, 31521, 29692, 26722, 22725, 17855, 12299, 6270,
	21407, 29692, 27969, 25172, 21407, 16819, 11585, 5906, 19266, 26722, 25172, 22654, 19266, 15137, 10426, 5315,
	16384, 22725, 21407, 19266, 16384, 12873, 8867, 4520, 12873, 17855, 16819, 15137, 12873, 10114, 6967, 3552,
	8867, 12299, 11585, 10426, 8867, 6967, 4799, 2446, 4520, 6270, 5906, 5315, 4520, 3552, 2446, 1247 };

inline void ifast_multipliers(const uint16_t* quant, int* mult)
{
	for (int i = 0; i < 64; i++) mult[i] = (int)(((int64_t)quant[i] * kAanScales[i] + (1 << 11)) >> 12);
}

inline int fmul(int v, int c) { return (v * c) >> 8; }
inline uint8_t range_limit_wrapped(int x)
{
	x &= 1023;                       // RANGE_MASK: values beyond +-512 wrap, as in the library's table
	if (x >= 512) x -= 1024;
	x += 128;
	return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x));
}

void idct_ifast(const int16_t* coef, const int* mult, uint8_t* out, int stride)
{
	const int F1_082 = 277, F1_414 = 362, F1_847 = 473, F2_613 = 669;
	int ws[64];
	for (int c = 0; c < 8; c++)
	{
		const int16_t* in = coef + c;
		const int* q = mult + c;
		int* w = ws + c;
		if (in[8] == 0 && in[16] == 0 && in[24] == 0 && in[32] == 0 && in[40] == 0 && in[48] == 0 && in[56] == 0)
		{
			const int dc = in[0] * q[0];
			for (int r = 0; r < 8; r++) w[r * 8] = dc;
			continue;
		}
		int tmp0 = in[0] * q[0], tmp1 = in[16] * q[16], tmp2 = in[32] * q[32], tmp3 = in[48] * q[48];
		int tmp10 = tmp0 + tmp2, tmp11 = tmp0 - tmp2;
		int tmp13 = tmp1 + tmp3;
		int tmp12 = fmul(tmp1 - tmp3, F1_414) - tmp13;
		tmp0 = tmp10 + tmp13; tmp3 = tmp10 - tmp13; tmp1 = tmp11 + tmp12; tmp2 = tmp11 - tmp12;
		int tmp4 = in[8] * q[8], tmp5 = in[24] * q[24], tmp6 = in[40] * q[40], tmp7 = in[56] * q[56];
		const int z13 = tmp6 + tmp5, z10 = tmp6 - tmp5, z11 = tmp4 + tmp7, z12 = tmp4 - tmp7;
		tmp7 = z11 + z13;
		tmp11 = fmul(z11 - z13, F1_414);
		const int z5 = fmul(z10 + z12, F1_847);
		tmp10 = fmul(z12, F1_082) - z5;
		tmp12 = fmul(z10, -F2_613) + z5;
		tmp6 = tmp12 - tmp7; tmp5 = tmp11 - tmp6; tmp4 = tmp10 + tmp5;
		w[0] = tmp0 + tmp7; w[56] = tmp0 - tmp7;
		w[8] = tmp1 + tmp6; w[48] = tmp1 - tmp6;
		w[16] = tmp2 + tmp5; w[40] = tmp2 - tmp5;
		w[32] = tmp3 + tmp4; w[24] = tmp3 - tmp4;
	}
	for (int r = 0; r < 8; r++)
	{
		const int* w = ws + r * 8;
		uint8_t* o = out + (size_t)r * stride;
		int tmp10 = w[0] + w[4], tmp11 = w[0] - w[4];
		int tmp13 = w[2] + w[6];
		int tmp12 = fmul(w[2] - w[6], F1_414) - tmp13;
		const int tmp0 = tmp10 + tmp13, tmp3 = tmp10 - tmp13, tmp1 = tmp11 + tmp12, tmp2 = tmp11 - tmp12;
		const int z13 = w[5] + w[3], z10 = w[5] - w[3], z11 = w[1] + w[7], z12 = w[1] - w[7];
		const int tmp7 = z11 + z13;
		tmp11 = fmul(z11 - z13, F1_414);
		const int z5 = fmul(z10 + z12, F1_847);
		tmp10 = fmul(z12, F1_082) - z5;
		tmp12 = fmul(z10, -F2_613) + z5;
		const int tmp6 = tmp12 - tmp7, tmp5 = tmp11 - tmp6, tmp4 = tmp10 + tmp5;
		o[0] = range_limit_wrapped((tmp0 + tmp7) >> 5); o[7] = range_limit_wrapped((tmp0 - tmp7) >> 5);
		o[1] = range_limit_wrapped((tmp1 + tmp6) >> 5); o[6] = range_limit_wrapped((tmp1 - tmp6) >> 5);
		o[2] = range_limit_wrapped((tmp2 + tmp5) >> 5); o[5] = range_limit_wrapped((tmp2 - tmp5) >> 5);
		o[4] = range_limit_wrapped((tmp3 + tmp4) >> 5); o[3] = range_limit_wrapped((tmp3 - tmp4) >> 5);
	}
}

uint16_t be16(const uint8_t* p) { return (uint16_t)((p[0] << 8) | p[1]); }

// one output row of a component at full resolution (jdsample.c)
void upsample_row(const Component& c, int hmax, int vmax, int y, int out_width, std::vector<uint8_t>& row, bool fancy)
{
	const int hs = hmax / c.h, vs = vmax / c.v;
	row.resize((size_t)out_width + 2);
	auto src_row = [&](int r) { r = r < 0 ? 0 : (r >= c.height ? c.height - 1 : r); return c.plane.data() + (size_t)r * c.stride; };
	if (hs == 1 && vs == 1) { memcpy(row.data(), src_row(y), (size_t)out_width); return; }
	const int n = c.width;
	if (fancy && hs == 2 && vs == 1)
	{
		const uint8_t* in = src_row(y);
		if (n <= 2) { for (int x = 0; x < out_width; x++) row[x] = in[(x / 2) < n ? x / 2 : n - 1]; return; }   // libjpeg: fancy only if width > 2
		std::vector<uint8_t> tmp((size_t)2 * n);
		tmp[0] = in[0];
		tmp[1] = (uint8_t)((in[0] * 3 + in[1] + 2) >> 2);
		for (int i = 1; i < n - 1; i++)
		{
			const int v = in[i] * 3;
			tmp[2 * i] = (uint8_t)((v + in[i - 1] + 1) >> 2);
			tmp[2 * i + 1] = (uint8_t)((v + in[i + 1] + 2) >> 2);
		}
		tmp[2 * n - 2] = (uint8_t)((in[n - 1] * 3 + in[n - 2] + 1) >> 2);
		tmp[2 * n - 1] = in[n - 1];
		memcpy(row.data(), tmp.data(), (size_t)out_width);
		return;
	}
	if (fancy && hs == 2 && vs == 2)
	{
		const int r = y >> 1;
		const uint8_t* in0 = src_row(r);
		const uint8_t* in1 = src_row((y & 1) ? r + 1 : r - 1);     // the nearer row weighs 3, the other neighbour 1
		if (n <= 2)
		{
			// libjpeg falls back to plain replication for tiny components
			for (int x = 0; x < out_width; x++) row[x] = in0[(x / 2) < n ? x / 2 : n - 1];
			return;
		}
		std::vector<uint8_t> tmp((size_t)2 * n);
		int thiscolsum = in0[0] * 3 + in1[0];
		int nextcolsum = in0[1] * 3 + in1[1];
		tmp[0] = (uint8_t)((thiscolsum * 4 + 8) >> 4);
		tmp[1] = (uint8_t)((thiscolsum * 3 + nextcolsum + 7) >> 4);
		int lastcolsum = thiscolsum;
		thiscolsum = nextcolsum;
		for (int i = 1; i < n - 1; i++)
		{
			nextcolsum = in0[i + 1] * 3 + in1[i + 1];
			tmp[2 * i] = (uint8_t)((thiscolsum * 3 + lastcolsum + 8) >> 4);
			tmp[2 * i + 1] = (uint8_t)((thiscolsum * 3 + nextcolsum + 7) >> 4);
			lastcolsum = thiscolsum;
			thiscolsum = nextcolsum;
		}
		tmp[2 * n - 2] = (uint8_t)((thiscolsum * 3 + lastcolsum + 8) >> 4);
		tmp[2 * n - 1] = (uint8_t)((thiscolsum * 4 + 7) >> 4);
		memcpy(row.data(), tmp.data(), (size_t)out_width);
		return;
	}
	// other integral factors: replication (int_upsample)
	const uint8_t* in = src_row(y / vs);
	for (int x = 0; x < out_width; x++) { int sx = x / hs; row[x] = in[sx < n ? sx : n - 1]; }
}

// One scan's entropy-coded segment (ITU T.81 Annex F sequential, Annex G progressive).  Blocks go to the coefficient
// arrays; a non-interleaved scan walks the component's own block grid, an interleaved one the MCU grid.
struct Scan
{
	int ns = 0;
	Component* comp[3] = { nullptr, nullptr, nullptr };
	int ss = 0, se = 63, ah = 0, al = 0;
	bool progressive = false;
};

bool decode_scan(const std::vector<uint8_t>& f, size_t& pos, const Scan& sc, const HuffTable* dc_tab, const HuffTable* ac_tab,
	int mcus_x, int mcus_y, int restart_interval)
{
	BitSource bs(f.data(), f.size(), pos);
	int eobrun = 0;
	const int total_x = sc.ns == 1 ? sc.comp[0]->bw : mcus_x, total_y = sc.ns == 1 ? sc.comp[0]->bh : mcus_y;
	int restarts_left = restart_interval;
	const int p1 = 1 << sc.al, m1 = -(1 << sc.al);
	for (int i = 0; i < sc.ns; i++) sc.comp[i]->dc_pred = 0;

	auto decode_block = [&](Component& c, int16_t* block) -> bool
	{
		if (!sc.progressive)
		{
			int s = huff_decode(bs, dc_tab[c.td]);
			if (s < 0 || s > 15) return false;
			c.dc_pred += s ? huff_extend(bs.get(s), s) : 0;
			block[0] = (int16_t)c.dc_pred;
			for (int k = 1; k < 64;)
			{
				const int rs = huff_decode(bs, ac_tab[c.ta]);
				if (rs < 0) return false;
				const int rr = rs >> 4, ss = rs & 15;
				if (ss == 0)
				{
					if (rr != 15) break;
					k += 16;
					continue;
				}
				k += rr;
				if (k > 63) return false;
				block[kZigzag[k]] = (int16_t)huff_extend(bs.get(ss), ss);
				k++;
			}
			return true;
		}
		if (sc.ss == 0)
		{
			if (sc.ah == 0)
			{
				const int s = huff_decode(bs, dc_tab[c.td]);
				if (s < 0 || s > 15) return false;
				c.dc_pred += s ? huff_extend(bs.get(s), s) : 0;
				block[0] = (int16_t)(c.dc_pred * (1 << sc.al));
			}
			else if (bs.get(1)) block[0] |= (int16_t)p1;
			return true;
		}
		if (sc.ah == 0)
		{
			if (eobrun > 0) { eobrun--; return true; }
			for (int k = sc.ss; k <= sc.se; k++)
			{
				const int rs = huff_decode(bs, ac_tab[c.ta]);
				if (rs < 0) return false;
				const int r = rs >> 4, s = rs & 15;
				if (s)
				{
					k += r;
					if (k > 63) return false;
					block[kZigzag[k]] = (int16_t)(huff_extend(bs.get(s), s) * (1 << sc.al));
				}
				else if (r == 15) k += 15;
				else
				{
					eobrun = 1 << r;
					if (r) eobrun += bs.get(r);
					eobrun--;
					break;
				}
			}
			return true;
		}
		// AC refinement (G.1.2.3): correction bits for coefficients already non-zero, new +-1 coefficients after runs of zeros
		int k = sc.ss;
		if (eobrun == 0)
		{
			for (; k <= sc.se; k++)
			{
				const int rs = huff_decode(bs, ac_tab[c.ta]);
				if (rs < 0) return false;
				int r = rs >> 4, s = rs & 15;
				if (s)
				{
					if (s != 1) return false;
					s = bs.get(1) ? p1 : m1;
				}
				else if (r != 15)
				{
					eobrun = 1 << r;
					if (r) eobrun += bs.get(r);
					break;
				}
				do
				{
					int16_t& coef = block[kZigzag[k]];
					if (coef != 0)
					{
						if (bs.get(1) && (coef & p1) == 0) coef = (int16_t)(coef + (coef >= 0 ? p1 : m1));
					}
					else if (--r < 0) break;
					k++;
				} while (k <= sc.se);
				if (s)
				{
					if (k > 63) return false;
					block[kZigzag[k]] = (int16_t)s;
				}
			}
		}
		if (eobrun > 0)
		{
			for (; k <= sc.se; k++)
			{
				int16_t& coef = block[kZigzag[k]];
				if (coef != 0 && bs.get(1) && (coef & p1) == 0) coef = (int16_t)(coef + (coef >= 0 ? p1 : m1));
			}
			eobrun--;
		}
		return true;
	};

	for (int my = 0; my < total_y; my++)
		for (int mx = 0; mx < total_x; mx++)
		{
			if (restart_interval && restarts_left == 0)
			{
				// byte-align, expect RSTn
				bs.reset();
				size_t p2 = bs.pos;
				while (p2 + 1 < f.size() && !(f[p2] == 0xff && f[p2 + 1] >= 0xd0 && f[p2 + 1] <= 0xd7)) p2++;
				if (p2 + 1 >= f.size()) return false;
				bs.pos = p2 + 2;
				for (int i = 0; i < sc.ns; i++) sc.comp[i]->dc_pred = 0;
				eobrun = 0;
				restarts_left = restart_interval;
			}
			if (sc.ns == 1)
			{
				Component& c = *sc.comp[0];
				if (!decode_block(c, c.coefs.data() + ((size_t)my * c.pbw + mx) * 64)) return false;
			}
			else
				for (int i = 0; i < sc.ns; i++)
				{
					Component& c = *sc.comp[i];
					for (int by = 0; by < c.v; by++)
						for (int bx = 0; bx < c.h; bx++)
							if (!decode_block(c, c.coefs.data() + ((size_t)(my * c.v + by) * c.pbw + (mx * c.h + bx)) * 64)) return false;
				}
			if (restart_interval) restarts_left--;
		}
	// the next marker: entropy-coded data holds only stuffed 0xff00 and RSTn
	size_t p2 = bs.pos;
	while (p2 + 1 < f.size() && !(f[p2] == 0xff && f[p2 + 1] != 0x00 && f[p2 + 1] != 0xff && !(f[p2 + 1] >= 0xd0 && f[p2 + 1] <= 0xd7))) p2++;
	pos = p2;
	return true;
}

int g_jpeg_mode = kJpegReference;

} // namespace

void set_jpeg_mode(int mode) { g_jpeg_mode = mode; }
int jpeg_mode() { return g_jpeg_mode; }

bool decode_jpeg(const std::vector<uint8_t>& f, Texture& out)
{
	const int mode = g_jpeg_mode;
	const bool fast = mode != kJpegAccurate;
	if (f.size() < 4 || f[0] != 0xff || f[1] != 0xd8) return false;
	uint16_t quant[4][64];
	bool quant_present[4] = { false, false, false, false };
	HuffTable dc_tab[4], ac_tab[4];
	Component comp[3];
	int n_comp = 0, width = 0, height = 0, hmax = 1, vmax = 1, restart_interval = 0, mcus_x = 0, mcus_y = 0;
	bool have_sof = false, progressive = false, adobe = false, saw_scan = false;
	int adobe_transform = -1;
	size_t pos = 2;
	while (pos + 4 <= f.size())
	{
		if (f[pos] != 0xff) { pos++; continue; }
		const int marker = f[pos + 1];
		if (marker == 0xff) { pos++; continue; }
		pos += 2;
		if (marker == 0xd8 || (marker >= 0xd0 && marker <= 0xd7) || marker == 0x01) continue;
		if (marker == 0xd9) break;
		if (pos + 2 > f.size()) return false;
		const size_t len = be16(&f[pos]);
		if (len < 2 || pos + len > f.size()) return false;
		const uint8_t* seg = &f[pos + 2];
		const size_t seg_len = len - 2;
		if (marker == 0xdb)
		{
			size_t o = 0;
			while (o < seg_len)
			{
				const int pq = seg[o] >> 4, tq = seg[o] & 15;
				o++;
				if (tq > 3 || o + (pq ? 128 : 64) > seg_len) return false;
				for (int i = 0; i < 64; i++) { quant[tq][kZigzag[i]] = pq ? be16(&seg[o + 2 * i]) : seg[o + i]; }
				o += pq ? 128 : 64;
				quant_present[tq] = true;
			}
		}
		else if (marker == 0xc0 || marker == 0xc1 || marker == 0xc2)
		{
			if (have_sof || seg_len < 6 || seg[0] != 8) return false;
			progressive = marker == 0xc2;
			height = be16(&seg[1]); width = be16(&seg[3]); n_comp = seg[5];
			if (width <= 0 || height <= 0 || (uint64_t)width * height > kMaxImagePixels || (n_comp != 1 && n_comp != 3) || seg_len < (size_t)(6 + 3 * n_comp)) return false;
			for (int i = 0; i < n_comp; i++)
			{
				comp[i].id = seg[6 + 3 * i];
				comp[i].h = seg[7 + 3 * i] >> 4; comp[i].v = seg[7 + 3 * i] & 15;
				comp[i].tq = seg[8 + 3 * i];
				if (comp[i].h < 1 || comp[i].h > 2 || comp[i].v < 1 || comp[i].v > 2 || comp[i].tq > 3) return false;
				hmax = comp[i].h > hmax ? comp[i].h : hmax;
				vmax = comp[i].v > vmax ? comp[i].v : vmax;
			}
			if (n_comp == 1) { comp[0].h = comp[0].v = 1; hmax = vmax = 1; }       // a single component is never interleaved
			const int mcu_w = 8 * hmax, mcu_h = 8 * vmax;
			mcus_x = (width + mcu_w - 1) / mcu_w; mcus_y = (height + mcu_h - 1) / mcu_h;
			for (int i = 0; i < n_comp; i++)
			{
				Component& c = comp[i];
				c.width = (width * c.h + hmax - 1) / hmax;
				c.height = (height * c.v + vmax - 1) / vmax;
				c.bw = (c.width + 7) / 8; c.bh = (c.height + 7) / 8;
				c.pbw = mcus_x * c.h; c.pbh = mcus_y * c.v;
				c.stride = c.pbw * 8; c.rows = c.pbh * 8;
				c.coefs.assign((size_t)c.pbw * c.pbh * 64, 0);
				memset(c.known_al, -1, sizeof(c.known_al));
			}
			have_sof = true;
		}
		else if (marker >= 0xc3 && marker <= 0xcf && marker != 0xc4 && marker != 0xc8 && marker != 0xcc) return false;   // lossless, hierarchical, arithmetic
		else if (marker == 0xc4)
		{
			size_t o = 0;
			while (o + 17 <= seg_len)
			{
				const int tc = seg[o] >> 4, th = seg[o] & 15;
				if (tc > 1 || th > 3) return false;
				HuffTable& t = tc ? ac_tab[th] : dc_tab[th];
				int total = 0;
				t.bits[0] = 0;
				for (int i = 1; i <= 16; i++) { t.bits[i] = seg[o + i]; total += t.bits[i]; }
				o += 17;
				if (total > 256 || o + total > seg_len) return false;
				memcpy(t.vals, &seg[o], (size_t)total);
				o += total;
				if (!t.build()) return false;
				t.present = true;
			}
		}
		else if (marker == 0xdd) { if (seg_len < 2) return false; restart_interval = be16(seg); }
		else if (marker == 0xee) { if (seg_len >= 12 && memcmp(seg, "Adobe", 5) == 0) { adobe = true; adobe_transform = seg[11]; } }
		else if (marker == 0xda)
		{
			if (!have_sof || seg_len < 1) return false;
			Scan sc;
			sc.ns = seg[0];
			sc.progressive = progressive;
			if (sc.ns < 1 || sc.ns > n_comp || seg_len < (size_t)(1 + 2 * sc.ns + 3)) return false;
			for (int i = 0; i < sc.ns; i++)
			{
				const int cid = seg[1 + 2 * i];
				int k = -1;
				for (int j = 0; j < n_comp; j++) if (comp[j].id == cid) k = j;
				if (k < 0) return false;
				for (int j = 0; j < i; j++) if (sc.comp[j] == &comp[k]) return false;
				comp[k].td = seg[2 + 2 * i] >> 4; comp[k].ta = seg[2 + 2 * i] & 15;
				if (comp[k].td > 3 || comp[k].ta > 3) return false;
				sc.comp[i] = &comp[k];
			}
			sc.ss = seg[1 + 2 * sc.ns]; sc.se = seg[2 + 2 * sc.ns];
			sc.ah = seg[3 + 2 * sc.ns] >> 4; sc.al = seg[3 + 2 * sc.ns] & 15;
			if (!progressive) { if (sc.ss != 0 || sc.se != 63 || sc.ah != 0 || sc.al != 0) return false; }
			else
			{
				if (sc.ss > sc.se || sc.se > 63 || sc.al > 13 || (sc.ah != 0 && sc.ah != sc.al + 1)) return false;
				if (sc.ss == 0 ? sc.se != 0 : sc.ns != 1) return false;
			}
			for (int i = 0; i < sc.ns; i++)
			{
				Component& c = *sc.comp[i];
				const bool need_dc = sc.ss == 0 && sc.ah == 0, need_ac = sc.se > 0;
				if ((need_dc && !dc_tab[c.td].present) || (need_ac && !ac_tab[c.ta].present)) return false;
				if (!c.seen)
				{
					if (!quant_present[c.tq]) return false;
					memcpy(c.quant, quant[c.tq], sizeof(c.quant));
					ifast_multipliers(c.quant, c.ifast_mult);
					c.seen = true;
				}
				// successive-approximation bookkeeping: a band must be introduced before it is refined, one bit at a time
				for (int k = sc.ss; k <= sc.se; k++)
				{
					if (progressive && (sc.ah == 0 ? c.known_al[k] >= 0 : c.known_al[k] != sc.ah)) return false;
					c.known_al[k] = (int8_t)sc.al;
				}
				if (progressive && sc.ss > 0 && c.known_al[0] < 0) return false;
			}
			if (adobe && n_comp == 3 && adobe_transform == 0) return false;     // RGB-coded JPEG: not produced by the reference's assets
			pos += len;
			if (!decode_scan(f, pos, sc, dc_tab, ac_tab, mcus_x, mcus_y, restart_interval)) return false;
			saw_scan = true;
			continue;
		}
		pos += len;
	}
	if (!have_sof || !saw_scan) return false;
	// every coefficient of every component fully coded (a progressive file cut short would need libjpeg's block smoothing)
	for (int i = 0; i < n_comp; i++)
		for (int k = 0; k < 64; k++)
			if (!comp[i].seen || comp[i].known_al[k] != 0) return false;

	// ---- inverse DCT of every block
	for (int i = 0; i < n_comp; i++)
	{
		Component& c = comp[i];
		c.plane.assign((size_t)c.stride * c.rows, 0);
		for (int by = 0; by < c.pbh; by++)
			for (int bx = 0; bx < c.pbw; bx++)
			{
				const int16_t* block = c.coefs.data() + ((size_t)by * c.pbw + bx) * 64;
				uint8_t* dst = c.plane.data() + (size_t)by * 8 * c.stride + (size_t)bx * 8;
				if (fast) idct_ifast(block, c.ifast_mult, dst, c.stride);
				else idct_islow(block, c.quant, dst, c.stride);
			}
		c.coefs.clear(); c.coefs.shrink_to_fit();
	}
	// ---- upsample + colour conversion
	out.width = width; out.height = height;
	out.rgba.assign((size_t)width * height * 4, 255);
	int cr_r[256], cb_b[256], cr_g[256], cb_g[256];
	const int fix_cb_g = mode == kJpegReference ? 22553 : 22554;   // FIX(0.344136286) in libjpeg 9a, FIX(0.34414) in 6b / turbo
	for (int i = 0; i < 256; i++)
	{
		const int x = i - 128;
		cr_r[i] = (91881 * x + 32768) >> 16;        // FIX(1.402)
		cb_b[i] = (116130 * x + 32768) >> 16;       // FIX(1.772)
		cr_g[i] = -46802 * x;                       // FIX(0.71414) == FIX(0.714136286)
		cb_g[i] = -fix_cb_g * x + 32768;
	}
	std::vector<uint8_t> r0, r1, r2;
	for (int y = 0; y < height; y++)
	{
		uint8_t* dst = &out.rgba[(size_t)y * width * 4];
		upsample_row(comp[0], hmax, vmax, y, width, r0, !fast);
		if (n_comp == 1)
		{
			for (int x = 0; x < width; x++) { dst[4 * x] = dst[4 * x + 1] = dst[4 * x + 2] = r0[x]; }
			continue;
		}
		upsample_row(comp[1], hmax, vmax, y, width, r1, !fast);
		upsample_row(comp[2], hmax, vmax, y, width, r2, !fast);
		for (int x = 0; x < width; x++)
		{
			const int yy = r0[x], cb = r1[x], cr = r2[x];
			int rr = yy + cr_r[cr], gg = yy + ((cb_g[cb] + cr_g[cr]) >> 16), bb = yy + cb_b[cb];
			dst[4 * x] = (uint8_t)(rr < 0 ? 0 : (rr > 255 ? 255 : rr));
			dst[4 * x + 1] = (uint8_t)(gg < 0 ? 0 : (gg > 255 ? 255 : gg));
			dst[4 * x + 2] = (uint8_t)(bb < 0 ? 0 : (bb > 255 ? 255 : bb));
		}
	}
	return true;
}

} // namespace ptb
