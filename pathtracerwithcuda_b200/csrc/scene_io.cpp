// Scene front-end of the ptb200 render path: reads the reference's config JSON, scene JSON and
// OBJ groups UNCHANGED and produces the flat HostScene the device side uploads.
//
// Behaviour follows (file:line under /root/reference/gpu_path_tracer/):
//   config       Core/config_parser.cpp:8-124 (all 23 keys mandatory, every value a string)
//   scene        Core/scene_parser.cpp:37-442 (Background/Texture/Material/Sphere/Mesh)
//   OBJ          Core/triangle_mesh.cpp:8-213 over tinyobjloader 1.1.0 semantics
//                (lib/tiny_obj_loader/tiny_obj_loader.h:498-611 number grammar, :747-800 index
//                triples, :985-1175 ear clipping, :1660-1960 shape splitting)
//   transforms   triangle_mesh.cpp:147-170,200-202,617-647 with glm 0.9.9's operation order
//                (lib/glm/gtc/matrix_transform.inl:10-87, detail/func_matrix.inl:297-355,
//                detail/type_mat4x4.inl:487-520) so world-space vertices are bit-identical
//   materials    Core/material.cpp:12-580, scene_parser.cpp:675-708
//   images       Others/image_loader.cpp:31-95 pixel contract (RGBA8, A=255, row 0 = top)
// This file must be compiled WITHOUT floating-point contraction (-ffp-contract=off): the
// reference's host code is plain IEEE binary32 with no fused multiply-adds.
#include "scene.h"
#include "json_min.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <new>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <thread>
#include <atomic>
#include <exception>
#include <mutex>
#include <chrono>
#include <dirent.h>
#include <fcntl.h>
#include <memory>
#include <type_traits>
#include <sys/stat.h>
#include <unistd.h>
#include <cerrno>

namespace ptb
{

std::string& last_error()
{
	static thread_local std::string e;
	return e;
}

void set_error(const std::string& msg)
{
	last_error() = msg;
}

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------

static std::string normalize_separators(std::string p)
{
	for (auto& c : p) if (c == '\\') c = '/';
	return p;
}

static std::string join_path(const std::string& root, const std::string& rel)
{
	std::string r = normalize_separators(rel);
	if (root.empty() || (!r.empty() && r[0] == '/')) return r;
	if (root.back() == '/') return root + r;
	return root + "/" + r;
}

// Host threads of the loader.  An exception in a worker (bad_alloc on a huge mesh) must reach the calling thread and, through it,
// the C boundary as a load error — never std::terminate; and no thread may be left joinable when one cannot be started.
class WorkerGroup
{
public:
	~WorkerGroup() { join(); }
	template <class F> void spawn(F body)
	{
		threads_.emplace_back([this, body]() mutable { try { body(); } catch (...) { capture(); } });
	}
	template <class F> void run_here(F body) { try { body(); } catch (...) { capture(); } }
	// joins the workers; rethrows the first exception any of them (or run_here) met
	void finish()
	{
		join();
		if (error_) { std::exception_ptr e = error_; error_ = nullptr; std::rethrow_exception(e); }
	}
private:
	void join() { for (auto& t : threads_) if (t.joinable()) t.join(); threads_.clear(); }
	void capture() { std::lock_guard<std::mutex> lock(mutex_); if (!error_) error_ = std::current_exception(); }
	std::vector<std::thread> threads_;
	std::mutex mutex_;
	std::exception_ptr error_;
};

static bool read_text_file(const std::string& path, std::string& out)
{
	// one sized read (mesh files run to hundreds of MB; a stringstream copies them twice)
	FILE* f = fopen(path.c_str(), "rb");
	if (!f) return false;
	bool ok = fseek(f, 0, SEEK_END) == 0;
	const long size = ok ? ftell(f) : -1;
	ok = ok && size >= 0 && fseek(f, 0, SEEK_SET) == 0;
	if (ok)
	{
		out.resize((size_t)size);
		ok = size == 0 || fread(&out[0], 1, (size_t)size, f) == (size_t)size;
	}
	fclose(f);
	return ok;
}

// `istringstream >> float` (config_parser.cpp:189-195): leading blanks skipped, 0 on failure.
static float parse_float(const std::string& text)
{
	const char* s = text.c_str();
	char* end = nullptr;
	float v = strtof(s, &end);
	return end == s ? 0.0f : v;
}

static int parse_int(const std::string& text)
{
	const char* s = text.c_str();
	char* end = nullptr;
	long v = strtol(s, &end, 10);
	if (end == s) return 0;
	if (v > 2147483647L) v = 2147483647L;
	if (v < -2147483647L - 1) v = -2147483647L - 1;
	return (int)v;
}

static bool parse_bool(const std::string& text)
{
	return text == "true";
}

static Vec3 parse_float3(const std::string& text)
{
	float v[3] = { 0.0f, 0.0f, 0.0f };
	const char* s = text.c_str();
	for (int k = 0; k < 3; k++)
	{
		char* end = nullptr;
		float f = strtof(s, &end);
		if (end == s) break;
		v[k] = f;
		s = end;
	}
	return Vec3{ v[0], v[1], v[2] };
}

static float clampf(float f, float a, float b)
{
	return fmaxf(a, fminf(f, b));
}

// ------------------------------------------------------------------------------------------
// config
// ------------------------------------------------------------------------------------------

static bool get_string(const JValue& obj, const char* category, const char* key, std::string& out)
{
	const JValue& v = obj[key];
	if (v.is_null())
	{
		set_error(std::string("[Error]") + category + " property <" + key + "> not defined!");
		return false;
	}
	if (!v.is_string())
	{
		set_error(std::string("[Error]") + category + " property <" + key + "> must be a string");
		return false;
	}
	out = v.str;
	return true;
}

bool load_config(const std::string& path, Config& c)
{
	std::string text, err;
	if (!read_text_file(normalize_separators(path), text)) { set_error("[Error]cannot open config file " + path); return false; }
	JValue root;
	if (!JParser(text).parse(root, err)) { set_error("[Error]config parse error: " + err); return false; }
	if (!root.is_object()) { set_error("[Error]config must be a JSON object"); return false; }

	static const char* keys[23] = {
		"Height", "Width", "FullScreen", "BlockSize", "MaxBlockSize", "MaxDepth", "BiasLength", "EnergyThreshold",
		"SSSThreshold", "Skybox", "BilinearSample", "Sky", "GammaCorrection", "AntiAlias", "FOV", "BvhLeafNodeTriangleNum",
		"BvhBucketMaxDivideInternalNum", "BvhBuildBlockSize", "BvhBuildMethod", "AirRefractionIndex", "AirAbsorptionCoef",
		"AirReducedScatteringCoef", "CUDAAcceleration" };
	std::map<std::string, std::string> s;
	for (const char* k : keys)
	{
		std::string v;
		if (!get_string(root, "Config", k, v)) return false;
		s[k] = v;
	}

	memset(&c, 0, sizeof(c));
	c.height = parse_int(s["Height"]);
	c.width = parse_int(s["Width"]);
	c.use_fullscreen = parse_bool(s["FullScreen"]);
	c.block_size = parse_int(s["BlockSize"]);
	c.max_block_size = parse_int(s["MaxBlockSize"]);
	c.max_tracer_depth = parse_int(s["MaxDepth"]);
	c.vector_bias_length = parse_float(s["BiasLength"]);
	c.energy_exist_threshold = parse_float(s["EnergyThreshold"]);
	c.sss_threshold = parse_float(s["SSSThreshold"]);
	c.use_sky_box = parse_bool(s["Skybox"]);
	c.use_bilinear = parse_bool(s["BilinearSample"]);
	c.use_sky = parse_bool(s["Sky"]);
	c.gamma_correction = parse_bool(s["GammaCorrection"]);
	c.use_anti_alias = parse_bool(s["AntiAlias"]);
	// config_parser.cpp:111 parses FOV with parse_bool -> 0.0 or 1.0; rendering never reads it.
	c.fov = parse_bool(s["FOV"]) ? 1.0f : 0.0f;
	c.bvh_leaf_node_triangle_num = parse_int(s["BvhLeafNodeTriangleNum"]);
	c.bvh_bucket_max_divide_internal_num = parse_int(s["BvhBucketMaxDivideInternalNum"]);
	c.bvh_build_block_size = parse_int(s["BvhBuildBlockSize"]);
	std::string method = s["BvhBuildMethod"];
	std::transform(method.begin(), method.end(), method.begin(), [](unsigned char ch) { return (char)tolower(ch); });
	c.bvh_build = method == "mortoncodecpu" ? 1 : (method == "mortoncodecuda" ? 2 : 0);
	c.air_refraction_index = parse_float(s["AirRefractionIndex"]);
	c.air_absorption_coef = parse_float3(s["AirAbsorptionCoef"]);
	c.air_reduced_scattering_coef = parse_float3(s["AirReducedScatteringCoef"]);
	c.cuda_acceleration = parse_bool(s["CUDAAcceleration"]);

	if (c.width <= 0 || c.height <= 0) { set_error("[Error]config Width/Height must be positive"); return false; }
	// path ids (pass slot * pixels + pixel, up to 64 slots) and the per-depth counters are ints: bound what sizes them
	if ((long long)c.width * (long long)c.height > (long long)(0x7fffffff / 64)) { set_error("[Error]config Width*Height must not exceed 33554431 pixels"); return false; }
	if (c.max_tracer_depth < 0 || c.max_tracer_depth > 4096) { set_error("[Error]config MaxDepth must be in 0..4096"); return false; }
	return true;
}

// ------------------------------------------------------------------------------------------
// camera (Core/camera.cpp:3-14,56-60,80-98)
// ------------------------------------------------------------------------------------------

void default_camera(float width, float height, float aperture, float focal, ptb_camera& out)
{
	const float PI_F = 3.1415926535897f;
	float yaw = 0.0f, pitch = 0.3f, radius = 14.0f;
	float aperture_radius = 0.0f, focal_distance = radius;
	float fov_x = 45.0f;
	// set_fov: degrees_to_radians = d / 180 * PI; radians_to_degrees = r * 180 / PI (Math/basic_math.hpp:41-49)
	float half = (fov_x / 180.0f * PI_F) * 0.5f;
	float fov_y = (atanf(tanf(half) * (height / width)) * 2.0f) * 180.0f / PI_F;
	if (focal >= 0.0f) focal_distance = fminf(fmaxf(focal, 0.0f), 2.0f * radius);
	if (aperture >= 0.0f) aperture_radius = fminf(fmaxf(aperture, 0.0f), 1.0f);

	float x = sinf(yaw) * cosf(pitch);
	float y = sinf(pitch);
	float z = cosf(yaw) * cosf(pitch);
	memset(&out, 0, sizeof(out));
	out.eye[0] = 0.0f + x * radius; out.eye[1] = 0.0f + y * radius; out.eye[2] = 0.0f + z * radius;
	out.view[0] = -1.0f * x; out.view[1] = -1.0f * y; out.view[2] = -1.0f * z;
	out.up[0] = 0.0f; out.up[1] = 1.0f; out.up[2] = 0.0f;
	out.resolution[0] = width; out.resolution[1] = height;
	out.fov[0] = fov_x; out.fov[1] = fov_y;
	out.focal_distance = focal_distance;
	out.aperture_radius = aperture_radius;
}

// ------------------------------------------------------------------------------------------
// built-in materials (Core/material.cpp)
// ------------------------------------------------------------------------------------------

static ptb_material make_material(float dr, float dg, float db, float er, float eg, float eb, float sr, float sg, float sb,
	bool transparent, float roughness, float n, float k, float ar, float ag, float ab, float s)
{
	ptb_material m;
	memset(&m, 0, sizeof(m));
	m.diffuse_color[0] = dr; m.diffuse_color[1] = dg; m.diffuse_color[2] = db;
	m.emission_color[0] = er; m.emission_color[1] = eg; m.emission_color[2] = eb;
	m.specular_color[0] = sr; m.specular_color[1] = sg; m.specular_color[2] = sb;
	m.is_transparent = transparent ? 1 : 0;
	m.roughness = roughness;
	m.refraction_index = n;
	m.extinction_coefficient = k;
	m.absorption_coefficient[0] = ar; m.absorption_coefficient[1] = ag; m.absorption_coefficient[2] = ab;
	m.reduced_scattering_coefficient[0] = s; m.reduced_scattering_coefficient[1] = s; m.reduced_scattering_coefficient[2] = s;
	m.diffuse_texture_id = -1;
	m.specular_texture_id = -1;
	return m;
}

static ptb_material metal(float sr, float sg, float sb, float n, float k)
{
	return make_material(0, 0, 0, 0, 0, 0, sr, sg, sb, false, 0.3f, n, k, 0, 0, 0, 0);
}

static ptb_material opaque(float dr, float dg, float db, float spec)
{
	return make_material(dr, dg, db, 0, 0, 0, spec, spec, spec, false, 0.01f, 1.491f, 0, 0, 0, 0, 0);
}

static const std::map<std::string, ptb_material>& builtin_table()
{
	static const std::map<std::string, ptb_material> table = {
		{ "titanium", metal(0.542f, 0.497f, 0.499f, 2.2670f, 3.0385f) },
		{ "chromium", metal(0.549f, 0.556f, 0.554f, 2.3230f, 3.1350f) },
		{ "iron", metal(0.562f, 0.556f, 0.578f, 2.5845f, 2.7670f) },
		{ "nickel", metal(0.662f, 0.609f, 0.526f, 1.7290f, 2.9435f) },
		{ "platinum", metal(0.673f, 0.637f, 0.585f, 1.3400f, 1.0300f) },
		{ "copper", metal(0.955f, 0.638f, 0.538f, 1.2404f, 2.3929f) },
		{ "palladium", metal(0.733f, 0.697f, 0.652f, 1.4080f, 3.2540f) },
		{ "zinc", metal(0.664f, 0.824f, 0.850f, 0.67767f, 4.01220f) },
		{ "gold", metal(1.022f, 0.782f, 0.344f, 0.89863f, 2.4584f) },
		{ "aluminum", metal(0.913f, 0.922f, 0.924f, 0.63324f, 5.4544f) },
		{ "silver", metal(0.972f, 0.960f, 0.915f, 0.04f, 2.6484f) },
		{ "glass", make_material(1, 1, 1, 0, 0, 0, 0.045f, 0.045f, 0.045f, true, 0.1f, 1.5319f, 0, 0, 0, 0, 0) },
		{ "green_glass", make_material(1, 1, 1, 0, 0, 0, 0.045f, 0.045f, 0.045f, true, 0.1f, 1.5319f, 0, 0.8f, 0.01f, 0.8f, 0) },
		{ "diamond", make_material(1, 1, 1, 0, 0, 0, 1, 1, 1, true, 0.01f, 2.4392f, 0, 0, 0, 0, 0) },
		{ "red", opaque(0.87f, 0.15f, 0.15f, 1.0f) },
		{ "green", opaque(0.15f, 0.87f, 0.15f, 1.0f) },
		{ "orange", opaque(0.93f, 0.33f, 0.04f, 1.0f) },
		{ "purple", opaque(0.5f, 0.1f, 0.9f, 1.0f) },
		{ "blue", opaque(0.4f, 0.6f, 0.8f, 1.0f) },
		{ "wall_blue", opaque(0.4f, 0.6f, 0.8f, 0.0f) },
		{ "wall_red", opaque(0.87f, 0.15f, 0.15f, 0.0f) },
		{ "wall_green", opaque(0.15f, 0.87f, 0.15f, 0.0f) },
		{ "wall_white", opaque(1.0f, 1.0f, 1.0f, 0.0f) },
		{ "marble", make_material(0, 0, 0, 0, 0, 0, 1, 1, 1, true, 0.01f, 1.486f, 0, 0.6f, 0.6f, 0.6f, 8.0f) },
		{ "something_blue", make_material(0, 0, 0, 0, 0, 0, 1, 1, 1, true, 0.01f, 1.333f, 0, 0.9f, 0.3f, 0.02f, 2.0f) },
		{ "something_red", make_material(0, 0, 0, 0, 0, 0, 1, 1, 1, true, 0.01f, 1.35f, 0, 0.02f, 5.1f, 5.7f, 9.0f) },
		{ "light", make_material(0, 0, 0, 13.0f, 13.0f, 11.0f, 0, 0, 0, false, 0.01f, 1.000293f, 0, 0, 0, 0, 0) },
	};
	return table;
}

bool builtin_material(const std::string& name, ptb_material& out)
{
	auto it = builtin_table().find(name);
	if (it == builtin_table().end()) return false;
	out = it->second;
	return true;
}

// ------------------------------------------------------------------------------------------
// images: RGBA8, A=255, row 0 = top
// ------------------------------------------------------------------------------------------

static bool read_binary_file(const std::string& path, std::vector<uint8_t>& out)
{
	FILE* fp = fopen(path.c_str(), "rb");
	if (!fp) return false;
	fseek(fp, 0, SEEK_END);
	long n = ftell(fp);
	fseek(fp, 0, SEEK_SET);
	out.resize(n > 0 ? (size_t)n : 0);
	size_t got = n > 0 ? fread(out.data(), 1, (size_t)n, fp) : 0;
	fclose(fp);
	return got == (size_t)(n > 0 ? n : 0);
}

static uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint32_t le16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

static bool decode_bmp(const std::vector<uint8_t>& f, Texture& out)
{
	if (f.size() < 54 || f[0] != 'B' || f[1] != 'M') return false;
	uint32_t data_offset = le32(&f[10]);
	int32_t w = (int32_t)le32(&f[18]);
	int32_t h = (int32_t)le32(&f[22]);
	uint32_t bits = le16(&f[28]);
	uint32_t compression = le32(&f[30]);
	if (w <= 0 || h == 0 || h == INT32_MIN || (bits != 24 && bits != 32) || (compression != 0 && compression != 3)) return false;
	bool bottom_up = h > 0;
	uint32_t rows = (uint32_t)(h > 0 ? h : -h);
	uint32_t bpp = bits / 8;
	size_t pitch = ((size_t)w * bpp + 3) & ~(size_t)3;
	if (f.size() < data_offset + pitch * rows) return false;
	out.width = w;
	out.height = (int)rows;
	out.rgba.resize((size_t)w * rows * 4);
	for (uint32_t y = 0; y < rows; y++)
	{
		const uint8_t* src = &f[data_offset + (size_t)(bottom_up ? rows - 1 - y : y) * pitch];
		uint8_t* dst = &out.rgba[(size_t)y * w * 4];
		for (int32_t x = 0; x < w; x++)
		{
			dst[x * 4 + 0] = src[x * bpp + 2];
			dst[x * 4 + 1] = src[x * bpp + 1];
			dst[x * 4 + 2] = src[x * bpp + 0];
			dst[x * 4 + 3] = 255;
		}
	}
	return true;
}

// Uncompressed / RLE true-colour and greyscale TGA (types 2, 3, 10, 11).
static bool decode_tga(const std::vector<uint8_t>& f, Texture& out)
{
	if (f.size() < 18) return false;
	uint32_t id_len = f[0], cmap_type = f[1], type = f[2];
	uint32_t w = le16(&f[12]), h = le16(&f[14]), bits = f[16], desc = f[17];
	if (cmap_type != 0 || (type != 2 && type != 3 && type != 10 && type != 11) || w == 0 || h == 0 || (uint64_t)w * h > kMaxImagePixels) return false;
	uint32_t bpp = bits / 8;
	if (bpp != 1 && bpp != 3 && bpp != 4) return false;
	size_t pos = 18 + id_len;
	std::vector<uint8_t> px((size_t)w * h * bpp);
	if (type == 2 || type == 3)
	{
		if (f.size() < pos + px.size()) return false;
		memcpy(px.data(), &f[pos], px.size());
	}
	else
	{
		size_t o = 0;
		while (o < px.size())
		{
			if (pos >= f.size()) return false;
			uint8_t hd = f[pos++];
			uint32_t count = (hd & 0x7F) + 1;
			if (hd & 0x80)
			{
				if (pos + bpp > f.size()) return false;
				for (uint32_t k = 0; k < count && o + bpp <= px.size(); k++) { memcpy(&px[o], &f[pos], bpp); o += bpp; }
				pos += bpp;
			}
			else
			{
				size_t n = (size_t)count * bpp;
				if (pos + n > f.size() || o + n > px.size()) return false;
				memcpy(&px[o], &f[pos], n); o += n; pos += n;
			}
		}
	}
	bool top_down = (desc & 0x20) != 0;
	bool right_left = (desc & 0x10) != 0;
	out.width = (int)w; out.height = (int)h;
	out.rgba.resize((size_t)w * h * 4);
	for (uint32_t y = 0; y < h; y++)
	{
		const uint8_t* src = &px[(size_t)(top_down ? y : h - 1 - y) * w * bpp];
		uint8_t* dst = &out.rgba[(size_t)y * w * 4];
		for (uint32_t x = 0; x < w; x++)
		{
			const uint8_t* p = src + (size_t)(right_left ? w - 1 - x : x) * bpp;
			if (bpp == 1) { dst[x * 4 + 0] = dst[x * 4 + 1] = dst[x * 4 + 2] = p[0]; }
			else { dst[x * 4 + 0] = p[2]; dst[x * 4 + 1] = p[1]; dst[x * 4 + 2] = p[0]; }
			dst[x * 4 + 3] = 255;
		}
	}
	return true;
}

// "<file>.rgba8" side-car (u32 width, u32 height, RGBA8 top-down): how a caller hands over
// formats this library does not decode itself (JPG, PNG); the reference used FreeImage for all.
// ------------------------------------------------------------------------------------------
// PNG (FreeImage_Load + FreeImage_ConvertTo24Bits of Others/image_loader.cpp:31-95): own inflate + unfilter.
// Lossless, so parity with the reference's decode is exact by construction: 8-bit RGB as stored, alpha
// dropped, grey replicated, palette expanded, 16-bit samples reduced to their high byte; output alpha = 255.
// ------------------------------------------------------------------------------------------
namespace
{

struct BitReader
{
	const uint8_t* p; size_t n, pos; uint32_t buf; int cnt;
	BitReader(const uint8_t* d, size_t len) : p(d), n(len), pos(0), buf(0), cnt(0) {}
	bool need(int k) { while (cnt < k) { if (pos >= n) return false; buf |= (uint32_t)p[pos++] << cnt; cnt += 8; } return true; }
	bool bits(int k, uint32_t& v) { if (k == 0) { v = 0; return true; } if (!need(k)) return false; v = buf & ((1u << k) - 1u); buf >>= k; cnt -= k; return true; }
	void align() { buf = 0; cnt = 0; }
};

struct Huffman
{
	uint16_t count[16], symbol[288];
	bool build(const uint8_t* lengths, int n)
	{
		memset(count, 0, sizeof(count));
		for (int i = 0; i < n; i++) count[lengths[i]]++;
		uint16_t offs[16];
		offs[1] = 0;
		for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + count[l];
		for (int i = 0; i < n; i++) if (lengths[i]) symbol[offs[lengths[i]]++] = (uint16_t)i;
		int left = 1;
		for (int l = 1; l <= 15; l++) { left <<= 1; left -= count[l]; if (left < 0) return false; }
		return true;
	}
	int decode(BitReader& br) const
	{
		int code = 0, first = 0, index = 0;
		for (int l = 1; l <= 15; l++)
		{
			uint32_t b;
			if (!br.bits(1, b)) return -1;
			code |= (int)b;
			int c = count[l];
			if (code - c < first) return symbol[index + (code - first)];
			index += c; first += c; first <<= 1; code <<= 1;
		}
		return -1;
	}
};

bool inflate_zlib(const std::vector<uint8_t>& z, std::vector<uint8_t>& out, size_t expect)
{
	if (z.size() < 6 || (z[0] & 0x0f) != 8 || ((z[0] << 8) | z[1]) % 31 != 0 || (z[1] & 0x20)) return false;
	static const uint16_t len_base[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
	static const uint8_t len_extra[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
	static const uint16_t dist_base[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
	static const uint8_t dist_extra[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };
	static const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
	BitReader br(z.data() + 2, z.size() - 2);
	out.clear();
	out.reserve(expect);
	uint32_t last = 0;
	do
	{
		uint32_t type;
		if (!br.bits(1, last) || !br.bits(2, type)) return false;
		if (type == 0)
		{
			br.align();
			if (br.pos + 4 > br.n) return false;
			uint32_t len = br.p[br.pos] | (br.p[br.pos + 1] << 8), nlen = br.p[br.pos + 2] | (br.p[br.pos + 3] << 8);
			br.pos += 4;
			if ((len ^ 0xffffu) != nlen || br.pos + len > br.n) return false;
			out.insert(out.end(), br.p + br.pos, br.p + br.pos + len);
			br.pos += len;
			continue;
		}
		if (type == 3) return false;
		Huffman lit, dist;
		uint8_t lengths[320];
		if (type == 1)
		{
			int i = 0;
			for (; i < 144; i++) lengths[i] = 8;
			for (; i < 256; i++) lengths[i] = 9;
			for (; i < 280; i++) lengths[i] = 7;
			for (; i < 288; i++) lengths[i] = 8;
			lit.build(lengths, 288);
			for (i = 0; i < 30; i++) lengths[i] = 5;
			dist.build(lengths, 30);
		}
		else
		{
			uint32_t hlit, hdist, hclen;
			if (!br.bits(5, hlit) || !br.bits(5, hdist) || !br.bits(4, hclen)) return false;
			hlit += 257; hdist += 1; hclen += 4;
			if (hlit > 286 || hdist > 30) return false;
			uint8_t cl[19] = { 0 };
			for (uint32_t i = 0; i < hclen; i++) { uint32_t v; if (!br.bits(3, v)) return false; cl[order[i]] = (uint8_t)v; }
			Huffman clh;
			if (!clh.build(cl, 19)) return false;
			uint32_t idx = 0;
			while (idx < hlit + hdist)
			{
				int sym = clh.decode(br);
				if (sym < 0) return false;
				if (sym < 16) lengths[idx++] = (uint8_t)sym;
				else
				{
					uint32_t rep, prev = 0;
					if (sym == 16) { if (idx == 0) return false; prev = lengths[idx - 1]; if (!br.bits(2, rep)) return false; rep += 3; }
					else if (sym == 17) { if (!br.bits(3, rep)) return false; rep += 3; }
					else { if (!br.bits(7, rep)) return false; rep += 11; }
					if (idx + rep > hlit + hdist) return false;
					while (rep--) lengths[idx++] = (uint8_t)prev;
				}
			}
			if (lengths[256] == 0) return false;
			if (!lit.build(lengths, (int)hlit)) return false;
			dist.build(lengths + hlit, (int)hdist);   // an incomplete distance code is legal (single code)
		}
		while (true)
		{
			int sym = lit.decode(br);
			if (sym < 0) return false;
			if (sym < 256) { out.push_back((uint8_t)sym); continue; }
			if (sym == 256) break;
			sym -= 257;
			if (sym >= 29) return false;
			uint32_t eb;
			if (!br.bits(len_extra[sym], eb)) return false;
			size_t len = len_base[sym] + eb;
			int ds = dist.decode(br);
			if (ds < 0 || ds >= 30) return false;
			if (!br.bits(dist_extra[ds], eb)) return false;
			size_t d = dist_base[ds] + eb;
			if (d > out.size()) return false;
			size_t from = out.size() - d;
			for (size_t k = 0; k < len; k++) out.push_back(out[from + k]);
		}
	} while (!last);
	return true;
}

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]; }

} // namespace

static bool decode_png(const std::vector<uint8_t>& f, Texture& out)
{
	static const uint8_t sig[8] = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
	if (f.size() < 33 || memcmp(f.data(), sig, 8) != 0) return false;
	uint32_t w = 0, h = 0;
	int depth = 0, ctype = 0, interlace = 0;
	std::vector<uint8_t> idat, palette;
	size_t pos = 8;
	bool have_ihdr = false, done = false;
	while (!done && pos + 12 <= f.size())
	{
		uint32_t n = be32(&f[pos]);
		const uint8_t* type = &f[pos + 4];
		if (pos + 12 + (size_t)n > f.size()) return false;
		const uint8_t* body = &f[pos + 8];
		if (!memcmp(type, "IHDR", 4))
		{
			if (n != 13) return false;
			w = be32(body); h = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
			if (body[10] != 0 || body[11] != 0) return false;
			have_ihdr = true;
		}
		else if (!memcmp(type, "PLTE", 4)) palette.assign(body, body + n);
		else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + n);
		else if (!memcmp(type, "IEND", 4)) done = true;
		pos += 12 + (size_t)n;
	}
	if (!have_ihdr || w == 0 || h == 0 || w > 65536 || h > 65536 || (uint64_t)w * h > kMaxImagePixels || interlace > 1) return false;
	int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
	if (channels == 0) return false;
	if (!(depth == 8 || depth == 16 || (depth < 8 && (ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4)))) return false;
	if (ctype == 3 && (depth > 8 || palette.empty())) return false;
	const size_t bits_pp = (size_t)channels * depth;
	const size_t bpp = bits_pp >= 8 ? bits_pp / 8 : 1;
	// one pass for plain files, the seven Adam7 passes (x0, y0, dx, dy) for interlaced ones: each pass is a reduced image with
	// its own filter bytes; empty passes carry no data
	static const int adam7[7][4] = { { 0, 0, 8, 8 }, { 4, 0, 8, 8 }, { 0, 4, 4, 8 }, { 2, 0, 4, 4 }, { 0, 2, 2, 4 }, { 1, 0, 2, 2 }, { 0, 1, 1, 2 } };
	static const int whole[1][4] = { { 0, 0, 1, 1 } };
	const int (*passes)[4] = interlace ? adam7 : whole;
	const int n_passes = interlace ? 7 : 1;
	size_t expected = 0;
	for (int k = 0; k < n_passes; k++)
	{
		const size_t pw = (w - passes[k][0] + passes[k][2] - 1) / passes[k][2], ph = (h - passes[k][1] + passes[k][3] - 1) / passes[k][3];
		if ((uint32_t)passes[k][0] >= w || (uint32_t)passes[k][1] >= h) continue;
		expected += ((pw * bits_pp + 7) / 8 + 1) * ph;
	}
	std::vector<uint8_t> raw;
	if (!inflate_zlib(idat, raw, expected) || raw.size() < expected) return false;
	out.width = (int)w; out.height = (int)h;
	out.rgba.assign((size_t)w * h * 4, 255);
	size_t offset = 0;
	for (int k = 0; k < n_passes; k++)
	{
		if ((uint32_t)passes[k][0] >= w || (uint32_t)passes[k][1] >= h) continue;
		const uint32_t pw = (w - passes[k][0] + passes[k][2] - 1) / passes[k][2], ph = (h - passes[k][1] + passes[k][3] - 1) / passes[k][3];
		const size_t row_bytes = (pw * bits_pp + 7) / 8;
		std::vector<uint8_t> prev(row_bytes, 0);
		for (uint32_t py = 0; py < ph; py++)
		{
			uint8_t* row = &raw[offset + (row_bytes + 1) * py + 1];
			const int filter = raw[offset + (row_bytes + 1) * py];
			for (size_t i = 0; i < row_bytes; i++)
			{
				const int a = i >= bpp ? row[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
				int x = row[i];
				switch (filter)
				{
				case 0: break;
				case 1: x += a; break;
				case 2: x += b; break;
				case 3: x += (a + b) >> 1; break;
				case 4: { int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c); x += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
				default: return false;
				}
				row[i] = (uint8_t)x;
			}
			memcpy(prev.data(), row, row_bytes);
			const uint32_t y = passes[k][1] + py * passes[k][3];
			for (uint32_t px_i = 0; px_i < pw; px_i++)
			{
				const uint32_t x = passes[k][0] + px_i * passes[k][2];
				uint8_t* dst = &out.rgba[((size_t)y * w + x) * 4];
				uint8_t r, g, b;
				if (depth < 8)
				{
					const size_t bit = (size_t)px_i * depth;
					const int v = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
					if (ctype == 3)
					{
						if ((size_t)v * 3 + 2 >= palette.size()) { r = g = b = 0; }
						else { r = palette[v * 3]; g = palette[v * 3 + 1]; b = palette[v * 3 + 2]; }
					}
					else r = g = b = (uint8_t)(v * 255 / ((1 << depth) - 1));
				}
				else
				{
					const size_t step = depth / 8;
					const uint8_t* px = row + (size_t)px_i * channels * step;
					if (ctype == 3)
					{
						const int v = px[0];
						if ((size_t)v * 3 + 2 >= palette.size()) { r = g = b = 0; }
						else { r = palette[v * 3]; g = palette[v * 3 + 1]; b = palette[v * 3 + 2]; }
					}
					else if (channels <= 2) r = g = b = px[0];
					else { r = px[0]; g = px[step]; b = px[2 * step]; }
				}
				dst[0] = r; dst[1] = g; dst[2] = b;
			}
		}
		offset += (row_bytes + 1) * ph;
	}
	return true;
}

static bool decode_sidecar(const std::vector<uint8_t>& f, Texture& out)
{
	if (f.size() < 8) return false;
	uint32_t w = le32(&f[0]), h = le32(&f[4]);
	if (w == 0 || h == 0 || f.size() != 8 + (size_t)w * h * 4) return false;
	out.width = (int)w; out.height = (int)h;
	out.rgba.assign(f.begin() + 8, f.end());
	for (size_t i = 3; i < out.rgba.size(); i += 4) out.rgba[i] = 255;
	return true;
}

bool load_image_rgba8(const std::string& path_in, Texture& out)
{
	std::string path = normalize_separators(path_in);
	std::vector<uint8_t> bytes;
	if (read_binary_file(path + ".rgba8", bytes) && decode_sidecar(bytes, out)) return true;
	if (!read_binary_file(path, bytes)) { set_error("[Error]Failed to load image file " + path); return false; }
	try
	{
		if (decode_bmp(bytes, out)) return true;
		if (decode_png(bytes, out)) return true;
		if (decode_jpeg(bytes, out)) return true;
		std::string lower = path;
		std::transform(lower.begin(), lower.end(), lower.begin(), [](unsigned char ch) { return (char)tolower(ch); });
		if (lower.size() > 4 && lower.substr(lower.size() - 4) == ".tga" && decode_tga(bytes, out)) return true;
	}
	catch (const std::bad_alloc&)
	{
		// a header announcing more pixels than memory holds (corrupt or hostile file): fail like any other undecodable file
		set_error("[Error]Image too large to decode: " + path);
		return false;
	}
	set_error("[Error]Unsupported image file format (BMP, TGA, PNG and Huffman-coded JPEG are decoded natively; provide a .rgba8 side-car): " + path);
	return false;
}

// ------------------------------------------------------------------------------------------
// glm-order float32 matrix helpers (column-major m[col][row])
// ------------------------------------------------------------------------------------------

struct V4 { float x, y, z, w; };
struct M4 { V4 c[4]; };

static inline V4 v4(float x, float y, float z, float w) { return V4{ x, y, z, w }; }
static inline V4 mul(const V4& a, float s) { return V4{ a.x * s, a.y * s, a.z * s, a.w * s }; }
static inline V4 mul(const V4& a, const V4& b) { return V4{ a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w }; }
static inline V4 add(const V4& a, const V4& b) { return V4{ a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w }; }
static inline V4 sub(const V4& a, const V4& b) { return V4{ a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w }; }
static inline float& at(V4& v, int i) { return (&v.x)[i]; }
static inline float at(const V4& v, int i) { return (&v.x)[i]; }

static M4 identity()
{
	M4 m;
	m.c[0] = v4(1, 0, 0, 0); m.c[1] = v4(0, 1, 0, 0); m.c[2] = v4(0, 0, 1, 0); m.c[3] = v4(0, 0, 0, 1);
	return m;
}

// glm::rotate(m, angle, axis) — matrix_transform.inl:19-47. `axis` must be a unit basis vector here
// (normalize() of a basis vector is exact), which is all triangle_mesh.cpp:148-150 passes.
static M4 rotate(const M4& m, float angle, float ax, float ay, float az)
{
	float c = cosf(angle);
	float s = sinf(angle);
	float inv_len = 1.0f / sqrtf(ax * ax + ay * ay + az * az);
	float axis[3] = { ax * inv_len, ay * inv_len, az * inv_len };
	float temp[3] = { (1.0f - c) * axis[0], (1.0f - c) * axis[1], (1.0f - c) * axis[2] };
	float R[3][3];
	R[0][0] = c + temp[0] * axis[0];
	R[0][1] = temp[0] * axis[1] + s * axis[2];
	R[0][2] = temp[0] * axis[2] - s * axis[1];
	R[1][0] = temp[1] * axis[0] - s * axis[2];
	R[1][1] = c + temp[1] * axis[1];
	R[1][2] = temp[1] * axis[2] + s * axis[0];
	R[2][0] = temp[2] * axis[0] + s * axis[1];
	R[2][1] = temp[2] * axis[1] - s * axis[0];
	R[2][2] = c + temp[2] * axis[2];
	M4 r;
	r.c[0] = add(add(mul(m.c[0], R[0][0]), mul(m.c[1], R[0][1])), mul(m.c[2], R[0][2]));
	r.c[1] = add(add(mul(m.c[0], R[1][0]), mul(m.c[1], R[1][1])), mul(m.c[2], R[1][2]));
	r.c[2] = add(add(mul(m.c[0], R[2][0]), mul(m.c[1], R[2][1])), mul(m.c[2], R[2][2]));
	r.c[3] = m.c[3];
	return r;
}

static M4 translate(const M4& m, float x, float y, float z)
{
	M4 r = m;
	r.c[3] = add(add(add(mul(m.c[0], x), mul(m.c[1], y)), mul(m.c[2], z)), m.c[3]);
	return r;
}

static M4 scale(const M4& m, float x, float y, float z)
{
	M4 r;
	r.c[0] = mul(m.c[0], x); r.c[1] = mul(m.c[1], y); r.c[2] = mul(m.c[2], z); r.c[3] = m.c[3];
	return r;
}

static M4 transpose(const M4& m)
{
	M4 r;
	for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) at(r.c[i], j) = at(m.c[j], i);
	return r;
}

// glm::inverse(mat4) — detail/func_matrix.inl:297-355
static M4 inverse(const M4& mm)
{
	auto m = [&](int col, int row) { return at(mm.c[col], row); };
	float Coef00 = m(2, 2) * m(3, 3) - m(3, 2) * m(2, 3);
	float Coef02 = m(1, 2) * m(3, 3) - m(3, 2) * m(1, 3);
	float Coef03 = m(1, 2) * m(2, 3) - m(2, 2) * m(1, 3);
	float Coef04 = m(2, 1) * m(3, 3) - m(3, 1) * m(2, 3);
	float Coef06 = m(1, 1) * m(3, 3) - m(3, 1) * m(1, 3);
	float Coef07 = m(1, 1) * m(2, 3) - m(2, 1) * m(1, 3);
	float Coef08 = m(2, 1) * m(3, 2) - m(3, 1) * m(2, 2);
	float Coef10 = m(1, 1) * m(3, 2) - m(3, 1) * m(1, 2);
	float Coef11 = m(1, 1) * m(2, 2) - m(2, 1) * m(1, 2);
	float Coef12 = m(2, 0) * m(3, 3) - m(3, 0) * m(2, 3);
	float Coef14 = m(1, 0) * m(3, 3) - m(3, 0) * m(1, 3);
	float Coef15 = m(1, 0) * m(2, 3) - m(2, 0) * m(1, 3);
	float Coef16 = m(2, 0) * m(3, 2) - m(3, 0) * m(2, 2);
	float Coef18 = m(1, 0) * m(3, 2) - m(3, 0) * m(1, 2);
	float Coef19 = m(1, 0) * m(2, 2) - m(2, 0) * m(1, 2);
	float Coef20 = m(2, 0) * m(3, 1) - m(3, 0) * m(2, 1);
	float Coef22 = m(1, 0) * m(3, 1) - m(3, 0) * m(1, 1);
	float Coef23 = m(1, 0) * m(2, 1) - m(2, 0) * m(1, 1);

	V4 Fac0 = v4(Coef00, Coef00, Coef02, Coef03);
	V4 Fac1 = v4(Coef04, Coef04, Coef06, Coef07);
	V4 Fac2 = v4(Coef08, Coef08, Coef10, Coef11);
	V4 Fac3 = v4(Coef12, Coef12, Coef14, Coef15);
	V4 Fac4 = v4(Coef16, Coef16, Coef18, Coef19);
	V4 Fac5 = v4(Coef20, Coef20, Coef22, Coef23);

	V4 Vec0 = v4(m(1, 0), m(0, 0), m(0, 0), m(0, 0));
	V4 Vec1 = v4(m(1, 1), m(0, 1), m(0, 1), m(0, 1));
	V4 Vec2 = v4(m(1, 2), m(0, 2), m(0, 2), m(0, 2));
	V4 Vec3_ = v4(m(1, 3), m(0, 3), m(0, 3), m(0, 3));

	V4 Inv0 = add(sub(mul(Vec1, Fac0), mul(Vec2, Fac1)), mul(Vec3_, Fac2));
	V4 Inv1 = add(sub(mul(Vec0, Fac0), mul(Vec2, Fac3)), mul(Vec3_, Fac4));
	V4 Inv2 = add(sub(mul(Vec0, Fac1), mul(Vec1, Fac3)), mul(Vec3_, Fac5));
	V4 Inv3 = add(sub(mul(Vec0, Fac2), mul(Vec1, Fac4)), mul(Vec2, Fac5));

	V4 SignA = v4(+1, -1, +1, -1);
	V4 SignB = v4(-1, +1, -1, +1);
	M4 Inverse;
	Inverse.c[0] = mul(Inv0, SignA); Inverse.c[1] = mul(Inv1, SignB); Inverse.c[2] = mul(Inv2, SignA); Inverse.c[3] = mul(Inv3, SignB);

	V4 Row0 = v4(Inverse.c[0].x, Inverse.c[1].x, Inverse.c[2].x, Inverse.c[3].x);
	V4 Dot0 = mul(mm.c[0], Row0);
	float Dot1 = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);
	float OneOverDeterminant = 1.0f / Dot1;
	M4 r;
	for (int i = 0; i < 4; i++) r.c[i] = mul(Inverse.c[i], OneOverDeterminant);
	return r;
}

// mat4 * vec4 — detail/type_mat4x4.inl:507-518: (m0*v0 + m1*v1) + (m2*v2 + m3*v3)
static V4 transform(const M4& m, const V4& v)
{
	V4 Add0 = add(mul(m.c[0], v.x), mul(m.c[1], v.y));
	V4 Add1 = add(mul(m.c[2], v.z), mul(m.c[3], v.w));
	return add(Add0, Add1);
}

// The reference normalises normals on the host with Math/cuda_math.hpp:1458-1462
// (v * rsqrtf(dot(v,v))). Under nvcc's host pass rsqrtf is CUDA's host fallback,
// (float)(1.0 / sqrt((double)x)); the parity baseline (oracle/_ref) is built that way.
static Vec3 normalize_host(float x, float y, float z)
{
	float d = x * x + y * y + z * z;
	float inv = (float)(1.0 / sqrt((double)d));
	return Vec3{ x * inv, y * inv, z * inv };
}

// ------------------------------------------------------------------------------------------
// OBJ (tinyobjloader 1.1.0 semantics)
// ------------------------------------------------------------------------------------------

namespace
{

struct ObjIndex { int v = -1, vt = -1, vn = -1; };

struct ObjShape
{
	std::vector<ObjIndex> indices; // 3 per triangle
};

struct ObjData
{
	std::vector<float> v, vn, vt;
	std::vector<ObjShape> shapes;
};

inline bool is_space(char c) { return c == ' ' || c == '\t'; }
inline bool is_digit(char c) { return (unsigned)(c - '0') < 10u; }
inline bool is_new_line(char c) { return c == '\r' || c == '\n' || c == '\0'; }

// tiny_obj_loader.h:498-611 — hand-rolled decimal grammar accumulated in double.
bool try_parse_double(const char* s, const char* s_end, double* result)
{
	if (s >= s_end) return false;
	double mantissa = 0.0;
	int exponent = 0;
	char sign = '+', exp_sign = '+';
	const char* curr = s;
	int read = 0;
	bool end_not_reached = false;

	if (*curr == '+' || *curr == '-') { sign = *curr; curr++; }
	else if (!is_digit(*curr)) return false;

	end_not_reached = (curr != s_end);
	while (end_not_reached && is_digit(*curr))
	{
		mantissa *= 10;
		mantissa += (int)(*curr - 0x30);
		curr++; read++;
		end_not_reached = (curr != s_end);
	}
	if (read == 0) return false;
	bool assemble = !end_not_reached;

	if (!assemble)
	{
		if (*curr == '.')
		{
			curr++;
			read = 1;
			end_not_reached = (curr != s_end);
			while (end_not_reached && is_digit(*curr))
			{
				static const double pow_lut[] = { 1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001 };
				const int lut_entries = 8;
				mantissa += (int)(*curr - 0x30) * (read < lut_entries ? pow_lut[read] : std::pow(10.0, -read));
				read++; curr++;
				end_not_reached = (curr != s_end);
			}
		}
		else if (*curr == 'e' || *curr == 'E') { }
		else assemble = true;
	}

	if (!assemble && end_not_reached && (*curr == 'e' || *curr == 'E'))
	{
		curr++;
		end_not_reached = (curr != s_end);
		if (end_not_reached && (*curr == '+' || *curr == '-')) { exp_sign = *curr; curr++; }
		else if (is_digit(*curr)) { }
		else return false;
		read = 0;
		end_not_reached = (curr != s_end);
		while (end_not_reached && is_digit(*curr))
		{
			exponent *= 10;
			exponent += (int)(*curr - 0x30);
			curr++; read++;
			end_not_reached = (curr != s_end);
		}
		exponent *= (exp_sign == '+' ? 1 : -1);
		if (read == 0) return false;
	}

	*result = (sign == '+' ? 1 : -1) * (exponent ? std::ldexp(mantissa * std::pow(5.0, exponent), exponent) : mantissa);
	return true;
}

// strspn(s, " \t") / strcspn(s, " \t\r") / strcspn(s, "/ \t\r") / atoi as inline loops: the libc calls were a third of the time the OBJ
// slices take to parse (a dozen of them per face line).  Same results on every input: the sets are spelled out, and the integer
// conversion is strtol's (leading white space, sign, digits, saturation) followed by atoi's cast.
inline const char* skip_blanks(const char* s) { while (*s == ' ' || *s == '\t') s++; return s; }
inline const char* skip_blanks_cr(const char* s) { while (*s == ' ' || *s == '\t' || *s == '\r') s++; return s; }
inline const char* token_end(const char* s) { while (*s != '\0' && *s != ' ' && *s != '\t' && *s != '\r') s++; return s; }
inline const char* index_end(const char* s) { while (*s != '\0' && *s != '/' && *s != ' ' && *s != '\t' && *s != '\r') s++; return s; }
inline int parse_int_like_atoi(const char* s)
{
	while (*s == ' ' || (*s >= '\t' && *s <= '\r')) s++;      // isspace in the C locale
	bool negative = false;
	if (*s == '+' || *s == '-') { negative = *s == '-'; s++; }
	unsigned long long acc = 0;
	const unsigned long long limit = negative ? 9223372036854775808ull : 9223372036854775807ull;
	bool saturated = false;
	for (; is_digit(*s); s++)
	{
		const unsigned d = (unsigned)(*s - '0');
		if (saturated || acc > (limit - d) / 10) { saturated = true; continue; }
		acc = acc * 10 + d;
	}
	if (saturated) acc = limit;
	const long long v = negative ? (long long)(0ull - acc) : (long long)acc;
	return (int)v;
}

float parse_real(const char** token, double default_value = 0.0)
{
	(*token) = skip_blanks(*token);
	const char* end = token_end(*token);
	double val = default_value;
	try_parse_double((*token), end, &val);
	(*token) = end;
	return (float)val;
}

bool fix_index(int idx, int n, int* ret)
{
	if (idx > 0) { *ret = idx - 1; return true; }
	if (idx == 0) return false;
	*ret = n + idx;
	return true;
}

// tiny_obj_loader.h:747-800
bool parse_triple(const char** token, int vsize, int vnsize, int vtsize, ObjIndex* ret)
{
	ObjIndex vi;
	if (!fix_index(parse_int_like_atoi(*token), vsize, &vi.v)) return false;
	(*token) = index_end(*token);
	if ((*token)[0] != '/') { *ret = vi; return true; }
	(*token)++;
	if ((*token)[0] == '/')
	{
		(*token)++;
		if (!fix_index(parse_int_like_atoi(*token), vnsize, &vi.vn)) return false;
		(*token) = index_end(*token);
		*ret = vi;
		return true;
	}
	if (!fix_index(parse_int_like_atoi(*token), vtsize, &vi.vt)) return false;
	(*token) = index_end(*token);
	if ((*token)[0] != '/') { *ret = vi; return true; }
	(*token)++;
	if (!fix_index(parse_int_like_atoi(*token), vnsize, &vi.vn)) return false;
	(*token) = index_end(*token);
	*ret = vi;
	return true;
}

int point_in_polygon(int nvert, const float* vertx, const float* verty, float testx, float testy)
{
	int c = 0;
	for (int i = 0, j = nvert - 1; i < nvert; j = i++)
	{
		if (((verty[i] > testy) != (verty[j] > testy)) &&
			(testx < (vertx[j] - vertx[i]) * (testy - verty[i]) / (verty[j] - verty[i]) + vertx[i]))
			c = !c;
	}
	return c;
}

// tiny_obj_loader.h:985-1175 with triangulate=true: project on the dominant plane of the first
// non-degenerate corner, then clip ears; a triangle passes through untouched.
bool emit_face(ObjShape& shape, const ObjIndex* face_begin, size_t face_size, const std::vector<float>& v, size_t n_vertices)
{
	// tinyobj triangulates with the vertices read so far (n_vertices of them) and does not check the indices; a face pointing
	// outside them (corrupt file, forward reference) is an error here instead of a wild read
	for (size_t k = 0; k < face_size; k++)
		if (face_begin[k].v < 0 || (size_t)face_begin[k].v >= n_vertices) { set_error("[Error]OBJ face references a vertex that is not defined"); return false; }
	if (face_size == 3)
	{
		// a triangle passes through the ear clipper untouched (nothing below has a side effect for three vertices)
		shape.indices.push_back(face_begin[0]); shape.indices.push_back(face_begin[1]); shape.indices.push_back(face_begin[2]);
		return true;
	}
	const std::vector<ObjIndex> face(face_begin, face_begin + face_size);
	size_t npolys = face.size();
	size_t axes[2] = { 1, 2 };
	for (size_t k = 0; k < npolys; ++k)
	{
		size_t vi0 = (size_t)face[(k + 0) % npolys].v, vi1 = (size_t)face[(k + 1) % npolys].v, vi2 = (size_t)face[(k + 2) % npolys].v;
		float e0x = v[vi1 * 3 + 0] - v[vi0 * 3 + 0], e0y = v[vi1 * 3 + 1] - v[vi0 * 3 + 1], e0z = v[vi1 * 3 + 2] - v[vi0 * 3 + 2];
		float e1x = v[vi2 * 3 + 0] - v[vi1 * 3 + 0], e1y = v[vi2 * 3 + 1] - v[vi1 * 3 + 1], e1z = v[vi2 * 3 + 2] - v[vi1 * 3 + 2];
		float cx = (float)fabs(e0y * e1z - e0z * e1y);
		float cy = (float)fabs(e0z * e1x - e0x * e1z);
		float cz = (float)fabs(e0x * e1y - e0y * e1x);
		const float epsilon = 0.0001f;
		if (cx > epsilon || cy > epsilon || cz > epsilon)
		{
			if (cx > cy && cx > cz) { }
			else
			{
				axes[0] = 0;
				if (cz > cx && cz > cy) axes[1] = 1;
			}
			break;
		}
	}

	float area = 0;
	for (size_t k = 0; k < npolys; ++k)
	{
		size_t vi0 = (size_t)face[(k + 0) % npolys].v, vi1 = (size_t)face[(k + 1) % npolys].v;
		float v0x = v[vi0 * 3 + axes[0]], v0y = v[vi0 * 3 + axes[1]];
		float v1x = v[vi1 * 3 + axes[0]], v1y = v[vi1 * 3 + axes[1]];
		area += (v0x * v1y - v0y * v1x) * 0.5f;
	}

	int max_rounds = 10;
	std::vector<ObjIndex> remaining = face;
	size_t guess_vert = 0;
	ObjIndex ind[3];
	float vx[3], vy[3];
	while (remaining.size() > 3 && max_rounds > 0)
	{
		npolys = remaining.size();
		if (guess_vert >= npolys) { max_rounds -= 1; guess_vert -= npolys; }
		for (size_t k = 0; k < 3; k++)
		{
			ind[k] = remaining[(guess_vert + k) % npolys];
			size_t vi = (size_t)ind[k].v;
			vx[k] = v[vi * 3 + axes[0]];
			vy[k] = v[vi * 3 + axes[1]];
		}
		float e0x = vx[1] - vx[0], e0y = vy[1] - vy[0];
		float e1x = vx[2] - vx[1], e1y = vy[2] - vy[1];
		float cross = e0x * e1y - e0y * e1x;
		if (cross * area < 0.0f) { guess_vert += 1; continue; }

		bool overlap = false;
		for (size_t other = 3; other < npolys; ++other)
		{
			size_t ovi = (size_t)remaining[(guess_vert + other) % npolys].v;
			if (point_in_polygon(3, vx, vy, v[ovi * 3 + axes[0]], v[ovi * 3 + axes[1]])) { overlap = true; break; }
		}
		if (overlap) { guess_vert += 1; continue; }

		shape.indices.push_back(ind[0]); shape.indices.push_back(ind[1]); shape.indices.push_back(ind[2]);

		size_t removed = (guess_vert + 1) % npolys;
		while (removed + 1 < npolys) { remaining[removed] = remaining[removed + 1]; removed += 1; }
		remaining.pop_back();
	}
	if (remaining.size() == 3)
	{
		shape.indices.push_back(remaining[0]); shape.indices.push_back(remaining[1]); shape.indices.push_back(remaining[2]);
	}
	return true;
}

#ifdef PTB_LOAD_TRACE   // phase timings of the loader on stderr (diagnostic builds only: tools/load_time.py notes)
struct LoadTrace
{
	std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
	void mark(const char* what)
	{
		auto n = std::chrono::steady_clock::now();
		fprintf(stderr, "[load] %-28s %.3f s\n", what, std::chrono::duration<double>(n - t).count());
		t = n;
	}
};
#define PTB_TRACE_BEGIN LoadTrace load_trace
#define PTB_TRACE(what) load_trace.mark(what)
#else
#define PTB_TRACE_BEGIN
#define PTB_TRACE(what)
#endif

int g_loader_threads = 0;   // 0: by file size and host cores (set_loader_threads / option "loader_threads")
int g_loader_mesh_lanes = 0;   // mesh files parsed at the same time; 0: host cores / threads per file (option "loader_mesh_lanes")
int g_loader_per_vertex = 1;   // append_mesh transforms every position / normal once instead of once per triangle corner (option "loader_per_vertex"; same output)

// One slice of the file (whole lines), parsed by one host thread.  Negative (relative) face indices need the number of
// v / vn / vt lines before the face, so a first pass counts them per slice and a prefix sum gives every slice its base.
struct ObjMark
{
	int kind;              // 1 `g`, 2 `o`, 3 parse error (message in the slice)
	size_t face_index;     // faces of this slice recorded before the line
	size_t n_vertices;     // vertices read before the line: what tinyobj's triangulation sees when the line flushes the group
};

struct ObjSlice
{
	size_t begin = 0, end = 0;
	size_t count_v = 0, count_vn = 0, count_vt = 0;     // pass 1
	size_t base_v = 0, base_vn = 0, base_vt = 0;
	std::vector<float> v, vn, vt;
	std::vector<ObjIndex> idx;
	std::vector<size_t> face_start;                     // face f = idx[face_start[f] .. face_start[f + 1]) (idx.size() after the last)
	std::vector<ObjMark> marks;
	std::string error;
};

// safeGetline: lines end at \n, \r\n or a lone \r.  A line is parsed where it lies: its terminator becomes a NUL (an embedded NUL
// simply ends the line early, as it did for a copied string's c_str()).  Returns the start of the line's first token or nullptr.
// The OBJ text: one uninitialised allocation of file size + 1 (terminating NUL), filled by pread from a few host threads — a 190 MB
// mesh file otherwise spends more time in the zero-fill of a std::string and its page faults than in the read itself.
struct TextBuffer
{
	std::unique_ptr<char[]> buf;
	size_t size = 0;      // bytes of the file; buf[size] = NUL
	char* data() { return buf.get(); }
	const char* data() const { return buf.get(); }
	char& operator[](size_t i) { return buf[i]; }
	const char& operator[](size_t i) const { return buf[i]; }
};

bool read_text_buffer(const std::string& path, TextBuffer& out)
{
	const int fd = open(path.c_str(), O_RDONLY);
	if (fd < 0) return false;
	struct stat st;
	if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size < 0) { close(fd); return false; }
	const size_t size = (size_t)st.st_size;
	try { out.buf.reset(new char[size + 1]); }
	catch (...) { close(fd); throw; }      // out of memory is reported as such at the C boundary, not as an unreadable file
	out.size = size;
	out.buf[size] = '\0';
	const size_t hw = std::max(1u, std::thread::hardware_concurrency());
	const size_t n_threads = size < (16u << 20) ? 1 : std::min<size_t>(8, hw);
	std::atomic<bool> ok(true);
	auto read_range = [&](size_t begin, size_t end)
	{
		while (begin < end)
		{
			const ssize_t got = pread(fd, out.buf.get() + begin, end - begin, (off_t)begin);
			if (got <= 0) { if (got < 0 && errno == EINTR) continue; ok = false; return; }     // 0: the file shrank under us
			begin += (size_t)got;
		}
	};
	{
		WorkerGroup group;
		for (size_t k = 1; k < n_threads; k++)
		{
			const size_t b = size * k / n_threads, e = size * (k + 1) / n_threads;
			group.spawn([&read_range, b, e] { read_range(b, e); });
		}
		group.run_here([&] { read_range(0, size / n_threads); });
		try { group.finish(); } catch (...) { close(fd); throw; }
	}
	close(fd);
	return ok;
}

inline const char* next_line(TextBuffer& text, size_t& pos, size_t end)
{
	size_t e = pos;
	while (e < end && text[e] != '\n' && text[e] != '\r') e++;
	const size_t line_begin = pos, line_len = e - pos;
	const bool crlf = e < end && text[e] == '\r' && e + 1 < end && text[e + 1] == '\n';
	text[e] = '\0';
	pos = e + (crlf ? 2 : 1);
	if (line_len == 0) return nullptr;
	const char* token = skip_blanks(text.data() + line_begin);
	if (token[0] == '\0' || token[0] == '#') return nullptr;
	return token;
}

void count_slice(const TextBuffer& text, ObjSlice& sl)
{
	// the same classification as parse_slice, without touching the text
	size_t pos = sl.begin;
	while (pos < sl.end)
	{
		size_t e = pos;
		while (e < sl.end && text[e] != '\n' && text[e] != '\r') e++;
		size_t t = pos;
		while (t < e && (text[t] == ' ' || text[t] == '\t')) t++;
		if (t + 1 < e + 1 && text[t] == 'v' && t < e)
		{
			const char c1 = t + 1 < e ? text[t + 1] : '\0', c2 = t + 2 < e ? text[t + 2] : '\0';
			if (is_space(c1)) sl.count_v++;
			else if (c1 == 'n' && is_space(c2)) sl.count_vn++;
			else if (c1 == 't' && is_space(c2)) sl.count_vt++;
		}
		const bool crlf = e < sl.end && text[e] == '\r' && e + 1 < sl.end && text[e + 1] == '\n';
		pos = e + (crlf ? 2 : 1);
	}
}

void parse_slice(TextBuffer& text, ObjSlice& sl, const std::string& path)
{
	size_t pos = sl.begin;
	sl.v.reserve(sl.count_v * 3); sl.vn.reserve(sl.count_vn * 3); sl.vt.reserve(sl.count_vt * 2);
	while (pos < sl.end)
	{
		const char* token = next_line(text, pos, sl.end);
		if (!token) continue;
		if (token[0] == 'v' && is_space(token[1]))
		{
			token += 2;
			float x = parse_real(&token), y = parse_real(&token), z = parse_real(&token);
			sl.v.push_back(x); sl.v.push_back(y); sl.v.push_back(z);
			continue;
		}
		if (token[0] == 'v' && token[1] == 'n' && is_space(token[2]))
		{
			token += 3;
			float x = parse_real(&token), y = parse_real(&token), z = parse_real(&token);
			sl.vn.push_back(x); sl.vn.push_back(y); sl.vn.push_back(z);
			continue;
		}
		if (token[0] == 'v' && token[1] == 't' && is_space(token[2]))
		{
			token += 3;
			float x = parse_real(&token), y = parse_real(&token);
			sl.vt.push_back(x); sl.vt.push_back(y);
			continue;
		}
		const size_t n_vertices = sl.base_v + sl.v.size() / 3;
		if (token[0] == 'f' && is_space(token[1]))
		{
			token += 2;
			token = skip_blanks(token);
			const size_t face_begin = sl.idx.size();
			while (!is_new_line(token[0]))
			{
				ObjIndex vi;
				if (!parse_triple(&token, (int)n_vertices, (int)(sl.base_vn + sl.vn.size() / 3), (int)(sl.base_vt + sl.vt.size() / 2), &vi))
				{
					sl.error = "[TinyObj]Failed parse `f' line(e.g. zero value for face index).";
					sl.idx.resize(face_begin);     // the last recorded face ends where the failed one began
					sl.marks.push_back(ObjMark{ 3, sl.face_start.size(), n_vertices });
					return;
				}
				sl.idx.push_back(vi);
				token = skip_blanks_cr(token);
			}
			const size_t face_size = sl.idx.size() - face_begin;
			if (face_size >= 3) sl.face_start.push_back(face_begin);
			else if (face_size != 0)
			{
				sl.error = "[Error]" + path + " has a face with fewer than 3 vertices";
				sl.idx.resize(face_begin);
				sl.marks.push_back(ObjMark{ 3, sl.face_start.size(), n_vertices });
				return;
			}
			continue;
		}
		// `usemtl`: no .mtl files ship, every name maps to material id -1, so the per-face material
		// never changes and the statement is a no-op (tiny_obj_loader.h:1779-1803).
		if (token[0] == 'g' && is_space(token[1])) { sl.marks.push_back(ObjMark{ 1, sl.face_start.size(), n_vertices }); continue; }
		if (token[0] == 'o' && is_space(token[1])) { sl.marks.push_back(ObjMark{ 2, sl.face_start.size(), n_vertices }); continue; }
	}
}

bool parse_obj(const std::string& path, ObjData& out)
{
	PTB_TRACE_BEGIN;
	TextBuffer text;                  // NUL-terminated: the last line is terminated like every other (lines are parsed in place)
	if (!read_text_buffer(path, text)) { set_error("[Info]Load file " + path + " failed: Cannot open file"); return false; }
	const size_t text_size = text.size;
	PTB_TRACE("obj read");

	// slices of whole lines, cut after a '\n' (a file of lone-'\r' line ends stays one slice)
	const size_t hw = std::max(1u, std::thread::hardware_concurrency());
	const size_t want = g_loader_threads > 0 ? (size_t)g_loader_threads : text_size < (4u << 20) ? 1 : std::min<size_t>(16, hw);
	std::vector<ObjSlice> slices;
	size_t cut = 0;
	for (size_t k = 1; k <= want && cut < text_size; k++)
	{
		size_t end = k == want ? text_size : text_size * k / want;
		if (end < cut) end = cut;
		while (end < text_size && text[end > 0 ? end - 1 : 0] != '\n') end++;
		if (end <= cut) continue;
		ObjSlice sl;
		sl.begin = cut; sl.end = end;
		slices.push_back(std::move(sl));
		cut = end;
	}
	auto run = [&](auto&& fn)
	{
		WorkerGroup group;
		for (size_t k = 1; k < slices.size(); k++) group.spawn([&, k] { fn(slices[k]); });
		if (!slices.empty()) group.run_here([&] { fn(slices[0]); });
		group.finish();
	};
	run([&](ObjSlice& sl) { count_slice(text, sl); });
	PTB_TRACE("obj count");
	for (size_t k = 1; k < slices.size(); k++)
	{
		slices[k].base_v = slices[k - 1].base_v + slices[k - 1].count_v;
		slices[k].base_vn = slices[k - 1].base_vn + slices[k - 1].count_vn;
		slices[k].base_vt = slices[k - 1].base_vt + slices[k - 1].count_vt;
	}
	run([&](ObjSlice& sl) { parse_slice(text, sl, path); });
	PTB_TRACE("obj parse slices");

	// a slice that stopped at a parse error read fewer v / vn / vt lines than counted; everything after it is discarded anyway
	for (const ObjSlice& sl : slices)
	{
		out.v.insert(out.v.end(), sl.v.begin(), sl.v.end());
		out.vn.insert(out.vn.end(), sl.vn.begin(), sl.vn.end());
		out.vt.insert(out.vt.end(), sl.vt.begin(), sl.vt.end());
		if (!sl.error.empty()) break;
	}
	PTB_TRACE("obj merge vertices");
	// Groups, shapes and triangulation in file order.  A group is flushed by the next `g` / `o` line or the end of the file and
	// is triangulated with the vertices read up to there; its faces may lie in several slices: one run per slice.  The runs are
	// triangulated by host threads (each face is independent), then appended to their shapes in order.
	// direct: every face of the run is a triangle, so its output IS the slice's index range [d0, d1) (checked, not copied)
	struct Run { size_t slice, f0, f1, n_vertices; ObjShape tris; bool failed = false; std::string error; bool direct = false; size_t d0 = 0, d1 = 0; };
	struct Group { size_t run0 = 0, run1 = 0; int kind = 0; };     // kind: the mark that flushed it (1 `g`, 2 `o`), 0 = end of file
	std::vector<Run> runs;
	std::vector<Group> groups;
	const std::string* parse_error = nullptr;     // a slice that stopped at a parse error ends the file there (the open group is never flushed)
	{
		Group g;
		size_t n_vertices = 0;
		for (size_t k = 0; k < slices.size() && !parse_error; k++)
		{
			const ObjSlice& sl = slices[k];
			size_t f_cur = 0;
			for (const ObjMark& m : sl.marks)
			{
				if (m.face_index > f_cur) { Run r; r.slice = k; r.f0 = f_cur; r.f1 = m.face_index; r.n_vertices = 0; runs.push_back(std::move(r)); f_cur = m.face_index; }
				if (m.kind == 3) { parse_error = &sl.error; break; }
				g.run1 = runs.size(); g.kind = m.kind;
				for (size_t r = g.run0; r < g.run1; r++) runs[r].n_vertices = m.n_vertices;
				groups.push_back(g);
				g = Group(); g.run0 = g.run1 = runs.size();
			}
			if (parse_error) break;
			if (sl.face_start.size() > f_cur) { Run r; r.slice = k; r.f0 = f_cur; r.f1 = sl.face_start.size(); r.n_vertices = 0; runs.push_back(std::move(r)); }
			n_vertices = sl.base_v + sl.v.size() / 3;
		}
		if (!parse_error)
		{
			g.run1 = runs.size(); g.kind = 0;
			for (size_t r = g.run0; r < g.run1; r++) runs[r].n_vertices = n_vertices;
			groups.push_back(g);
		}
	}
	const size_t flushed_runs = groups.empty() ? 0 : groups.back().run1;
	{
		std::atomic<size_t> next(0);
		auto work = [&]
		{
			for (size_t i = next++; i < flushed_runs; i = next++)
			{
				Run& r = runs[i];
				const ObjSlice& sl = slices[r.slice];
				{
					// faces have >= 3 corners each: a run with exactly 3 per face on average holds triangles only, which the ear clipper
					// passes through untouched — nothing to compute, only the vertex references to check
					const size_t b0 = sl.face_start[r.f0], e1 = r.f1 < sl.face_start.size() ? sl.face_start[r.f1] : sl.idx.size();
					if (e1 - b0 == 3 * (r.f1 - r.f0))
					{
						bool ok = true;
						for (size_t k = b0; k < e1 && ok; k++) ok = sl.idx[k].v >= 0 && (size_t)sl.idx[k].v < r.n_vertices;
						if (!ok) { r.failed = true; r.error = "[Error]OBJ face references a vertex that is not defined"; }
						r.direct = true; r.d0 = b0; r.d1 = e1;
						continue;
					}
				}
				r.tris.indices.reserve((sl.face_start[r.f1 - 1] - sl.face_start[r.f0]) + 3);     // exact for triangles; polygons grow it
				for (size_t f = r.f0; f < r.f1; f++)
				{
					const size_t b = sl.face_start[f], e = f + 1 < sl.face_start.size() ? sl.face_start[f + 1] : sl.idx.size();
					if (!emit_face(r.tris, sl.idx.data() + b, e - b, out.v, r.n_vertices)) { r.failed = true; r.error = last_error(); break; }
				}
			}
		};
		WorkerGroup group;
		for (size_t k = 1; k < slices.size() && k < flushed_runs; k++) group.spawn([&work] { work(); });
		group.run_here([&work] { work(); });
		group.finish();
	}
	ObjShape shape;
	for (const Group& g : groups)
	{
		size_t group_indices = 0;
		for (size_t r = g.run0; r < g.run1; r++) group_indices += runs[r].direct ? runs[r].d1 - runs[r].d0 : runs[r].tris.indices.size();
		for (size_t r = g.run0; r < g.run1; r++)
		{
			if (runs[r].failed) { set_error(runs[r].error); return false; }
			if (runs[r].direct)
			{
				const std::vector<ObjIndex>& idx = slices[runs[r].slice].idx;
				if (shape.indices.empty()) shape.indices.reserve(group_indices);
				shape.indices.insert(shape.indices.end(), idx.begin() + (ptrdiff_t)runs[r].d0, idx.begin() + (ptrdiff_t)runs[r].d1);
				continue;
			}
			if (shape.indices.empty()) shape.indices = std::move(runs[r].tris.indices);
			else shape.indices.insert(shape.indices.end(), runs[r].tris.indices.begin(), runs[r].tris.indices.end());
			runs[r].tris = ObjShape();
		}
		const bool had_faces = g.run1 > g.run0;
		if (g.kind == 1) { if (!shape.indices.empty()) out.shapes.push_back(std::move(shape)); shape = ObjShape(); }
		else if (g.kind == 2) { if (had_faces) out.shapes.push_back(std::move(shape)); shape = ObjShape(); }
		else if (had_faces || !shape.indices.empty()) out.shapes.push_back(std::move(shape));
	}
	if (parse_error) { set_error(*parse_error); return false; }
	PTB_TRACE("obj groups + triangulation");
	return true;
}

} // namespace

// triangle_mesh::load_obj + create_mesh_device_data for one mesh (triangle_mesh.cpp:8-213,558-655)
static bool mesh_in_range(const HostScene& scene, const MeshInfo& m);

// A mesh between its two load steps: the OBJ text parsed (read_mesh), triangles not yet appended to the scene (append_mesh).
// load_scene reads every mesh first so the scene's triangle arrays are allocated once, at their final size.
struct PendingMesh
{
	ObjData obj;
	bool ok = false;
	std::string error;       // of the failed step; reported when the meshes before this one have been appended
	size_t triangle_count() const { size_t n = 0; for (const ObjShape& sh : obj.shapes) n += sh.indices.size() / 3; return n; }
};

static void read_mesh(const std::string& path, size_t mat_num, PendingMesh& out)
{
	out.ok = false;
	if (mat_num == 0) { out.error = "[Error]Mesh has no material"; return; }
	PTB_TRACE_BEGIN;
	if (!parse_obj(path, out.obj)) { out.error = last_error(); return; }
	PTB_TRACE("parse_obj total");
	out.ok = true;
}

static bool append_mesh(HostScene& scene, const std::string& path, const Vec3& position, const Vec3& scale_v, const Vec3& rotate_v,
	std::vector<ptb_material> mats, const PendingMesh& pending)
{
	if (!pending.ok) { set_error(pending.error); return false; }
	const ObjData& obj = pending.obj;
	PTB_TRACE_BEGIN;
	if (obj.vn.empty()) { set_error("[Error]Mesh does not have normal! (" + path + ")"); return false; }
	const int mat_num = (int)mats.size();
	const bool has_uv = !obj.vt.empty();

	const float deg2rad = (float)0.01745329251994329576923690768489;
	M4 rot = identity();
	rot = rotate(rot, rotate_v.z * deg2rad, 0.0f, 0.0f, 1.0f);
	rot = rotate(rot, rotate_v.y * deg2rad, 0.0f, 1.0f, 0.0f);
	rot = rotate(rot, rotate_v.x * deg2rad, 1.0f, 0.0f, 0.0f);
	M4 rot_it = transpose(inverse(rot));

	M4 xf = identity();
	xf = translate(xf, position.x, position.y, position.z);
	xf = scale(xf, scale_v.x, scale_v.y, scale_v.z);
	M4 xf_it = transpose(inverse(xf));

	const int material_base = (int)scene.materials.size();
	const int triangle_base = (int)scene.triangles.size();
	int mesh_triangles = 0;
	const size_t nv = obj.v.size() / 3, nn = obj.vn.size() / 3, nt = obj.vt.size() / 2;
	// every triangle is independent: sized once, filled by a few host threads (same arithmetic, same order of results)
	std::vector<size_t> shape_base(obj.shapes.size() + 1, 0);
	for (size_t si = 0; si < obj.shapes.size(); si++) shape_base[si + 1] = shape_base[si] + obj.shapes[si].indices.size() / 3;
	const size_t total = shape_base.back();
	scene.triangles.resize((size_t)triangle_base + total);
	scene.local_triangles.resize((size_t)triangle_base + total);
	scene.triangle_material.resize((size_t)triangle_base + total);
	PTB_TRACE("mesh resize");
	std::atomic<bool> bad_index(false);
	// A vertex is shared by about six triangles of a closed mesh: when the file has fewer positions + normals than triangle corners, every
	// position and normal is transformed ONCE (same arithmetic per element, so the triangles are bit-identical to transforming per corner)
	// and the triangles only gather.
	const bool per_vertex = g_loader_per_vertex > 0 && (total >= 65536 || g_loader_per_vertex > 1) && nv + nn <= total * 2;      // 2: also for small meshes (tests)
	std::vector<Vec3, DefaultInitAllocator<Vec3>> local_p, world_p, local_n, world_n;
	if (per_vertex)
	{
		local_p.resize(nv); world_p.resize(nv); local_n.resize(nn); world_n.resize(nn);
		auto xform = [&](size_t b, size_t e, bool normals)
		{
			for (size_t i = b; i < e; i++)
			{
				if (!normals)
				{
					V4 p = transform(rot, v4(obj.v[i * 3], obj.v[i * 3 + 1], obj.v[i * 3 + 2], 1.0f));
					V4 pw = transform(xf, v4(p.x, p.y, p.z, 1.0f));
					local_p[i] = Vec3{ p.x, p.y, p.z }; world_p[i] = Vec3{ pw.x, pw.y, pw.z };
				}
				else
				{
					V4 n = transform(rot_it, v4(obj.vn[i * 3], obj.vn[i * 3 + 1], obj.vn[i * 3 + 2], 0.0f));
					Vec3 n1 = normalize_host(n.x, n.y, n.z);
					V4 nw = transform(xf_it, v4(n1.x, n1.y, n1.z, 0.0f));
					local_n[i] = n1; world_n[i] = normalize_host(nw.x, nw.y, nw.z);
				}
			}
		};
		const size_t hw = std::max(1u, std::thread::hardware_concurrency());
		const size_t n_threads = std::min<size_t>(16, hw);
		WorkerGroup group;
		for (size_t w = 1; w < n_threads; w++)
			group.spawn([&xform, w, n_threads, nv, nn] { xform(nv * w / n_threads, nv * (w + 1) / n_threads, false); xform(nn * w / n_threads, nn * (w + 1) / n_threads, true); });
		group.run_here([&] { xform(0, nv / n_threads, false); xform(0, nn / n_threads, true); });
		group.finish();
	}
	std::atomic<bool> out_of_range(false);      // mesh_in_range's test, made while the triangle is at hand instead of in a second, single-threaded pass
	auto fill = [&](size_t t_begin, size_t t_end)
	{
		bool range_ok = true;
		size_t si = 0;
		for (size_t t = t_begin; t < t_end; t++)
		{
			while (t >= shape_base[si + 1]) si++;
			const ObjShape& sh = obj.shapes[si];
			const size_t f = (t - shape_base[si]) * 3;
			const int mat_index = (int)si < mat_num ? (int)si : mat_num - 1;
			Triangle tri, local;
			Vec3* vv[3] = { &tri.v0, &tri.v1, &tri.v2 };
			Vec3* nn3[3] = { &tri.n0, &tri.n1, &tri.n2 };
			Vec2* uu[3] = { &tri.uv0, &tri.uv1, &tri.uv2 };
			Vec3* lv[3] = { &local.v0, &local.v1, &local.v2 };
			Vec3* ln[3] = { &local.n0, &local.n1, &local.n2 };
			Vec2* lu[3] = { &local.uv0, &local.uv1, &local.uv2 };
			for (int k = 0; k < 3; k++)
			{
				const ObjIndex& ix = sh.indices[f + k];
				if (ix.v < 0 || (size_t)ix.v >= nv || ix.vn < 0 || (size_t)ix.vn >= nn || (has_uv && (ix.vt < 0 || (size_t)ix.vt >= nt)))
				{
					bad_index = true;
					return;
				}
				if (per_vertex)
				{
					*vv[k] = world_p[ix.v]; *nn3[k] = world_n[ix.vn];
					*uu[k] = has_uv ? Vec2{ obj.vt[ix.vt * 2], obj.vt[ix.vt * 2 + 1] } : Vec2{ 0.0f, 0.0f };
					*lv[k] = local_p[ix.v]; *ln[k] = local_n[ix.vn]; *lu[k] = *uu[k];
					continue;
				}
				V4 p = transform(rot, v4(obj.v[ix.v * 3], obj.v[ix.v * 3 + 1], obj.v[ix.v * 3 + 2], 1.0f));
				V4 n = transform(rot_it, v4(obj.vn[ix.vn * 3], obj.vn[ix.vn * 3 + 1], obj.vn[ix.vn * 3 + 2], 0.0f));
				Vec3 n1 = normalize_host(n.x, n.y, n.z);
				// upload step: world = T*S, normals through its inverse transpose, re-normalised
				V4 pw = transform(xf, v4(p.x, p.y, p.z, 1.0f));
				V4 nw = transform(xf_it, v4(n1.x, n1.y, n1.z, 0.0f));
				*vv[k] = Vec3{ pw.x, pw.y, pw.z };
				*nn3[k] = normalize_host(nw.x, nw.y, nw.z);
				*uu[k] = has_uv ? Vec2{ obj.vt[ix.vt * 2], obj.vt[ix.vt * 2 + 1] } : Vec2{ 0.0f, 0.0f };
				*lv[k] = Vec3{ p.x, p.y, p.z }; *ln[k] = n1; *lu[k] = *uu[k];
			}
			{
				const float c[9] = { tri.v0.x, tri.v0.y, tri.v0.z, tri.v1.x, tri.v1.y, tri.v1.z, tri.v2.x, tri.v2.y, tri.v2.z };
				for (float x : c) if (!(std::fabs(x) <= 1e18f)) range_ok = false;     // also false for NaN
			}
			scene.triangles[(size_t)triangle_base + t] = tri;
			scene.local_triangles[(size_t)triangle_base + t] = local;
			scene.triangle_material[(size_t)triangle_base + t] = material_base + mat_index;
		}
		if (!range_ok) out_of_range = true;
	};
	{
		const size_t hw = std::max(1u, std::thread::hardware_concurrency());
		const size_t n_threads = total < 65536 ? 1 : std::min<size_t>(16, hw);
		WorkerGroup group;
		for (size_t w = 1; w < n_threads; w++)
		{
			const size_t b = total * w / n_threads, e = total * (w + 1) / n_threads;
			group.spawn([&fill, b, e] { fill(b, e); });
		}
		group.run_here([&] { fill(0, total / n_threads); });
		group.finish();
	}
	PTB_TRACE("mesh fill world triangles");
	if (bad_index)
	{
		scene.triangles.resize((size_t)triangle_base); scene.local_triangles.resize((size_t)triangle_base); scene.triangle_material.resize((size_t)triangle_base);
		set_error("[Error]" + path + ": face index out of range / missing normal or texcoord index");
		return false;
	}
	// a mesh without texture coordinates cannot carry a diffuse texture (triangle_mesh.cpp:131-137): applies to every material a triangle uses
	if (!has_uv)
		for (size_t si = 0; si < obj.shapes.size(); si++)
			if (shape_base[si + 1] > shape_base[si]) mats[(int)si < mat_num ? (int)si : mat_num - 1].diffuse_texture_id = -1;
	mesh_triangles = (int)total;
	scene.materials.insert(scene.materials.end(), mats.begin(), mats.end());
	scene.mesh_triangle_count.push_back(mesh_triangles);
	scene.mesh_material_count.push_back(mat_num);
	MeshInfo info;
	info.first_triangle = triangle_base; info.triangle_count = mesh_triangles;
	info.first_material = material_base; info.material_count = mat_num;
	info.position = position; info.scale = scale_v; info.rotate = rotate_v; info.rotate_applied = rotate_v;
	if (out_of_range) { set_error("[Error]Mesh " + path + " has vertices outside the supported range (non-finite or beyond 1e18)"); return false; }
	scene.meshes.push_back(info);
	PTB_TRACE("mesh range check");
	return true;
}

// World-space positions must be finite and small enough that box extents and surface areas stay finite in float: the tree
// builders (host and device) bin by centroid and compare areas, and neither defines a result for NaN / inf input.  The reference
// builds garbage from such a mesh; here it is an error.
static bool mesh_in_range(const HostScene& scene, const MeshInfo& m)
{
	const float limit = 1e18f;
	for (int i = m.first_triangle; i < m.first_triangle + m.triangle_count; i++)
	{
		const Triangle& t = scene.triangles[i];
		const float c[9] = { t.v0.x, t.v0.y, t.v0.z, t.v1.x, t.v1.y, t.v1.z, t.v2.x, t.v2.y, t.v2.z };
		for (float x : c) if (!(std::fabs(x) <= limit)) return false;     // also false for NaN
	}
	return true;
}

// world = (T * S) * local, normals through the inverse transpose, re-normalised — the upload step of
// triangle_mesh::create_mesh_device_data and the loop of set_transform_device (triangle_mesh.cpp:294-325)
static void place_mesh(HostScene& scene, const MeshInfo& m)
{
	M4 xf = identity();
	xf = translate(xf, m.position.x, m.position.y, m.position.z);
	xf = scale(xf, m.scale.x, m.scale.y, m.scale.z);
	M4 xf_it = transpose(inverse(xf));
	for (int i = m.first_triangle; i < m.first_triangle + m.triangle_count; i++)
	{
		const Triangle& l = scene.local_triangles[i];
		Triangle& t = scene.triangles[i];
		const Vec3* lv[3] = { &l.v0, &l.v1, &l.v2 };
		const Vec3* ln[3] = { &l.n0, &l.n1, &l.n2 };
		Vec3* tv[3] = { &t.v0, &t.v1, &t.v2 };
		Vec3* tn[3] = { &t.n0, &t.n1, &t.n2 };
		for (int k = 0; k < 3; k++)
		{
			V4 pw = transform(xf, v4(lv[k]->x, lv[k]->y, lv[k]->z, 1.0f));
			V4 nw = transform(xf_it, v4(ln[k]->x, ln[k]->y, ln[k]->z, 0.0f));
			*tv[k] = Vec3{ pw.x, pw.y, pw.z };
			*tn[k] = normalize_host(nw.x, nw.y, nw.z);
		}
	}
}

bool set_mesh_transform(HostScene& scene, int mesh, const Vec3& position, const Vec3& scale_v)
{
	if (mesh < 0 || mesh >= (int)scene.meshes.size()) { set_error("[Error]mesh index out of range"); return false; }
	MeshInfo& m = scene.meshes[mesh];
	const Vec3 old_position = m.position, old_scale = m.scale;
	m.position = position;
	m.scale = scale_v;
	place_mesh(scene, m);
	if (!mesh_in_range(scene, m))
	{
		m.position = old_position; m.scale = old_scale;
		place_mesh(scene, m);
		set_error("[Error]mesh transform puts vertices outside the supported range (non-finite or beyond 1e18)");
		return false;
	}
	return true;
}

bool apply_mesh_rotate(HostScene& scene, int mesh, const Vec3& rotate_v)
{
	if (mesh < 0 || mesh >= (int)scene.meshes.size()) { set_error("[Error]mesh index out of range"); return false; }
	MeshInfo& m = scene.meshes[mesh];
	if (!std::isfinite(rotate_v.x) || !std::isfinite(rotate_v.y) || !std::isfinite(rotate_v.z)) { set_error("[Error]mesh rotation is not finite"); return false; }
	m.rotate = rotate_v;
	// the rotation still to apply, about z, then y, then x (triangle_mesh.cpp:364-368)
	const float deg2rad = (float)0.01745329251994329576923690768489;
	M4 rot = identity();
	rot = rotate(rot, (m.rotate.z - m.rotate_applied.z) * deg2rad, 0.0f, 0.0f, 1.0f);
	rot = rotate(rot, (m.rotate.y - m.rotate_applied.y) * deg2rad, 0.0f, 1.0f, 0.0f);
	rot = rotate(rot, (m.rotate.x - m.rotate_applied.x) * deg2rad, 1.0f, 0.0f, 0.0f);
	M4 rot_it = transpose(inverse(rot));
	M4 xf = identity();
	xf = translate(xf, m.position.x, m.position.y, m.position.z);
	xf = scale(xf, m.scale.x, m.scale.y, m.scale.z);
	M4 xf_it = transpose(inverse(xf));
	for (int i = m.first_triangle; i < m.first_triangle + m.triangle_count; i++)
	{
		Triangle& l = scene.local_triangles[i];
		Triangle& t = scene.triangles[i];
		Vec3* lv[3] = { &l.v0, &l.v1, &l.v2 };
		Vec3* ln[3] = { &l.n0, &l.n1, &l.n2 };
		Vec3* tv[3] = { &t.v0, &t.v1, &t.v2 };
		Vec3* tn[3] = { &t.n0, &t.n1, &t.n2 };
		for (int k = 0; k < 3; k++)
		{
			V4 p = transform(rot, v4(lv[k]->x, lv[k]->y, lv[k]->z, 1.0f));
			V4 n = transform(rot_it, v4(ln[k]->x, ln[k]->y, ln[k]->z, 0.0f));
			*lv[k] = Vec3{ p.x, p.y, p.z };
			*ln[k] = normalize_host(n.x, n.y, n.z);
			// apply_rotate carries the rotated vec4s straight into the world transform: the world normal is
			// normalize(xf_it * (rot_it * n)) WITHOUT the intermediate normalisation the stored local normal gets
			// (triangle_mesh.cpp:370-408), unlike set_transform_device which starts from the stored one
			V4 pw = transform(xf, p);
			V4 nw = transform(xf_it, n);
			*tv[k] = Vec3{ pw.x, pw.y, pw.z };
			*tn[k] = normalize_host(nw.x, nw.y, nw.z);
		}
	}
	m.rotate_applied = m.rotate;
	return true;
}

// ------------------------------------------------------------------------------------------
// scene JSON
// ------------------------------------------------------------------------------------------

void set_loader_threads(int n) { g_loader_threads = n < 0 ? 0 : n > 64 ? 64 : n; }
void set_loader_mesh_lanes(int n) { g_loader_mesh_lanes = n < 0 ? 0 : n > 8 ? 8 : n; }
void set_loader_per_vertex(int mode) { g_loader_per_vertex = mode < 0 ? 0 : mode > 2 ? 2 : mode; }

bool load_scene(const std::string& scene_json_path, const std::string& asset_root, HostScene& scene)
{
	scene = HostScene();
	std::string text, err;
	if (!read_text_file(normalize_separators(scene_json_path), text)) { set_error("[Error]cannot open scene file " + scene_json_path); return false; }
	JValue root;
	if (!JParser(text).parse(root, err)) { set_error("[Error]scene parse error: " + err); return false; }

	std::map<std::string, ptb_material> materials = builtin_table();
	std::vector<std::string> texture_paths;

	// Background (scene_parser.cpp:80-110)
	const JValue& background = root["Background"];
	if (background.is_null()) { set_error("[Error]Background not defined!"); return false; }
	if (background.is_array()) { set_error("[Error]Background can not be array!"); return false; }
	std::string bg_name, bg_path, bg_format;
	if (!get_string(background, "Background", "Name", bg_name)) return false;
	if (!get_string(background, "Background", "Path", bg_path)) return false;
	if (!get_string(background, "Background", "Format", bg_format)) return false;
	static const char* face_names[6] = { "xpos", "xneg", "ypos", "yneg", "zpos", "zneg" };

	// Texture (scene_parser.cpp:112-126)
	const JValue& textures = root["Texture"];
	if (!textures.is_null())
	{
		if (!textures.is_array()) { set_error("[Error]Texture must be array!"); return false; }
		for (auto& t : textures.arr)
		{
			if (!t.is_string()) { set_error("[Error]Texture entries must be strings"); return false; }
			texture_paths.push_back(t.str);
		}
	}

	// Material (scene_parser.cpp:128-234)
	const JValue& mats = root["Material"];
	if (!mats.is_null())
	{
		if (!mats.is_array()) { set_error("[Error]Material must be array!"); return false; }
		for (auto& m : mats.arr)
		{
			std::string name, diffuse, emission, specular, transparent, roughness, n, k, sa, ss;
			if (!get_string(m, "Material", "Name", name) || !get_string(m, "Material", "Diffuse", diffuse) ||
				!get_string(m, "Material", "Emission", emission) || !get_string(m, "Material", "Specular", specular) ||
				!get_string(m, "Material", "Transparent", transparent) || !get_string(m, "Material", "Roughness", roughness) ||
				!get_string(m, "Material", "RefractionIndex", n) || !get_string(m, "Material", "ExtinctionCoef", k) ||
				!get_string(m, "Material", "AbsorptionCoef", sa) || !get_string(m, "Material", "ReducedScatteringCoef", ss))
				return false;
			ptb_material mat;
			memset(&mat, 0, sizeof(mat));
			Vec3 d = parse_float3(diffuse), e = parse_float3(emission), s = parse_float3(specular), a = parse_float3(sa), r = parse_float3(ss);
			mat.diffuse_color[0] = d.x; mat.diffuse_color[1] = d.y; mat.diffuse_color[2] = d.z;
			mat.emission_color[0] = e.x; mat.emission_color[1] = e.y; mat.emission_color[2] = e.z;
			mat.specular_color[0] = s.x; mat.specular_color[1] = s.y; mat.specular_color[2] = s.z;
			mat.is_transparent = parse_bool(transparent) ? 1 : 0;
			mat.roughness = clampf(parse_float(roughness), 0.0f, 1.0f);
			mat.refraction_index = parse_float(n);
			mat.extinction_coefficient = parse_float(k);
			mat.absorption_coefficient[0] = a.x; mat.absorption_coefficient[1] = a.y; mat.absorption_coefficient[2] = a.z;
			mat.reduced_scattering_coefficient[0] = r.x; mat.reduced_scattering_coefficient[1] = r.y; mat.reduced_scattering_coefficient[2] = r.z;
			mat.diffuse_texture_id = -1;
			mat.specular_texture_id = -1;
			const JValue* ids[2] = { &m["DiffuseTextureId"], &m["SpecularTextureId"] };
			int32_t* dst[2] = { &mat.diffuse_texture_id, &mat.specular_texture_id };
			for (int q = 0; q < 2; q++)
			{
				if (ids[q]->is_null()) continue;
				if (!ids[q]->is_string()) { set_error("[Error]Materail <" + name + ">: texture id must be a string"); return false; }
				int id = parse_int(ids[q]->str);
				if (id != -1 && (id >= (int)texture_paths.size() || id < 0))
				{
					set_error("[Error]Materail <" + name + ">: Texture index out of range!");
					return false;
				}
				*dst[q] = id;
			}
			if (mat.is_transparent && mat.extinction_coefficient > 0.0f)
			{
				set_error("[Error]Materail <" + name + ">: Extinction coefficient of transparent material should be zero!");
				return false;
			}
			materials[name] = mat;
		}
	}

	// Sphere (scene_parser.cpp:236-268)
	std::vector<std::string> sphere_materials;
	const JValue& spheres = root["Sphere"];
	if (!spheres.is_null())
	{
		if (!spheres.is_array()) { set_error("[Error]Sphere must be array!"); return false; }
		for (auto& s : spheres.arr)
		{
			std::string center, radius, material;
			if (!get_string(s, "Sphere", "Center", center) || !get_string(s, "Sphere", "Radius", radius) || !get_string(s, "Sphere", "Material", material))
				return false;
			Sphere sp;
			memset(&sp, 0, sizeof(sp));
			sp.center = parse_float3(center);
			sp.radius = clampf(parse_float(radius), 0.0f, INFINITY);
			scene.spheres.push_back(sp);
			sphere_materials.push_back(material);
		}
	}

	// Mesh (scene_parser.cpp:270-322)
	struct MeshDecl { std::string path; std::vector<std::string> mats; Vec3 position, scale, rotate; };
	std::vector<MeshDecl> meshes;
	const JValue& mesh_array = root["Mesh"];
	if (!mesh_array.is_null())
	{
		if (!mesh_array.is_array()) { set_error("[Error]Mesh must be array!"); return false; }
		for (auto& m : mesh_array.arr)
		{
			MeshDecl d;
			std::string position, scale_s, rotate_s;
			if (!get_string(m, "Mesh", "Path", d.path)) return false;
			const JValue& mm = m["Material"];
			if (mm.is_null()) { set_error("[Error]Mesh property <Material> not defined!"); return false; }
			if (!get_string(m, "Mesh", "Position", position) || !get_string(m, "Mesh", "Scale", scale_s) || !get_string(m, "Mesh", "Rotate", rotate_s))
				return false;
			if (!mm.is_array()) { set_error("[Error]Material of mesh must be array!"); return false; }
			for (auto& name : mm.arr)
			{
				if (!name.is_string()) { set_error("[Error]Material of mesh must be an array of strings"); return false; }
				d.mats.push_back(name.str);
			}
			d.position = parse_float3(position);
			Vec3 sc = parse_float3(scale_s);
			d.scale = Vec3{ clampf(sc.x, 0.0f, INFINITY), clampf(sc.y, 0.0f, INFINITY), clampf(sc.z, 0.0f, INFINITY) };
			d.rotate = parse_float3(rotate_s);
			meshes.push_back(d);
		}
	}

	// Step 2 (scene_parser.cpp:324-440): cube map, textures, material names, meshes, spheres.
	// The image files decode independently (a 2048^2 JPEG cube face takes 0.3 s): by host threads, checked afterwards in file order
	// so the outcome and the error reported are those of loading them one after the other.
	PTB_TRACE_BEGIN;
	std::vector<std::string> image_paths;
	for (int f = 0; f < 6; f++) image_paths.push_back(join_path(asset_root, bg_path + bg_name + "\\" + face_names[f] + "." + bg_format));
	for (auto& tp : texture_paths) image_paths.push_back(join_path(asset_root, tp));
	scene.textures.assign(texture_paths.size(), Texture());
	std::vector<char> image_ok(image_paths.size(), 0);
	{
		std::atomic<size_t> next(0);
		auto work = [&]
		{
			for (size_t i = next++; i < image_paths.size(); i = next++)
			{
				Texture& out = i < 6 ? scene.cube_faces[i] : scene.textures[i - 6];
				bool ok = false;
				try { ok = load_image_rgba8(image_paths[i], out); } catch (...) { ok = false; }      // bad_alloc in a decoder is a load error
				image_ok[i] = ok ? 1 : 0;
			}
		};
		const size_t hw = std::max(1u, std::thread::hardware_concurrency());
		const size_t n_threads = g_loader_threads > 0 ? (size_t)g_loader_threads : std::min<size_t>(8, hw);
		WorkerGroup group;
		for (size_t k = 1; k < n_threads && k < image_paths.size(); k++) group.spawn([&work] { work(); });
		group.run_here([&work] { work(); });
		group.finish();
	}
	PTB_TRACE("images decoded");
	for (int f = 0; f < 6; f++)
	{
		const std::string& p = image_paths[f];
		if (!image_ok[f])
		{
			set_error("[Error]Background load fail, please check the <Path> and <Name>! (" + p + ")");
			return false;
		}
		const Texture& t = scene.cube_faces[f];
		if (t.width != t.height || t.width != scene.cube_faces[0].width)
		{
			set_error("[Error]Background load fail: cube map faces must be square and equal (" + p + ")");
			return false;
		}
	}
	scene.cube_length = scene.cube_faces[0].width;

	for (size_t i = 0; i < texture_paths.size(); i++)
		if (!image_ok[6 + i]) { set_error("[Error]Texture " + texture_paths[i] + " load fail."); return false; }

	bool missing = false;
	std::string missing_names;
	for (auto& n : sphere_materials) if (!materials.count(n)) { missing = true; missing_names += " <" + n + ">"; }
	for (auto& m : meshes) for (auto& n : m.mats) if (!materials.count(n)) { missing = true; missing_names += " <" + n + ">"; }
	if (missing) { set_error("[Error]Material" + missing_names + " not found!"); return false; }

	// every mesh file is parsed first (stopping at the first that fails), then the triangle arrays are sized once and the meshes
	// appended in order: results and the error reported are those of loading the meshes one after the other
	// Each file is parsed by up to 16 host threads, and several FILES are parsed at the same time (the 8-GPU box has 32 cores: the two
	// 190 MB meshes of the 4K workload side by side).  Files are handed out in order from a shared cursor; what is reported is decided
	// afterwards, in mesh order, so a failure behaves as if the files had been read one after the other (the files behind it were read in vain).
	std::vector<PendingMesh> pending(meshes.size());
	size_t total_triangles = 0;
	{
		const size_t hw = std::max(1u, std::thread::hardware_concurrency());
		const size_t per_file = g_loader_threads > 0 ? (size_t)g_loader_threads : std::min<size_t>(16, hw);
		// at least two: one file's serial stretches (merging the slices' vertices, appending the runs) overlap the other's parallel ones even
		// on a host whose cores one parser already fills (8 cores, two 190 MB files: 0.64 -> 0.55 s)
		size_t lanes = g_loader_mesh_lanes > 0 ? (size_t)g_loader_mesh_lanes : std::max<size_t>(2, hw / std::max<size_t>(1, per_file));
		lanes = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(lanes, 8), meshes.size()));
		std::atomic<size_t> next(0);
		auto work = [&]
		{
			for (size_t i = next++; i < meshes.size(); i = next++)
				read_mesh(join_path(asset_root, meshes[i].path), meshes[i].mats.size(), pending[i]);
		};
		if (lanes <= 1)
		{
			for (size_t i = 0; i < meshes.size(); i++)
			{
				read_mesh(join_path(asset_root, meshes[i].path), meshes[i].mats.size(), pending[i]);
				if (!pending[i].ok) break;
			}
		}
		else
		{
			WorkerGroup group;
			for (size_t k = 1; k < lanes; k++) group.spawn([&work] { work(); });
			group.run_here([&work] { work(); });
			group.finish();
		}
	}
	for (size_t i = 0; i < meshes.size(); i++)
	{
		if (!pending[i].ok) break;
		total_triangles += pending[i].triangle_count();
	}
	scene.triangles.reserve(total_triangles); scene.local_triangles.reserve(total_triangles); scene.triangle_material.reserve(total_triangles);
	for (size_t i = 0; i < meshes.size(); i++)
	{
		auto& m = meshes[i];
		std::vector<ptb_material> mesh_mats;
		for (auto& n : m.mats) mesh_mats.push_back(materials[n]);
		if (!append_mesh(scene, join_path(asset_root, m.path), m.position, m.scale, m.rotate, mesh_mats, pending[i])) return false;
		pending[i].obj = ObjData();     // the parsed text of a 2.5 M-triangle mesh holds ~150 MB
	}
	for (size_t i = 0; i < scene.spheres.size(); i++) scene.spheres[i].mat = materials[sphere_materials[i]];
	return true;
}

int list_scenes(const std::string& dir_in, std::vector<std::string>& out)
{
	std::string dir = normalize_separators(dir_in);
	DIR* d = opendir(dir.c_str());
	if (!d) return -1;
	while (dirent* e = readdir(d))
	{
		std::string n = e->d_name;
		if (n.size() > 5 && n.substr(n.size() - 5) == ".json") out.push_back(join_path(dir, n));
	}
	closedir(d);
	std::sort(out.begin(), out.end());
	return (int)out.size();
}

} // namespace ptb
