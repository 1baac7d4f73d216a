// Host-side scene model of the ptb200 render path: what the reference keeps in
// scene_parser / triangle_mesh / cube_map_loader / config_parser, flattened.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/ptb200.h"

namespace ptb
{

struct Vec3 { float x, y, z; };
struct Vec2 { float x, y; };

// Field-for-field mirror of the reference `configuration` (Core/configuration.h:9-34, 96 bytes).
struct Config
{
	int width;
	int height;
	bool use_fullscreen;
	int block_size;
	int max_block_size;
	int max_tracer_depth;
	float vector_bias_length;
	float energy_exist_threshold;
	float sss_threshold;
	bool use_sky_box;
	bool use_sky;
	bool use_bilinear;
	bool gamma_correction;
	bool use_anti_alias;
	float fov;
	int bvh_leaf_node_triangle_num;
	int bvh_bucket_max_divide_internal_num;
	int bvh_build_block_size;
	int bvh_build;                 // 0 NaiveCPU, 1 MortonCodeCPU, 2 MortonCodeCUDA (Bvh/bvh_build_config.h)
	float air_refraction_index;
	Vec3 air_absorption_coef;
	Vec3 air_reduced_scattering_coef;
	bool cuda_acceleration;
};
static_assert(sizeof(Config) == 96, "Config must match the reference configuration layout");
static_assert(sizeof(ptb_material) == 84, "ptb_material must match the reference material layout");
static_assert(sizeof(ptb_camera) == 64, "ptb_camera must match the reference render_camera layout");

// One triangle as the reference stores it after upload (Core/triangle.h:11-25), minus the pointer.
struct Triangle
{
	Vec3 v0, v1, v2;
	Vec3 n0, n1, n2;
	Vec2 uv0, uv1, uv2;
};
static_assert(sizeof(Triangle) == 96, "24 floats");

// Triangle arrays run to hundreds of MB: resize() must not zero-fill them on one thread before the loader's host threads
// write every element anyway (default-initialising allocator: `resize(n)` leaves new elements uninitialised, `resize(n, x)`,
// push_back and copies behave as usual).
template <class T>
struct DefaultInitAllocator : std::allocator<T>
{
	template <class U> struct rebind { using other = DefaultInitAllocator<U>; };
	using std::allocator<T>::allocator;
	template <class U> void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new (static_cast<void*>(p)) U; }
	template <class U, class... Args> void construct(U* p, Args&&... args) { ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...); }
};
using TriangleArray = std::vector<Triangle, DefaultInitAllocator<Triangle>>;

struct Sphere
{
	Vec3 center;
	float radius;
	ptb_material mat;
};
static_assert(sizeof(Sphere) == 100, "Sphere must match the reference sphere layout");

struct Texture
{
	int width = 0, height = 0;
	std::vector<uint8_t> rgba;     // RGBA8, row 0 = top (Others/image_loader.cpp:31-95)
};

// Per-mesh placement state the reference keeps for live edits (triangle_mesh.h: m_mesh_position,
// m_mesh_scale, m_mesh_rotate, m_mesh_rotate_applied, m_mesh_triangles_num, m_mesh_material_num).
struct MeshInfo
{
	int first_triangle = 0, triangle_count = 0;
	int first_material = 0, material_count = 0;
	Vec3 position{ 0, 0, 0 }, scale{ 1, 1, 1 }, rotate{ 0, 0, 0 }, rotate_applied{ 0, 0, 0 };
};

struct HostScene
{
	TriangleArray triangles;               // global order: meshes in JSON order, shapes, faces
	TriangleArray local_triangles;         // the reference's m_triangles: rotation applied, translate/scale not (triangle_mesh.cpp:150-185)
	std::vector<MeshInfo> meshes;
	std::vector<int32_t> triangle_material; // index into `materials`
	std::vector<ptb_material> materials;    // per-mesh private copies, concatenated
	std::vector<int> mesh_triangle_count;
	std::vector<int> mesh_material_count;
	std::vector<Sphere> spheres;
	std::vector<Texture> textures;
	Texture cube_faces[6];                  // +x -x +y -y +z -z
	int cube_length = 0;
};

// Error text of the last failing loader call on this thread.
std::string& last_error();
void set_error(const std::string& msg);

bool load_config(const std::string& path, Config& out);
bool load_scene(const std::string& scene_json_path, const std::string& asset_root, HostScene& out);
bool load_image_rgba8(const std::string& path, Texture& out);
// decoders refuse headers announcing more than this many pixels (1 GiB of RGBA8) before allocating anything
const uint64_t kMaxImagePixels = 1ull << 28;
// baseline / progressive JPEG (csrc/jpeg_decode.cpp); false for arithmetic / CMYK / corrupt / incomplete files
bool decode_jpeg(const std::vector<uint8_t>& file, Texture& out);
// process-wide JPEG decode mode: reference = what FreeImage does for the reference's loads (IFAST IDCT, replicated chroma,
// libjpeg 9a colour constants), fast = libjpeg-turbo's fast decode, accurate = libjpeg-turbo's default decode (PIL)
enum { kJpegReference = 0, kJpegFast = 1, kJpegAccurate = 2 };
void set_jpeg_mode(int mode);
int jpeg_mode();
// OBJ text is parsed in this many slices by host threads (0 = chosen from the file size and the host's cores)
void set_loader_threads(int n);
void set_loader_mesh_lanes(int n);      // mesh files parsed at the same time (0: by host cores)
void set_loader_per_vertex(int mode);   // 0: per triangle corner, 1 (default): per vertex for meshes of >= 65536 triangles, 2: per vertex always
// live edits with the reference's arithmetic: triangle_mesh::set_transform_device (triangle_mesh.cpp:271-328)
// and set_rotate + apply_rotate (:330-426).  They rewrite `triangles` (and `local_triangles`) of one mesh.
bool set_mesh_transform(HostScene& scene, int mesh, const Vec3& position, const Vec3& scale);
bool apply_mesh_rotate(HostScene& scene, int mesh, const Vec3& rotate);
bool builtin_material(const std::string& name, ptb_material& out);
void default_camera(float width, float height, float aperture, float focal, ptb_camera& out);
int list_scenes(const std::string& dir, std::vector<std::string>& out);

} // namespace ptb
