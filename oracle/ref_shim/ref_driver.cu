/* TEST INFRASTRUCTURE — headless harness around the UNMODIFIED reference sources.
 *
 * This translation unit textually includes the reference's integrator
 * (/root/reference/gpu_path_tracer/Kernel/path_tracer_kernel.cu, where it lies) and adds
 *   - a C ABI that drives it the way Main/window.cpp:338-360,381-391 and
 *     Core/path_tracer.cpp:18-99,371-395 do (config -> camera -> scene -> passes), without
 *     GLFW/ImGui;
 *   - ref_trace_batch: a kernel that runs the reference's OWN closest-hit code
 *     (sphere loop + intersect_triangle_mesh_bvh, path_tracer_kernel.cu:431-454) on a caller
 *     supplied ray batch and returns primitive ids — the source of truth for prim-ID parity;
 *   - ref_pass_instrumented: the reference's pass loop (path_tracer_kernel.cu:706-779)
 *     re-issued launch by launch with CUDA events around trace_ray_kernel and a ray-segment
 *     counter, and optional capture of the live ray batch at a chosen depth.
 * Nothing here ships in the product library. Built only by oracle/build_ref.py into
 * oracle/_ref/libptref.so (git-ignored). */
#define private public
#include "Core\scene_parser.h"
#include "Core\config_parser.h"
#undef private
#include "Kernel\path_tracer_kernel.cu"

#include <chrono>
#include <string>
#include <vector>

namespace
{
	config_parser* g_config = nullptr;
	scene_parser* g_scene = nullptr;
	view_camera* g_view_cam = nullptr;
	render_camera* g_render_cam = nullptr;
	image* g_image = nullptr;
	color* g_not_absorbed = nullptr;
	color* g_accumulated = nullptr;
	ray* g_rays = nullptr;
	int* g_energy_exist = nullptr;
	scattering* g_scatterings = nullptr;
	bool g_scene_ready = false;
	double g_last_trace_ms = 0.0, g_last_shrink_ms = 0.0, g_last_pass_ms = 0.0;
	long long g_last_segments = 0;
}

__global__ void ref_trace_batch_kernel(
	int mesh_num, bvh_node_device** bvh_nodes, triangle* triangles,
	int sphere_num, sphere* spheres, configuration* config,
	const ray* rays, int n, int* out_prim, float* out_t)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	ray tracing_ray = rays[i];

	/* same statement sequence as path_tracer_kernel.cu:418-454 */
	float hit_t, hit_t1, hit_t2;
	float3 hit_point, hit_normal;
	int hit_triangle_index;
	float min_t = INFINITY;
	int prim = -1;

	for (int s = 0; s < sphere_num; s++)
	{
		if (spheres[s].intersect(tracing_ray, hit_point, hit_normal, hit_t) && hit_t < min_t && hit_t > 0.0f)
		{
			min_t = hit_t;
			prim = -(s + 2);
		}
	}
	for (int mesh_index = 0; mesh_index < mesh_num; mesh_index++)
	{
		if (intersect_triangle_mesh_bvh(triangles, bvh_nodes, mesh_index, tracing_ray, config, hit_t, hit_t1, hit_t2, hit_triangle_index) && hit_t < min_t && hit_t > 0.0f)
		{
			min_t = hit_t;
			prim = hit_triangle_index;
		}
	}
	out_prim[i] = prim;
	out_t[i] = min_t;
}

extern "C"
{

int ref_close();

/* config_json / scene_dir use the reference's own relative, backslash-separated
 * conventions; the process CWD must be the scratch asset root. */
int ref_open(const char* config_json, const char* scene_dir, const char* scene_name)
{
	ref_close();
	g_config = new config_parser();
	if (!g_config->load_config(config_json)) { return 1; }
	g_config->create_config_device_data();
	configuration* cfg = g_config->get_config_device_ptr();

	g_view_cam = new view_camera();
	g_render_cam = new render_camera();
	g_view_cam->set_resolution((float)cfg->width, (float)cfg->height);
	g_view_cam->set_fov(45.0f);
	g_view_cam->get_render_camera(g_render_cam);

	g_image = create_image(cfg->width, cfg->height);
	path_tracer_kernel_memory_allocate(&g_not_absorbed, &g_accumulated, &g_rays, &g_energy_exist, &g_scatterings, g_image->pixel_count);

	bvh_build_config::bvh_leaf_node_triangle_num = cfg->bvh_leaf_node_triangle_num;
	bvh_build_config::bvh_bucket_max_divide_internal_num = cfg->bvh_bucket_max_divide_internal_num;
	bvh_build_config::bvh_build_block_size = cfg->bvh_build_block_size;

	g_scene = new scene_parser();
	std::vector<std::string> files = g_scene->set_scene_file_directory(scene_dir);
	int index = -1;
	std::string want = std::string(scene_name);
	for (size_t i = 0; i < files.size(); i++)
	{
		std::string f = files[i];
		size_t p = f.find_last_of("\\/");
		std::string base = p == std::string::npos ? f : f.substr(p + 1);
		if (base == want || base == want + ".json") { index = (int)i; }
	}
	if (index < 0) { return 2; }
	if (!g_scene->load_scene(index)) { return 3; }
	if (!g_scene->create_scene_data_device(cfg->bvh_build)) { return 4; }
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) { return 5; }
	g_scene_ready = true;
	return 0;
}

int ref_close()
{
	if (g_scene) { delete g_scene; g_scene = nullptr; }
	if (g_image)
	{
		path_tracer_kernel_memory_free(g_not_absorbed, g_accumulated, g_rays, g_energy_exist, g_scatterings);
		release_image(g_image); g_image = nullptr;
	}
	if (g_config) { g_config->release_config_device_data(); delete g_config; g_config = nullptr; }
	if (g_view_cam) { delete g_view_cam; g_view_cam = nullptr; }
	if (g_render_cam) { delete g_render_cam; g_render_cam = nullptr; }
	g_scene_ready = false;
	return 0;
}

/* 16 floats: eye3 view3 up3 res2 fov2 aperture focal (Core/camera.h:14-23) */
void ref_get_camera(float* cam16) { memcpy(cam16, g_render_cam, sizeof(render_camera)); }
void ref_set_camera(const float* cam16) { memcpy(g_render_cam, cam16, sizeof(render_camera)); }

int ref_width() { return g_image ? g_image->width : 0; }
int ref_height() { return g_image ? g_image->height : 0; }
int ref_pass_counter() { return g_image ? g_image->pass_counter : 0; }
void ref_clear() { if (g_image) reset_image(g_image); }

static void call_reference_pass()
{
	/* Core/path_tracer.cpp:44-67 */
	g_image->pass_counter++;
	path_tracer_kernel(
		g_scene->get_mesh_num(), g_scene->get_bvh_node_device_ptr(), g_scene->get_triangles_device_ptr(),
		g_scene->get_sphere_num(), g_scene->get_sphere_device_ptr(),
		g_image->pixel_count, g_image->pixels_device, g_image->pixels_256_device, g_image->pass_counter,
		g_render_cam, g_scene->get_cube_map_device_ptr(),
		g_not_absorbed, g_accumulated, g_rays, g_energy_exist, g_scatterings,
		g_scene->get_mesh_texture_device_ptr(), g_config->get_config_device_ptr());
}

/* returns wall seconds spent in n synchronous reference passes */
double ref_render(int n_passes)
{
	if (!g_scene_ready) return -1.0;
	auto t0 = std::chrono::steady_clock::now();
	for (int i = 0; i < n_passes; i++) call_reference_pass();
	auto t1 = std::chrono::steady_clock::now();
	return std::chrono::duration<double>(t1 - t0).count();
}

/* wall seconds of each of n synchronous reference passes (diagnostic for the step-time spread of the reference arm) */
int ref_render_per_pass(int n_passes, double* out_seconds)
{
	if (!g_scene_ready) return 1;
	for (int i = 0; i < n_passes; i++)
	{
		auto t0 = std::chrono::steady_clock::now();
		call_reference_pass();
		out_seconds[i] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	}
	return 0;
}

/* the 18 arguments the reference hands to path_tracer_kernel (Core/path_tracer.cpp:48-67), as raw
 * pointers/ints, so a test can call ANOTHER implementation of the same symbol on the reference's own
 * managed scene buffers.  out[] order = argument order; ints are stored as intptr_t. */
void ref_kernel_args(void** out18)
{
	out18[0] = (void*)(intptr_t)g_scene->get_mesh_num();
	out18[1] = (void*)g_scene->get_bvh_node_device_ptr();
	out18[2] = (void*)g_scene->get_triangles_device_ptr();
	out18[3] = (void*)(intptr_t)g_scene->get_sphere_num();
	out18[4] = (void*)g_scene->get_sphere_device_ptr();
	out18[5] = (void*)(intptr_t)g_image->pixel_count;
	out18[6] = (void*)g_image->pixels_device;
	out18[7] = (void*)g_image->pixels_256_device;
	out18[8] = (void*)(intptr_t)g_image->pass_counter;
	out18[9] = (void*)g_render_cam;
	out18[10] = (void*)g_scene->get_cube_map_device_ptr();
	out18[11] = (void*)g_not_absorbed;
	out18[12] = (void*)g_accumulated;
	out18[13] = (void*)g_rays;
	out18[14] = (void*)g_energy_exist;
	out18[15] = (void*)g_scatterings;
	out18[16] = (void*)g_scene->get_mesh_texture_device_ptr();
	out18[17] = (void*)g_config->get_config_device_ptr();
}

void ref_set_pass_counter(int v) { if (g_image) g_image->pass_counter = v; }

int ref_image_f32(float* out_rgb_sum)
{
	cudaDeviceSynchronize();
	return cudaMemcpy(out_rgb_sum, g_image->pixels_device, (size_t)g_image->pixel_count * sizeof(color), cudaMemcpyDefault) != cudaSuccess;
}

int ref_image_u8(unsigned char* out_rgb)
{
	cudaDeviceSynchronize();
	return cudaMemcpy(out_rgb, g_image->pixels_256_device, (size_t)g_image->pixel_count * sizeof(color256), cudaMemcpyDefault) != cudaSuccess;
}

/* last pass's un-clamped per-pixel radiance (accumulated_colors work buffer) */
int ref_last_pass_f32(float* out_rgb)
{
	cudaDeviceSynchronize();
	return cudaMemcpy(out_rgb, g_accumulated, (size_t)g_image->pixel_count * sizeof(color), cudaMemcpyDefault) != cudaSuccess;
}

int ref_num_triangles() { return g_scene ? g_scene->get_total_triangles_num() : 0; }
int ref_num_spheres() { return g_scene ? g_scene->get_sphere_num() : 0; }
int ref_num_meshes() { return g_scene ? g_scene->get_mesh_num() : 0; }

/* per triangle 24 floats (v0 v1 v2 n0 n1 n2 uv0 uv1 uv2) + material index into the mesh material table */
int ref_get_triangles(float* out24, int* out_mat)
{
	cudaDeviceSynchronize();
	int n = ref_num_triangles();
	triangle* t = g_scene->get_triangles_device_ptr();
	material* base = g_scene->m_triangle_mesh.m_mat_device;
	for (int i = 0; i < n; i++)
	{
		memcpy(out24 + (size_t)i * 24, &t[i], 24 * sizeof(float));
		out_mat[i] = (int)(t[i].mat - base);
	}
	return 0;
}

int ref_num_mesh_materials() { return g_scene ? (int)g_scene->m_triangle_mesh.m_mesh_material.size() : 0; }

/* 21 ints/floats per material = the 84-byte struct verbatim (Core/material.h:49-78) */
int ref_get_mesh_materials(void* out84)
{
	cudaDeviceSynchronize();
	memcpy(out84, g_scene->m_triangle_mesh.m_mat_device, (size_t)ref_num_mesh_materials() * sizeof(material));
	return 0;
}

/* 25 words per sphere = the 100-byte struct verbatim (Core/sphere.h:11-16) */
int ref_get_spheres(void* out100)
{
	cudaDeviceSynchronize();
	if (ref_num_spheres() > 0) memcpy(out100, g_scene->get_sphere_device_ptr(), (size_t)ref_num_spheres() * sizeof(sphere));
	return 0;
}

/* ---- live scene edits, driven the way path_tracer::render_ui does (Core/path_tracer.cpp:109-369):
 * scene_parser::set_sphere_device / set_mesh_material_device / set_mesh_transform_device (with the
 * builder's own bvh update function and the UI's scale clamp) / set_mesh_rotate + apply_mesh_rotate,
 * each followed by reset_image like path_tracer::clear(). */
int ref_set_sphere(int index, const void* sphere100)
{
	if (!g_scene || index < 0 || index >= g_scene->get_sphere_num()) return 1;
	cudaDeviceSynchronize();
	sphere s;
	memcpy(&s, sphere100, sizeof(sphere));
	g_scene->set_sphere_device(index, s);
	if (g_image) reset_image(g_image);
	return 0;
}

int ref_set_mesh_materials(int index, const void* mats84, int n)
{
	if (!g_scene || index < 0 || index >= g_scene->get_mesh_num()) return 1;
	cudaDeviceSynchronize();
	std::vector<material> mats(n);
	memcpy(mats.data(), mats84, (size_t)n * sizeof(material));
	g_scene->set_mesh_material_device(index, mats);
	if (g_image) reset_image(g_image);
	return 0;
}

int ref_set_mesh_transform(int index, const float* position3, const float* scale3)
{
	if (!g_scene || index < 0 || index >= g_scene->get_mesh_num()) return 1;
	cudaDeviceSynchronize();
	bvh_build_method build_method = g_config->get_config_device_ptr()->bvh_build;
	std::function<void(const glm::mat4&, const glm::mat4&, bvh_node_device*, bvh_node_device*)> bvh_update_function = auto_bvh_update(build_method);
	float3 position = make_float3(position3[0], position3[1], position3[2]);
	float3 scale = make_float3(scale3[0], scale3[1], scale3[2]);
	g_scene->set_mesh_transform_device(index, position,
		clamp(scale, make_float3(0.000001f, 0.000001f, 0.000001f), make_float3(INFINITY, INFINITY, INFINITY)), bvh_update_function);
	if (g_image) reset_image(g_image);
	return 0;
}

int ref_apply_mesh_rotate(int index, const float* rotate3)
{
	if (!g_scene || index < 0 || index >= g_scene->get_mesh_num()) return 1;
	cudaDeviceSynchronize();
	g_scene->set_mesh_rotate(index, make_float3(rotate3[0], rotate3[1], rotate3[2]));
	g_scene->apply_mesh_rotate(index);
	if (g_image) reset_image(g_image);
	return 0;
}

int ref_mesh_material_counts(int* out)
{
	if (!g_scene) return 1;
	for (int i = 0; i < g_scene->get_mesh_num(); i++) out[i] = g_scene->m_triangle_mesh.m_mesh_material_num[i];
	return 0;
}

int ref_struct_sizes(int* out8)
{
	out8[0] = sizeof(ray); out8[1] = sizeof(material); out8[2] = sizeof(sphere); out8[3] = sizeof(triangle);
	out8[4] = sizeof(bvh_node_device); out8[5] = sizeof(render_camera); out8[6] = sizeof(configuration); out8[7] = sizeof(cube_map);
	return 0;
}

/* total BVH nodes over all meshes (bvh_node_device[0].next_node_index per mesh) */
long long ref_num_bvh_nodes()
{
	cudaDeviceSynchronize();
	long long n = 0;
	bvh_node_device** nodes = g_scene->get_bvh_node_device_ptr();
	for (int m = 0; m < g_scene->get_mesh_num(); m++) n += nodes[m][0].next_node_index;
	return n;
}

/* Move every managed scene / work buffer to the device so timing excludes unified-memory page
 * migration (second timing).  The reference reads *config_device on the HOST every
 * pass (path_tracer_kernel.cu:706), which bounces that page (and its small managed neighbours)
 * between CPU and GPU; with advise != 0 the harness marks the config read-mostly so both sides
 * keep a copy — the most favourable steady state for the reference, applied from OUTSIDE its
 * sources. */
static void prefetch_one(const void* p, size_t bytes, int dev)
{
	if (p && bytes) cudaMemPrefetchAsync(p, bytes, dev, 0);
}

int ref_prefetch(int advise)
{
	int dev = 0; cudaGetDevice(&dev);
	size_t px = g_image->pixel_count;
	int ntri = ref_num_triangles();
	configuration* cfg = g_config->get_config_device_ptr();
	if (advise) cudaMemAdvise(cfg, sizeof(configuration), cudaMemAdviseSetReadMostly, dev);
	prefetch_one(cfg, sizeof(configuration), dev);
	prefetch_one(g_scene->get_triangles_device_ptr(), (size_t)ntri * sizeof(triangle), dev);
	prefetch_one(g_scene->m_triangle_mesh.m_mat_device, (size_t)ref_num_mesh_materials() * sizeof(material), dev);
	int nmesh = g_scene->get_mesh_num();
	bvh_node_device** init = g_scene->m_triangle_mesh.m_mesh_bvh_initial_device;
	if (nmesh > 0)
	{
		std::vector<bvh_node_device*> roots(init, init + nmesh);   /* read the pointer arrays before moving them */
		for (int m = 0; m < nmesh; m++)
		{
			int nn = roots[m][0].next_node_index;
			int leaves = 0; int* slab = nullptr;
			for (int k = 0; k < nn; k++) if (roots[m][k].is_leaf) { if (!slab) slab = roots[m][k].triangle_indices; leaves++; }
			prefetch_one(slab, (size_t)leaves * cfg->bvh_leaf_node_triangle_num * sizeof(int), dev);
			prefetch_one(roots[m], (size_t)nn * 2 * sizeof(bvh_node_device), dev);
		}
		prefetch_one(init, nmesh * sizeof(bvh_node_device*), dev);
		prefetch_one(g_scene->m_triangle_mesh.m_mesh_bvh_transformed_device, nmesh * sizeof(bvh_node_device*), dev);
	}
	prefetch_one(g_scene->get_sphere_device_ptr(), (size_t)ref_num_spheres() * sizeof(sphere), dev);
	texture_wrapper* tex = g_scene->get_mesh_texture_device_ptr();
	for (int i = 0; i < g_scene->m_textures_num; i++) prefetch_one(tex[i].pixels, (size_t)tex[i].width * tex[i].height * 4, dev);
	prefetch_one(tex, (size_t)g_scene->m_textures_num * sizeof(texture_wrapper), dev);
	cube_map* cm = g_scene->get_cube_map_device_ptr();
	if (cm)
	{
		size_t face = (size_t)cm->length * cm->length * 4;
		uchar* faces[6] = { cm->m_x_positive_map, cm->m_x_negative_map, cm->m_y_positive_map, cm->m_y_negative_map, cm->m_z_positive_map, cm->m_z_negative_map };
		for (int f = 0; f < 6; f++) prefetch_one(faces[f], face, dev);
		prefetch_one(cm, sizeof(cube_map), dev);
	}
	prefetch_one(g_not_absorbed, px * sizeof(color), dev);
	prefetch_one(g_accumulated, px * sizeof(color), dev);
	prefetch_one(g_rays, px * sizeof(ray), dev);
	prefetch_one(g_energy_exist, px * sizeof(int), dev);
	prefetch_one(g_scatterings, px * sizeof(scattering), dev);
	prefetch_one(g_image->pixels_device, px * sizeof(color), dev);
	prefetch_one(g_image->pixels_256_device, px * sizeof(color256), dev);
	return cudaDeviceSynchronize() != cudaSuccess;
}

/* rays: n * 6 floats (origin, direction). out_prim: triangle id >= 0, sphere s -> -(s+2), none -> -1 */
int ref_trace_batch(const float* rays6, int n, int* out_prim, float* out_t)
{
	if (!g_scene_ready) return 1;
	ray* d_rays; int* d_prim; float* d_t;
	cudaMalloc(&d_rays, (size_t)n * sizeof(ray));
	cudaMalloc(&d_prim, (size_t)n * sizeof(int));
	cudaMalloc(&d_t, (size_t)n * sizeof(float));
	cudaMemcpy(d_rays, rays6, (size_t)n * sizeof(ray), cudaMemcpyHostToDevice);
	int block = 64;
	ref_trace_batch_kernel<<<(n + block - 1) / block, block>>>(
		g_scene->get_mesh_num(), g_scene->get_bvh_node_device_ptr(), g_scene->get_triangles_device_ptr(),
		g_scene->get_sphere_num(), g_scene->get_sphere_device_ptr(), g_config->get_config_device_ptr(),
		d_rays, n, d_prim, d_t);
	cudaError_t e = cudaDeviceSynchronize();
	cudaMemcpy(out_prim, d_prim, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost);
	cudaMemcpy(out_t, d_t, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost);
	cudaFree(d_rays); cudaFree(d_prim); cudaFree(d_t);
	return e != cudaSuccess;
}

/* One pass issued launch-by-launch exactly as path_tracer_kernel.cu:706-779 does, with
 * events around trace_ray_kernel. If capture_depth >= 0 the live rays entering that depth
 * are copied out (pixel ids + 6 floats each) and the pass stops there WITHOUT touching the
 * image. Otherwise the pass completes and accumulates like a normal pass.
 * Returns the number of captured rays (capture mode) or total ray segments (normal mode). */
long long ref_pass_instrumented(int pass_counter, int capture_depth, int* out_pixels, float* out_rays6, int max_out)
{
	if (!g_scene_ready) return -1;
	configuration* config_device = g_config->get_config_device_ptr();
	configuration config = *config_device;
	int pixel_count = g_image->pixel_count;
	int threads = config.block_size;
	int blocks = (pixel_count + threads - 1) / threads;
	int count = pixel_count;
	int seed = pass_counter;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	double trace_ms = 0.0; long long segments = 0;
	g_last_shrink_ms = 0.0;
	auto pass_t0 = std::chrono::steady_clock::now();

	init_data_kernel<<<blocks, threads>>>(pixel_count, g_energy_exist, g_not_absorbed, g_accumulated, g_scatterings, config_device);
	generate_ray_kernel<<<blocks, threads>>>(g_render_cam->eye, g_render_cam->view, g_render_cam->up, g_render_cam->resolution,
		g_render_cam->fov, g_render_cam->aperture_radius, g_render_cam->focal_distance, pixel_count, g_rays, seed, config_device);

	for (int depth = 0; depth < config.max_tracer_depth; depth++)
	{
		if (count == 0) break;
		if (depth == capture_depth)
		{
			cudaDeviceSynchronize();
			int n = count < max_out ? count : max_out;
			for (int i = 0; i < n; i++)
			{
				int p = g_energy_exist[i];
				out_pixels[i] = p;
				memcpy(out_rays6 + (size_t)i * 6, &g_rays[p], sizeof(ray));
			}
			cudaEventDestroy(e0); cudaEventDestroy(e1);
			return n;
		}
		int used = (count + threads - 1) / threads;
		segments += count;
		cudaEventRecord(e0);
		trace_ray_kernel<<<used, threads>>>(g_scene->get_mesh_num(), g_scene->get_bvh_node_device_ptr(), g_scene->get_triangles_device_ptr(),
			g_scene->get_sphere_num(), g_scene->get_sphere_device_ptr(), pixel_count, depth, count, g_energy_exist, g_rays, g_scatterings,
			g_not_absorbed, g_accumulated, g_scene->get_cube_map_device_ptr(), g_scene->get_mesh_texture_device_ptr(), seed, config_device);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms = 0; cudaEventElapsedTime(&ms, e0, e1); trace_ms += ms;
		auto s0 = std::chrono::steady_clock::now();
		count = thread_shrink(g_energy_exist, count);   /* thrust::remove_if + its temporary cudaMalloc / cudaFree + the implicit synchronisation */
		g_last_shrink_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - s0).count();
	}
	if (capture_depth >= 0) { cudaEventDestroy(e0); cudaEventDestroy(e1); return 0; }
	g_image->pass_counter = pass_counter;
	pixel_256_transform_gamma_corrected_kernel<<<blocks, threads>>>(g_accumulated, g_image->pixels_device, g_image->pixels_256_device, pixel_count, pass_counter, config_device);
	cudaDeviceSynchronize();
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	g_last_trace_ms = trace_ms; g_last_segments = segments;
	g_last_pass_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - pass_t0).count();
	return segments;
}

double ref_last_trace_ms() { return g_last_trace_ms; }
double ref_last_shrink_ms() { return g_last_shrink_ms; }
double ref_last_pass_ms() { return g_last_pass_ms; }

/* ---- known-answer helpers: the reference's own header functions evaluated on the HOST
 * (they are __host__ __device__), and its device-only hash() evaluated on the device ---- */

int ref_kat_triangle(const float* v9, const float* ray6, float* out_t_t1_t2)
{
	triangle t; memset(&t, 0, sizeof(t));
	memcpy(&t.vertex0, v9, 9 * sizeof(float));
	ray r; memcpy(&r, ray6, sizeof(ray));
	float a = 0, b = 0, c = 0;
	bool hit = t.intersect(r, a, b, c);
	out_t_t1_t2[0] = a; out_t_t1_t2[1] = b; out_t_t1_t2[2] = c;
	return hit ? 1 : 0;
}

int ref_kat_sphere(const float* center_radius4, const float* ray6, float* out_t_p_n7)
{
	sphere s; memset(&s, 0, sizeof(s));
	s.center = make_float3(center_radius4[0], center_radius4[1], center_radius4[2]); s.radius = center_radius4[3];
	ray r; memcpy(&r, ray6, sizeof(ray));
	float3 p = make_float3(0, 0, 0), n = make_float3(0, 0, 0); float t = 0;
	bool hit = s.intersect(r, p, n, t);
	out_t_p_n7[0] = t; memcpy(out_t_p_n7 + 1, &p, 12); memcpy(out_t_p_n7 + 4, &n, 12);
	return hit ? 1 : 0;
}

int ref_kat_box(const float* lo_hi6, const float* ray6, float* inout_t)
{
	bounding_box b(make_float3(lo_hi6[0], lo_hi6[1], lo_hi6[2]), make_float3(lo_hi6[3], lo_hi6[4], lo_hi6[5]));
	ray r; memcpy(&r, ray6, sizeof(ray));
	return b.intersect_bounding_box(r, *inout_t) ? 1 : 0;
}

float ref_kat_fresnel_dielectric(const float* n3, const float* d3, float n_in, float n_out, const float* refl3, const float* refr3)
{
	return fresnel::get_fresnel_dielectrics(make_float3(n3[0], n3[1], n3[2]), make_float3(d3[0], d3[1], d3[2]), n_in, n_out,
		make_float3(refl3[0], refl3[1], refl3[2]), make_float3(refr3[0], refr3[1], refr3[2])).reflection_index;
}

float ref_kat_fresnel_conductor(const float* n3, const float* d3, float n, float k)
{
	return fresnel::get_fresnel_conductors(make_float3(n3[0], n3[1], n3[2]), make_float3(d3[0], d3[1], d3[2]), n, k).reflection_index;
}

int ref_kat_cube_uv(float x, float y, float z, float* out_uv2)
{
	int index = -1; float u = 0, v = 0;
	convert_xyz_to_cube_uv(x, y, z, index, u, v);
	out_uv2[0] = u; out_uv2[1] = v;
	return index;
}

void ref_kat_texture(int w, int h, unsigned char* rgba, float u, float v, int bilinear, float* out3)
{
	texture_wrapper t; t.width = w; t.height = h; t.pixels = rgba;
	float3 c = t.sample_texture(make_float2(u, v), bilinear != 0);
	memcpy(out3, &c, 12);
}

void ref_kat_background(int length, unsigned char** faces6, const float* d3, int use_sky_box, int use_sky, int bilinear, float* out3)
{
	cube_map m; memset(&m, 0, sizeof(m));
	if (faces6)
	{
		m.m_x_positive_map = faces6[0]; m.m_x_negative_map = faces6[1]; m.m_y_positive_map = faces6[2];
		m.m_y_negative_map = faces6[3]; m.m_z_positive_map = faces6[4]; m.m_z_negative_map = faces6[5];
	}
	m.length = length;
	float3 c = m.get_background_color(make_float3(d3[0], d3[1], d3[2]), use_sky_box != 0, use_sky != 0, bilinear != 0);
	memcpy(out3, &c, 12);
}

/* thrust::default_random_engine + uniform_real_distribution<float>(lo, hi), as constructed at
 * path_tracer_kernel.cu:324-325,415-416, evaluated on the host from a given 32-bit seed */
void ref_kat_rng(unsigned int seed, float lo, float hi, int n, float* out)
{
	thrust::default_random_engine random_engine(seed);
	thrust::uniform_real_distribution<float> uniform_distribution(lo, hi);
	for (int i = 0; i < n; i++) out[i] = uniform_distribution(random_engine);
}

/* 84-byte material struct of a built-in material name (Core/material.cpp, scene_parser.cpp:675-708) */
int ref_builtin_material(const char* name, void* out84)
{
	scene_parser tmp;
	std::map<std::string, material> table;
	tmp.init_default_material(table);
	auto it = table.find(name);
	if (it == table.end()) return 1;
	memcpy(out84, &it->second, sizeof(material));
	return 0;
}

int ref_builtin_material_names(char* out, int cap)
{
	scene_parser tmp;
	std::map<std::string, material> table;
	tmp.init_default_material(table);
	std::string all;
	for (auto& kv : table) { all += kv.first; all += "\n"; }
	if ((int)all.size() + 1 > cap) return -1;
	memcpy(out, all.c_str(), all.size() + 1);
	return (int)table.size();
}

/* Core/camera.cpp:3-14,56-66,80-98: default orbit camera for a given resolution */
void ref_default_camera(float width, float height, float aperture, float focal, float* cam16)
{
	view_camera vc;
	vc.set_resolution(width, height);
	vc.set_fov(45.0f);
	if (focal >= 0.0f) vc.set_focal_distance(focal);
	if (aperture >= 0.0f) vc.set_aperture_radius(aperture);
	render_camera rc; memset(&rc, 0, sizeof(rc));
	vc.get_render_camera(&rc);
	memcpy(cam16, &rc, sizeof(rc));
}

} /* extern "C" */

__global__ void ref_hash_kernel(const int* in, int n, int* out)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = hash(in[i]);
}

extern "C" int ref_device_hash(const int* in, int n, int* out)
{
	int *d_in, *d_out;
	if (cudaMalloc(&d_in, n * sizeof(int)) != cudaSuccess) return 1;
	cudaMalloc(&d_out, n * sizeof(int));
	cudaMemcpy(d_in, in, n * sizeof(int), cudaMemcpyHostToDevice);
	ref_hash_kernel<<<(n + 127) / 128, 128>>>(d_in, n, d_out);
	cudaError_t e = cudaDeviceSynchronize();
	cudaMemcpy(out, d_out, n * sizeof(int), cudaMemcpyDeviceToHost);
	cudaFree(d_in); cudaFree(d_out);
	return e != cudaSuccess;
}
