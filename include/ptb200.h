/* ptb200 — C ABI of the B200-native render hot path (drop-in for PathTracerWithCuda's
 * Kernel/ path-trace loop).  Plain pointers and sizes only; no C++/torch types.
 *
 * Every entry point names the reference interface it replaces (paths are relative to
 * /root/reference/gpu_path_tracer/).  The reference-side binding a maintainer would add is
 * shown in INTEGRATION.md.
 *
 * Conventions: functions returning int return 0 on success, non-zero on failure with a
 * message retrievable through ptb_last_error() (the reference itself only prints
 * "[Error]..." / "[Cuda]Error..." and carries on — Others/utilities.hpp:10-18).  A renderer
 * handle is bound to one CUDA device and is not re-entrant; all calls are synchronous unless
 * stated otherwise (the reference returns after cudaDeviceSynchronize,
 * Kernel/path_tracer_kernel.cu:779).  There is NO CPU fallback: without a CUDA device every
 * compute entry point fails. */
#ifndef PTB200_H
#define PTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ptb_renderer ptb_renderer;

/* Same 64-byte layout as the reference's `render_camera` (Core/camera.h:14-23): three float3,
 * 4 bytes of padding (float2 is 8-byte aligned), resolution, fov in DEGREES, aperture, focal. */
typedef struct ptb_camera
{
	float eye[3];
	float view[3];
	float up[3];
	float _pad;
	float resolution[2];
	float fov[2];
	float aperture_radius;
	float focal_distance;
} ptb_camera;

/* Same 84-byte layout as the reference's `material` (Core/material.h:49-78). */
typedef struct ptb_material
{
	float diffuse_color[3];
	float emission_color[3];
	float specular_color[3];
	uint8_t is_transparent; uint8_t _pad[3];
	float roughness;
	float refraction_index;
	float extinction_coefficient;
	float absorption_coefficient[3];
	float reduced_scattering_coefficient[3];
	int32_t diffuse_texture_id;
	int32_t specular_texture_id;
} ptb_material;

/* Counters of the last ptb_render* call (M1/M2 inputs). */
typedef struct ptb_stats
{
	int64_t passes;            /* passes rendered by the call */
	int64_t ray_segments;      /* sum over depth of live paths entering the extend stage */
	int64_t kernel_launches;   /* CUDA kernels launched by the call */
	double gpu_ms_total;       /* CUDA-event time of the whole call on the render stream */
	double gpu_ms_extend;      /* CUDA-event time spent in the extend (closest-hit) kernel; 0 unless profiling enabled */
	int64_t bvh_nodes;         /* wide-BVH node count of the loaded scene */
	int64_t bvh_bytes;         /* bytes of node + triangle intersection data */
	int64_t nodes_visited;     /* instrumented builds only (ptb_set_option "count_traversal" = 1) */
	int64_t tris_tested;
	int64_t wide_nodes_visited; /* 80-byte compressed 8-wide nodes visited (layout 8 / hybrid bounce rays); nodes_visited counts 64-byte binary nodes */
} ptb_stats;

const char* ptb_last_error(void);
int ptb_version(void);
/* number of visible CUDA devices (0 = the library cannot render) */
int ptb_device_count(void);

/* ---- lifecycle -------------------------------------------------------------------------
 * ptb_create        = config_parser::load_config + create_config_device_data
 *                     (Core/config_parser.cpp:8-124,159-187) followed by path_tracer::init's
 *                     buffer allocation (Core/path_tracer.cpp:18-38,
 *                     Kernel/path_tracer_kernel.cu:782-796) and the default orbit camera of
 *                     Main/window.cpp:356-360 / Core/camera.cpp:3-14.
 * ptb_destroy       = path_tracer::~path_tracer (Core/path_tracer.cpp:3-16). */
ptb_renderer* ptb_create(const char* config_json_path, int cuda_device);
void ptb_destroy(ptb_renderer* r);

/* scene_parser::set_scene_file_directory (Core/scene_parser.cpp:9-35): '\n'-separated list of
 * "*.json" under scene_dir, sorted; returns the count or -1. */
int ptb_list_scenes(const char* scene_dir, char* out, int cap);

/* path_tracer::init_scene_device_data (Core/path_tracer.cpp:371-395): parse the scene JSON
 * (Core/scene_parser.cpp:37-442), load OBJ groups (Core/triangle_mesh.cpp:8-213), upload and
 * build the acceleration structure (triangle_mesh.cpp:498-655).  Paths inside the scene file
 * are resolved against asset_root (the reference resolves them against its CWD); both '\\' and
 * '/' separators are accepted.  Resets the accumulation like a scene switch does. */
int ptb_load_scene(ptb_renderer* r, const char* scene_json_path, const char* asset_root);
/* path_tracer::release_scene_device_data (Core/path_tracer.cpp:397-406) */
int ptb_release_scene(ptb_renderer* r);

/* view_camera::get_render_camera with the constructor defaults (Core/camera.cpp:3-14,80-98)
 * for a given resolution; aperture/focal < 0 keep the defaults (0 and radius=14). */
int ptb_default_camera(float width, float height, float aperture_radius, float focal_distance, ptb_camera* out);
int ptb_set_camera(ptb_renderer* r, const ptb_camera* cam);
int ptb_get_camera(ptb_renderer* r, ptb_camera* out);

/* ---- rendering ---------------------------------------------------------------------------
 * ptb_render(r, n)  = n consecutive path_tracer::render() calls (Core/path_tracer.cpp:40-99):
 *                     pass_counter advances by n; pass k uses seed k
 *                     (Kernel/path_tracer_kernel.cu:712) and adds clamp(L_k, 0, 2*MaxDepth) to
 *                     the float accumulation image (:644-651).
 * ptb_clear         = path_tracer::clear -> reset_image (Core/image.cpp:36-40).
 * ptb_render_strided: multi-GPU sharding hook — renders passes first, first+stride, ... (n of
 *                     them) into the local accumulation WITHOUT touching pass_counter's
 *                     meaning for other ranks; see ptb_finalize. */
int ptb_render(ptb_renderer* r, int n_passes);
int ptb_render_strided(ptb_renderer* r, int first_pass, int stride, int n_passes);
int ptb_clear(ptb_renderer* r);
int ptb_pass_counter(ptb_renderer* r);
int ptb_width(ptb_renderer* r);
int ptb_height(ptb_renderer* r);

/* image::pixels_device (float3 running sum, Core/image.h:16) and image::pixels_256_device
 * (gamma-corrected 8-bit, Kernel/path_tracer_kernel.cu:653-680) copied to HOST buffers of
 * width*height*3 elements. */
int ptb_image_f32(ptb_renderer* r, float* out_rgb_sum, int* out_passes);
int ptb_image_u8(ptb_renderer* r, uint8_t* out_rgb);
/* un-clamped radiance of the most recent pass (the reference's accumulated_colors work buffer) */
int ptb_last_pass_f32(ptb_renderer* r, float* out_rgb);

/* Device pointer of the float3 accumulation image (3*width*height floats) so a caller can
 * sum it across ranks with NCCL, then ptb_finalize(total_passes) runs the mean/gamma/8-bit
 * step of pixel_256_transform_gamma_corrected_kernel (:653-680) on the reduced sum. */
void* ptb_image_device_ptr(ptb_renderer* r);
int ptb_finalize(ptb_renderer* r, int total_passes);
/* enqueue work without the trailing synchronize / wait for it (benchmarks use CUDA events) */
int ptb_render_async(ptb_renderer* r, int n_passes);
int ptb_synchronize(ptb_renderer* r);
void* ptb_stream(ptb_renderer* r);

/* Closest hit of a caller-supplied ray batch (n rays x 6 floats: origin, direction), same
 * acceptance rules as Kernel/path_tracer_kernel.cu:418-454.  out_prim: global triangle index
 * >= 0, sphere s -> -(s+2), miss -> -1; out_t: hit distance (+inf on miss); out_bary may be
 * NULL or n x 2 floats (t1, t2).  Host pointers. */
int ptb_trace_batch(ptb_renderer* r, const float* rays6, int n, int32_t* out_prim, float* out_t, float* out_bary);
/* same query answered by a brute-force scan over every primitive (test hook) */
int ptb_trace_batch_bruteforce(ptb_renderer* r, const float* rays6, int n, int32_t* out_prim, float* out_t);
/* rays the camera stage generates for pass `pass` (Kernel/path_tracer_kernel.cu:299-379): n = w*h x 6 floats */
int ptb_generate_rays(ptb_renderer* r, int pass, float* out_rays6);

/* Test hook: the live ray batch entering `depth` of pass `pass` (pixel ids ascending + 6 floats
 * each) — what the reference holds in rays_device[energy_exist_pixels[i]] at that point of
 * Kernel/path_tracer_kernel.cu:738-768.  Does not touch the image. Returns the count or -1. */
int ptb_capture_rays(ptb_renderer* r, int pass, int depth, int32_t* out_pixels, float* out_rays6, int max_out);

int ptb_get_stats(ptb_renderer* r, ptb_stats* out);
/* per-bounce breakdown of the last synchronous render call: live paths entering each depth and
 * (with option "profile_stages"=1) CUDA-event milliseconds spent in the extend kernel there.
 * Returns the number of depths written (<= max_entries) or -1. */
int ptb_get_depth_profile(ptb_renderer* r, int max_entries, int64_t* out_segments, double* out_extend_ms);
/* Acceleration structure of the loaded scene (replaces the BVH producers Bvh/bvh.cpp:185-219,667-780,
 * 862-1047 + Kernel/bvh_morton_code_kernel.cu:298-346): facts about the last build and a structural
 * check of the DEVICE arrays (binary layout).  out_i[10]: node records, reachable inner nodes, leaves,
 * depth, valid (every triangle in exactly one leaf, vertices inside the leaf box, child boxes inside the
 * parent's), built on the GPU (1/0), level-synchronous rounds, sub-trees finished in shared memory, compressed 8-wide
 * nodes of the hybrid bounce-ray tree, wide tree collapsed on the GPU (1/0).
 * out_d[4]: builder milliseconds (CUDA events for the GPU builder), SAH cost, whole scene upload ms,
 * violations found.  ptb_bvh_leaf_labels: per triangle the smallest triangle id sharing its leaf. */
int ptb_bvh_info(ptb_renderer* r, int64_t* out_i, double* out_d);
int ptb_bvh_leaf_labels(ptb_renderer* r, int32_t* out_label);
/* raw host copy of the device arrays: out_i[0] node records x 16 floats (bvh.h layout #1), and the triangle id per leaf slot */
int ptb_bvh_download(ptb_renderer* r, float* out_nodes16, int32_t* out_leaf_order);
/* with option "count_traversal" = 1: out25[0] = most node visits of one ray in the last render, out25[1 + k] = rays with
 * floor(log2(visits + 1)) == k (binary-tree kernel) */
int ptb_get_traversal_histogram(ptb_renderer* r, int64_t* out25);
/* String options (every key ptb_set_option accepts; unknown keys fail).  None of them changes WHICH paths are traced except the
 * three estimator keys; images are bit-identical across all scheduling / tree options (tests/test_gpu_parity.py).
 *  scheduling
 *   "passes_in_flight"   "1".."64" (default 4)  passes traced as one wavefront batch; re-allocates the path state, resets the image
 *   "streams_in_flight"  "1".."8"  (default 4)  batches overlapped on separate CUDA streams (own path-state buffers each)
 *   "active_streams"     "0" = all (default) | n  use only the first n stream contexts ("1": serial batches, clean per-kernel timing)
 *   "tile_order"         "1" (default) | "0"    camera rays enter the first queue in 8x4 pixel tiles / row-major
 *   "octant_order"       "0" (default) | "1"    next-depth queue grouped by ray-direction octant per block (measured slower)
 *   "sort_by_material"   "0" (default) | "1"    block-local material sort of the shade queue (measured 3-7 % slower, off)
 *  acceleration structure (take effect at the next ptb_load_scene / geometry edit)
 *   "bvh_builder"        "gpu_sah" (default, csrc/bvh_build.cu) | "host_sah" (csrc/bvh_host.cpp, cross-check / fallback)
 *   "bvh_layout"         "2" (default: 64-byte binary nodes) | "8" (80-byte compressed 8-wide nodes only)
 *   "bvh_hybrid"         "1" (default) | "0"    layout 2 only: also build the compressed wide tree for the deeper bounces
 *   "hybrid_from_depth"  "2" (default)          first bounce depth traced over the wide tree (immediate)
 *   "small_tree_bytes"   bytes (default 64 MiB) binary nodes + sibling records + triangles up to this size: every bounce stays on the binary
 *                                               tree with leaf starts ("upwalk"); an explicit hybrid_from_depth wins; "0" = off
 *   "bvh_collapse"       "gpu" (default) | "host"  where the binary tree is collapsed to 8-wide nodes
 *   "bvh_max_leaf"       "1".."8" (default 8)    triangles per leaf of the binary tree
 *   "bvh_intersect_cost" float > 0 (default 0.8) SAH cost of a triangle test relative to a node visit
 *   "l2_persist"         "0" (default) | "1"    nodes + leaf-order triangles in one allocation under a persisting L2 access-policy
 *                                               window on every render stream (measured neutral: the tree is never evicted)
 *  where a closest-hit search starts (same hits bit for bit, tests/test_gpu_entry.py; binary tree + persistent kernels only)
 *   "entry_cuts"         "1" (default) | "0"    camera rays start at the sub-trees their pixel tile's shaft touches instead of at the
 *                                               root (csrc/kernels_entry.cuh: k_entry_cut, rebuilt when camera / geometry change;
 *                                               cameras it does not cover — focal distance <= 0, fov >= 175 degrees — use the root)
 *   "entry_k"            "1".."31" (default 15) sub-trees per tile; "entry_tile" "WxH" powers of two (default "8x4")
 *   "entry_min_passes"   (default 4)            a batch of fewer passes builds the lists only for a camera the previous batch had too: a
 *                                               host that moves the camera with every single-pass call is served from the root (building
 *                                               the lists costs ~0.4 ms at 1080p, they save ~0.15 ms per pass)
 *   "sky_fast"           "1" (default) | "0"    with entry cuts: camera rays of tiles whose cut is EMPTY are finished by k_generate with the
 *                                               background colour (never queued, searched or shaded); needs tile_order, 8x4 tiles, a
 *                                               resolution that is a multiple of them, no spheres and air that does not participate
 *   "upwalk"             "1" (default) | "0"    a bounce ray that leaves a triangle starts at that triangle's leaf and collects the
 *                                               siblings of the leaf's ancestors it hits walking UP (k_up_level / k_up_pair records;
 *                                               takes effect at the next ptb_load_scene / geometry edit); "upwalk_min_nodes" (64):
 *                                               smaller trees are searched from the root
 *   "fused_upwalk"       "0" (default) | "1"    scattering media: the whole subsurface walk on the binary tree with every search
 *                                               started at the leaf the path entered through (measured slower than the wide tree)
 *  closest-hit kernels (immediate; tuning knobs of tools/sweep_*.py, defaults are the measured optima)
 *   "extend_persistent"  "1" (default) | "0"    persistent warp-voting kernels / one ray per thread
 *   "extend_variant"     "0" (default) | "1" postponed leaves | "2" + leaf prefetch | "3" top of the tree in shared memory
 *                                               (all measured slower than 0 on c2, profiles/r02_experiments.md; same hits)
 *   "treelet_block", "treelet_nodes"            variant 3: threads per block (1024), nodes held in shared memory (1023)
 *   "tune_refill", "tune_leaf", "tune_reps"     binary-tree kernel: refill when >= N lanes idle (20), leaf phase when >= N lanes
 *                                               wait (6), node steps per node phase (6); "unroll_reps" "1" | "0"
 *   "tune_refill8", "tune_leaf8"                wide-tree kernel (12, 6)
 *   "tune_refill_e", "tune_leaf_e", "tune_reps_e"  camera-ray kernel k_extend_entry (28, 6, 6)
 *   "tune_refill_u", "tune_leaf_u", "tune_reps_u"  bounce-ray kernel k_extend_upwalk (20, 6, 6); "tune_refill_f", "tune_leaf_f": with fused_upwalk
 *   "inline_scatter"     "1" (default) | "0"    medium scatter events performed inside the closest-hit kernel of the deeper bounces;
 *                                               "fused_from_depth" (-1 = where the wide tree takes over), "tune_scatter" (8 lanes)
 *   "persistent_grid", "persistent_grid8"       blocks of the persistent launches (default: one resident wave)
 *  estimator (NOT parity modes: they change the samples, same expectation; tests/test_gpu_nee.py, test_gpu_rr.py, test_gpu_estimators.py)
 *   "estimator"          "reference" (default) | "nee"   next-event estimation + shadow rays (binary tree)
 *   "russian_roulette"   "0" (default) | "1"    from bounce 3 on
 *   "pass_clamp"         float (default -1 = the reference's 2 * MaxDepth)   diagnostic: per-pass clamp of the accumulation
 *   "sampler"            "reference" (default: hash-product seeds + minstd, Kernel/path_tracer_kernel.cu:35-44,324-325,415-416) | "pcg"
 *   "sss"                "reference" (default: free flight from sigma_s'.x only, :456-464) | "per_channel" (uniform channel pick +
 *                                               single-sample MIS over sigma_s'.xyz; equals the reference mode when the channels agree)
 *  filtering (takes effect at the next ptb_load_scene)
 *   "texture_filter"     "software" (default: the reference's filter, Core/texture.h:15-79, bit-compatible) | "hardware" (bilinear lookups
 *                                               of textures and cube-map faces by the texture unit on cudaArray copies; 9-bit weights)
 *  measurement
 *   "profile_stages"     "0" | "1"   CUDA events around every closest-hit launch (ptb_get_depth_profile, ptb_stats.gpu_ms_extend)
 *   "count_traversal"    "0" | "1"   instrumented kernels: node visits / triangle tests (ptb_stats, ptb_get_traversal_histogram)
 *  loader (process-wide)
 *   "loader_threads"     "0".."64"   slices an OBJ file is parsed in; 0 = by file size and host cores
 *   "loader_mesh_lanes"  "0".."8"    mesh files parsed at the same time; 0 = max(2, host cores / threads per file)
 *   "loader_per_vertex"  "1" (default) | "0" | "2"   world-space triangles from positions / normals transformed once per VERTEX (meshes of
 *                                    >= 65536 triangles; "2": every mesh) or once per triangle corner ("0"); the output is the same bit for bit */
int ptb_set_option(ptb_renderer* r, const char* key, const char* value);

/* ---- multi-GPU --------------------------------------------------------------------------------
 * The reference picks one device (Main/window.cpp:281-295) and calls path_tracer::render() once per pass
 * (Core/path_tracer.cpp:40-99).  A pass depends only on (scene, camera, configuration, pass number), so the box shards by pass
 * index: device k of G renders passes P+1+k, P+1+k+G, ... (each keeps its single-GPU seed), and ONE ncclReduce(sum, float,
 * 3*W*H) per image lands on the root, which holds the MERGED image (its own accumulation stays untouched, so rendering more
 * and reducing again is correct).  NCCL is dlopen'ed at first use (libnccl.so.2; override with the environment variable
 * PTB200_NCCL_LIB): no link-time dependency, single-GPU hosts never load it.  csrc/multi.inc.
 *
 * (a) ONE process, every GPU of the box — what the reference's C++ host would call instead of path_tracer::render():
 *     ptb_multi_create(config, n, devices-or-NULL) = n renderers + ncclCommInitAll; ptb_multi_load_scene parses the scene files
 *     ONCE and uploads / builds the BVH on every device concurrently (one host thread per device); ptb_multi_render(total)
 *     renders the next `total` passes of the image across the devices (blocking) and merges; ptb_multi_image_* read the merged
 *     image; ptb_multi_renderer(i) exposes the per-device handle (options, stats, edits). */
typedef struct ptb_multi ptb_multi;
ptb_multi* ptb_multi_create(const char* config_json_path, int n_devices, const int* devices);
void ptb_multi_destroy(ptb_multi* m);
int ptb_multi_device_count(ptb_multi* m);
ptb_renderer* ptb_multi_renderer(ptb_multi* m, int index);
int ptb_multi_set_option(ptb_multi* m, const char* key, const char* value);
int ptb_multi_load_scene(ptb_multi* m, const char* scene_json_path, const char* asset_root);
int ptb_multi_set_camera(ptb_multi* m, const ptb_camera* cam);
int ptb_multi_render(ptb_multi* m, int total_passes);
int ptb_multi_clear(ptb_multi* m);
int ptb_multi_pass_counter(ptb_multi* m);
int ptb_multi_image_f32(ptb_multi* m, float* out_rgb_sum, int* out_passes);
int ptb_multi_image_u8(ptb_multi* m, uint8_t* out_rgb);
/* (b) one process per GPU (MPI / torchrun): rank 0 calls ptb_dist_unique_id, the host hands the 128 bytes to every rank however
 *     it likes, every rank calls ptb_dist_init(r, rank, world, id).  ptb_dist_broadcast_scene(r, root): the root has loaded the
 *     scene, every other rank receives the PARSED scene (world-space triangles, materials, spheres, textures, cube map) through
 *     one ncclBroadcast and builds its own BVH — no rank but the root reads the scene files.  ptb_dist_render(r, total): this
 *     rank's share of the next `total` passes (same `total` on every rank).  ptb_dist_reduce(r, root): collective; afterwards
 *     ptb_merged_image_f32 / _u8 on the root return the merged sum / the finished 8-bit image (out_passes = passes of all ranks).
 *     ptb_nccl_version: NCCL's version code, 0 if it cannot be loaded. */
int ptb_nccl_version(void);
int ptb_dist_unique_id(void* out_id128);
int ptb_dist_init(ptb_renderer* r, int rank, int world_size, const void* id128);
int ptb_dist_shutdown(ptb_renderer* r);
int ptb_dist_broadcast_scene(ptb_renderer* r, int root);
/* collective load: `root` reads and parses the files (the other ranks' path arguments are ignored), then every rank — the root included —
 * uploads and builds its BVH from the broadcast device copy at the same time */
int ptb_dist_load_scene(ptb_renderer* r, const char* scene_json_path, const char* asset_root, int root);
/* wall ms of the last broadcast on this rank: staging on the root, ncclBroadcast, device->host unpack, upload + BVH build, whole call; out6[5] = bytes */
int ptb_dist_broadcast_timing(ptb_renderer* r, double* out6);
int ptb_dist_render(ptb_renderer* r, int total_passes);
int ptb_dist_reduce(ptb_renderer* r, int root);
int ptb_dist_clear(ptb_renderer* r);
int ptb_merged_image_f32(ptb_renderer* r, float* out_rgb_sum, int* out_passes);
int ptb_merged_image_u8(ptb_renderer* r, uint8_t* out_rgb);
/* test hook: the loaded scene through the broadcast format and back on this device; 0 = byte-identical */
int ptb_test_scene_blob_roundtrip(ptb_renderer* r);
/* test hook (host only): first pass index and pass count of rank `rank` of `world` for the next `total` passes after `global_done` */
int ptb_test_shard(int global_done, int total, int rank, int world, int* out_first, int* out_count);

/* ---- output side -----------------------------------------------------------------------------
 * ptb_save_png        = screenshot() of Main/window.cpp:712-740 (lodepng::encode of the displayed RGBA8 image),
 *                       without the GL read-back: the 8-bit image of pixel_256_transform_gamma_corrected_kernel.
 * ptb_save_pfm        = the float mean image (accumulation / pass_counter) as a little-endian colour PFM.
 * ptb_save_checkpoint / ptb_load_checkpoint = the reference `image` state that defines a render in progress
 *                       (Core/image.h:10-23: pixels + pass_counter) with a CRC, so a render continues with pass
 *                       pass_counter+1 after a restart and produces the SAME image as an uninterrupted run; loading
 *                       fails if resolution or MaxDepth differ.  restore_camera != 0 also restores the camera.
 * ptb_write_png_rgb8 / ptb_write_pfm: the same writers on caller-supplied HOST buffers (row 0 = top). */
int ptb_save_png(ptb_renderer* r, const char* path);
int ptb_save_pfm(ptb_renderer* r, const char* path);
int ptb_save_checkpoint(ptb_renderer* r, const char* path);
int ptb_load_checkpoint(ptb_renderer* r, const char* path, int restore_camera);
int ptb_write_png_rgb8(const char* path, const uint8_t* rgb, int width, int height);
int ptb_write_pfm(const char* path, const float* rgb, int width, int height, float scale);

/* ---- live scene edits -----------------------------------------------------------------------
 * What path_tracer::render_ui does between passes (Core/path_tracer.cpp:109-369) through
 * scene_parser::set_sphere_device / set_mesh_material_device / set_mesh_transform_device /
 * set_mesh_rotate + apply_mesh_rotate (Core/scene_parser.cpp:645-673, Core/triangle_mesh.cpp:252-426),
 * followed by clear().  Same arithmetic as the reference (world = T*S * rotated-local, normals through the
 * inverse transpose, re-normalised); the acceleration structure is REBUILT on the device instead of
 * re-transforming stale boxes (Bvh/bvh.cpp:332-356).  sphere100: center[3], radius, ptb_material.
 * ptb_set_mesh_material ignores a list whose length differs from the mesh's material count, like the
 * reference.  The scale is clamped to >= 1e-6 like the UI does.  Every edit resets the accumulation. */
int ptb_set_sphere(ptb_renderer* r, int index, const void* sphere100);
int ptb_set_mesh_material(ptb_renderer* r, int mesh, const ptb_material* mats, int n);
int ptb_set_mesh_transform(ptb_renderer* r, int mesh, const float* position3, const float* scale3);
int ptb_apply_mesh_rotate(ptb_renderer* r, int mesh, const float* rotate_degrees3);
/* live config toggles: the reference's UI edits the managed `configuration` (Core/configuration.h:9-34) between passes and clears;
 * config96 = that 96-byte struct (ptb_get_config returns the current one).  Width / Height / MaxDepth may not change. */
int ptb_set_config(ptb_renderer* r, const void* config96);
int ptb_get_mesh_placement(ptb_renderer* r, int mesh, float* out_position3, float* out_scale3, float* out_rotate3, int* out_first_triangle, int* out_triangle_count, int* out_first_material, int* out_material_count);

/* ---- loaded-scene introspection (flat host copies; used by the parity tests) ------------- */
int ptb_scene_counts(ptb_renderer* r, int* n_triangles, int* n_materials, int* n_spheres, int* n_textures, int* cube_length, int* n_meshes);
/* per triangle 24 floats: v0 v1 v2 n0 n1 n2 uv0 uv1 uv2 (Core/triangle.h:11-25) + material index */
int ptb_scene_triangles(ptb_renderer* r, float* out24, int32_t* out_material);
int ptb_scene_materials(ptb_renderer* r, ptb_material* out);
/* per sphere: center[3], radius, then ptb_material (Core/sphere.h:11-16) = 100 bytes */
int ptb_scene_spheres(ptb_renderer* r, void* out100);
int ptb_scene_texture(ptb_renderer* r, int index, int* width, int* height, uint8_t* out_rgba /* may be NULL */);
int ptb_scene_cubemap_face(ptb_renderer* r, int face, uint8_t* out_rgba);
/* image_loader::load_image (Others/image_loader.cpp:31-95: FreeImage_Load + ConvertTo24Bits, alpha forced to 255,
 * row 0 = top) on one file.  Decoded natively: BMP (24/32-bit), TGA, PNG (plain or Adam7-interlaced; grey / RGB / palette /
 * alpha dropped / 16-bit reduced to the high byte), baseline and progressive JPEG (1 or 3 components, 4:4:4 / 4:2:2 / 4:2:0);
 * a "<file>.rgba8" side-car (u32 width, u32 height, RGBA8) takes precedence when present and serves every other format.
 * out_rgba may be NULL to query the size. */
int ptb_decode_image(const char* path, int* width, int* height, uint8_t* out_rgba);
/* JPEG decode parameters, process-wide, for every later load:
 *   "reference" (default) what the reference gets from FreeImage_Load(fif, name, 0): flags 0 = JPEG_FAST
 *               (lib/free_image/FreeImage.h:693-695) -> libjpeg 9a with the "ifast" IDCT and replicated chroma;
 *   "fast"      libjpeg-turbo with dct_method=JDCT_IFAST, do_fancy_upsampling=FALSE (bit-identical; differs from "reference" by
 *               one colour-table constant);
 *   "accurate"  libjpeg-turbo's default decode (bit-identical to PIL). */
int ptb_set_jpeg_decode(const char* mode);
/* the parsed configuration as the reference's 96-byte `configuration` (Core/configuration.h:9-34) */
int ptb_get_config(ptb_renderer* r, void* out96);

/* ---- reference-signature compatibility entries --------------------------------------------
 * The three extern "C" symbols Core/path_tracer.cpp links against (Core/path_tracer_kernel.h:18-54;
 * Kernel/path_tracer_kernel.cu:685,782,798), with the reference's own argument lists and struct
 * layouts, so its host code can call this library unchanged: the AoS managed
 * buffers (triangle 104 B, bvh_node_device 56 B, sphere 100 B, material 84 B, cube_map 56 B,
 * texture_wrapper 16 B, configuration 96 B) are ingested and cached, the pass with seed
 * `pass_counter` is traced by the wavefront kernels, and image_pixels (float3 running sum, "=" on pass
 * 1, "+=" otherwise), image_pixels_256 and accumulated_colors are written like
 * pixel_256_transform_gamma_corrected_kernel does.  Synchronous for the pass asked for; logs in the
 * reference's "[Cuda]Error ..." format and returns on failure.  Passes are rendered ahead of the calls
 * (environment variables PTB_COMPAT_LOOKAHEAD = passes per batch, default 8, 1 = off; PTB_COMPAT_CONTEXTS = batches in flight, default 4): a call for the
 * next pass number with unchanged camera / configuration / scene is answered from a finished batch
 * while later ones render; any other call starts over from the caller's image_pixels, so every call
 * leaves what one pass per call would have left (csrc/compat.inc).  Declared with opaque pointers here. */
#ifndef PTB200_BUILDING_LIBRARY
void path_tracer_kernel(
	int mesh_num, void** bvh_nodes_device, void* triangles_device, int sphere_num, void* spheres_device,
	int pixel_count, float* image_pixels, uint8_t* image_pixels_256, int pass_counter,
	const ptb_camera* render_camera_device, void* sky_cube_map_device,
	float* not_absorbed_colors_device, float* accumulated_colors_device, void* rays_device,
	int* energy_exist_pixels_device, void* scatterings_device, void* mesh_textures_device, void* config_device);
void path_tracer_kernel_memory_allocate(float** not_absorbed_colors_device, float** accumulated_colors_device, void** rays_device,
	int** energy_exist_pixels_device, void** scatterings_device, int pixel_count);
void path_tracer_kernel_memory_free(float* not_absorbed_colors_device, float* accumulated_colors_device, void* rays_device,
	int* energy_exist_pixels_device, void* scatterings_device);
#endif

#ifdef __cplusplus
}
#endif
#endif /* PTB200_H */
