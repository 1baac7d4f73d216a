// ptb200 wavefront integrator: kernels, renderer state and the C ABI (include/ptb200.h).
//
// The reference runs, per pass, init -> generate -> { trace_ray_kernel ; thrust::remove_if +
// host sync } x depth -> tonemap (Kernel/path_tracer_kernel.cu:706-779).  Here one batch of
// `passes_in_flight` passes is traced as a single wavefront:
//     k_generate -> { k_extend ; k_shade } x depth -> k_accumulate        (no host round trips)
// with SoA path state, device-side queue counters, warp-ballot compaction into the next-depth
// queue and per-(pass, pixel, depth) stateless RNG identical to the reference's, so every
// sample's value is the one the reference computes for that (pass, pixel).
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#define PTB200_BUILDING_LIBRARY 1
#include "../../include/ptb200.h"
#include "scene.h"
#include "bvh.h"
#include "bvh_gpu.h"
#include "image_out.h"
#include "kernels.cuh"

using namespace ptbdev;

namespace ptb
{

#define PTB_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	set_error(std::string("[Cuda]Error in file '") + __FILE__ + "' in line " + std::to_string(__LINE__) + " : " + cudaGetErrorString(e_)); return 1; } } while (0)

} // namespace ptb

#include "kernels_generate.cuh"
#include "kernels_entry.cuh"
#include "kernels_extend.cuh"
#include "kernels_shade.cuh"

#ifdef PTB_EXPERIMENT_SORT
// EXPERIMENT ONLY (tools/build_variants.sh "sortexp", never in the shipped library): upper bound of what re-ordering a bounce queue by
// origin cell + direction octant can buy the closest-hit kernel — a full radix sort by a library, its cost ignored.
#include <cub/cub.cuh>
namespace ptb
{
__device__ __forceinline__ unsigned spread3(unsigned v) { v &= 0x3ffu; v = (v | (v << 16)) & 0x030000ffu; v = (v | (v << 8)) & 0x0300f00fu; v = (v | (v << 4)) & 0x030c30c3u; v = (v | (v << 2)) & 0x09249249u; return v; }
__global__ void k_sort_keys(PathState st, const int* __restrict__ queue, const int* __restrict__ count_ptr, unsigned* __restrict__ keys, int total, float3 lo, float3 inv_extent, int cell_bits, int use_octant)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= total) return;
	if (i >= *count_ptr) { keys[i] = 0xffffffffu; return; }
	const int id = queue[i];
	const float4 o = st.ray_o[id], d = st.ray_d[id];
	const unsigned scale = 1u << cell_bits;
	const unsigned cx = min(scale - 1u, (unsigned)max(0.0f, (o.x - lo.x) * inv_extent.x * scale));
	const unsigned cy = min(scale - 1u, (unsigned)max(0.0f, (o.y - lo.y) * inv_extent.y * scale));
	const unsigned cz = min(scale - 1u, (unsigned)max(0.0f, (o.z - lo.z) * inv_extent.z * scale));
	const unsigned morton = spread3(cx) | (spread3(cy) << 1) | (spread3(cz) << 2);
	const unsigned oct = use_octant ? ((d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u)) : 0u;
	keys[i] = use_octant == 2 ? ((morton << 3) | oct) : ((oct << (3 * cell_bits)) | morton);
}
}
#endif

// ==========================================================================================
// Renderer
// ==========================================================================================

using namespace ptb;

#define PTB_NODE_REPS 6

struct ptb_renderer
{
	Config cfg;
	int device = -1;
	bool host_only = false;
	bool scene_loaded = false;
	HostScene scene;
	ptb_camera cam;
	int pass_counter = 0;

	// options
	int passes_in_flight = 4;
	int profile_stages = 0;
	int count_traversal = 0;
	std::string bvh_builder = "gpu_sah";   // "gpu_sah" (csrc/bvh_build.cu) | "host_sah" (csrc/bvh_host.cpp)
	int bvh_layout = 2;
	int bvh_hybrid = 1;                    // layout 2 only: also build the compressed 8-wide tree and trace depth >= hybrid_from_depth with it
	int hybrid_from_depth = 2;             // measured: c2 -0.4 %, c3 +1.2 %, c4 +8.4 % against binary-only (from depth 1: c2 -2.8 %)
	bool hybrid_from_user = false;         // hybrid_from_depth was set by the caller: small_tree_bytes does not override it
	int64_t small_tree_bytes = 64 << 20;   // binary nodes + sibling records + triangles up to this size: EVERY bounce stays on the binary tree with leaf starts
	                                       // (c2, 21 MB: +1.8 %; c3, 43 MB: +1 %; c5, 720 MB: -7 % — the compact wide tree wins once the tree leaves the caches); 0 = off
	int64_t bvh8_nodes = 0;
	int bvh_max_leaf = 8;                  // binary layout; the wide layout holds <= 3 per leaf slot
	float bvh_intersect_cost = 0.8f;       // SAH cost of a triangle test relative to a node visit (measured optimum on c2, profiles/r01_experiments.md)
	int russian_roulette = 0;              // estimator option (absent in the reference): roulette from bounce 3 on, see kernels_shade.cuh
	float pass_clamp = -1.0f;              // < 0: the reference's per-pass clamp 2 * MaxDepth (path_tracer_kernel.cu:644-651)
	int nee = 0;                           // estimator: 0 = the reference's (default, parity mode), 1 = next-event estimation
	int sort_by_material = 0;              // block-local material sort in k_shade (measured: profiles/r01_experiments.md)
	int octant_order = 0;                  // next-depth queue grouped by ray-direction octant per block (measured: profiles/r01_experiments.md)
	int tile_order = 1;                    // camera rays enter the first queue in 8x4 pixel tiles
	int inline_scatter = 1;                // medium scatter events performed inside the wide-tree closest-hit kernel (bit-identical images; off: one wavefront round trip per event)
	int fused_from_depth = -1;             // first loop depth with inline scatter events (-1: where the wide tree takes over)
	int tune_scatter = 8;                  // ... as a voted phase once >= N lanes wait for one
	bool scene_has_medium = false;         // some material (or the air) scatters: sigma_s'.x > 0
	int sampler = 0;                       // estimator option: 0 = the reference's hash-product + minstd streams, 1 = pcg (pt_device.cuh)
	int sss_mode = 0;                      // estimator option: 1 = per-channel subsurface scattering (kernels_shade.cuh)
	int hw_textures = 0;                   // texture_filter=hardware: bilinear lookups by the texture unit on cudaArray copies (takes effect at load)
	std::vector<cudaArray_t> texture_arrays; std::vector<cudaTextureObject_t> texture_objects;
	int l2_persist = 0;                    // 1: binary nodes + leaf-order triangles in one arena under a persisting L2 access-policy window (measured: profiles/r02_experiments.md)
	void* l2_arena = nullptr; size_t l2_arena_bytes = 0;
	// facts about the last acceleration-structure build (ptb_bvh_info)
	int bvh_collapsed_on_gpu = 0;
	std::string bvh_collapse = "gpu";      // "gpu" (csrc/bvh_build.cu: k_collapse8) | "host" (csrc/bvh_host.cpp: build_bvh8)
	int bvh_built_on_gpu = 0, bvh_levels = 0, bvh_small_tasks = 0, bvh_max_depth = 0;
	double bvh_build_ms = 0.0, scene_upload_ms = 0.0;
	std::string bvh_note;
	int extend_persistent = 1;
	// entry cuts (kernels_entry.cuh): camera rays start at the sub-trees their 8x4 pixel tile's shaft touches instead of at the root.
	// Same hits bit for bit; the lists depend on camera + geometry only and are rebuilt when either changes.
	int upwalk_min_nodes = 64;
	int upwalk = 1;                        // bounce rays on the binary tree start at the leaf of the triangle they leave (kernels_entry.cuh: k_up_level); same hits
	int entry_cuts = 1;
	int sky_fast = 1;                      // camera rays of tiles with an empty entry cut are finished by k_generate (background colour), never queued
	int entry_k = 15;                      // sub-trees per tile (<= PTB_ENTRY_STRIDE - 1)
	int entry_tile_w = 8, entry_tile_h = 4; // pixels per tile (powers of two); 8x4 = the 32 lanes of a warp under tile_order
	int fused_upwalk = 0;                  // scattering media: the whole walk on the binary tree, every search started at the leaf of the triangle the path entered through (measured SLOWER than the wide tree from the root: c4 807 -> 716-746 Msamples/s, profiles/r02_experiments.md)
	int tune_refill_f = 20, tune_leaf_f = 6;   // voting thresholds of k_extend_persistent_fused<.., UPWALK>
	int tune_refill_u = 20, tune_leaf_u = 6, tune_reps_u = 6;   // voting thresholds of k_extend_upwalk
	int tune_refill_e = 28, tune_leaf_e = 6, tune_reps_e = 6;   // voting thresholds of k_extend_entry (short searches: refills batched harder than in the other kernels; swept in profiles/r02_experiments.md)
	int2* entry_buf = nullptr; size_t entry_buf_slots = 0; int entry_stride = 16;
	int* entry_rank = nullptr; size_t entry_rank_tiles = 0;   // k_tile_rank: [n_tiles] ranks + 1 int (non-empty tiles)
	bool entry_valid = false;
	size_t entry_n_tiles = 0;
	unsigned char entry_key[96] = { 0 };   // camera, resolution, k and geometry version the lists were built for
	unsigned char entry_last_key[96] = { 0 };   // ... of the previous batch (ensure_entry_cuts)
	int entry_min_passes = 4;              // a batch of fewer passes builds the lists only for a camera it has seen before
	uint64_t geometry_version = 0;
	int tune_refill4 = 8;                  // extend_variant 4: pop staged rays when >= N lanes are idle
	int treelet_block = 1024, treelet_nodes = 1023;   // extend_variant 3: threads per block and top-of-tree nodes held in shared memory (64 B each)
	int sort_depth_mask = 0, sort_cell_bits = 5, sort_octant = 1;   // PTB_EXPERIMENT_SORT builds only
	unsigned* sort_keys[2] = { nullptr, nullptr }; int* sort_vals = nullptr; void* sort_tmp = nullptr; size_t sort_tmp_bytes = 0;
	float scene_lo[3] = { 0, 0, 0 }, scene_hi[3] = { 1, 1, 1 };
	int extend_variant = 0;                // binary-tree kernel: 0 = k_extend_persistent, 1 = k_extend_speculative (postponed leaves), 2 = + leaf prefetch
	int persistent_grid = 148 * 4;
	int persistent_grid8 = 148 * 4;
	// voting thresholds of the persistent kernels (tools/sweep_tune.py, profiles/r01_experiments.md): refill when >= N lanes are idle,
	// run a leaf / triangle phase when >= N lanes wait for one, node steps per node phase
	int tune_refill = 20, tune_leaf = 6, tune_reps = 6;   // binary-tree kernel (camera rays + first bounce)
	int unroll_reps = 1;
	int tune_refill8 = 12, tune_leaf8 = 6;                 // wide-tree kernel (deep bounces)

	cudaStream_t stream = nullptr;
	int sm_count = 148;

	// scene on device
	DeviceScene dscene;
	std::vector<void*> scene_allocs;      // textures, cube map
	std::vector<void*> geometry_allocs;   // triangles, BVH, shading attributes (rebuilt by mesh edits)
	std::vector<void*> material_allocs;   // materials + spheres (rewritten by material / sphere edits)
	std::vector<void*> light_allocs;      // emissive-triangle list of the "nee" estimator
	int64_t bvh_nodes = 0, bvh_bytes = 0;

	// work buffers
	int pixel_count = 0;
	size_t capacity = 0;       // paths = pixel_count * passes_in_flight
	PathState st;
	int* queue[2] = { nullptr, nullptr };
	int* counts = nullptr;     // max_depth + 2 ints
	unsigned long long* counts_host = nullptr; // pinned copy of segment_totals
	unsigned long long* counters = nullptr;
	float* image_sum = nullptr;
	float* last_pass = nullptr;
	uint8_t* image_u8 = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	std::vector<cudaEvent_t> stage_events;
	// Batches are issued round-robin over `streams_in_flight` contexts (own stream + own path-state
	// buffers) so the long-tail rays and tiny deep-bounce launches of one batch overlap with the
	// bulk of another.  Context 0 is (stream, st, queue, counts) above.
	struct BatchContext
	{
		cudaStream_t stream = nullptr;
		PathState st;
		int* queue[2] = { nullptr, nullptr };
		int* counts = nullptr;
		cudaEvent_t accumulated = nullptr;   // recorded after the context's k_accumulate
	};
	int streams_in_flight = 4;
	int active_streams = 0;                  // 0 = all contexts; 1 = serialise batches (clean per-kernel timing)
	std::vector<BatchContext> contexts;      // size streams_in_flight; [0] aliases the members above
	unsigned long long* segment_totals = nullptr;   // per-depth live-path totals of the current call (device)

	// multi-GPU (csrc/multi.inc): NCCL communicator of this device, its rank, and the merged image a reduce leaves on the root
	void* nccl_comm = nullptr;
	int dist_rank = 0, dist_world = 1, dist_global_passes = 0;
	float* merged_sum = nullptr;
	uint8_t* merged_u8 = nullptr;
	int merged_passes = 0;
	unsigned long long* pass_count_dev = nullptr;
	const float* staged_tris24 = nullptr; size_t staged_tris_count = 0;   // device copy of the world-space triangles upload_geometry may adopt
	bool root_parse_ok = false;            // ptb_dist_load_scene: this rank parsed the scene files and has not uploaded them yet
	double bcast_ms[5] = { 0, 0, 0, 0, 0 }, bcast_bytes = 0.0;   // timing of the last ptb_dist_broadcast_scene on this rank

	ptb_stats stats;
	int64_t traversal_histogram[32] = { 0 };  // raw device counters of the last count_traversal call
	std::vector<double> depth_extend_ms;   // per-depth sums of the last call (profile_stages)
	std::vector<int64_t> depth_segments;
};

namespace
{

int alloc_path_state(ptb_renderer* r, PathState& st, int* queue[2], int** counts)
{
	const int n_counts = r->cfg.max_tracer_depth + 2;
	PTB_CUDA(cudaMalloc(&st.ray_o, r->capacity * sizeof(float4)));
	PTB_CUDA(cudaMalloc(&st.ray_d, r->capacity * sizeof(float4)));
	PTB_CUDA(cudaMalloc(&st.throughput, r->capacity * sizeof(float4)));
	PTB_CUDA(cudaMalloc(&st.radiance, r->capacity * sizeof(float4)));
	PTB_CUDA(cudaMalloc(&st.hit, r->capacity * sizeof(float4)));
	PTB_CUDA(cudaMalloc(&queue[0], r->capacity * sizeof(int)));
	PTB_CUDA(cudaMalloc(&queue[1], r->capacity * sizeof(int)));
	PTB_CUDA(cudaMalloc(counts, 3 * n_counts * sizeof(int)));
	if (r->nee)
	{
		PTB_CUDA(cudaMalloc(&st.shadow_o, r->capacity * sizeof(float4)));
		PTB_CUDA(cudaMalloc(&st.shadow_d, r->capacity * sizeof(float4)));
		PTB_CUDA(cudaMalloc(&st.shadow_c, r->capacity * sizeof(float4)));
		PTB_CUDA(cudaMalloc(&st.shadow_queue, r->capacity * sizeof(int)));
	}
	return 0;
}

void free_path_state(PathState& st, int* queue[2], int** counts)
{
	cudaFree(st.ray_o); cudaFree(st.ray_d); cudaFree(st.throughput); cudaFree(st.radiance); cudaFree(st.hit);
	cudaFree(st.shadow_o); cudaFree(st.shadow_d); cudaFree(st.shadow_c); cudaFree(st.shadow_queue);
	cudaFree(queue[0]); cudaFree(queue[1]); cudaFree(*counts);
	st = PathState(); queue[0] = queue[1] = nullptr; *counts = nullptr;
}

void apply_l2_window(ptb_renderer* r);

int alloc_work_buffers(ptb_renderer* r)
{
	r->pixel_count = r->cfg.width * r->cfg.height;
	r->capacity = (size_t)r->pixel_count * r->passes_in_flight;
	const int n_counts = r->cfg.max_tracer_depth + 2;
	if (alloc_path_state(r, r->st, r->queue, &r->counts)) return 1;
	r->contexts.assign(std::max(1, r->streams_in_flight), ptb_renderer::BatchContext());
	for (size_t c = 0; c < r->contexts.size(); c++)
	{
		ptb_renderer::BatchContext& ctx = r->contexts[c];
		if (c == 0)
		{
			ctx.stream = r->stream; ctx.st = r->st; ctx.queue[0] = r->queue[0]; ctx.queue[1] = r->queue[1]; ctx.counts = r->counts;
		}
		else
		{
			PTB_CUDA(cudaStreamCreateWithFlags(&ctx.stream, cudaStreamNonBlocking));
			if (alloc_path_state(r, ctx.st, ctx.queue, &ctx.counts)) return 1;
		}
		PTB_CUDA(cudaEventCreateWithFlags(&ctx.accumulated, cudaEventDisableTiming));
	}
	PTB_CUDA(cudaMallocHost(&r->counts_host, n_counts * sizeof(unsigned long long)));
	PTB_CUDA(cudaMalloc(&r->segment_totals, n_counts * sizeof(unsigned long long)));
	PTB_CUDA(cudaMalloc(&r->counters, 32 * sizeof(unsigned long long)));
	PTB_CUDA(cudaMalloc(&r->image_sum, (size_t)r->pixel_count * 3 * sizeof(float)));
	PTB_CUDA(cudaMalloc(&r->last_pass, (size_t)r->pixel_count * 3 * sizeof(float)));
	PTB_CUDA(cudaMalloc(&r->image_u8, (size_t)r->pixel_count * 3));
	PTB_CUDA(cudaMemsetAsync(r->image_sum, 0, (size_t)r->pixel_count * 3 * sizeof(float), r->stream));
	PTB_CUDA(cudaMemsetAsync(r->last_pass, 0, (size_t)r->pixel_count * 3 * sizeof(float), r->stream));
	PTB_CUDA(cudaMemsetAsync(r->image_u8, 0, (size_t)r->pixel_count * 3, r->stream));
	PTB_CUDA(cudaMemsetAsync(r->counters, 0, 32 * sizeof(unsigned long long), r->stream));
	PTB_CUDA(cudaStreamSynchronize(r->stream));
	apply_l2_window(r);   // new streams: re-apply the access-policy window of option l2_persist
	return 0;
}

void free_work_buffers(ptb_renderer* r)
{
	for (size_t c = 0; c < r->contexts.size(); c++)
	{
		ptb_renderer::BatchContext& ctx = r->contexts[c];
		if (ctx.accumulated) cudaEventDestroy(ctx.accumulated);
		if (c == 0) continue;
		if (ctx.stream) { cudaStreamSynchronize(ctx.stream); cudaStreamDestroy(ctx.stream); }
		free_path_state(ctx.st, ctx.queue, &ctx.counts);
	}
	r->contexts.clear();
	free_path_state(r->st, r->queue, &r->counts);
	cudaFree(r->counters); cudaFree(r->segment_totals);
	if (r->counts_host) cudaFreeHost(r->counts_host);
	cudaFree(r->image_sum); cudaFree(r->last_pass); cudaFree(r->image_u8);
	cudaFree(r->entry_buf); r->entry_buf = nullptr; r->entry_buf_slots = 0; r->entry_valid = false;
	cudaFree(r->entry_rank); r->entry_rank = nullptr; r->entry_rank_tiles = 0;
	cudaFree(r->merged_sum); cudaFree(r->merged_u8); cudaFree(r->pass_count_dev);
	r->merged_sum = nullptr; r->merged_u8 = nullptr; r->pass_count_dev = nullptr; r->merged_passes = 0;
	r->counts_host = nullptr; r->counters = nullptr; r->segment_totals = nullptr;
	r->image_sum = nullptr; r->last_pass = nullptr; r->image_u8 = nullptr;
}

// Large host <-> device copies of PAGEABLE memory (the parsed scene lives in std::vectors).  The driver stages such copies through
// its own small pinned buffer on one thread (measured on the B200 boxes: 0.6-1 GB/s for the 480 MB triangle array of c5); here the copy
// is cut into 32 MB chunks that host threads memcpy into / out of two pinned buffers while the previous chunk's DMA runs.
struct PinnedStager
{
	static const size_t kChunk = 32u << 20;
	char* buf[2] = { nullptr, nullptr };
	cudaEvent_t ev[2] = { nullptr, nullptr };
	bool ok = false;
	bool init()
	{
		if (ok) return true;
		for (int i = 0; i < 2; i++)
		{
			if (cudaMallocHost(&buf[i], kChunk) != cudaSuccess || cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) { release(); cudaGetLastError(); return false; }
		}
		ok = true;
		return true;
	}
	void release()
	{
		for (int i = 0; i < 2; i++) { if (buf[i]) cudaFreeHost(buf[i]); if (ev[i]) cudaEventDestroy(ev[i]); buf[i] = nullptr; ev[i] = nullptr; }
		ok = false;
	}
	~PinnedStager() { release(); }
};

void parallel_memcpy(char* dst, const char* src, size_t bytes)
{
	const size_t hw = std::max(1u, std::thread::hardware_concurrency());
	const size_t n = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(8, hw), bytes / (4u << 20)));
	if (n == 1) { memcpy(dst, src, bytes); return; }
	std::vector<std::thread> th;
	const size_t part = ((bytes + n - 1) / n + 4095) & ~(size_t)4095;
	for (size_t k = 1; k < n; k++)
	{
		const size_t a = std::min(bytes, k * part), b = std::min(bytes, (k + 1) * part);
		if (b > a) th.emplace_back([=] { memcpy(dst + a, src + a, b - a); });
	}
	memcpy(dst, src, std::min(bytes, part));
	for (auto& t : th) t.join();
}

// to_device: host (pageable) -> device; else device -> host (pageable).  Synchronous with respect to the host buffer; ordered on `stream`.
int staged_copy(void* dst, const void* src, size_t bytes, bool to_device, cudaStream_t stream)
{
	static thread_local PinnedStager stager;
	if (bytes < (8u << 20) || !stager.init())
	{
		PTB_CUDA(cudaMemcpyAsync(dst, src, bytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, stream));
		PTB_CUDA(cudaStreamSynchronize(stream));
		return 0;
	}
	const size_t chunk = PinnedStager::kChunk;
	const size_t n_chunks = (bytes + chunk - 1) / chunk;
	if (to_device)
	{
		for (size_t c = 0; c < n_chunks; c++)
		{
			const int b = (int)(c & 1);
			const size_t off = c * chunk, n = std::min(chunk, bytes - off);
			if (c >= 2) PTB_CUDA(cudaEventSynchronize(stager.ev[b]));      // the DMA that last read this buffer is done
			parallel_memcpy(stager.buf[b], (const char*)src + off, n);
			PTB_CUDA(cudaMemcpyAsync((char*)dst + off, stager.buf[b], n, cudaMemcpyHostToDevice, stream));
			PTB_CUDA(cudaEventRecord(stager.ev[b], stream));
		}
		PTB_CUDA(cudaStreamSynchronize(stream));
	}
	else
	{
		// chunk c + 1 is in flight on the copy engine while the host threads drain chunk c
		for (size_t c = 0; c <= n_chunks; c++)
		{
			if (c < n_chunks)
			{
				const int b = (int)(c & 1);
				const size_t off = c * chunk, n = std::min(chunk, bytes - off);
				PTB_CUDA(cudaMemcpyAsync(stager.buf[b], (const char*)src + off, n, cudaMemcpyDeviceToHost, stream));
				PTB_CUDA(cudaEventRecord(stager.ev[b], stream));
			}
			if (c >= 1)
			{
				const int b = (int)((c - 1) & 1);
				const size_t off = (c - 1) * chunk, n = std::min(chunk, bytes - off);
				PTB_CUDA(cudaEventSynchronize(stager.ev[b]));
				parallel_memcpy((char*)dst + off, stager.buf[b], n);
			}
		}
	}
	return 0;
}

template <class T>
int upload(ptb_renderer* r, const T* host, size_t count, const T** out, std::vector<void*>* group = nullptr)
{
	void* d = nullptr;
	size_t bytes = std::max<size_t>(count * sizeof(T), 16);
	PTB_CUDA(cudaMalloc(&d, bytes));
	(group ? *group : r->scene_allocs).push_back(d);
	// staged through pinned chunks from 48 MB on: below that the driver's own pageable copy takes a few milliseconds, while pinning this
	// thread's 64 MB of staging memory for the first time took 0.1-0.8 s on some boxes (c2's 14 MB of triangles: load 0.09 -> 0.9 s there)
	if (count * sizeof(T) >= (48u << 20)) { if (staged_copy(d, host, count * sizeof(T), true, r->stream)) return 1; }
	else if (count) PTB_CUDA(cudaMemcpyAsync(d, host, count * sizeof(T), cudaMemcpyHostToDevice, r->stream));
	*out = (const T*)d;
	return 0;
}

void release_scene_device(ptb_renderer* r)
{
	for (void* p : r->scene_allocs) cudaFree(p);
	for (void* p : r->geometry_allocs) cudaFree(p);
	for (void* p : r->material_allocs) cudaFree(p);
	for (void* p : r->light_allocs) cudaFree(p);
	for (cudaTextureObject_t t : r->texture_objects) cudaDestroyTextureObject(t);
	for (cudaArray_t a : r->texture_arrays) cudaFreeArray(a);
	r->texture_objects.clear(); r->texture_arrays.clear();
	r->scene_allocs.clear(); r->geometry_allocs.clear(); r->material_allocs.clear(); r->light_allocs.clear();
	memset(&r->dscene, 0, sizeof(r->dscene));
	r->bvh_nodes = r->bvh_bytes = 0;
	r->entry_valid = false;
}

DeviceMaterial pack_material(const ptb_material& m)
{
	DeviceMaterial d;
	int transparent = m.is_transparent ? 1 : 0;
	float tbits, dbits, sbits;
	memcpy(&tbits, &transparent, 4); memcpy(&dbits, &m.diffuse_texture_id, 4); memcpy(&sbits, &m.specular_texture_id, 4);
	d.a = make_float4(m.diffuse_color[0], m.diffuse_color[1], m.diffuse_color[2], m.roughness);
	d.b = make_float4(m.emission_color[0], m.emission_color[1], m.emission_color[2], m.refraction_index);
	d.c = make_float4(m.specular_color[0], m.specular_color[1], m.specular_color[2], m.extinction_coefficient);
	d.d = make_float4(m.absorption_coefficient[0], m.absorption_coefficient[1], m.absorption_coefficient[2], m.reduced_scattering_coefficient[0]);
	d.e = make_float4(m.reduced_scattering_coefficient[1], m.reduced_scattering_coefficient[2], tbits, dbits);
	d.f = make_float4(sbits, 0.0f, 0.0f, 0.0f);
	return d;
}

// Host copy of a device-built binary tree in the Bvh2 form build_bvh8 consumes (the wide layout's
// collapse + quantisation still runs on the host).  Node boxes are the padded child boxes stored in
// the parents; the root's box is the union of its children.
int download_bvh2(const GpuBuildOutput& gb, Bvh2& out)
{
	std::vector<float> nodes((size_t)gb.n_nodes * 16);
	out.prim_order.resize(gb.n_prims);
	PTB_CUDA(cudaMemcpy(nodes.data(), gb.nodes, nodes.size() * sizeof(float), cudaMemcpyDeviceToHost));
	PTB_CUDA(cudaMemcpy(out.prim_order.data(), gb.prim_order, (size_t)gb.n_prims * sizeof(int), cudaMemcpyDeviceToHost));
	out.nodes.clear();
	out.nodes.reserve((size_t)gb.n_prims);
	struct Item { int dev_ref; int host_index; };
	std::vector<Item> stack;
	out.nodes.emplace_back();
	for (int a = 0; a < 3; a++) { out.nodes[0].box.lo[a] = INFINITY; out.nodes[0].box.hi[a] = -INFINITY; }
	stack.push_back({ 0, 0 });
	bool root = true;
	while (!stack.empty())
	{
		Item it = stack.back();
		stack.pop_back();
		if (it.dev_ref < 0)
		{
			int ref = ~it.dev_ref;
			out.nodes[it.host_index].first = ref >> 3;
			out.nodes[it.host_index].count = (ref & 7) + 1;
			continue;
		}
		const float* d = &nodes[(size_t)it.dev_ref * 16];
		Aabb cb[2];
		cb[0].lo[0] = d[0]; cb[0].hi[0] = d[1]; cb[0].lo[1] = d[2]; cb[0].hi[1] = d[3]; cb[0].lo[2] = d[8]; cb[0].hi[2] = d[9];
		cb[1].lo[0] = d[4]; cb[1].hi[0] = d[5]; cb[1].lo[1] = d[6]; cb[1].hi[1] = d[7]; cb[1].lo[2] = d[10]; cb[1].hi[2] = d[11];
		int refs[2];
		memcpy(refs, &d[12], 8);
		const bool second_empty = cb[1].lo[0] > cb[1].hi[0];
		if (root && second_empty)
		{
			// single-leaf tree (wrapper node): the Bvh2 root is that leaf
			out.nodes[0].box = cb[0];
			stack.push_back({ refs[0], 0 });
			root = false;
			continue;
		}
		const int l = (int)out.nodes.size();
		out.nodes.emplace_back();
		out.nodes.emplace_back();
		out.nodes[l].box = cb[0]; out.nodes[l + 1].box = cb[1];
		out.nodes[it.host_index].left = l; out.nodes[it.host_index].right = l + 1; out.nodes[it.host_index].count = 0;
		if (root)
			for (int a = 0; a < 3; a++) { out.nodes[0].box.lo[a] = std::min(cb[0].lo[a], cb[1].lo[a]); out.nodes[0].box.hi[a] = std::max(cb[0].hi[a], cb[1].hi[a]); }
		root = false;
		stack.push_back({ refs[1], l + 1 });
		stack.push_back({ refs[0], l });
	}
	return 0;
}

// Option l2_persist: the binary tree's nodes and its leaf-order triangles are moved into ONE allocation and every render stream gets a
// persisting access-policy window over it (cudaStreamAttributeAccessPolicyWindow), everything else in those streams is "streaming":
// the tree the traversal kernels gather from is protected against the path-state records streaming through L2.
void apply_l2_window(ptb_renderer* r)
{
	if (r->host_only) return;
	cudaStreamAttrValue attr;
	memset(&attr, 0, sizeof(attr));
	if (r->l2_persist && r->l2_arena && r->l2_arena_bytes)
	{
		int max_window = 0, max_persist = 0;
		cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, r->device);
		cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, r->device);
		const size_t bytes = std::min<size_t>(r->l2_arena_bytes, (size_t)std::max(max_window, 0));
		cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>(bytes, (size_t)std::max(max_persist, 0)));
		attr.accessPolicyWindow.base_ptr = r->l2_arena;
		attr.accessPolicyWindow.num_bytes = bytes;
		attr.accessPolicyWindow.hitRatio = max_persist > 0 ? std::min(1.0f, (float)max_persist / (float)std::max<size_t>(bytes, 1)) : 0.0f;
		attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
		attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
	}
	else
	{
		attr.accessPolicyWindow.num_bytes = 0;
		attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
		attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
	}
	if (r->stream) cudaStreamSetAttribute(r->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
	for (size_t c = 1; c < r->contexts.size(); c++)
		if (r->contexts[c].stream) cudaStreamSetAttribute(r->contexts[c].stream, cudaStreamAttributeAccessPolicyWindow, &attr);
	if (!(r->l2_persist && r->l2_arena)) cudaCtxResetPersistingL2Cache();
	cudaGetLastError();
}

// triangles, acceleration structure and shading attributes (re-run by mesh edits: the BVH is REBUILT on the
// device — 5-65 ms for 0.15-5 M triangles — where the reference only re-transforms the old boxes,
// Bvh/bvh.cpp:332-356)
int upload_geometry(ptb_renderer* r)
{
	r->geometry_version++;        // entry cuts of the camera rays refer to the old tree
	r->entry_valid = false;
	for (void* p : r->geometry_allocs) cudaFree(p);
	r->geometry_allocs.clear();
	const HostScene& s = r->scene;
	DeviceScene& ds = r->dscene;
	ds.bvh_nodes = nullptr; ds.tri_isect = nullptr; ds.tri_shade = nullptr;
	std::vector<void*>* G = &r->geometry_allocs;

	// raw triangles + material indices on the device: the builder, the leaf-order emitter and the
	// shading-attribute packer all read them there
	const int n_tris = (int)s.triangles.size();
#ifdef PTB_EXPERIMENT_SORT
	for (int a = 0; a < 3; a++) { r->scene_lo[a] = INFINITY; r->scene_hi[a] = -INFINITY; }
	for (int i = 0; i < n_tris; i++)
	{
		const float* v = (const float*)&s.triangles[i];
		for (int k = 0; k < 3; k++) for (int a = 0; a < 3; a++) { r->scene_lo[a] = std::min(r->scene_lo[a], v[k * 3 + a]); r->scene_hi[a] = std::max(r->scene_hi[a], v[k * 3 + a]); }
	}
#endif
	const float* d_tris24 = nullptr;
	const int* d_material = nullptr;
	if (r->staged_tris24 && r->staged_tris_count == (size_t)n_tris && n_tris > 0)
	{
		// the triangles are already on this device (a scene received by ptb_dist_broadcast_scene): device -> device instead of host -> device
		float* d = nullptr;
		PTB_CUDA(cudaMalloc(&d, (size_t)n_tris * 96));
		G->push_back(d);
		PTB_CUDA(cudaMemcpyAsync(d, r->staged_tris24, (size_t)n_tris * 96, cudaMemcpyDeviceToDevice, r->stream));
		d_tris24 = d;
	}
	else if (upload(r, (const float*)s.triangles.data(), (size_t)n_tris * 24, &d_tris24, G)) return 1;
	if (upload(r, s.triangle_material.data(), (size_t)n_tris, &d_material, G)) return 1;
	ds.tris24 = d_tris24;
	r->bvh_built_on_gpu = 0; r->bvh_levels = 0; r->bvh_small_tasks = 0; r->bvh_max_depth = 0; r->bvh_build_ms = 0.0; r->bvh_note.clear();
	// the host builders / fallbacks read the HOST triangles: a scene that arrived on the device first (staged_tris24, csrc/multi.inc) is
	// drained to the host before the first of them runs
	bool host_triangles_ready = r->staged_tris24 == nullptr;
	auto host_triangles = [&]() -> int {
		if (host_triangles_ready || n_tris == 0) return 0;
		host_triangles_ready = true;
		return staged_copy(r->scene.triangles.data(), r->staged_tris24, (size_t)n_tris * 96, false, r->stream);
	};

	// acceleration structure over all meshes' world-space triangles
	const int max_leaf = r->bvh_layout == 8 ? 3 : r->bvh_max_leaf;   // <= 3 triangles per leaf slot of a wide node
	Bvh2 bvh;
	bool have_device_bvh2 = false, have_device_bvh8 = false;
	r->bvh_collapsed_on_gpu = 0;
	if (r->bvh_builder != "host_sah" && n_tris > 0)
	{
		GpuBuildOutput gb;
		std::string why;
		if (build_bvh2_gpu(d_tris24, n_tris, max_leaf, r->bvh_intersect_cost, r->stream, gb, why) == 0)
		{
			r->bvh_built_on_gpu = 1; r->bvh_levels = gb.levels; r->bvh_small_tasks = gb.small_tasks; r->bvh_max_depth = gb.max_depth; r->bvh_build_ms = gb.build_ms;
			if (gb.max_depth >= PTB_STACK_SIZE)
			{
				cudaFree(gb.nodes); cudaFree(gb.tri_isect); cudaFree(gb.prim_order);
				set_error("[Error]BVH too deep for the traversal stack");
				return 1;
			}
			if (r->bvh_layout == 8)
			{
				// wide layout: the binary tree and its 8-wide collapse + quantisation both come from the device
				GpuWideOutput gw;
				std::string why8;
				if (r->bvh_collapse != "host" && collapse_bvh8_gpu(gb, d_tris24, r->stream, gw, why8) == 0)
				{
					cudaFree(gb.nodes); cudaFree(gb.tri_isect); cudaFree(gb.prim_order);
					if (gw.max_depth > PTB_STACK_SIZE8) { cudaFree(gw.nodes); cudaFree(gw.tris); set_error("[Error]BVH8 too deep for the traversal stack"); return 1; }
					G->push_back(gw.nodes); G->push_back(gw.tris);
					ds.bvh_nodes = gw.nodes; ds.tri_isect = gw.tris;
					r->bvh_nodes = gw.n_nodes;
					r->bvh_bytes = (int64_t)gw.n_nodes * 80 + (int64_t)gw.n_tris * 48;
					r->bvh_build_ms += gw.collapse_ms;
					r->bvh_collapsed_on_gpu = 1;
					have_device_bvh8 = true;
				}
				else
				{
					cudaGetLastError();
					int rc = download_bvh2(gb, bvh);
					cudaFree(gb.nodes); cudaFree(gb.tri_isect); cudaFree(gb.prim_order);
					if (rc) return 1;
				}
			}
			else
			{
				G->push_back(gb.nodes); G->push_back(gb.tri_isect);
				cudaFree(gb.prim_order);
				ds.bvh_nodes = gb.nodes; ds.tri_isect = gb.tri_isect;
				r->bvh_nodes = gb.n_nodes;
				r->bvh_bytes = (int64_t)gb.n_nodes * 64 + (int64_t)n_tris * 48;
				have_device_bvh2 = true;
			}
		}
		else
		{
			// adversarial input (capacity) or out of memory: say so and use the host builder
			r->bvh_note = why;
			fprintf(stderr, "[Warn]GPU BVH build failed (%s); using the host SAH builder\n", why.c_str());
			cudaGetLastError();
		}
	}
	if (!have_device_bvh2 && !(r->bvh_built_on_gpu && r->bvh_layout == 8))
	{
		if (host_triangles()) return 1;
		const auto t0 = std::chrono::steady_clock::now();
		build_bvh2_sah(s.triangles, max_leaf, bvh, r->bvh_intersect_cost);
		r->bvh_build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		r->bvh_max_depth = bvh.max_depth;
		if (bvh.max_depth >= PTB_STACK_SIZE) { set_error("[Error]BVH too deep for the traversal stack"); return 1; }
	}
	if (r->bvh_layout == 8 && !have_device_bvh8)
	{
		GpuBvh8 wide;
		if (host_triangles()) return 1;
		build_bvh8(bvh, s.triangles, wide);
		if (wide.max_depth > PTB_STACK_SIZE8) { set_error("[Error]BVH8 too deep for the traversal stack"); return 1; }
		if (upload(r, (const float4*)wide.nodes.data(), wide.nodes.size() / 4, &ds.bvh_nodes, G)) return 1;
		if (upload(r, (const float4*)wide.tris.data(), wide.tris.size() / 4, &ds.tri_isect, G)) return 1;
		r->bvh_nodes = (int64_t)wide.nodes.size() / 20;
		r->bvh_bytes = (int64_t)(wide.nodes.size() + wide.tris.size()) * 4;
	}
	else if (r->bvh_layout != 8 && !have_device_bvh2)
	{
		GpuBvh2 flat;
		if (host_triangles()) return 1;
		flatten_bvh2(bvh, s.triangles, flat);
		if (upload(r, (const float4*)flat.nodes.data(), flat.nodes.size() / 4, &ds.bvh_nodes, G)) return 1;
		if (upload(r, (const float4*)flat.tris.data(), flat.tris.size() / 4, &ds.tri_isect, G)) return 1;
		r->bvh_nodes = (int64_t)flat.nodes.size() / 16;
		r->bvh_bytes = (int64_t)(flat.nodes.size() + flat.tris.size()) * 4;
	}
	ds.bvh_layout = r->bvh_layout;
	ds.n_triangles = n_tris;
	ds.root_ref = 0;
	ds.bvh8_nodes = nullptr; ds.tri_isect8 = nullptr;
	r->bvh8_nodes = 0;
	if (r->bvh_layout == 2 && r->bvh_hybrid && n_tris > 0)
	{
		// Hybrid: coherent camera rays are issue-bound and fastest on the binary tree; bounce rays are bound by L1
		// wavefronts and fastest on the compressed wide tree, which moves half the bytes per ray (measured on c2:
		// depth 0 8.2 vs 10.8 ms, depth 1 10.3 vs 8.7 ms per 32 passes; profiles/r01_experiments.md).  Both trees
		// return the same hits (conservative culling, the winner is decided by the same triangle arithmetic).
		Bvh2 tree3;
		GpuBuildOutput gb;
		std::string why;
		bool have_tree = false, done = false;
		if (r->bvh_builder != "host_sah" && build_bvh2_gpu(d_tris24, n_tris, 3, r->bvh_intersect_cost, r->stream, gb, why) == 0)
		{
			r->bvh_build_ms += gb.build_ms;
			// collapse + quantisation on the device; the host collapse (bvh_host.cpp) stays as fallback and cross-check
			GpuWideOutput gw;
			if (r->bvh_collapse != "host" && collapse_bvh8_gpu(gb, d_tris24, r->stream, gw, why) == 0)
			{
				if (gw.max_depth <= PTB_STACK_SIZE8)
				{
					G->push_back(gw.nodes); G->push_back(gw.tris);
					ds.bvh8_nodes = gw.nodes; ds.tri_isect8 = gw.tris;
					r->bvh8_nodes = gw.n_nodes;
					r->bvh_bytes += (int64_t)gw.n_nodes * 80 + (int64_t)gw.n_tris * 48;
					r->bvh_build_ms += gw.collapse_ms;
					r->bvh_collapsed_on_gpu = 1;
				}
				else { cudaFree(gw.nodes); cudaFree(gw.tris); }
				done = true;
			}
			else
			{
				cudaGetLastError();
				const int rc = download_bvh2(gb, tree3);
				if (rc) { cudaFree(gb.nodes); cudaFree(gb.tri_isect); cudaFree(gb.prim_order); return 1; }
				have_tree = true;
			}
			cudaFree(gb.nodes); cudaFree(gb.tri_isect); cudaFree(gb.prim_order);
		}
		if (!done)
		{
			if (!have_tree)
			{
				cudaGetLastError();
				if (host_triangles()) return 1;
				build_bvh2_sah(s.triangles, 3, tree3, r->bvh_intersect_cost);
			}
			GpuBvh8 wide;
			if (host_triangles()) return 1;
			build_bvh8(tree3, s.triangles, wide);
			if (wide.max_depth <= PTB_STACK_SIZE8)
			{
				if (upload(r, (const float4*)wide.nodes.data(), wide.nodes.size() / 4, &ds.bvh8_nodes, G)) return 1;
				if (upload(r, (const float4*)wide.tris.data(), wide.tris.size() / 4, &ds.tri_isect8, G)) return 1;
				r->bvh8_nodes = (int64_t)wide.nodes.size() / 20;
				r->bvh_bytes += (int64_t)(wide.nodes.size() + wide.tris.size()) * 4;
			}
		}
	}

	// shading attributes by global triangle id, packed on the device
	{
		float4* d_shade = nullptr;
		PTB_CUDA(cudaMalloc(&d_shade, std::max<size_t>((size_t)n_tris * 64, 16)));
		G->push_back(d_shade);
		pack_tri_shade_gpu(d_tris24, d_material, n_tris, d_shade, r->stream);
		ds.tri_shade = d_shade;
	}
	r->l2_arena = nullptr; r->l2_arena_bytes = 0;
	if (r->l2_persist && ds.bvh_layout == 2 && ds.bvh_nodes && ds.tri_isect && n_tris > 0)
	{
		const size_t node_bytes = ((size_t)r->bvh_nodes * 64 + 255) & ~(size_t)255, tri_bytes = (size_t)n_tris * 48;
		char* arena = nullptr;
		PTB_CUDA(cudaMalloc(&arena, node_bytes + tri_bytes));
		PTB_CUDA(cudaMemcpyAsync(arena, ds.bvh_nodes, (size_t)r->bvh_nodes * 64, cudaMemcpyDeviceToDevice, r->stream));
		PTB_CUDA(cudaMemcpyAsync(arena + node_bytes, ds.tri_isect, tri_bytes, cudaMemcpyDeviceToDevice, r->stream));
		PTB_CUDA(cudaStreamSynchronize(r->stream));
		for (const void* old_ptr : { (const void*)ds.bvh_nodes, (const void*)ds.tri_isect })
		{
			auto it = std::find(G->begin(), G->end(), (void*)old_ptr);
			if (it != G->end()) { cudaFree(*it); G->erase(it); }
		}
		G->push_back(arena);
		ds.bvh_nodes = (const float4*)arena; ds.tri_isect = (const float4*)(arena + node_bytes);
		r->l2_arena = arena; r->l2_arena_bytes = node_bytes + tri_bytes;
	}
	apply_l2_window(r);
	ds.up_records = nullptr; ds.tri_slot = nullptr;
	if (r->upwalk && ds.bvh_layout == 2 && ds.bvh_nodes && ds.tri_isect && n_tris > 0 && r->bvh_nodes >= r->upwalk_min_nodes)   // a tree of a few nodes is as quick from the root (c1: -1 %)
	{
		// leaf starts of the bounce rays: sibling records per child slot + the slot of every triangle's leaf, built top-down from the root
		const int n_nodes = (int)r->bvh_nodes;
		float4* up = nullptr; int* tri_slot = nullptr; int2* frontier[2] = { nullptr, nullptr }; int* level_counts = nullptr;
		const int levels = PTB_STACK_SIZE + 2;
		// (an accelerator, not a requirement: without memory for the records the bounce rays start at the root)
		float4* up1 = nullptr;      // one level per record (k_up_level), paired into `up` below
		bool have = cudaMalloc(&up1, (size_t)n_nodes * 64) == cudaSuccess && cudaMalloc(&up, (size_t)n_nodes * 128) == cudaSuccess &&
			cudaMalloc(&tri_slot, (size_t)n_tris * sizeof(int)) == cudaSuccess && cudaMalloc(&frontier[0], (size_t)n_nodes * sizeof(int2)) == cudaSuccess &&
			cudaMalloc(&frontier[1], (size_t)n_nodes * sizeof(int2)) == cudaSuccess && cudaMalloc(&level_counts, (levels + 1) * sizeof(int)) == cudaSuccess;
		if (have)
		{
			PTB_CUDA(cudaMemsetAsync(up1, 0xff, (size_t)n_nodes * 64, r->stream));      // unused pool slots of the node array read "no parent"
			PTB_CUDA(cudaMemsetAsync(tri_slot, 0xff, (size_t)n_tris * sizeof(int), r->stream));
			PTB_CUDA(cudaMemsetAsync(level_counts, 0, (levels + 1) * sizeof(int), r->stream));
			const int2 root_entry = make_int2(ds.root_ref, -1);
			const int one = 1;
			PTB_CUDA(cudaMemcpyAsync(frontier[0], &root_entry, sizeof(int2), cudaMemcpyHostToDevice, r->stream));
			PTB_CUDA(cudaMemcpyAsync(level_counts, &one, sizeof(int), cudaMemcpyHostToDevice, r->stream));
			PTB_CUDA(cudaStreamSynchronize(r->stream));   // the two host words above are stack variables
			const int grid = std::max(1, std::min((n_nodes + 127) / 128, r->sm_count * 8));
			for (int level = 0; level < levels; level++)
				k_up_level<<<grid, 128, 0, r->stream>>>(ds.bvh_nodes, ds.tri_isect, frontier[level & 1], frontier[(level + 1) & 1], level_counts, level, n_nodes, up1, tri_slot, n_tris);
			k_up_pair<<<(n_nodes * 2 + 127) / 128, 128, 0, r->stream>>>(up1, up, n_nodes * 2);
			PTB_CUDA(cudaStreamSynchronize(r->stream));
			PTB_CUDA(cudaGetLastError());
		}
		else cudaGetLastError();
		cudaFree(frontier[0]); cudaFree(frontier[1]); cudaFree(level_counts); cudaFree(up1);
		if (!have) { cudaFree(up); cudaFree(tri_slot); up = nullptr; tri_slot = nullptr; }
		else { G->push_back(up); G->push_back(tri_slot); }
		ds.up_records = up; ds.tri_slot = tri_slot;
		if (have) r->bvh_bytes += (int64_t)n_nodes * 128 + (int64_t)n_tris * 4;
	}
	PTB_CUDA(cudaStreamSynchronize(r->stream));
	PTB_CUDA(cudaGetLastError());
	return 0;
}

// emissive, non-transparent triangles with their area CDF (estimator "nee"); depends on geometry AND materials
int upload_lights(ptb_renderer* r)
{
	for (void* p : r->light_allocs) cudaFree(p);
	r->light_allocs.clear();
	const HostScene& s = r->scene;
	DeviceScene& ds = r->dscene;
	std::vector<int> ids;
	std::vector<double> area;
	double total = 0.0;
	for (size_t i = 0; i < s.triangles.size(); i++)
	{
		const ptb_material& m = s.materials[s.triangle_material[i]];
		// a transparent emitter never reaches the emission branch in the reference (path_tracer_kernel.cu:573-616)
		if (m.is_transparent || !(m.emission_color[0] > 0.0f || m.emission_color[1] > 0.0f || m.emission_color[2] > 0.0f)) continue;
		const Triangle& t = s.triangles[i];
		const double e1[3] = { (double)t.v1.x - t.v0.x, (double)t.v1.y - t.v0.y, (double)t.v1.z - t.v0.z };
		const double e2[3] = { (double)t.v2.x - t.v0.x, (double)t.v2.y - t.v0.y, (double)t.v2.z - t.v0.z };
		const double cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2], cz = e1[0] * e2[1] - e1[1] * e2[0];
		const double a = 0.5 * std::sqrt(cx * cx + cy * cy + cz * cz);
		if (!(a > 0.0)) continue;
		ids.push_back((int)i); area.push_back(a); total += a;
	}
	std::vector<float> cdf(ids.size());
	double run = 0.0;
	for (size_t k = 0; k < ids.size(); k++) { run += area[k]; cdf[k] = (float)(run / total); }
	if (!cdf.empty()) cdf.back() = 1.0f;
	if (upload(r, ids.data(), ids.size(), &ds.light_tri, &r->light_allocs)) return 1;
	if (upload(r, cdf.data(), cdf.size(), &ds.light_cdf, &r->light_allocs)) return 1;
	ds.n_lights = (int)ids.size();
	ds.light_area = (float)total;
	PTB_CUDA(cudaStreamSynchronize(r->stream));
	return 0;
}

// mesh materials followed by one material per sphere, and the sphere array (re-run by material / sphere edits)
int upload_materials(ptb_renderer* r)
{
	for (void* p : r->material_allocs) cudaFree(p);
	r->material_allocs.clear();
	const HostScene& s = r->scene;
	DeviceScene& ds = r->dscene;
	std::vector<DeviceMaterial> mats;
	for (auto& m : s.materials) mats.push_back(pack_material(m));
	r->scene_has_medium = false;
	for (auto& m : s.materials) if (m.reduced_scattering_coefficient[0] > 0.0f) r->scene_has_medium = true;
	for (auto& sp : s.spheres) if (sp.mat.reduced_scattering_coefficient[0] > 0.0f) r->scene_has_medium = true;
	ds.sphere_material_base = (int)mats.size();
	std::vector<float4> spheres;
	for (auto& sp : s.spheres)
	{
		mats.push_back(pack_material(sp.mat));
		spheres.push_back(make_float4(sp.center.x, sp.center.y, sp.center.z, sp.radius));
	}
	if (upload(r, mats.data(), mats.size(), &ds.materials, &r->material_allocs)) return 1;
	if (upload(r, spheres.data(), spheres.size(), &ds.spheres, &r->material_allocs)) return 1;
	ds.n_spheres = (int)spheres.size();
	PTB_CUDA(cudaStreamSynchronize(r->stream));
	return 0;
}

// texture_filter=hardware: an RGBA8 image as a cudaArray + texture object (unnormalised coordinates, clamp, linear filter, bytes read as
// normalised floats) — the reference filters in software (Core/texture.h:15-79, Core/cube_map.h:20-119); this is the optional fast path
int make_texture_object(ptb_renderer* r, const Texture& t, cudaTextureObject_t* out)
{
	*out = 0;
	if (t.width <= 0 || t.height <= 0 || t.rgba.size() < (size_t)t.width * t.height * 4) return 0;
	cudaChannelFormatDesc desc = cudaCreateChannelDesc<uchar4>();
	cudaArray_t arr = nullptr;
	PTB_CUDA(cudaMallocArray(&arr, &desc, (size_t)t.width, (size_t)t.height));
	r->texture_arrays.push_back(arr);
	PTB_CUDA(cudaMemcpy2DToArray(arr, 0, 0, t.rgba.data(), (size_t)t.width * 4, (size_t)t.width * 4, (size_t)t.height, cudaMemcpyHostToDevice));
	cudaResourceDesc res;
	memset(&res, 0, sizeof(res));
	res.resType = cudaResourceTypeArray;
	res.res.array.array = arr;
	cudaTextureDesc td;
	memset(&td, 0, sizeof(td));
	td.addressMode[0] = cudaAddressModeClamp; td.addressMode[1] = cudaAddressModeClamp;
	td.filterMode = cudaFilterModeLinear;
	td.readMode = cudaReadModeNormalizedFloat;
	td.normalizedCoords = 0;
	cudaTextureObject_t obj = 0;
	PTB_CUDA(cudaCreateTextureObject(&obj, &res, &td, nullptr));
	r->texture_objects.push_back(obj);
	*out = obj;
	return 0;
}

int upload_scene(ptb_renderer* r)
{
	release_scene_device(r);
	const HostScene& s = r->scene;
	DeviceScene& ds = r->dscene;
	memset(&ds, 0, sizeof(ds));
	const auto t_upload0 = std::chrono::steady_clock::now();
	if (upload_geometry(r)) return 1;
	if (upload_materials(r)) return 1;
	if (upload_lights(r)) return 1;

	std::vector<DeviceTexture> textures;
	for (auto& t : s.textures)
	{
		DeviceTexture dt;
		if (upload(r, t.rgba.data(), t.rgba.size(), &dt.pixels)) return 1;
		dt.width = t.width; dt.height = t.height;
		dt.tex = 0;
		if (r->hw_textures && make_texture_object(r, t, &dt.tex)) return 1;
		textures.push_back(dt);
	}
	if (upload(r, textures.data(), textures.size(), &ds.textures)) return 1;
	ds.n_textures = (int)textures.size();

	for (int f = 0; f < 6; f++)
	{
		if (upload(r, s.cube_faces[f].rgba.data(), s.cube_faces[f].rgba.size(), &ds.sky.faces[f])) return 1;
		ds.sky.face_tex[f] = 0;
		if (r->hw_textures && make_texture_object(r, s.cube_faces[f], &ds.sky.face_tex[f])) return 1;
	}
	ds.sky.length = s.cube_length;
	ds.sky.use_sky_box = r->cfg.use_sky_box ? 1 : 0;
	ds.sky.use_sky = r->cfg.use_sky ? 1 : 0;
	ds.sky.use_bilinear = r->cfg.use_bilinear ? 1 : 0;
	PTB_CUDA(cudaStreamSynchronize(r->stream));
	PTB_CUDA(cudaGetLastError());
	r->scene_upload_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_upload0).count();
	return 0;
}

DeviceConfig device_config(const ptb_renderer* r)
{
	DeviceConfig c;
	c.max_depth = r->cfg.max_tracer_depth;
	c.bias_length = r->cfg.vector_bias_length;
	c.energy_threshold = r->cfg.energy_exist_threshold;
	c.sss_threshold = r->cfg.sss_threshold;
	c.use_bilinear = r->cfg.use_bilinear ? 1 : 0;
	c.gamma_correction = r->cfg.gamma_correction ? 1 : 0;
	c.use_anti_alias = r->cfg.use_anti_alias ? 1 : 0;
	c.air_n = r->cfg.air_refraction_index;
	c.air_sigma_a = make_float3(r->cfg.air_absorption_coef.x, r->cfg.air_absorption_coef.y, r->cfg.air_absorption_coef.z);
	c.air_sigma_s = make_float3(r->cfg.air_reduced_scattering_coef.x, r->cfg.air_reduced_scattering_coef.y, r->cfg.air_reduced_scattering_coef.z);
	c.sampler = r->sampler;
	c.sss_mode = r->sss_mode;
	return c;
}

CameraParams camera_params(const ptb_camera& c)
{
	CameraParams p;
	p.eye = make_float3(c.eye[0], c.eye[1], c.eye[2]);
	p.view = make_float3(c.view[0], c.view[1], c.view[2]);
	p.up = make_float3(c.up[0], c.up[1], c.up[2]);
	p.resolution = make_float2(c.resolution[0], c.resolution[1]);
	p.fov = make_float2(c.fov[0], c.fov[1]);
	p.aperture_radius = c.aperture_radius;
	p.focal_distance = c.focal_distance;
	return p;
}

int grid_for(const ptb_renderer* r, size_t items, int block, int blocks_per_sm)
{
	size_t need = (items + block - 1) / block;
	size_t cap = (size_t)r->sm_count * blocks_per_sm;
	return (int)std::max<size_t>(1, std::min(need, cap));
}

// Entry cuts of the camera rays (kernels_entry.cuh): true when this render can use them (binary tree traced by the persistent kernel,
// a camera whose rays the shaft construction covers).
bool entry_cuts_usable(const ptb_renderer* r)
{
	if (!r->entry_cuts || r->dscene.bvh_layout != 2 || !r->extend_persistent || r->extend_variant != 0 || r->dscene.n_triangles <= 0) return false;
	const ptb_camera& c = r->cam;
	const float vl = std::sqrt(c.view[0] * c.view[0] + c.view[1] * c.view[1] + c.view[2] * c.view[2]);
	if (!(vl > 1e-20f) || !std::isfinite(vl)) return false;
	if (!(c.focal_distance > 1e-6f) || !std::isfinite(c.focal_distance)) return false;     // a non-positive focal distance flips the rays
	if (!(c.fov[0] > 0.01f && c.fov[0] < 175.0f && c.fov[1] > 0.01f && c.fov[1] < 175.0f)) return false;
	if (c.resolution[0] != (float)r->cfg.width || c.resolution[1] != (float)r->cfg.height || r->cfg.width < 2 || r->cfg.height < 2) return false;
	if (!(c.aperture_radius < 0.25f * c.focal_distance)) return false;                      // (also rejects NaN)
	for (int k = 0; k < 3; k++) if (!std::isfinite(c.eye[k]) || !std::isfinite(c.up[k])) return false;
	// The generator forms eye + view + offsets and eye + direction * focal_distance in binary32 and subtracts the eye again
	// (path_tracer_kernel.cu:299-379): a camera far from the origin relative to |view| / the focal distance loses direction bits there.  The
	// shafts carry 1/16 pixel of slack for that; beyond 1/20 pixel of worst-case rounding the rays are searched from the root.
	const float eye_max = std::max(std::fabs(c.eye[0]), std::max(std::fabs(c.eye[1]), std::fabs(c.eye[2])));
	const float tan_x = std::tan(c.fov[0] * 0.5f * 3.14159265f / 180.0f), tan_y = std::tan(c.fov[1] * 0.5f * 3.14159265f / 180.0f);
	const float pixel = std::min(2.0f * tan_x / (float)(r->cfg.width - 1), 2.0f * tan_y / (float)(r->cfg.height - 1));      // on a canvas at distance 1
	const float ulp2 = 2.4e-7f;      // 2^-22
	if (!(ulp2 * (eye_max + vl * (1.0f + tan_x + tan_y)) <= pixel * vl / 20.0f)) return false;
	if (!(ulp2 * (eye_max + c.focal_distance) <= pixel * c.focal_distance / 20.0f)) return false;
	return true;
}

// (Re)builds the lists when camera, resolution, k or geometry changed since the last build.  Rare (a camera move, a scene edit): waits
// for every stream that may still read the old lists, builds on `stream`, and makes the other streams wait for the build.
// `n_slots`: passes of the batch that asks.  Building the lists costs ~0.4 ms at 1080p (k_entry_cut + k_tile_rank + two waits) and saves
// ~0.15 ms per pass, so a camera that changes with EVERY pass — a host dragging the view, one pass per call — is served from the root: the
// lists are built when a batch brings at least entry_min_passes passes or when the camera of the previous batch is seen again.
// Returns 0 and leaves r->entry_valid false in that case.
int ensure_entry_cuts(ptb_renderer* r, cudaStream_t stream, int n_slots)
{
	unsigned char key[sizeof(r->entry_key)];
	memset(key, 0, sizeof(key));
	static_assert(sizeof(ptb_camera) + 4 * sizeof(int) + sizeof(uint64_t) <= sizeof(key), "entry key too small");
	memcpy(key, &r->cam, sizeof(ptb_camera));
	const int ints[4] = { r->cfg.width, r->cfg.height, r->entry_k, r->entry_tile_w * 256 + r->entry_tile_h };
	memcpy(key + sizeof(ptb_camera), ints, sizeof(ints));
	memcpy(key + sizeof(ptb_camera) + sizeof(ints), &r->geometry_version, sizeof(uint64_t));
	if (r->entry_valid && memcmp(key, r->entry_key, sizeof(key)) == 0) return 0;
	const bool seen_last_batch = memcmp(key, r->entry_last_key, sizeof(key)) == 0;
	memcpy(r->entry_last_key, key, sizeof(key));
	if (n_slots < r->entry_min_passes && !seen_last_batch)
	{
		// no lists for this camera yet; the stale ones must not be used (enqueue_batch checks entry_valid).  Batches still in flight keep
		// reading them: nothing writes the buffer before the next build, which waits for the device first.
		r->entry_valid = false;
		return 0;
	}
	const int tiles_x = (r->cfg.width + r->entry_tile_w - 1) / r->entry_tile_w, tiles_y = (r->cfg.height + r->entry_tile_h - 1) / r->entry_tile_h;
	const size_t n_tiles = (size_t)tiles_x * tiles_y;
	PTB_CUDA(cudaDeviceSynchronize());
	r->entry_stride = 2;
	while (r->entry_stride < r->entry_k + 1) r->entry_stride *= 2;
	// the lists are an accelerator, not a requirement: without memory for them the searches start at the root
	if (n_tiles * r->entry_stride > r->entry_buf_slots)
	{
		cudaFree(r->entry_buf); r->entry_buf = nullptr; r->entry_buf_slots = 0;
		if (cudaMalloc(&r->entry_buf, n_tiles * r->entry_stride * sizeof(int2)) != cudaSuccess) { cudaGetLastError(); r->entry_buf = nullptr; r->entry_valid = false; return 0; }
		r->entry_buf_slots = n_tiles * r->entry_stride;
	}
	if (n_tiles > r->entry_rank_tiles)
	{
		cudaFree(r->entry_rank); r->entry_rank = nullptr; r->entry_rank_tiles = 0;
		if (cudaMalloc(&r->entry_rank, (n_tiles + 1) * sizeof(int)) != cudaSuccess) { cudaGetLastError(); r->entry_rank = nullptr; r->entry_valid = false; return 0; }
		r->entry_rank_tiles = n_tiles;
	}
	k_entry_cut<<<(int)((n_tiles + 127) / 128), 128, 0, stream>>>(r->dscene, camera_params(r->cam), r->cfg.width, r->cfg.height, tiles_x, (int)n_tiles, r->entry_k, r->entry_tile_w, r->entry_tile_h, r->entry_stride, r->entry_buf);
	k_tile_rank<<<1, 1024, 0, stream>>>(r->entry_buf, r->entry_stride, (int)n_tiles, r->entry_rank, r->entry_rank + n_tiles);
	r->entry_n_tiles = n_tiles;
	r->stats.kernel_launches += 2;
	PTB_CUDA(cudaGetLastError());
	PTB_CUDA(cudaStreamSynchronize(stream));
	memcpy(r->entry_key, key, sizeof(key));
	r->entry_valid = true;
	return 0;
}

void launch_extend(ptb_renderer* r, cudaStream_t stream, size_t items, const PathState& st, const int* queue, const int* count_ptr, int* work_counter, int depth = 0, const FusedArgs* fused = nullptr, bool entry = false, bool upwalk = false, int hybrid_from = -1)
{
	bool wide = r->dscene.bvh_layout == 8;
	if (hybrid_from < 0) hybrid_from = r->hybrid_from_depth;
	if (entry && !fused && r->entry_valid)
	{
		// camera rays: every ray starts at its tile's entry cut
		EntryArgs ea;
		ea.cuts = r->entry_buf; ea.pixel_count = r->pixel_count; ea.width = r->cfg.width; ea.tiles_x = (r->cfg.width + r->entry_tile_w - 1) / r->entry_tile_w;
		ea.tile_w_shift = 0; ea.tile_h_shift = 0; ea.stride_shift = 0;
		while ((1 << ea.stride_shift) < r->entry_stride) ea.stride_shift++;
		while ((1 << ea.tile_w_shift) < r->entry_tile_w) ea.tile_w_shift++;
		while ((1 << ea.tile_h_shift) < r->entry_tile_h) ea.tile_h_shift++;
		int grid = std::max(1, std::min(r->persistent_grid, (int)((items + 127) / 128)));
		if (r->count_traversal) k_extend_entry<true, 0><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill_e, r->tune_leaf_e, r->tune_reps_e, ea);
		else if (r->tune_reps_e == PTB_NODE_REPS && r->unroll_reps) k_extend_entry<false, PTB_NODE_REPS><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill_e, r->tune_leaf_e, r->tune_reps_e, ea);
		else k_extend_entry<false, 0><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill_e, r->tune_leaf_e, r->tune_reps_e, ea);
		return;
	}
	if (r->dscene.bvh_layout == 2 && r->dscene.bvh8_nodes && r->extend_persistent && depth >= hybrid_from)
	{
		// hybrid: bounce rays over the compressed wide tree
		DeviceScene sc8 = r->dscene;
		sc8.bvh_nodes = r->dscene.bvh8_nodes; sc8.tri_isect = r->dscene.tri_isect8; sc8.bvh_layout = 8;
		int grid = std::max(1, std::min(r->persistent_grid8, (int)((items + 127) / 128)));
		if (fused)
		{
			if (r->count_traversal) k_extend_persistent8<true, true><<<grid, 128, 0, stream>>>(sc8, st, queue, count_ptr, work_counter, r->counters, r->tune_refill8, r->tune_leaf8, *fused);
			else k_extend_persistent8<false, true><<<grid, 128, 0, stream>>>(sc8, st, queue, count_ptr, work_counter, r->counters, r->tune_refill8, r->tune_leaf8, *fused);
		}
		else if (r->count_traversal) k_extend_persistent8<true><<<grid, 128, 0, stream>>>(sc8, st, queue, count_ptr, work_counter, r->counters, r->tune_refill8, r->tune_leaf8);
		else k_extend_persistent8<false><<<grid, 128, 0, stream>>>(sc8, st, queue, count_ptr, work_counter, r->counters, r->tune_refill8, r->tune_leaf8);
		return;
	}
	if (wide && r->extend_persistent)
	{
		int grid = std::max(1, std::min(r->persistent_grid8, (int)((items + 127) / 128)));
		if (fused)
		{
			if (r->count_traversal) k_extend_persistent8<true, true><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill8, r->tune_leaf8, *fused);
			else k_extend_persistent8<false, true><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill8, r->tune_leaf8, *fused);
		}
		else if (r->count_traversal) k_extend_persistent8<true><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill8, r->tune_leaf8);
		else k_extend_persistent8<false><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill8, r->tune_leaf8);
		return;
	}
	if (!wide && r->extend_persistent)
	{
		// persistent warps: one resident wave, sized from the occupancy the kernel actually gets
		int grid = std::max(1, std::min(r->persistent_grid, (int)((items + 127) / 128)));
		if (fused)
		{
			// inline scatter events over the binary tree (fused_from_depth <= depth < hybrid_from_depth)
			if (upwalk && r->dscene.up_records)
			{
				// ... every search of the walk started at the leaf of the triangle the path entered the medium through
				if (r->count_traversal) k_extend_persistent_fused<true, true><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill_f, r->tune_leaf_f, *fused);
				else k_extend_persistent_fused<false, true><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill_f, r->tune_leaf_f, *fused);
			}
			else if (r->count_traversal) k_extend_persistent_fused<true><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill, r->tune_leaf, *fused);
			else k_extend_persistent_fused<false><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill, r->tune_leaf, *fused);
			return;
		}
		if (r->extend_variant == 4)
		{
			// staged refill (kernels_extend.cuh: k_extend_staged): rays set up 32 at a time into shared memory, idle lanes pop ready rays
			if (r->count_traversal) k_extend_staged<true><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill4, r->tune_leaf);
			else k_extend_staged<false><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill4, r->tune_leaf);
			return;
		}
		if (r->extend_variant == 3)
		{
			// top of the tree in shared memory, blocks as large as the SM allows (kernels_extend.cuh: k_extend_treelet)
			const int n_top = (int)std::max<int64_t>(0, std::min<int64_t>(r->treelet_nodes, r->bvh_nodes));
			const int block = std::max(128, std::min(1024, r->treelet_block)) & ~31;
			const size_t smem = (size_t)n_top * 64;
			static bool attr_set = false;
			if (!attr_set)
			{
				cudaFuncSetAttribute(k_extend_treelet<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
				cudaFuncSetAttribute(k_extend_treelet<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
				attr_set = true;
			}
			int per_sm = 1;
			cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_extend_treelet<false>, block, smem);
			const int tgrid = std::max(1, std::min(r->sm_count * std::max(per_sm, 1), (int)((items + block - 1) / block)));
			if (r->count_traversal) k_extend_treelet<true><<<tgrid, block, smem, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill, r->tune_leaf, n_top);
			else k_extend_treelet<false><<<tgrid, block, smem, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill, r->tune_leaf, n_top);
			return;
		}
		if (r->extend_variant > 0)
		{
			if (r->count_traversal) k_extend_speculative<true, false><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill, r->tune_leaf);
			else if (r->extend_variant == 2) k_extend_speculative<false, true><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill, r->tune_leaf);
			else k_extend_speculative<false, false><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill, r->tune_leaf);
			return;
		}
		if (upwalk && r->dscene.up_records)
		{
			// bounce rays: start at the leaf of the triangle the ray leaves
			if (r->count_traversal) k_extend_upwalk<true, 0><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill_u, r->tune_leaf_u, r->tune_reps_u);
			else if (r->tune_reps_u == PTB_NODE_REPS && r->unroll_reps) k_extend_upwalk<false, PTB_NODE_REPS><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill_u, r->tune_leaf_u, r->tune_reps_u);
			else k_extend_upwalk<false, 0><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill_u, r->tune_leaf_u, r->tune_reps_u);
			return;
		}
		if (r->count_traversal) k_extend_persistent<true, 0><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill, r->tune_leaf, r->tune_reps);
		else if (r->tune_reps == PTB_NODE_REPS && r->unroll_reps) k_extend_persistent<false, PTB_NODE_REPS><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill, r->tune_leaf, r->tune_reps);
		else k_extend_persistent<false, 0><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, work_counter, r->counters, r->tune_refill, r->tune_leaf, r->tune_reps);
		return;
	}
	int grid = grid_for(r, items, 128, 16);
	if (r->count_traversal)
	{
		if (wide) k_extend<true, true><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, r->counters);
		else k_extend<true, false><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, r->counters);
	}
	else
	{
		if (wide) k_extend<false, true><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, r->counters);
		else k_extend<false, false><<<grid, 128, 0, stream>>>(r->dscene, st, queue, count_ptr, r->counters);
	}
}

// enqueue one batch of n_slots passes (first, first+stride, ...) on a context's stream.
// `prev_accumulated`: event of the previous batch's k_accumulate — the running sum is updated in
// pass order whatever the overlap, so the image is bit-identical to a serial run.
int enqueue_batch(ptb_renderer* r, ptb_renderer::BatchContext& ctx, cudaEvent_t prev_accumulated, int first_pass, int stride, int n_slots)
{
	const int px = r->pixel_count;
	const size_t total = (size_t)px * n_slots;
	DeviceConfig dc = device_config(r);
	CameraParams cp = camera_params(r->cam);
	const int n_counts = r->cfg.max_tracer_depth + 2;
	const bool prof = r->profile_stages != 0;
	cudaStream_t stream = ctx.stream;
	const int tiles_x = (r->tile_order && r->cfg.width % 8 == 0 && r->cfg.height % 4 == 0) ? r->cfg.width / 8 : 0;
	const bool alt = r->sampler != 0 || r->sss_mode != 0;   // estimator options: the ALT instantiations of k_generate / k_shade
	// inline scatter events (kernels_extend.cuh: k_extend_persistent8<.., FUSED>): the reference's estimator only, and only where a medium exists
	const bool fused = r->inline_scatter && r->cfg.max_tracer_depth <= 255 && (r->scene_has_medium || r->cfg.air_reduced_scattering_coef.x > 0.0f) && !alt && !r->nee && !r->russian_roulette && !r->sort_by_material && r->extend_persistent && r->extend_variant == 0;
	FusedArgs fa;
	fa.cfg = dc; fa.pixel_count = px; fa.first_pass = first_pass; fa.pass_stride = stride; fa.scatter_min = r->tune_scatter;
	fa.depth_segments = r->segment_totals; fa.n_depth_slots = n_counts;
	// loop depths traced by the fused kernel tally their own searches per actual depth; the queue sizes only count below that
	// first loop depth whose closest-hit launch scatters inline: fused_from_depth (default: where the wide tree takes over; lower values run
	// the walk on the binary-tree kernel until then)
	// scattering media with leaf starts (fused_upwalk): every bounce stays on the binary tree — a search that starts at the leaf of the
	// triangle the walk entered through costs one 32-byte record per tree level and few node visits, less than descending the wide tree
	// from the root for a segment one mean free path long
	const bool have_up = r->upwalk && !r->nee && r->dscene.bvh_layout == 2 && r->extend_persistent && r->extend_variant == 0 && r->dscene.up_records != nullptr;
	const bool fused_up = fused && have_up && r->fused_upwalk;
	const bool small_tree = have_up && !fused && !r->hybrid_from_user && r->small_tree_bytes > 0 &&
		r->bvh_nodes * (int64_t)(64 + 128) + (int64_t)r->dscene.n_triangles * 48 <= r->small_tree_bytes;
	const int hybrid_from = (fused_up || small_tree) ? 0x7fffffff : r->hybrid_from_depth;
	const int fused_from = r->fused_from_depth >= 0 ? r->fused_from_depth : (fused_up ? (r->cfg.air_reduced_scattering_coef.x > 0.0f ? 0 : 1) : (r->dscene.bvh_layout == 8 ? 0 : std::max(r->hybrid_from_depth, 0)));
	const int tally_counts = fused ? std::min(fused_from, n_counts) : n_counts;
	const bool entry = entry_cuts_usable(r) && hybrid_from > 0 && !(fused && fused_from <= 0);
	if (entry && ensure_entry_cuts(r, stream, n_slots)) return 1;
	// bounce rays start at the leaf they leave: k_shade leaves the triangle in ray_o.w (kernels.cuh: PTB_FROM_BITS)
	const bool upwalk = have_up;
	// camera rays of tiles with an EMPTY entry cut end at the background: k_generate finishes them (kernels_generate.cuh, SKY) — valid when
	// nothing but triangles can be hit and the air does not take part
	const bool sky_fast = entry && r->entry_valid && r->sky_fast && tiles_x > 0 && r->entry_tile_w == 8 && r->entry_tile_h == 4 && r->dscene.n_spheres == 0 &&
		r->cfg.max_tracer_depth >= 1 && !(fused && fused_from <= 0) &&
		!(dc.air_sigma_s.x > 0.0f || dc.air_sigma_s.y > 0.0f || dc.air_sigma_s.z > 0.0f ||
		  std::sqrt(dc.air_sigma_a.x * dc.air_sigma_a.x + dc.air_sigma_a.y * dc.air_sigma_a.y + dc.air_sigma_a.z * dc.air_sigma_a.z) > dc.sss_threshold);
	if (sky_fast)
	{
		SkyArgs sa;
		sa.sky = r->dscene.sky; sa.tile_rank = r->entry_rank; sa.n_nonempty = r->entry_rank + r->entry_n_tiles;
		if (alt) k_generate<true, true><<<grid_for(r, total, 256, 8), 256, 0, stream>>>(ctx.st, ctx.queue[0], ctx.counts, n_counts, cp, dc, px, n_slots, first_pass, stride, tiles_x, sa);
		else k_generate<false, true><<<grid_for(r, total, 256, 8), 256, 0, stream>>>(ctx.st, ctx.queue[0], ctx.counts, n_counts, cp, dc, px, n_slots, first_pass, stride, tiles_x, sa);
	}
	else if (alt) k_generate<true><<<grid_for(r, total, 256, 8), 256, 0, stream>>>(ctx.st, ctx.queue[0], ctx.counts, n_counts, cp, dc, px, n_slots, first_pass, stride, tiles_x);
	else k_generate<false><<<grid_for(r, total, 256, 8), 256, 0, stream>>>(ctx.st, ctx.queue[0], ctx.counts, n_counts, cp, dc, px, n_slots, first_pass, stride, tiles_x);
	r->stats.kernel_launches++;
	for (int depth = 0; depth < r->cfg.max_tracer_depth; depth++)
	{
		int* qin = ctx.queue[depth & 1];
		int* qout = ctx.queue[(depth + 1) & 1];
		cudaEvent_t e0 = nullptr, e1 = nullptr;
		if (prof)
		{
			cudaEventCreate(&e0); cudaEventCreate(&e1);
			r->stage_events.push_back(e0); r->stage_events.push_back(e1);
			cudaEventRecord(e0, stream);
		}
		fa.loop_depth = depth;
		launch_extend(r, stream, total, ctx.st, qin, ctx.counts + depth, ctx.counts + n_counts + depth, depth, (fused && depth >= fused_from) ? &fa : nullptr, entry && depth == 0, upwalk && depth >= 1 && (!(fused && depth >= fused_from) || fused_up), hybrid_from);
		if (prof) cudaEventRecord(e1, stream);
		int* shadow_count = ctx.counts + 2 * n_counts + depth;
		const int sgrid = grid_for(r, total, 128, 16);
#define PTB_SHADE_ARGS r->dscene, ctx.st, dc, depth, px, first_pass, stride, qin, ctx.counts + depth, qout, ctx.counts + depth + 1, shadow_count, r->octant_order
		if (r->nee)
		{
			if (alt && r->russian_roulette) k_shade<false, true, true, true><<<sgrid, 128, 0, stream>>>(PTB_SHADE_ARGS);
			else if (alt) k_shade<false, true, false, true><<<sgrid, 128, 0, stream>>>(PTB_SHADE_ARGS);
			else if (r->russian_roulette) k_shade<false, true, true><<<sgrid, 128, 0, stream>>>(PTB_SHADE_ARGS);
			else if (r->sort_by_material) k_shade<true, true><<<sgrid, 128, 0, stream>>>(PTB_SHADE_ARGS);
			else k_shade<false, true><<<sgrid, 128, 0, stream>>>(PTB_SHADE_ARGS);
			if (depth + 1 < r->cfg.max_tracer_depth)
			{
				k_shadow<<<sgrid, 128, 0, stream>>>(r->dscene, ctx.st, shadow_count);
				r->stats.kernel_launches++;
			}
		}
		else if (fused) k_shade<false, false, false, false, true><<<sgrid, 128, 0, stream>>>(PTB_SHADE_ARGS);
		else if (alt && r->russian_roulette) k_shade<false, false, true, true><<<sgrid, 128, 0, stream>>>(PTB_SHADE_ARGS);
		else if (alt) k_shade<false, false, false, true><<<sgrid, 128, 0, stream>>>(PTB_SHADE_ARGS);
		else if (r->russian_roulette) k_shade<false, false, true><<<sgrid, 128, 0, stream>>>(PTB_SHADE_ARGS);
		else if (r->sort_by_material) k_shade<true, false><<<sgrid, 128, 0, stream>>>(PTB_SHADE_ARGS);
		else k_shade<false, false><<<sgrid, 128, 0, stream>>>(PTB_SHADE_ARGS);
#undef PTB_SHADE_ARGS
		r->stats.kernel_launches += 2;
#ifdef PTB_EXPERIMENT_SORT
		if ((r->sort_depth_mask >> (depth + 1)) & 1)
		{
			const int tot = (int)total;
			if (!r->sort_keys[0])
			{
				cudaMalloc(&r->sort_keys[0], total * 4); cudaMalloc(&r->sort_keys[1], total * 4); cudaMalloc(&r->sort_vals, total * 4);
				cub::DeviceRadixSort::SortPairs(nullptr, r->sort_tmp_bytes, r->sort_keys[0], r->sort_keys[1], qout, r->sort_vals, tot);
				cudaMalloc(&r->sort_tmp, r->sort_tmp_bytes);
			}
			const float3 lo = make_float3(r->scene_lo[0], r->scene_lo[1], r->scene_lo[2]);
			const float3 inv = make_float3(1.0f / (r->scene_hi[0] - r->scene_lo[0]), 1.0f / (r->scene_hi[1] - r->scene_lo[1]), 1.0f / (r->scene_hi[2] - r->scene_lo[2]));
			k_sort_keys<<<(tot + 255) / 256, 256, 0, stream>>>(ctx.st, qout, ctx.counts + depth + 1, r->sort_keys[0], tot, lo, inv, r->sort_cell_bits, r->sort_octant);
			cub::DeviceRadixSort::SortPairs(r->sort_tmp, r->sort_tmp_bytes, r->sort_keys[0], r->sort_keys[1], qout, r->sort_vals, tot, 0, 3 * r->sort_cell_bits + 3, stream);
			cudaMemcpyAsync(qout, r->sort_vals, (size_t)tot * 4, cudaMemcpyDeviceToDevice, stream);
		}
#endif
	}
	if (prev_accumulated) PTB_CUDA(cudaStreamWaitEvent(stream, prev_accumulated, 0));
	k_accumulate<<<(px + 255) / 256, 256, 0, stream>>>(ctx.st.radiance, r->image_sum, r->last_pass, ctx.counts, r->segment_totals, tally_counts, px, n_slots, r->pass_clamp >= 0.0f ? r->pass_clamp : (float)r->cfg.max_tracer_depth * 2.0f, sky_fast ? (int)total : 0);
	r->stats.kernel_launches++;
	PTB_CUDA(cudaEventRecord(ctx.accumulated, stream));
	PTB_CUDA(cudaGetLastError());
	return 0;
}

int render_impl(ptb_renderer* r, int first_pass, int stride, int n_passes, bool advance_counter, bool synchronous)
{
	if (!r || r->host_only) { set_error("[Error]renderer has no CUDA device (host-only handle): rendering is unavailable, there is no CPU fallback"); return 1; }
	if (!r->scene_loaded) { set_error("[Error]no scene loaded"); return 1; }
	if (n_passes <= 0) return 0;
	PTB_CUDA(cudaSetDevice(r->device));   // the caller's thread may drive several devices (csrc/multi.inc)
	if (r->nee && r->dscene.bvh_layout != 2) { set_error("[Error]estimator nee needs the binary tree (bvh_layout 2): its shadow stage traverses it"); return 1; }
	memset(&r->stats, 0, sizeof(r->stats));
	r->stats.bvh_nodes = r->bvh_nodes; r->stats.bvh_bytes = r->bvh_bytes;
	const int n_counts = r->cfg.max_tracer_depth + 2;
	PTB_CUDA(cudaMemsetAsync(r->counters, 0, 32 * sizeof(unsigned long long), r->stream));
	PTB_CUDA(cudaMemsetAsync(r->segment_totals, 0, n_counts * sizeof(unsigned long long), r->stream));
	PTB_CUDA(cudaEventRecord(r->ev0, r->stream));
	const int n_ctx = r->active_streams > 0 ? std::min(r->active_streams, (int)r->contexts.size()) : (int)r->contexts.size();
	for (int c = 1; c < n_ctx; c++) PTB_CUDA(cudaStreamWaitEvent(r->contexts[c].stream, r->ev0, 0));
	int done = 0, batch = 0;
	cudaEvent_t prev_acc = nullptr;
	while (done < n_passes)
	{
		int nb = std::min(r->passes_in_flight, n_passes - done);
		ptb_renderer::BatchContext& ctx = r->contexts[batch % n_ctx];
		if (enqueue_batch(r, ctx, prev_acc, first_pass + done * stride, stride, nb)) return 1;
		prev_acc = ctx.accumulated;
		done += nb;
		batch++;
	}
	// the last accumulate depends (transitively) on every earlier batch: wait for it on the main stream
	if (prev_acc && batch > 0 && r->contexts[(batch - 1) % n_ctx].stream != r->stream) PTB_CUDA(cudaStreamWaitEvent(r->stream, prev_acc, 0));
	if (advance_counter) r->pass_counter += n_passes;
	// strided calls (spp sharding; ptb_render_strided advances the counter after the call) accumulate across calls: the displayed
	// image divides by every pass summed since the last clear, not by this call's count
	int total_passes = advance_counter ? r->pass_counter : r->pass_counter + n_passes;
	k_tonemap<<<(r->pixel_count + 255) / 256, 256, 0, r->stream>>>(r->image_sum, r->image_u8, r->pixel_count, std::max(total_passes, 1), r->cfg.gamma_correction ? 1 : 0);
	r->stats.kernel_launches++;
	PTB_CUDA(cudaMemcpyAsync(r->counts_host, r->segment_totals, n_counts * sizeof(unsigned long long), cudaMemcpyDeviceToHost, r->stream));
	PTB_CUDA(cudaEventRecord(r->ev1, r->stream));
	r->stats.passes = n_passes;
	if (synchronous)
	{
		PTB_CUDA(cudaStreamSynchronize(r->stream));
		for (int c = 1; c < n_ctx; c++) PTB_CUDA(cudaStreamSynchronize(r->contexts[c].stream));
		r->depth_segments.assign(r->cfg.max_tracer_depth, 0);
		for (int d = 0; d < r->cfg.max_tracer_depth; d++)
		{
			r->stats.ray_segments += (int64_t)r->counts_host[d];
			r->depth_segments[d] = (int64_t)r->counts_host[d];
		}

		float ms = 0.0f;
		cudaEventElapsedTime(&ms, r->ev0, r->ev1);
		r->stats.gpu_ms_total = ms;
		double ext = 0.0;
		r->depth_extend_ms.assign(r->cfg.max_tracer_depth, 0.0);
		for (size_t i = 0; i + 1 < r->stage_events.size(); i += 2)
		{
			float m = 0.0f;
			cudaEventElapsedTime(&m, r->stage_events[i], r->stage_events[i + 1]);
			ext += m;
			r->depth_extend_ms[(i / 2) % r->cfg.max_tracer_depth] += m;
		}
		for (auto e : r->stage_events) cudaEventDestroy(e);
		r->stage_events.clear();
		r->stats.gpu_ms_extend = ext;
		if (r->count_traversal)
		{
			unsigned long long c[32];
			cudaMemcpy(c, r->counters, sizeof(c), cudaMemcpyDeviceToHost);
			r->stats.nodes_visited = (int64_t)c[0]; r->stats.tris_tested = (int64_t)c[1]; r->stats.wide_nodes_visited = (int64_t)c[3];
			for (int k = 0; k < 32; k++) r->traversal_histogram[k] = (int64_t)c[k];
		}
	}
	return 0;
}

// shared by ptb_trace_batch / ptb_trace_batch_bruteforce
int trace_impl(ptb_renderer* r, const float* rays6, int n, int32_t* out_prim, float* out_t, float* out_bary, bool brute)
{
	if (!r || r->host_only) { set_error("[Error]renderer has no CUDA device (host-only handle): tracing is unavailable, there is no CPU fallback"); return 1; }
	if (!r->scene_loaded) { set_error("[Error]no scene loaded"); return 1; }
	if (n <= 0) return 0;
	float4 *d_o = nullptr, *d_d = nullptr, *d_hit = nullptr;
	int* d_q = nullptr; int* d_count = nullptr;
	std::vector<float4> ho(n), hd(n);
	for (int i = 0; i < n; i++)
	{
		ho[i] = make_float4(rays6[i * 6 + 0], rays6[i * 6 + 1], rays6[i * 6 + 2], 0.0f);
		hd[i] = make_float4(rays6[i * 6 + 3], rays6[i * 6 + 4], rays6[i * 6 + 5], INFINITY);
	}
	PTB_CUDA(cudaMalloc(&d_o, (size_t)n * sizeof(float4)));
	PTB_CUDA(cudaMalloc(&d_d, (size_t)n * sizeof(float4)));
	PTB_CUDA(cudaMalloc(&d_hit, (size_t)n * sizeof(float4)));
	PTB_CUDA(cudaMalloc(&d_q, (size_t)n * sizeof(int)));
	PTB_CUDA(cudaMalloc(&d_count, 2 * sizeof(int)));
	PTB_CUDA(cudaMemsetAsync(d_count, 0, 2 * sizeof(int), r->stream));
	PTB_CUDA(cudaMemcpyAsync(d_o, ho.data(), (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, r->stream));
	PTB_CUDA(cudaMemcpyAsync(d_d, hd.data(), (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, r->stream));
	if (brute)
	{
		k_bruteforce<<<(n + 127) / 128, 128, 0, r->stream>>>(r->dscene, d_o, d_d, d_hit, n);
	}
	else
	{
		k_iota<<<(n + 255) / 256, 256, 0, r->stream>>>(d_q, d_count, n);
		PathState st;
		memset(&st, 0, sizeof(st));
		st.ray_o = d_o; st.ray_d = d_d; st.hit = d_hit;
		launch_extend(r, r->stream, n, st, d_q, d_count, d_count + 1);
	}
	std::vector<float4> hh(n);
	PTB_CUDA(cudaMemcpyAsync(hh.data(), d_hit, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, r->stream));
	PTB_CUDA(cudaStreamSynchronize(r->stream));
	PTB_CUDA(cudaGetLastError());
	for (int i = 0; i < n; i++)
	{
		int prim;
		memcpy(&prim, &hh[i].w, 4);
		out_prim[i] = prim;
		if (out_t) out_t[i] = hh[i].x;
		if (out_bary) { out_bary[i * 2] = hh[i].y; out_bary[i * 2 + 1] = hh[i].z; }
	}
	cudaFree(d_o); cudaFree(d_d); cudaFree(d_hit); cudaFree(d_q); cudaFree(d_count);
	return 0;
}

} // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================

extern "C"
{

const char* ptb_last_error(void) { return last_error().c_str(); }
int ptb_version(void) { return 100; }

int ptb_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

ptb_renderer* ptb_create(const char* config_json_path, int cuda_device)
{
	ptb_renderer* r = new ptb_renderer();
	memset(&r->stats, 0, sizeof(r->stats));
	memset(&r->dscene, 0, sizeof(r->dscene));
	memset(&r->st, 0, sizeof(r->st));
	bool config_ok = false;
	// no exception may cross the C boundary: a parser running out of memory is a load error like any other
	try { config_ok = config_json_path && load_config(config_json_path, r->cfg); }
	catch (const std::exception& e) { set_error(std::string("[Error]") + e.what()); }
	catch (...) { set_error("[Error]unknown failure while reading the configuration"); }
	if (!config_ok) { delete r; return nullptr; }
	default_camera((float)r->cfg.width, (float)r->cfg.height, -1.0f, -1.0f, r->cam);
	r->device = cuda_device;
	if (cuda_device < 0)
	{
		// host-only handle: scene loading and introspection work, every compute entry point fails
		r->host_only = true;
		return r;
	}
	cudaError_t e = cudaSetDevice(cuda_device);
	if (e != cudaSuccess)
	{
		set_error(std::string("[Cuda]cannot select device ") + std::to_string(cuda_device) + ": " + cudaGetErrorString(e) + " (there is no CPU fallback)");
		cudaGetLastError();
		delete r;
		return nullptr;
	}
	cudaDeviceProp prop = {};
	if (cudaGetDeviceProperties(&prop, cuda_device) == cudaSuccess) r->sm_count = prop.multiProcessorCount;
	if (cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&r->ev0) != cudaSuccess || cudaEventCreate(&r->ev1) != cudaSuccess)
	{
		set_error("[Cuda]cannot create stream/events");
		delete r;
		return nullptr;
	}
	{
		// persistent grids = one resident wave; the shared-memory part of the traversal stack asks for an explicit L1 / shared split
		// (the rest of the 256 KB stays L1 for the node and triangle gathers)
		auto setup = [&](auto kernel, int* grid)
		{
			int per_sm = 0;
			if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 128, 0) != cudaSuccess || per_sm <= 0) { cudaGetLastError(); return; }
			if (grid) *grid = r->sm_count * per_sm;
			cudaFuncAttributes fa;
			if (cudaFuncGetAttributes(&fa, kernel) == cudaSuccess && fa.sharedSizeBytes > 0)
			{
				const size_t need = (size_t)per_sm * (fa.sharedSizeBytes + 1024);
				const int pct = (int)std::min<size_t>(100, (need * 100 + prop.sharedMemPerMultiprocessor - 1) / std::max<size_t>(prop.sharedMemPerMultiprocessor, 1));
				cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
			}
			cudaGetLastError();
		};
		setup(k_extend_persistent<false, 0>, &r->persistent_grid);
		setup(k_extend_persistent<false, PTB_NODE_REPS>, nullptr);
		setup(k_extend_persistent<true, 0>, nullptr);
		setup(k_extend_persistent_fused<false, true>, nullptr);
		setup(k_extend_persistent_fused<true, true>, nullptr);
		setup(k_extend_upwalk<false, PTB_NODE_REPS>, nullptr);
		setup(k_extend_upwalk<false, 0>, nullptr);
		setup(k_extend_upwalk<true, 0>, nullptr);
		setup(k_extend_entry<false, PTB_NODE_REPS>, nullptr);
		setup(k_extend_entry<false, 0>, nullptr);
		setup(k_extend_entry<true, 0>, nullptr);
		setup(k_extend_persistent8<false>, &r->persistent_grid8);
		setup(k_extend_persistent8<true>, nullptr);
	}
	if (alloc_work_buffers(r)) { free_work_buffers(r); delete r; return nullptr; }
	return r;
}

void ptb_destroy(ptb_renderer* r)
{
	if (!r) return;
	if (!r->host_only)
	{
		cudaSetDevice(r->device);
		cudaStreamSynchronize(r->stream);
		release_scene_device(r);
		free_work_buffers(r);
		if (r->ev0) cudaEventDestroy(r->ev0);
		if (r->ev1) cudaEventDestroy(r->ev1);
		if (r->stream) cudaStreamDestroy(r->stream);
	}
	delete r;
}

int ptb_list_scenes(const char* scene_dir, char* out, int cap)
{
	std::vector<std::string> files;
	int n = list_scenes(scene_dir ? scene_dir : "", files);
	if (n < 0) { set_error("[Warn]There exists no scene file!"); return -1; }
	std::string all;
	for (auto& f : files) { all += f; all += "\n"; }
	if ((int)all.size() + 1 > cap) { set_error("[Error]ptb_list_scenes: buffer too small"); return -1; }
	memcpy(out, all.c_str(), all.size() + 1);
	return n;
}

int ptb_load_scene(ptb_renderer* r, const char* scene_json_path, const char* asset_root)
{
	if (!r) { set_error("[Error]null renderer"); return 1; }
	ptb_release_scene(r);
	bool loaded = false;
	// no exception may cross the C boundary (see ptb_create)
	try { loaded = load_scene(scene_json_path ? scene_json_path : "", asset_root ? asset_root : "", r->scene); }
	catch (const std::exception& e) { set_error(std::string("[Error]") + e.what()); }
	catch (...) { set_error("[Error]unknown failure while reading the scene"); }
	if (!loaded)
	{
		r->scene = HostScene();
		return 1;
	}
	if (!r->host_only)
	{
		cudaSetDevice(r->device);
		if (upload_scene(r)) { release_scene_device(r); r->scene = HostScene(); return 1; }
	}
	r->scene_loaded = true;
	return ptb_clear(r);
}

int ptb_release_scene(ptb_renderer* r)
{
	if (!r) return 1;
	if (!r->host_only)
	{
		cudaSetDevice(r->device);
		for (size_t c = 1; c < r->contexts.size(); c++) cudaStreamSynchronize(r->contexts[c].stream);
		cudaStreamSynchronize(r->stream);
		release_scene_device(r);
	}
	r->scene = HostScene();
	r->scene_loaded = false;
	return 0;
}

int ptb_default_camera(float width, float height, float aperture_radius, float focal_distance, ptb_camera* out)
{
	if (!out) return 1;
	default_camera(width, height, aperture_radius, focal_distance, *out);
	return 0;
}

int ptb_set_camera(ptb_renderer* r, const ptb_camera* cam) { if (!r || !cam) return 1; r->cam = *cam; return 0; }
int ptb_get_camera(ptb_renderer* r, ptb_camera* out) { if (!r || !out) return 1; *out = r->cam; return 0; }

int ptb_render(ptb_renderer* r, int n_passes)
{
	if (!r) { set_error("[Error]null renderer"); return 1; }
	return render_impl(r, r->pass_counter + 1, 1, n_passes, true, true);
}

int ptb_render_async(ptb_renderer* r, int n_passes)
{
	if (!r) { set_error("[Error]null renderer"); return 1; }
	return render_impl(r, r->pass_counter + 1, 1, n_passes, true, false);
}

int ptb_render_strided(ptb_renderer* r, int first_pass, int stride, int n_passes)
{
	if (!r) { set_error("[Error]null renderer"); return 1; }
	int rc = render_impl(r, first_pass, stride, n_passes, false, true);
	if (rc == 0) r->pass_counter += n_passes;
	return rc;
}

static int sync_all_streams(ptb_renderer* r)
{
	for (size_t c = 1; c < r->contexts.size(); c++) PTB_CUDA(cudaStreamSynchronize(r->contexts[c].stream));
	PTB_CUDA(cudaStreamSynchronize(r->stream));
	return 0;
}

int ptb_synchronize(ptb_renderer* r)
{
	if (!r || r->host_only) return 1;
	return sync_all_streams(r);
}

void* ptb_stream(ptb_renderer* r) { return r ? (void*)r->stream : nullptr; }

int ptb_clear(ptb_renderer* r)
{
	if (!r) return 1;
	r->pass_counter = 0;
	if (r->host_only) return 0;
	cudaSetDevice(r->device);
	for (size_t c = 1; c < r->contexts.size(); c++) PTB_CUDA(cudaStreamSynchronize(r->contexts[c].stream));
	PTB_CUDA(cudaMemsetAsync(r->image_sum, 0, (size_t)r->pixel_count * 3 * sizeof(float), r->stream));
	PTB_CUDA(cudaStreamSynchronize(r->stream));
	return 0;
}

// ---- live scene edits (the reference mutates managed scene memory between passes,
// Core/path_tracer.cpp:109-369).  Each edit resets the accumulation like path_tracer::render_ui -> clear().
static int edit_prologue(ptb_renderer* r)
{
	if (!r) { set_error("[Error]null renderer"); return 1; }
	if (!r->scene_loaded) { set_error("[Error]no scene loaded"); return 1; }
	if (!r->host_only)
	{
		cudaSetDevice(r->device);
		if (sync_all_streams(r)) return 1;
	}
	return 0;
}

// A geometry edit whose device rebuild fails (out of memory, "BVH too deep") has already changed the host triangles and freed the old
// tree: the scene is unloaded, so the next ptb_render answers "[Error]no scene loaded" instead of traversing freed memory.
static int geometry_edit_failed(ptb_renderer* r)
{
	const std::string why = last_error();
	cudaGetLastError();
	release_scene_device(r);
	r->scene_loaded = false;
	r->scene = HostScene();
	set_error(why + " (the scene was unloaded: reload it)");
	return 1;
}

int ptb_set_sphere(ptb_renderer* r, int index, const void* sphere100)
{
	if (edit_prologue(r)) return 1;
	if (index < 0 || index >= (int)r->scene.spheres.size() || !sphere100) { set_error("[Error]sphere index out of range"); return 1; }
	memcpy(&r->scene.spheres[index], sphere100, sizeof(Sphere));
	if (!r->host_only && (upload_materials(r) || upload_lights(r))) return 1;
	return ptb_clear(r);
}

int ptb_set_mesh_material(ptb_renderer* r, int mesh, const ptb_material* mats, int n)
{
	if (edit_prologue(r)) return 1;
	if (mesh < 0 || mesh >= (int)r->scene.meshes.size() || !mats) { set_error("[Error]mesh index out of range"); return 1; }
	const MeshInfo& m = r->scene.meshes[mesh];
	if (n != m.material_count) return 0;   // triangle_mesh::set_material_device ignores a list of the wrong length (triangle_mesh.cpp:254-257)
	for (int i = 0; i < n; i++) r->scene.materials[m.first_material + i] = mats[i];
	if (!r->host_only && (upload_materials(r) || upload_lights(r))) return 1;
	return ptb_clear(r);
}

int ptb_set_mesh_transform(ptb_renderer* r, int mesh, const float* position3, const float* scale3)
{
	if (edit_prologue(r)) return 1;
	if (!position3 || !scale3) { set_error("[Error]null argument"); return 1; }
	// the UI clamps the scale to >= 1e-6 before the call (Core/path_tracer.cpp:346-351)
	Vec3 sc{ std::max(scale3[0], 0.000001f), std::max(scale3[1], 0.000001f), std::max(scale3[2], 0.000001f) };
	if (!set_mesh_transform(r->scene, mesh, Vec3{ position3[0], position3[1], position3[2] }, sc)) return 1;
	if (!r->host_only && (upload_geometry(r) || upload_lights(r))) return geometry_edit_failed(r);
	return ptb_clear(r);
}

int ptb_apply_mesh_rotate(ptb_renderer* r, int mesh, const float* rotate3)
{
	if (edit_prologue(r)) return 1;
	if (!rotate3) { set_error("[Error]null argument"); return 1; }
	if (!apply_mesh_rotate(r->scene, mesh, Vec3{ rotate3[0], rotate3[1], rotate3[2] })) return 1;
	if (!r->host_only && (upload_geometry(r) || upload_lights(r))) return geometry_edit_failed(r);
	return ptb_clear(r);
}

int ptb_get_mesh_placement(ptb_renderer* r, int mesh, float* out_position3, float* out_scale3, float* out_rotate3, int* out_first_triangle, int* out_triangle_count, int* out_first_material, int* out_material_count)
{
	if (!r || !r->scene_loaded || mesh < 0 || mesh >= (int)r->scene.meshes.size()) { set_error("[Error]mesh index out of range"); return 1; }
	const MeshInfo& m = r->scene.meshes[mesh];
	if (out_position3) { out_position3[0] = m.position.x; out_position3[1] = m.position.y; out_position3[2] = m.position.z; }
	if (out_scale3) { out_scale3[0] = m.scale.x; out_scale3[1] = m.scale.y; out_scale3[2] = m.scale.z; }
	if (out_rotate3) { out_rotate3[0] = m.rotate_applied.x; out_rotate3[1] = m.rotate_applied.y; out_rotate3[2] = m.rotate_applied.z; }
	if (out_first_triangle) *out_first_triangle = m.first_triangle;
	if (out_triangle_count) *out_triangle_count = m.triangle_count;
	if (out_first_material) *out_first_material = m.first_material;
	if (out_material_count) *out_material_count = m.material_count;
	return 0;
}

// ---- output side ----------------------------------------------------------------
int ptb_write_png_rgb8(const char* path, const uint8_t* rgb, int width, int height)
{
	std::string err;
	if (!path || !write_png_rgb8(path, rgb, width, height, err)) { set_error(err.empty() ? "[Error]null path" : err); return 1; }
	return 0;
}

int ptb_write_pfm(const char* path, const float* rgb, int width, int height, float scale)
{
	std::string err;
	if (!path || !write_pfm_rgb(path, rgb, width, height, scale, err)) { set_error(err.empty() ? "[Error]null path" : err); return 1; }
	return 0;
}

int ptb_save_png(ptb_renderer* r, const char* path)
{
	if (!r || r->host_only) { set_error("[Error]no CUDA device"); return 1; }
	std::vector<uint8_t> u8((size_t)r->pixel_count * 3);
	if (ptb_image_u8(r, u8.data())) return 1;
	return ptb_write_png_rgb8(path, u8.data(), r->cfg.width, r->cfg.height);
}

int ptb_save_pfm(ptb_renderer* r, const char* path)
{
	if (!r || r->host_only) { set_error("[Error]no CUDA device"); return 1; }
	std::vector<float> sum((size_t)r->pixel_count * 3);
	int passes = 0;
	if (ptb_image_f32(r, sum.data(), &passes)) return 1;
	return ptb_write_pfm(path, sum.data(), r->cfg.width, r->cfg.height, passes > 0 ? 1.0f / (float)passes : 1.0f);
}

int ptb_save_checkpoint(ptb_renderer* r, const char* path)
{
	if (!r || r->host_only) { set_error("[Error]no CUDA device"); return 1; }
	if (!path) { set_error("[Error]null path"); return 1; }
	std::vector<float> sum((size_t)r->pixel_count * 3);
	int passes = 0;
	if (sync_all_streams(r) || ptb_image_f32(r, sum.data(), &passes)) return 1;
	CheckpointHeader h;
	h.width = r->cfg.width; h.height = r->cfg.height; h.pass_counter = r->pass_counter; h.max_depth = r->cfg.max_tracer_depth;
	memcpy(h.camera, &r->cam, sizeof(h.camera));
	std::string err;
	if (!write_checkpoint(path, h, sum.data(), err)) { set_error(err); return 1; }
	return 0;
}

int ptb_load_checkpoint(ptb_renderer* r, const char* path, int restore_camera)
{
	if (!r || r->host_only) { set_error("[Error]no CUDA device"); return 1; }
	if (!path) { set_error("[Error]null path"); return 1; }
	CheckpointHeader h;
	std::vector<float> sum;
	std::string err;
	if (!read_checkpoint(path, h, sum, err)) { set_error(err); return 1; }
	if (h.width != r->cfg.width || h.height != r->cfg.height) { set_error("[Error]checkpoint: resolution differs from the renderer's configuration"); return 1; }
	if (h.max_depth != r->cfg.max_tracer_depth) { set_error("[Error]checkpoint: MaxDepth differs from the renderer's configuration"); return 1; }
	cudaSetDevice(r->device);
	if (sync_all_streams(r)) return 1;
	PTB_CUDA(cudaMemcpyAsync(r->image_sum, sum.data(), sum.size() * sizeof(float), cudaMemcpyHostToDevice, r->stream));
	PTB_CUDA(cudaStreamSynchronize(r->stream));
	r->pass_counter = h.pass_counter;
	if (restore_camera) memcpy(&r->cam, h.camera, sizeof(h.camera));
	const int rc = ptb_finalize(r, std::max(h.pass_counter, 1));   // refresh the 8-bit image from the restored sum
	r->pass_counter = h.pass_counter;
	return rc;
}

// Live config toggles: the reference's UI writes into the managed `configuration` between passes
// (Main/window.cpp; fields of Core/configuration.h:9-34 read by the kernels every pass) and restarts the accumulation.
// Everything the path reads per pass can change; the image size and MaxDepth size the work buffers and need a new handle.
int ptb_set_config(ptb_renderer* r, const void* config96)
{
	if (!r || !config96) { set_error("[Error]null argument"); return 1; }
	Config c;
	memcpy(&c, config96, sizeof(Config));
	if (c.width != r->cfg.width || c.height != r->cfg.height || c.max_tracer_depth != r->cfg.max_tracer_depth)
	{
		set_error("[Error]ptb_set_config: Width / Height / MaxDepth size the work buffers; create a new renderer for them");
		return 1;
	}
	if (!r->host_only)
	{
		cudaSetDevice(r->device);
		if (sync_all_streams(r)) return 1;
	}
	r->cfg = c;
	r->dscene.sky.use_sky_box = c.use_sky_box ? 1 : 0;
	r->dscene.sky.use_sky = c.use_sky ? 1 : 0;
	r->dscene.sky.use_bilinear = c.use_bilinear ? 1 : 0;
	return ptb_clear(r);
}

int ptb_pass_counter(ptb_renderer* r) { return r ? r->pass_counter : 0; }
int ptb_width(ptb_renderer* r) { return r ? r->cfg.width : 0; }
int ptb_height(ptb_renderer* r) { return r ? r->cfg.height : 0; }

int ptb_image_f32(ptb_renderer* r, float* out_rgb_sum, int* out_passes)
{
	if (!r || r->host_only) { set_error("[Error]no CUDA device"); return 1; }
	PTB_CUDA(cudaMemcpyAsync(out_rgb_sum, r->image_sum, (size_t)r->pixel_count * 3 * sizeof(float), cudaMemcpyDeviceToHost, r->stream));
	PTB_CUDA(cudaStreamSynchronize(r->stream));
	if (out_passes) *out_passes = r->pass_counter;
	return 0;
}

int ptb_image_u8(ptb_renderer* r, uint8_t* out_rgb)
{
	if (!r || r->host_only) { set_error("[Error]no CUDA device"); return 1; }
	PTB_CUDA(cudaMemcpyAsync(out_rgb, r->image_u8, (size_t)r->pixel_count * 3, cudaMemcpyDeviceToHost, r->stream));
	PTB_CUDA(cudaStreamSynchronize(r->stream));
	return 0;
}

int ptb_last_pass_f32(ptb_renderer* r, float* out_rgb)
{
	if (!r || r->host_only) { set_error("[Error]no CUDA device"); return 1; }
	PTB_CUDA(cudaMemcpyAsync(out_rgb, r->last_pass, (size_t)r->pixel_count * 3 * sizeof(float), cudaMemcpyDeviceToHost, r->stream));
	PTB_CUDA(cudaStreamSynchronize(r->stream));
	return 0;
}

void* ptb_image_device_ptr(ptb_renderer* r) { return (r && !r->host_only) ? (void*)r->image_sum : nullptr; }

int ptb_finalize(ptb_renderer* r, int total_passes)
{
	if (!r || r->host_only) { set_error("[Error]no CUDA device"); return 1; }
	k_tonemap<<<(r->pixel_count + 255) / 256, 256, 0, r->stream>>>(r->image_sum, r->image_u8, r->pixel_count, std::max(total_passes, 1), r->cfg.gamma_correction ? 1 : 0);
	PTB_CUDA(cudaStreamSynchronize(r->stream));
	r->pass_counter = total_passes;
	return 0;
}

int ptb_trace_batch(ptb_renderer* r, const float* rays6, int n, int32_t* out_prim, float* out_t, float* out_bary)
{
	return trace_impl(r, rays6, n, out_prim, out_t, out_bary, false);
}

int ptb_trace_batch_bruteforce(ptb_renderer* r, const float* rays6, int n, int32_t* out_prim, float* out_t)
{
	return trace_impl(r, rays6, n, out_prim, out_t, nullptr, true);
}

int ptb_generate_rays(ptb_renderer* r, int pass, float* out_rays6)
{
	if (!r || r->host_only) { set_error("[Error]no CUDA device"); return 1; }
	const int px = r->pixel_count;
	k_generate<false><<<grid_for(r, px, 256, 8), 256, 0, r->stream>>>(r->st, r->queue[0], r->counts, r->cfg.max_tracer_depth + 2, camera_params(r->cam), device_config(r), px, 1, pass, 1, 0);
	std::vector<float4> o(px), d(px);
	PTB_CUDA(cudaMemcpyAsync(o.data(), r->st.ray_o, (size_t)px * sizeof(float4), cudaMemcpyDeviceToHost, r->stream));
	PTB_CUDA(cudaMemcpyAsync(d.data(), r->st.ray_d, (size_t)px * sizeof(float4), cudaMemcpyDeviceToHost, r->stream));
	PTB_CUDA(cudaStreamSynchronize(r->stream));
	for (int i = 0; i < px; i++)
	{
		out_rays6[i * 6 + 0] = o[i].x; out_rays6[i * 6 + 1] = o[i].y; out_rays6[i * 6 + 2] = o[i].z;
		out_rays6[i * 6 + 3] = d[i].x; out_rays6[i * 6 + 4] = d[i].y; out_rays6[i * 6 + 5] = d[i].z;
	}
	return 0;
}

int ptb_capture_rays(ptb_renderer* r, int pass, int depth, int32_t* out_pixels, float* out_rays6, int max_out)
{
	if (!r || r->host_only) { set_error("[Error]no CUDA device"); return -1; }
	if (!r->scene_loaded) { set_error("[Error]no scene loaded"); return -1; }
	const int px = r->pixel_count;
	DeviceConfig dc = device_config(r);
	k_generate<false><<<grid_for(r, px, 256, 8), 256, 0, r->stream>>>(r->st, r->queue[0], r->counts, r->cfg.max_tracer_depth + 2, camera_params(r->cam), dc, px, 1, pass, 1, 0);
	int d = 0;
	for (; d < depth && d < r->cfg.max_tracer_depth; d++)
	{
		launch_extend(r, r->stream, px, r->st, r->queue[d & 1], r->counts + d, r->counts + (r->cfg.max_tracer_depth + 2) + d, d);
		k_shade<false, false><<<grid_for(r, px, 128, 16), 128, 0, r->stream>>>(r->dscene, r->st, dc, d, px, pass, 1, r->queue[d & 1], r->counts + d, r->queue[(d + 1) & 1], r->counts + d + 1, nullptr, 0);
	}
	int count = 0;
	if (cudaMemcpyAsync(&count, r->counts + d, sizeof(int), cudaMemcpyDeviceToHost, r->stream) != cudaSuccess || cudaStreamSynchronize(r->stream) != cudaSuccess)
	{
		set_error(std::string("[Cuda]") + cudaGetErrorString(cudaGetLastError()));
		return -1;
	}
	std::vector<int> q(count);
	std::vector<float4> o(px), dd(px);
	cudaMemcpy(q.data(), r->queue[d & 1], (size_t)count * sizeof(int), cudaMemcpyDeviceToHost);
	cudaMemcpy(o.data(), r->st.ray_o, (size_t)px * sizeof(float4), cudaMemcpyDeviceToHost);
	cudaMemcpy(dd.data(), r->st.ray_d, (size_t)px * sizeof(float4), cudaMemcpyDeviceToHost);
	// the wavefront's queue order is not the reference's stable order: report sorted by pixel
	std::sort(q.begin(), q.end());
	int n = std::min(count, max_out);
	for (int i = 0; i < n; i++)
	{
		int id = q[i];
		out_pixels[i] = id;
		out_rays6[i * 6 + 0] = o[id].x; out_rays6[i * 6 + 1] = o[id].y; out_rays6[i * 6 + 2] = o[id].z;
		out_rays6[i * 6 + 3] = dd[id].x; out_rays6[i * 6 + 4] = dd[id].y; out_rays6[i * 6 + 5] = dd[id].z;
	}
	return n;
}

int ptb_get_stats(ptb_renderer* r, ptb_stats* out) { if (!r || !out) return 1; *out = r->stats; return 0; }

int ptb_get_depth_profile(ptb_renderer* r, int max_entries, int64_t* out_segments, double* out_extend_ms)
{
	if (!r) return -1;
	int n = std::min(max_entries, r->cfg.max_tracer_depth);
	for (int d = 0; d < n; d++)
	{
		out_segments[d] = d < (int)r->depth_segments.size() ? r->depth_segments[d] : 0;
		out_extend_ms[d] = d < (int)r->depth_extend_ms.size() ? r->depth_extend_ms[d] : 0.0;
	}
	return n;
}

// ---- acceleration-structure introspection (test / diagnostic hooks; binary layout only) ----
// Walks a host copy of the DEVICE arrays the traversal kernels read, so it validates what was actually
// built whichever builder produced it.
int ptb_bvh_info(ptb_renderer* r, int64_t* out_i, double* out_d)
{
	if (!r || r->host_only) { set_error("[Error]no CUDA device"); return 1; }
	if (!r->scene_loaded) { set_error("[Error]no scene loaded"); return 1; }
	if (r->dscene.bvh_layout != 2) { set_error("[Error]ptb_bvh_info: binary layout only"); return 1; }
	cudaSetDevice(r->device);
	const int n = r->dscene.n_triangles;
	const int64_t n_nodes = r->bvh_nodes;
	std::vector<float> nodes((size_t)n_nodes * 16), tris((size_t)n * 12);
	if (n_nodes) PTB_CUDA(cudaMemcpy(nodes.data(), r->dscene.bvh_nodes, nodes.size() * sizeof(float), cudaMemcpyDeviceToHost));
	if (n) PTB_CUDA(cudaMemcpy(tris.data(), r->dscene.tri_isect, tris.size() * sizeof(float), cudaMemcpyDeviceToHost));
	std::vector<int> seen(n, 0);
	int64_t inner = 0, leaves = 0, max_depth = 0, bad = 0;
	double cost = 0.0, root_area = 0.0;
	auto area = [](const float* lo, const float* hi) { double dx = (double)hi[0] - lo[0], dy = (double)hi[1] - lo[1], dz = (double)hi[2] - lo[2]; return (dx < 0 || dy < 0 || dz < 0) ? 0.0 : dx * dy + dy * dz + dz * dx; };
	struct Item { int ref; int depth; float lo[3], hi[3]; };
	std::vector<Item> stack;
	if (n > 0 && n_nodes > 0)
	{
		Item root; root.ref = r->dscene.root_ref; root.depth = 0;
		const float* d = &nodes[(size_t)root.ref * 16];
		root.lo[0] = std::min(d[0], d[4]); root.hi[0] = std::max(d[1], d[5]);
		root.lo[1] = std::min(d[2], d[6]); root.hi[1] = std::max(d[3], d[7]);
		root.lo[2] = std::min(d[8], d[10]); root.hi[2] = std::max(d[9], d[11]);
		if (d[4] > d[5]) { root.lo[0] = d[0]; root.hi[0] = d[1]; root.lo[1] = d[2]; root.hi[1] = d[3]; root.lo[2] = d[8]; root.hi[2] = d[9]; }
		root_area = area(root.lo, root.hi);
		stack.push_back(root);
	}
	while (!stack.empty())
	{
		Item it = stack.back();
		stack.pop_back();
		max_depth = std::max<int64_t>(max_depth, it.depth);
		if (it.ref < 0)
		{
			const int ref = ~it.ref, first = ref >> 3, cnt = (ref & 7) + 1;
			leaves++;
			cost += (root_area > 0 ? area(it.lo, it.hi) / root_area : 0.0) * 1.5 * cnt;
			for (int k = 0; k < cnt; k++)
			{
				if (first + k >= n) { bad++; continue; }
				const float* t = &tris[(size_t)(first + k) * 12];
				int id; memcpy(&id, &t[3], 4);
				if (id < 0 || id >= n) { bad++; continue; }
				seen[id]++;
				// the three vertices must lie inside the (padded) leaf box
				for (int v = 0; v < 3; v++)
					for (int a = 0; a < 3; a++)
					{
						float x = t[a] + (v == 1 ? t[4 + a] : (v == 2 ? t[8 + a] : 0.0f));
						float slack = 1e-5f * (std::fabs(x) + std::fabs(t[a])) + 1e-30f;   // v0 + e1 re-rounds
						if (!(x >= it.lo[a] - slack && x <= it.hi[a] + slack)) bad++;
					}
			}
			continue;
		}
		if (it.ref >= n_nodes || it.depth > 4096) { bad++; continue; }
		inner++;
		cost += root_area > 0 ? area(it.lo, it.hi) / root_area : 0.0;
		const float* d = &nodes[(size_t)it.ref * 16];
		int refs[2]; memcpy(refs, &d[12], 8);
		Item c0, c1;
		c0.ref = refs[0]; c1.ref = refs[1]; c0.depth = c1.depth = it.depth + 1;
		c0.lo[0] = d[0]; c0.hi[0] = d[1]; c0.lo[1] = d[2]; c0.hi[1] = d[3]; c0.lo[2] = d[8]; c0.hi[2] = d[9];
		c1.lo[0] = d[4]; c1.hi[0] = d[5]; c1.lo[1] = d[6]; c1.hi[1] = d[7]; c1.lo[2] = d[10]; c1.hi[2] = d[11];
		const bool c1_empty = c1.lo[0] > c1.hi[0];
		for (int a = 0; a < 3; a++)
		{
			if (c0.lo[a] < it.lo[a] || c0.hi[a] > it.hi[a]) bad++;
			if (!c1_empty && (c1.lo[a] < it.lo[a] || c1.hi[a] > it.hi[a])) bad++;
		}
		if (!c1_empty) stack.push_back(c1);
		stack.push_back(c0);
	}
	for (int i = 0; i < n; i++) if (seen[i] != 1) bad++;
	if (out_i)
	{
		out_i[0] = n_nodes; out_i[1] = inner; out_i[2] = leaves; out_i[3] = max_depth; out_i[4] = bad == 0 ? 1 : 0;
		out_i[5] = r->bvh_built_on_gpu; out_i[6] = r->bvh_levels; out_i[7] = r->bvh_small_tasks;
		out_i[8] = r->bvh8_nodes; out_i[9] = r->bvh_collapsed_on_gpu;
	}
	if (out_d) { out_d[0] = r->bvh_build_ms; out_d[1] = cost; out_d[2] = r->scene_upload_ms; out_d[3] = (double)bad; }
	return 0;
}

// label[t] = smallest triangle id sharing t's leaf: two builds have the same leaf partition iff the labels agree
int ptb_bvh_leaf_labels(ptb_renderer* r, int32_t* out_label)
{
	if (!r || r->host_only) { set_error("[Error]no CUDA device"); return 1; }
	if (!r->scene_loaded) { set_error("[Error]no scene loaded"); return 1; }
	if (r->dscene.bvh_layout != 2) { set_error("[Error]ptb_bvh_leaf_labels: binary layout only"); return 1; }
	cudaSetDevice(r->device);
	const int n = r->dscene.n_triangles;
	const int64_t n_nodes = r->bvh_nodes;
	std::vector<float> nodes((size_t)n_nodes * 16), tris((size_t)n * 12);
	if (n_nodes) PTB_CUDA(cudaMemcpy(nodes.data(), r->dscene.bvh_nodes, nodes.size() * sizeof(float), cudaMemcpyDeviceToHost));
	if (n) PTB_CUDA(cudaMemcpy(tris.data(), r->dscene.tri_isect, tris.size() * sizeof(float), cudaMemcpyDeviceToHost));
	for (int i = 0; i < n; i++) out_label[i] = -1;
	std::vector<int> stack;
	if (n > 0 && n_nodes > 0) stack.push_back(r->dscene.root_ref);
	size_t guard = 0;
	while (!stack.empty() && guard++ < (size_t)4 * n + 16)
	{
		int ref = stack.back();
		stack.pop_back();
		if (ref < 0)
		{
			const int lr = ~ref, first = lr >> 3, cnt = (lr & 7) + 1;
			int lo = 0x7fffffff;
			for (int k = 0; k < cnt && first + k < n; k++) { int id; memcpy(&id, &tris[(size_t)(first + k) * 12 + 3], 4); lo = std::min(lo, id); }
			for (int k = 0; k < cnt && first + k < n; k++) { int id; memcpy(&id, &tris[(size_t)(first + k) * 12 + 3], 4); if (id >= 0 && id < n) out_label[id] = lo; }
			continue;
		}
		if (ref >= n_nodes) continue;
		const float* d = &nodes[(size_t)ref * 16];
		int refs[2]; memcpy(refs, &d[12], 8);
		if (!(d[4] > d[5])) stack.push_back(refs[1]);
		stack.push_back(refs[0]);
	}
	return 0;
}

// raw copy of the device arrays (binary layout): 16 floats per node record, triangle id per leaf slot
int ptb_bvh_download(ptb_renderer* r, float* out_nodes16, int32_t* out_leaf_order)
{
	if (!r || r->host_only) { set_error("[Error]no CUDA device"); return 1; }
	if (!r->scene_loaded) { set_error("[Error]no scene loaded"); return 1; }
	if (r->dscene.bvh_layout != 2) { set_error("[Error]ptb_bvh_download: binary layout only"); return 1; }
	cudaSetDevice(r->device);
	const int n = r->dscene.n_triangles;
	if (out_nodes16 && r->bvh_nodes) PTB_CUDA(cudaMemcpy(out_nodes16, r->dscene.bvh_nodes, (size_t)r->bvh_nodes * 64, cudaMemcpyDeviceToHost));
	if (out_leaf_order && n)
	{
		std::vector<float> tris((size_t)n * 12);
		PTB_CUDA(cudaMemcpy(tris.data(), r->dscene.tri_isect, tris.size() * sizeof(float), cudaMemcpyDeviceToHost));
		for (int i = 0; i < n; i++) memcpy(&out_leaf_order[i], &tris[(size_t)i * 12 + 3], 4);
	}
	return 0;
}

// node visits per ray of the last count_traversal render (binary-tree kernel): out[0] = most visits of one ray,
// out[1 + k] = rays with floor(log2(visits + 1)) == k, k < 24
int ptb_get_traversal_histogram(ptb_renderer* r, int64_t* out25)
{
	if (!r || !out25) return 1;
	out25[0] = r->traversal_histogram[2];
	for (int k = 0; k < 24; k++) out25[1 + k] = r->traversal_histogram[4 + k];
	return 0;
}

int ptb_set_option(ptb_renderer* r, const char* key, const char* value)
{
	if (!r || !key || !value) return 1;
	std::string k = key, v = value;
	if (k == "passes_in_flight")
	{
		int n = atoi(value);
		if (n < 1 || n > 64) { set_error("[Error]passes_in_flight must be in 1..64"); return 1; }
		if (n != r->passes_in_flight)
		{
			r->passes_in_flight = n;
			if (!r->host_only)
			{
				cudaSetDevice(r->device);
				cudaStreamSynchronize(r->stream);
				free_work_buffers(r);
				if (alloc_work_buffers(r)) return 1;
				r->pass_counter = 0;
			}
		}
		return 0;
	}
	if (k == "streams_in_flight")
	{
		int n = atoi(value);
		if (n < 1 || n > 8) { set_error("[Error]streams_in_flight must be in 1..8"); return 1; }
		if (n != r->streams_in_flight)
		{
			r->streams_in_flight = n;
			if (!r->host_only)
			{
				cudaSetDevice(r->device);
				cudaDeviceSynchronize();
				free_work_buffers(r);
				if (alloc_work_buffers(r)) return 1;
				r->pass_counter = 0;
			}
		}
		return 0;
	}
	if (k == "active_streams") { r->active_streams = atoi(value); return 0; }
	if (k == "profile_stages") { r->profile_stages = atoi(value); return 0; }
	if (k == "count_traversal") { r->count_traversal = atoi(value); return 0; }
	if (k == "bvh_builder")
	{
		if (v != "gpu_sah" && v != "host_sah") { set_error("[Error]bvh_builder must be gpu_sah or host_sah"); return 1; }
		r->bvh_builder = v;   // takes effect at the next ptb_load_scene
		return 0;
	}
	if (k == "bvh_max_leaf") { int n = atoi(value); if (n < 1 || n > 8) { set_error("[Error]bvh_max_leaf must be in 1..8"); return 1; } r->bvh_max_leaf = n; return 0; }
	if (k == "bvh_intersect_cost") { float c = (float)atof(value); if (!(c > 0.0f)) { set_error("[Error]bvh_intersect_cost must be > 0"); return 1; } r->bvh_intersect_cost = c; return 0; }
	if (k == "loader_threads") { set_loader_threads(atoi(value)); return 0; }   // process-wide, like ptb_set_jpeg_decode
	if (k == "loader_mesh_lanes") { set_loader_mesh_lanes(atoi(value)); return 0; }   // process-wide
	if (k == "loader_per_vertex") { set_loader_per_vertex(atoi(value)); return 0; }   // process-wide
	if (k == "tile_order") { r->tile_order = atoi(value); return 0; }
	if (k == "l2_persist") { r->l2_persist = atoi(value) != 0; return 0; }      // takes effect at the next ptb_load_scene / geometry edit
	if (k == "octant_order") { r->octant_order = atoi(value); return 0; }
	if (k == "sort_by_material") { r->sort_by_material = atoi(value); return 0; }
	if (k == "russian_roulette") { r->russian_roulette = atoi(value) != 0; return ptb_clear(r); }
	if (k == "inline_scatter") { r->inline_scatter = atoi(value) != 0; return 0; }
	if (k == "fused_from_depth") { r->fused_from_depth = atoi(value); return 0; }
	if (k == "tune_scatter") { r->tune_scatter = atoi(value); return 0; }
	if (k == "sampler")
	{
		if (v != "reference" && v != "pcg") { set_error("[Error]sampler must be reference or pcg"); return 1; }
		r->sampler = v == "pcg" ? 1 : 0;
		return ptb_clear(r);
	}
	if (k == "sss")
	{
		if (v != "reference" && v != "per_channel") { set_error("[Error]sss must be reference or per_channel"); return 1; }
		r->sss_mode = v == "per_channel" ? 1 : 0;
		return ptb_clear(r);
	}
	if (k == "texture_filter")
	{
		if (v != "software" && v != "hardware") { set_error("[Error]texture_filter must be software or hardware"); return 1; }
		r->hw_textures = v == "hardware" ? 1 : 0;   // takes effect at the next ptb_load_scene
		return 0;
	}
	if (k == "pass_clamp") { r->pass_clamp = (float)atof(value); return 0; }   // diagnostic: per-pass clamp of the accumulation (default: the reference's)
	if (k == "estimator")
	{
		if (v != "reference" && v != "nee") { set_error("[Error]estimator must be reference or nee"); return 1; }
		const int want = v == "nee" ? 1 : 0;
		if (want != r->nee)
		{
			r->nee = want;
			if (!r->host_only)
			{
				cudaSetDevice(r->device);
				cudaDeviceSynchronize();
				free_work_buffers(r);
				if (alloc_work_buffers(r)) return 1;
				r->pass_counter = 0;
			}
		}
		return 0;
	}
	if (k == "extend_persistent") { r->extend_persistent = atoi(value); return 0; }
	if (k == "extend_variant") { r->extend_variant = atoi(value); return 0; }
	if (k == "fused_upwalk") { r->fused_upwalk = atoi(value) != 0; return 0; }
	if (k == "tune_refill_f") { r->tune_refill_f = atoi(value); return 0; }
	if (k == "tune_leaf_f") { r->tune_leaf_f = atoi(value); return 0; }
	if (k == "upwalk_min_nodes") { r->upwalk_min_nodes = std::max(1, atoi(value)); return 0; }
	if (k == "upwalk") { r->upwalk = atoi(value) != 0; return 0; }      // takes effect at the next scene load / geometry edit
	if (k == "tune_refill_u") { r->tune_refill_u = atoi(value); return 0; }
	if (k == "tune_leaf_u") { r->tune_leaf_u = atoi(value); return 0; }
	if (k == "tune_reps_u") { r->tune_reps_u = std::max(1, atoi(value)); return 0; }
	if (k == "entry_min_passes") { r->entry_min_passes = std::max(1, atoi(value)); return 0; }
	if (k == "sky_fast") { r->sky_fast = atoi(value) != 0; return 0; }
	if (k == "entry_cuts") { r->entry_cuts = atoi(value) != 0; return 0; }
	if (k == "entry_k") { r->entry_k = std::max(1, std::min(atoi(value), PTB_ENTRY_STRIDE - 1)); return 0; }
	if (k == "entry_tile")
	{
		// "WxH", powers of two
		int tw = 0, th = 0;
		if (sscanf(value, "%dx%d", &tw, &th) != 2 || tw < 1 || th < 1 || tw > 64 || th > 64 || (tw & (tw - 1)) || (th & (th - 1))) { set_error("[Error]entry_tile: WxH with powers of two up to 64"); return 1; }
		r->entry_tile_w = tw; r->entry_tile_h = th;
		return 0;
	}
	if (k == "tune_refill_e") { r->tune_refill_e = atoi(value); return 0; }
	if (k == "tune_leaf_e") { r->tune_leaf_e = atoi(value); return 0; }
	if (k == "tune_reps_e") { r->tune_reps_e = std::max(1, atoi(value)); return 0; }
	if (k == "tune_refill4") { r->tune_refill4 = atoi(value); return 0; }
	if (k == "treelet_block") { r->treelet_block = atoi(value); return 0; }
	if (k == "treelet_nodes") { r->treelet_nodes = atoi(value); return 0; }
	if (k == "sort_depth_mask") { r->sort_depth_mask = atoi(value); return 0; }
	if (k == "sort_cell_bits") { r->sort_cell_bits = atoi(value); return 0; }
	if (k == "sort_octant") { r->sort_octant = atoi(value); return 0; }
	if (k == "tune_refill") { r->tune_refill = atoi(value); return 0; }
	if (k == "tune_leaf") { r->tune_leaf = atoi(value); return 0; }
	if (k == "tune_reps") { r->tune_reps = atoi(value); return 0; }
	if (k == "unroll_reps") { r->unroll_reps = atoi(value); return 0; }
	if (k == "tune_refill8") { r->tune_refill8 = atoi(value); return 0; }
	if (k == "tune_leaf8") { r->tune_leaf8 = atoi(value); return 0; }
	if (k == "persistent_grid") { r->persistent_grid = atoi(value); return 0; }
	if (k == "persistent_grid8") { r->persistent_grid8 = atoi(value); return 0; }
	if (k == "bvh_collapse") { if (v != "gpu" && v != "host") { set_error("[Error]bvh_collapse must be gpu or host"); return 1; } r->bvh_collapse = v; return 0; }
	if (k == "bvh_hybrid") { r->bvh_hybrid = atoi(value); return 0; }             // takes effect at the next ptb_load_scene
	if (k == "hybrid_from_depth") { r->hybrid_from_depth = atoi(value); r->hybrid_from_user = true; return 0; }
	if (k == "small_tree_bytes") { r->small_tree_bytes = atoll(value); return 0; }
	if (k == "bvh_layout")
	{
		int n = atoi(value);
		if (n != 2 && n != 8) { set_error("[Error]bvh_layout must be 2 or 8"); return 1; }
		r->bvh_layout = n;   // takes effect at the next ptb_load_scene
		return 0;
	}
	set_error("[Error]unknown option " + k);
	return 1;
}

int ptb_scene_counts(ptb_renderer* r, int* n_triangles, int* n_materials, int* n_spheres, int* n_textures, int* cube_length, int* n_meshes)
{
	if (!r || !r->scene_loaded) { set_error("[Error]no scene loaded"); return 1; }
	if (n_triangles) *n_triangles = (int)r->scene.triangles.size();
	if (n_materials) *n_materials = (int)r->scene.materials.size();
	if (n_spheres) *n_spheres = (int)r->scene.spheres.size();
	if (n_textures) *n_textures = (int)r->scene.textures.size();
	if (cube_length) *cube_length = r->scene.cube_length;
	if (n_meshes) *n_meshes = (int)r->scene.mesh_triangle_count.size();
	return 0;
}

int ptb_scene_triangles(ptb_renderer* r, float* out24, int32_t* out_material)
{
	if (!r || !r->scene_loaded) { set_error("[Error]no scene loaded"); return 1; }
	if (out24) memcpy(out24, r->scene.triangles.data(), r->scene.triangles.size() * sizeof(Triangle));
	if (out_material) memcpy(out_material, r->scene.triangle_material.data(), r->scene.triangle_material.size() * sizeof(int32_t));
	return 0;
}

int ptb_scene_materials(ptb_renderer* r, ptb_material* out)
{
	if (!r || !r->scene_loaded) { set_error("[Error]no scene loaded"); return 1; }
	memcpy(out, r->scene.materials.data(), r->scene.materials.size() * sizeof(ptb_material));
	return 0;
}

int ptb_scene_spheres(ptb_renderer* r, void* out100)
{
	if (!r || !r->scene_loaded) { set_error("[Error]no scene loaded"); return 1; }
	memcpy(out100, r->scene.spheres.data(), r->scene.spheres.size() * sizeof(Sphere));
	return 0;
}

int ptb_scene_texture(ptb_renderer* r, int index, int* width, int* height, uint8_t* out_rgba)
{
	if (!r || !r->scene_loaded || index < 0 || index >= (int)r->scene.textures.size()) { set_error("[Error]bad texture index"); return 1; }
	const Texture& t = r->scene.textures[index];
	if (width) *width = t.width;
	if (height) *height = t.height;
	if (out_rgba) memcpy(out_rgba, t.rgba.data(), t.rgba.size());
	return 0;
}

// image_loader::load_image (Others/image_loader.cpp:31-95) on one file: RGBA8, row 0 = top, alpha 255
int ptb_decode_image(const char* path, int* width, int* height, uint8_t* out_rgba)
{
	Texture t;
	if (!path || !load_image_rgba8(path, t)) return 1;
	if (width) *width = t.width;
	if (height) *height = t.height;
	if (out_rgba) memcpy(out_rgba, t.rgba.data(), t.rgba.size());
	return 0;
}

// which libjpeg parameters the JPEG decoder reproduces (process-wide; read by every later load)
int ptb_set_jpeg_decode(const char* mode)
{
	const std::string m = mode ? mode : "";
	if (m == "reference") set_jpeg_mode(kJpegReference);
	else if (m == "fast") set_jpeg_mode(kJpegFast);
	else if (m == "accurate") set_jpeg_mode(kJpegAccurate);
	else { set_error("[Error]jpeg decode mode must be reference, fast or accurate"); return 1; }
	return 0;
}

int ptb_scene_cubemap_face(ptb_renderer* r, int face, uint8_t* out_rgba)
{
	if (!r || !r->scene_loaded || face < 0 || face > 5) { set_error("[Error]bad cube face"); return 1; }
	memcpy(out_rgba, r->scene.cube_faces[face].rgba.data(), r->scene.cube_faces[face].rgba.size());
	return 0;
}

int ptb_get_config(ptb_renderer* r, void* out96)
{
	if (!r || !out96) return 1;
	memcpy(out96, &r->cfg, sizeof(Config));
	return 0;
}

} // extern "C"

#include "compat.inc"
#include "multi.inc"
