// ptb200 wavefront integrator, stage 1: k_generate (camera rays + path-state reset) and the free-flight bound shared with k_shade.
// Included by render.cu only (one translation unit).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include "kernels.cuh"

namespace ptb
{

using namespace ptbdev;

// Upper bound on the hit distance the NEXT bounce can use (stored in ray_d.w).  In a scattering
// medium the reference draws a free-flight distance d = -__logf(u0) / sigma_s'.x first thing in the
// bounce and, when d < t_hit, scatters WITHOUT looking at the hit (path_tracer_kernel.cu:460-486).
// The RNG stream of a bounce depends only on (pass, pixel, depth), so d is known before the ray is
// traced: the closest-hit search can stop at d.  Subsurface random walks (mean free path << object
// size) then cost a handful of node visits instead of a full traversal, with identical results.
// Per-channel subsurface scattering (option sss=per_channel; the reference's TODO at path_tracer_kernel.cu:456-464,
// which samples the free flight from sigma_s'.x alone): the channel the free-flight distance is drawn from, one uniform pick per bounce from
// its own stream.  Used by next_bounce_bound (which bounds the NEXT closest-hit search by that distance) and by k_shade — same pick.
__device__ __forceinline__ float sss_sampling_sigma(const DeviceConfig& cfg, float3 sigma_s, int seed, int pixel_index, int depth)
{
	if (cfg.sss_mode == 0) return sigma_s.x;
	Rng pick;
	pick.seed3(cfg.sampler, seed, pixel_index, depth, 0x51ed270bu, 0.0f, 1.0f);
	const float u = pick.next();
	return u < (1.0f / 3.0f) ? sigma_s.x : (u < (2.0f / 3.0f) ? sigma_s.y : sigma_s.z);
}

__device__ __forceinline__ bool medium_participates(const DeviceConfig& cfg, float3 sigma_a, float3 sigma_s)
{
	if (cfg.sss_mode == 0) return sigma_s.x > 0.0f || length(sigma_a) > cfg.sss_threshold;
	return fmaxf(sigma_s.x, fmaxf(sigma_s.y, sigma_s.z)) > 0.0f || length(sigma_a) > cfg.sss_threshold;
}

__device__ __forceinline__ float next_bounce_bound(const DeviceConfig& cfg, float3 sigma_a, float3 sigma_s, int seed, int pixel_index, int depth)
{
	if (!medium_participates(cfg, sigma_a, sigma_s)) return CUDART_INF_F;
	Rng rng;
	rng.seed3(cfg.sampler, seed, pixel_index, depth, 0u, 0.0f, 1.0f);
	const float scattering_distance = -__logf(rng.next()) / sss_sampling_sigma(cfg, sigma_s, seed, pixel_index, depth);
	// hits with t <= d are still needed (the test is d < t_hit): bound = next float above d.
	// NaN (0/0) never compares less than t_hit -> no clipping.
	if (!(scattering_distance == scattering_distance)) return CUDART_INF_F;
	if (scattering_distance < 0.0f || scattering_distance == 0.0f) return 1e-37f;
	if (scattering_distance >= 3.0e38f) return CUDART_INF_F;
	return __uint_as_float(__float_as_uint(scattering_distance) + 1u);
}

// the same bound for the reference's sampler with hash(seed) * hash(pixel) already multiplied (k_extend_persistent8<.., FUSED> computes
// it once per scatter event for this bounce's stream and the next one's)
__device__ __forceinline__ float next_bounce_bound_hashed(const DeviceConfig& cfg, float3 sigma_a, float3 sigma_s, int hash_seed_pixel, int depth)
{
	if (!(sigma_s.x > 0.0f || length(sigma_a) > cfg.sss_threshold)) return CUDART_INF_F;
	Rng rng;
	rng.seed((uint32_t)(hash_seed_pixel * hash_ref(depth)), 0.0f, 1.0f);
	const float scattering_distance = -__logf(rng.next()) / sigma_s.x;
	if (!(scattering_distance == scattering_distance)) return CUDART_INF_F;
	if (scattering_distance < 0.0f || scattering_distance == 0.0f) return 1e-37f;
	if (scattering_distance >= 3.0e38f) return CUDART_INF_F;
	return __uint_as_float(__float_as_uint(scattering_distance) + 1u);
}

// ------------------------------------------------------------------------------------------
// k_generate — init_data_kernel + generate_ray_kernel fused (path_tracer_kernel.cu:275-379)
// ------------------------------------------------------------------------------------------
// SKY (enqueue_batch, with valid entry cuts): a camera ray whose 8x4 pixel tile has an EMPTY entry cut (kernels_entry.cuh: no box of the
// tree overlaps the tile's shaft) cannot hit a triangle.  Without spheres and with air that does not participate, its bounce at depth 0 is the
// miss branch of k_shade: radiance = (1, 1, 1) * background(d) added to 0 — the background colour, exactly.  Such a path is finished
// here: it is never written, queued, searched or shaded (on c2 43 % of the camera rays).  The queue keeps its tile order without any
// atomic: k_tile_rank (kernels_entry.cuh, once per camera) numbers the non-empty tiles, tile t of pass slot s owns the 32 queue entries
// from (s * n_nonempty + rank[t]) * 32.  counts[0] = n_slots * n_nonempty * 32 rays queued; the others (total - counts[0]) are added to the
// depth-0 segment total by k_accumulate.
struct SkyArgs
{
	SkyParams sky;
	const int* tile_rank;     // per 8x4 tile (queue order): index among the non-empty tiles, -1 = empty entry cut
	const int* n_nonempty;
};

template <bool ALT = false, bool SKY = false>
__global__ void __launch_bounds__(256) k_generate(PathState st, int* __restrict__ queue, int* __restrict__ counts, int n_counts,
	CameraParams cam, DeviceConfig cfg, int pixel_count, int n_slots, int first_pass, int pass_stride, int tiles_x, SkyArgs sa = SkyArgs())
{
	if (!ALT) { cfg.sampler = 0; cfg.sss_mode = 0; }   // the default instantiation is the reference's sampler, folded at compile time
	const int total = pixel_count * n_slots;
	int tid = blockIdx.x * blockDim.x + threadIdx.x;
	const int n_nonempty = SKY ? __ldg(sa.n_nonempty) : 0;
	const int queued_total = SKY ? n_slots * n_nonempty * 32 : total;
	// counts[0..n_counts) = live paths per depth; counts[n_counts..2*n_counts) = per-depth work-fetch cursors of the persistent
	// extend kernel; counts[2*n_counts..3*n_counts) = shadow rays per depth (estimator "nee")
	for (int c = tid; c < 3 * n_counts; c += gridDim.x * blockDim.x) counts[c] = c == 0 ? queued_total : 0;   // any grid size, any MaxDepth
	for (int i = tid; i < total; i += gridDim.x * blockDim.x)
	{
		int slot = i / pixel_count;
		int pixel = i - slot * pixel_count;
		int q = i;                 // queue position
		bool finished = false;
		if (tiles_x > 0)
		{
			// queue position -> pixel in 8x4 tiles: the 32 rays of a warp cover a compact screen patch, so
			// they visit nearly the same nodes (higher L1 hit rate, lanes finish together)
			const int tile = pixel >> 5, within = pixel & 31;
			const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
			pixel = (ty * 4 + (within >> 3)) * (tiles_x * 8) + tx * 8 + (within & 7);
			if (SKY)
			{
				const int rank = __ldg(sa.tile_rank + tile);
				finished = rank < 0;
				q = (slot * n_nonempty + rank) * 32 + within;
			}
		}
		const int id = slot * pixel_count + pixel;
		int seed = first_pass + slot * pass_stride;
		float3 o, d;
		generate_camera_ray(cam, pixel, seed, cfg.use_anti_alias != 0, o, d, cfg.sampler);
		if (SKY && finished)
		{
			const float3 bg = background_color(sa.sky, d);
			st.radiance[id] = make_float4(bg.x, bg.y, bg.z, 0.0f);
			continue;
		}
		st.ray_o[id] = make_float4(o.x, o.y, o.z, 0.0f);
		st.ray_d[id] = make_float4(d.x, d.y, d.z, next_bounce_bound(cfg, cfg.air_sigma_a, cfg.air_sigma_s, seed, pixel, 0));
		st.radiance[id] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		queue[q] = id;
	}
}

} // namespace ptb
