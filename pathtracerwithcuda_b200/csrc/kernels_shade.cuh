// ptb200 wavefront integrator, stages 3-4: k_shade (medium, material, Fresnel branch, sky, compaction, optional light sampling),
// k_shadow (visibility of the light samples), k_accumulate / k_tonemap.  Included by render.cu only, after kernels_extend.cuh.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include "kernels.cuh"

namespace ptb
{

using namespace ptbdev;

// ------------------------------------------------------------------------------------------
// k_shade — medium, material, Fresnel, branch, sky (path_tracer_kernel.cu:456-624)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 sample_texture(const DeviceTexture& tex, float2 uv, bool use_bilinear)
{
	float ux = uv.x - floorf(uv.x);
	float uy = uv.y - floorf(uv.y);
	// option texture_filter=hardware: the bilinear filter of texture.h:22-67 done by the texture unit on a cudaArray copy
	// of the same RGBA8 image — texel i sits at i + 0.5 in unnormalised texture coordinates, clamp addressing like the software filter's
	// clamped floor / ceil indices.  9-bit filter weights: NOT bit-identical to the software filter, which stays the parity mode (default).
	if (use_bilinear && tex.tex) return sample_image_hw(tex.tex, tex.width, tex.height, ux, 1.0f - uy);
	return sample_image(tex.pixels, tex.width, tex.height, ux, 1.0f - uy, use_bilinear);
}

// SORT: the 128 queue entries a block handles per iteration are first reordered in shared memory by
// (miss | sphere | triangle material), so warps shade runs of one material ("shade-by-material"): the
// material fetch, texture sampling and the conductor / dielectric Fresnel paths stop diverging inside a
// warp.  The stochastic reflect / refract / diffuse choice still diverges — it is decided inside.
// NEE (estimator "nee", SURVEY.md 8f rank 4 — NOT the reference's estimator, off by default): at a diffuse bounce the
// direct light of the emissive triangles is estimated by one area sample + shadow ray (k_shadow), and the
// continuing path does not add the emission of an emissive triangle it hits next.  Expected value as in the
// reference: the reference picks emission up with probability (1 - F) at the light (Fresnel branch first,
// path_tracer_kernel.cu:529-616), so the light sample carries that factor; same bounce limit and energy cut.
#ifndef PTB_SHADE_MIN_BLOCKS
#define PTB_SHADE_MIN_BLOCKS 10   // 48 registers: measured optimum (profiles/r01_experiments.md); the kernel is latency / HBM bound, occupancy pays
#endif
// Russian roulette (estimator option; absent in the reference): from bounce PTB_RR_START_DEPTH on a path
// survives with probability p = clamp(max(throughput), 0.05, 1) and its throughput is divided by p — unbiased, fewer deep
// segments.  Own random stream, so the surviving path is still the one the reference estimator follows.
#ifndef PTB_RR_START_DEPTH
#define PTB_RR_START_DEPTH 3
#endif
__device__ __forceinline__ bool russian_roulette(float3& not_absorbed, int seed, int pixel_index, int depth, int sampler)
{
	Rng rr;
	rr.seed3(sampler, seed, pixel_index, depth, 0x3c6ef372u, 0.0f, 1.0f);
	const float p = fminf(fmaxf(fmaxf(not_absorbed.x, fmaxf(not_absorbed.y, not_absorbed.z)), 0.05f), 1.0f);
	if (rr.next() >= p) return false;
	not_absorbed = not_absorbed * (1.0f / p);
	return true;
}

// spectral single-sample MIS weights of option sss=per_channel (balance heuristic over the uniformly picked channel)
__device__ __forceinline__ float3 sss_scatter_weight(float3 sigma_s, float d)
{
	const float3 p = make_float3(sigma_s.x * __expf(-sigma_s.x * d), sigma_s.y * __expf(-sigma_s.y * d), sigma_s.z * __expf(-sigma_s.z * d));
	const float mean = (p.x + p.y + p.z) * (1.0f / 3.0f);
	return mean > 0.0f ? make_float3(p.x / mean, p.y / mean, p.z / mean) : make_float3(0.0f, 0.0f, 0.0f);
}
__device__ __forceinline__ float3 sss_survive_weight(float3 sigma_s, float t)
{
	// sigma = 0: the channel never scatters (and 0 * inf must not appear when the ray misses everything, t = inf)
	const float3 p = make_float3(sigma_s.x > 0.0f ? __expf(-sigma_s.x * t) : 1.0f, sigma_s.y > 0.0f ? __expf(-sigma_s.y * t) : 1.0f,
		sigma_s.z > 0.0f ? __expf(-sigma_s.z * t) : 1.0f);
	const float mean = (p.x + p.y + p.z) * (1.0f / 3.0f);
	return mean > 0.0f ? make_float3(p.x / mean, p.y / mean, p.z / mean) : make_float3(0.0f, 0.0f, 0.0f);
}

// ALT: the instantiation that honours the estimator options sampler / sss; the default instantiations fold both to the reference's
// behaviour at compile time, so the parity path carries none of their branches (measured: a run-time switch cost k_shade 2.4 % on c2).
// FUSED (option inline_scatter): the closest-hit kernel of the deeper bounces performs the medium's scatter events itself
// (k_extend_persistent8<.., true>, kernels_extend.cuh), so a path may be several bounces AHEAD of the wavefront's loop depth when it
// arrives here: ray_o.w carries that lead (an int), every depth-keyed quantity uses loop depth + lead, a path the extend stage
// finished off carries the primitive PTB_PRIM_DEAD, and the bounce limit is checked per path.
// survivors are queued with one atomic per BLOCK and iteration (1) or per warp (0): c2 +1.1 %, c3 +1.7 % (profiles/r02_experiments.md) — fewer
// same-address atomics, and 128 consecutive paths stay together in the next queue
#ifndef PTB_SHADE_BLOCK_COMPACT
#define PTB_SHADE_BLOCK_COMPACT 1
#endif
template <bool SORT, bool NEE, bool RR = false, bool ALT = false, bool FUSED = false>
__global__ void __launch_bounds__(128, PTB_SHADE_MIN_BLOCKS) k_shade(DeviceScene sc, PathState st, DeviceConfig cfg, int loop_depth, int pixel_count, int first_pass, int pass_stride,
	const int* __restrict__ queue_in, const int* __restrict__ count_in, int* __restrict__ queue_out, int* __restrict__ count_out, int* __restrict__ shadow_count,
	int octant_order)
{
	if (!ALT) { cfg.sampler = 0; cfg.sss_mode = 0; }
	__shared__ int s_oct[10];
#if PTB_SHADE_BLOCK_COMPACT
	__shared__ int s_cmp[5];
#endif
	__shared__ int s_ids[SORT ? 128 : 1];
	__shared__ int s_hist[SORT ? 16 : 1];
	const int count = *count_in;
	const unsigned lane = threadIdx.x & 31;
	const float pixel_count_inv = 1.0f / (float)pixel_count;
	// block-uniform trip count (barriers in the SORT path), hence warp-uniform: the ballots below are convergent
	for (int block_base = blockIdx.x * blockDim.x; block_base < count; block_base += gridDim.x * blockDim.x)
	{
		int i = block_base + threadIdx.x;
		bool valid = i < count;
		bool alive = false;
		bool want_shadow = false;
		int oct_key = 0;
		int id = 0;
		int path_lead = 0;
		if (SORT)
		{
			int key = 15;
			if (threadIdx.x < 16) s_hist[threadIdx.x] = 0;
			__syncthreads();
			if (valid)
			{
				id = queue_in[i];
				const int prim = __float_as_int(st.hit[id].w);
				key = prim == -1 ? 0 : (prim < -1 ? 1 : 2 + min(__float_as_int(__ldg(&sc.tri_shade[(size_t)prim * 4 + 3]).w), 12));
			}
			// counting sort over 16 keys: rank inside the key by warp-aggregated atomics
			const unsigned peers = __match_any_sync(0xffffffffu, key);
			const int leader = __ffs(peers) - 1;
			int base_in_key = 0;
			if ((int)lane == leader) base_in_key = atomicAdd(&s_hist[key], __popc(peers));
			base_in_key = __shfl_sync(0xffffffffu, base_in_key, leader) + __popc(peers & ((1u << lane) - 1u));
			__syncthreads();
			int before = 0;
			for (int k = 0; k < key; k++) before += s_hist[k];
			s_ids[before + base_in_key] = valid ? id : -1;
			__syncthreads();
			id = s_ids[threadIdx.x];
			valid = id >= 0;
			if (!valid) id = 0;
			__syncthreads();
		}
		else if (valid) id = queue_in[i];
		if (valid)
		{
			// id / pixel_count with a float estimate and one correction (the quotient is a batch slot, < 2^10; exact for every id)
			int slot = (int)((float)id * pixel_count_inv);
			int pixel_index = id - slot * pixel_count;
			if (pixel_index < 0) { slot--; pixel_index += pixel_count; }
			else if (pixel_index >= pixel_count) { slot++; pixel_index -= pixel_count; }
			int seed = first_pass + slot * pass_stride;

			// at depth 0 the throughput is (1, 1, 1) in air by construction (init_data_kernel :275-297): k_generate does not
			// write it and it is not read here
			float4 o4 = st.ray_o[id], d4 = st.ray_d[id], h4 = st.hit[id];
			const int lead = FUSED ? PTB_LEAD_OF(__float_as_int(o4.w)) : 0;
			path_lead = lead;
			const int depth = loop_depth + lead;
			float4 t4 = depth == 0 ? make_float4(1.0f, 1.0f, 1.0f, __int_as_float(-1)) : st.throughput[id];
			float3 ray_o = make_float3(o4.x, o4.y, o4.z);
			float3 ray_d = make_float3(d4.x, d4.y, d4.z);
			float3 not_absorbed = make_float3(t4.x, t4.y, t4.z);
			int medium_index = __float_as_int(t4.w);
			float min_t = h4.x, min_t1 = h4.y, min_t2 = h4.z;
			int prim = __float_as_int(h4.w);

			Rng rng;
			rng.seed3(cfg.sampler, seed, pixel_index, depth, 0u, 0.0f, 1.0f);

			float3 sigma_a = cfg.air_sigma_a, sigma_s = cfg.air_sigma_s;
			if (medium_index >= 0)
			{
				float4 md = __ldg(&sc.materials[medium_index].d), me = __ldg(&sc.materials[medium_index].e);
				sigma_a = make_float3(md.x, md.y, md.z);
				sigma_s = make_float3(md.w, me.x, me.y);
			}

			bool done = false;
			alive = true;
			if (FUSED && prim == PTB_PRIM_DEAD) { done = true; alive = false; }   // ended inside the closest-hit kernel (energy cut or bounce limit)
			else if (medium_participates(cfg, sigma_a, sigma_s))
			{
				float rand = rng.next();
				float scattering_distance = -__logf(rand) / sss_sampling_sigma(cfg, sigma_s, seed, pixel_index, depth);
				if (scattering_distance < min_t)
				{
					float rand1 = rng.next();
					float rand2 = rng.next();
					float3 next_o = ray_o + ray_d * scattering_distance;
					float3 next_d = sample_on_sphere(rand1, rand2);
					not_absorbed = not_absorbed * absorption_through_medium(sigma_a, scattering_distance);
					// per-channel mode: the distance came from ONE channel's sigma_s' picked uniformly; single-sample MIS over the three
					// channels weights channel c by sigma_c exp(-sigma_c d) / mean_j(sigma_j exp(-sigma_j d)) (1 when the three are equal)
					if (cfg.sss_mode) not_absorbed = not_absorbed * sss_scatter_weight(sigma_s, scattering_distance);
					st.ray_o[id] = make_float4(next_o.x, next_o.y, next_o.z, NEE ? 0.0f : __int_as_float(lead | PTB_FROM_BITS_OF(__float_as_int(o4.w))));   // not on a surface: the walk keeps the triangle it entered through as the start of its searches
					st.ray_d[id] = make_float4(next_d.x, next_d.y, next_d.z, next_bounce_bound(cfg, sigma_a, sigma_s, seed, pixel_index, depth + 1));
					oct_key = (next_d.x < 0.0f ? 4 : 0) | (next_d.y < 0.0f ? 2 : 0) | (next_d.z < 0.0f ? 1 : 0);
					st.throughput[id] = make_float4(not_absorbed.x, not_absorbed.y, not_absorbed.z, t4.w);
					if (length(not_absorbed) <= cfg.energy_threshold) alive = false;
					if (RR && alive && depth >= PTB_RR_START_DEPTH)
					{
						alive = russian_roulette(not_absorbed, seed, pixel_index, depth, cfg.sampler);
						st.throughput[id] = make_float4(not_absorbed.x, not_absorbed.y, not_absorbed.z, t4.w);
					}
					done = true;
				}
				else
				{
					not_absorbed = not_absorbed * absorption_through_medium(sigma_a, min_t);
					// per-channel mode: reaching the surface unscattered has probability mean_j exp(-sigma_j t) under the channel pick
					if (cfg.sss_mode) not_absorbed = not_absorbed * sss_survive_weight(sigma_s, min_t);
				}
			}

			if (!done)
			{
				if (prim != -1)
				{
					float3 diffuse_color, emission_color, specular_color, min_normal, min_point;
					float roughness, mat_n, mat_k;
					float2 uv0 = make_float2(0.0f, 0.0f), uv1 = uv0, uv2 = uv0;
					bool is_transparent;
					int material_index;
					if (prim < -1)
					{
						int s = -(prim + 2);
						material_index = sc.sphere_material_base + s;
						float4 sp = __ldg(&sc.spheres[s]);
						min_point = ray_o + ray_d * min_t;
						min_normal = normalize(min_point - make_float3(sp.x, sp.y, sp.z));
					}
					else
					{
						const float4* sh = sc.tri_shade + (size_t)prim * 4;
						float4 s0 = __ldg(sh + 0), s1 = __ldg(sh + 1), s2 = __ldg(sh + 2), s3 = __ldg(sh + 3);
						material_index = __float_as_int(s3.w);
						float3 normal0 = make_float3(s0.x, s0.y, s0.z), normal1 = make_float3(s0.w, s1.x, s1.y), normal2 = make_float3(s1.z, s1.w, s2.x);
						min_normal = normal0 * (1.0f - min_t1 - min_t2) + normal1 * min_t1 + normal2 * min_t2;
						min_point = ray_o + ray_d * min_t;
						uv0 = make_float2(s2.y, s2.z); uv1 = make_float2(s2.w, s3.x); uv2 = make_float2(s3.y, s3.z);
					}
					const DeviceMaterial* mp = &sc.materials[material_index];
					float4 ma = __ldg(&mp->a), mb = __ldg(&mp->b), mc = __ldg(&mp->c), me = __ldg(&mp->e), mf = __ldg(&mp->f);
					diffuse_color = make_float3(ma.x, ma.y, ma.z);
					emission_color = make_float3(mb.x, mb.y, mb.z);
					specular_color = make_float3(mc.x, mc.y, mc.z);
					is_transparent = __float_as_int(me.z) != 0;
					if (prim >= 0)
					{
						int diffuse_tex = __float_as_int(me.w), specular_tex = __float_as_int(mf.x);
						if (diffuse_tex != -1 || specular_tex != -1)
						{
							float2 uv = uv0 * (1.0f - min_t1 - min_t2) + uv1 * min_t1 + uv2 * min_t2;
							if (diffuse_tex != -1) diffuse_color = diffuse_color * sample_texture(sc.textures[diffuse_tex], uv, cfg.use_bilinear != 0);
							if (specular_tex != -1) specular_color = specular_color * sample_texture(sc.textures[specular_tex], uv, cfg.use_bilinear != 0);
						}
					}
					roughness = ma.w; mat_n = mb.w; mat_k = mc.w;

					float3 in_direction = ray_d;
					float in_n = cfg.air_n, out_n = mat_n;
					float out_k = mat_k;
					int in_medium = -1, out_medium = material_index;

					bool is_hit_on_back = dot(in_direction, min_normal) > 0;
					if (is_hit_on_back)
					{
						min_normal = min_normal * -1.0f;
						if (is_transparent)
						{
							float tn = in_n; in_n = out_n; out_n = tn;
							int tm = in_medium; in_medium = out_medium; out_medium = tm;
							out_k = 0.0f;
						}
					}

					float3 reflection_direction = reflection(min_normal, in_direction);
					float3 refraction_direction = refraction(min_normal, in_direction, in_n, out_n);
					float3 bias_vector = cfg.bias_length * min_normal;
					(void)reflection_direction;

					float fresnel_reflection;
					if (mat_k == 0 || is_transparent) fresnel_reflection = fresnel_dielectric(min_normal, in_direction, in_n, out_n, refraction_direction);
					else fresnel_reflection = fresnel_conductor(min_normal, in_direction, out_n, out_k);

					float rand = rng.next();
					float3 next_o, next_d;
					float medium_bits = t4.w;
					bool nee_candidate = false;
					float nee_flag = 0.0f;
					if (rand < fresnel_reflection)
					{
						float rand1 = rng.next();
						float rand2 = rng.next();
						float remap_roughness = __powf(roughness, 1.85f) * 0.238f;
						float3 micro_normal = sample_on_hemisphere_ggx_weight(min_normal, remap_roughness, rand1, rand2);
						float3 micro_reflection_direction = reflection(micro_normal, in_direction);
						float self_shadowing = ggx_shadowing_masking(remap_roughness, min_normal, micro_normal, ray_d) *
							ggx_shadowing_masking(remap_roughness, min_normal, micro_normal, micro_reflection_direction);
						next_o = min_point + bias_vector;
						next_d = micro_reflection_direction;
						not_absorbed = not_absorbed * (specular_color * self_shadowing);
					}
					else if (is_transparent)
					{
						next_o = min_point - bias_vector;
						next_d = refraction_direction;
						medium_bits = __int_as_float(out_medium);
						not_absorbed = not_absorbed * __powf((out_n / in_n), 2.0f);
					}
					else
					{
						// NEE: an emissive triangle reached straight after a bounce that sampled the lights is already counted
						const bool counted_by_nee = NEE && o4.w != 0.0f && prim >= 0 && (emission_color.x != 0.0f || emission_color.y != 0.0f || emission_color.z != 0.0f);
						if (!counted_by_nee)
						{
							// accumulated += not_absorbed * emission (:608-609); adding an exact zero changes nothing, so the
							// read-modify-write is skipped for non-emissive hits (a NaN / inf product still goes through)
							const float3 add = not_absorbed * emission_color;
							if (!(add.x == 0.0f && add.y == 0.0f && add.z == 0.0f))
							{
								const float4 r4 = st.radiance[id];
								st.radiance[id] = make_float4(r4.x + add.x, r4.y + add.y, r4.z + add.z, 0.0f);
							}
						}
						not_absorbed = not_absorbed * diffuse_color;
						float rand1 = rng.next();
						float rand2 = rng.next();
						next_o = min_point + bias_vector;
						next_d = sample_on_hemisphere_cosine_weight(min_normal, rand1, rand2);
						nee_candidate = true;
					}
					float3 nsa = sigma_a, nss = sigma_s;
					{
						// medium the next segment travels in: unchanged unless the path refracted
						const int next_medium = __float_as_int(medium_bits);
						if (next_medium != medium_index)
						{
							nsa = cfg.air_sigma_a; nss = cfg.air_sigma_s;
							if (next_medium >= 0)
							{
								float4 md = __ldg(&sc.materials[next_medium].d), me2 = __ldg(&sc.materials[next_medium].e);
								nsa = make_float3(md.x, md.y, md.z);
								nss = make_float3(md.w, me2.x, me2.y);
							}
						}
						st.ray_d[id] = make_float4(next_d.x, next_d.y, next_d.z, next_bounce_bound(cfg, nsa, nss, seed, pixel_index, depth + 1));
						oct_key = (next_d.x < 0.0f ? 4 : 0) | (next_d.y < 0.0f ? 2 : 0) | (next_d.z < 0.0f ? 1 : 0);
					}
					if (length(not_absorbed) <= cfg.energy_threshold) alive = false;
					if (NEE && nee_candidate && alive && depth + 1 < cfg.max_depth && sc.n_lights > 0 && !(nss.x > 0.0f || length(nsa) > cfg.sss_threshold))
					{
						nee_flag = 1.0f;
						// own random stream: the path itself is the one the reference estimator follows
						Rng lrng;
						lrng.seed3(cfg.sampler, seed, pixel_index, depth, 0x68bc21ebu, 0.0f, 1.0f);
						const float u0 = lrng.next(), u1 = lrng.next(), u2 = lrng.next();
						int lo = 0, hi = sc.n_lights - 1;
						while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(&sc.light_cdf[mid]) < u0) lo = mid + 1; else hi = mid; }
						const int lt = __ldg(&sc.light_tri[lo]);
						const float* tv = sc.tris24 + (size_t)lt * 24;
						const float3 lv0 = make_float3(__ldg(tv + 0), __ldg(tv + 1), __ldg(tv + 2));
						const float3 lv1 = make_float3(__ldg(tv + 3), __ldg(tv + 4), __ldg(tv + 5));
						const float3 lv2 = make_float3(__ldg(tv + 6), __ldg(tv + 7), __ldg(tv + 8));
						const float su = sqrtf(u1);
						const float b0 = 1.0f - su, b1 = su * (1.0f - u2), b2 = su * u2;
						const float3 y = lv0 * b0 + lv1 * b1 + lv2 * b2;
						const float3 to = y - next_o;
						const float r2 = dot(to, to);
						const float rr = sqrtf(r2);
						const float3 wi = to * (1.0f / rr);
						const float cos_x = dot(normalize(min_normal), wi);
						const float3 ng = cross(lv1 - lv0, lv2 - lv0);
						const float ng_len = length(ng);
						const float cos_l = fabsf(dot(ng, wi)) / ng_len;
						if (cos_x > 0.0f && cos_l > 0.0f && rr > 0.0f && ng_len > 0.0f)
						{
							const float4* lsh = sc.tri_shade + (size_t)lt * 4;
							const float4 l0 = __ldg(lsh + 0), l1 = __ldg(lsh + 1), l2 = __ldg(lsh + 2), l3 = __ldg(lsh + 3);
							float3 ln = make_float3(l0.x, l0.y, l0.z) * b0 + make_float3(l0.w, l1.x, l1.y) * b1 + make_float3(l1.z, l1.w, l2.x) * b2;
							if (dot(wi, ln) > 0) ln = ln * -1.0f;
							const DeviceMaterial* lm = &sc.materials[__float_as_int(l3.w)];
							const float4 lmb = __ldg(&lm->b), lmc = __ldg(&lm->c);
							const float3 refr = refraction(ln, wi, cfg.air_n, lmb.w);
							const float fl = lmc.w == 0 ? fresnel_dielectric(ln, wi, cfg.air_n, lmb.w, refr) : fresnel_conductor(ln, wi, lmb.w, lmc.w);
							const float keep = fminf(fmaxf(1.0f - fl, 0.0f), 1.0f);      // probability of the emission branch at the light
							const float w = keep * cos_x * cos_l * sc.light_area / (3.14159265358979f * r2);
							const float3 c = not_absorbed * make_float3(lmb.x, lmb.y, lmb.z) * w;
							if (c.x > 0.0f || c.y > 0.0f || c.z > 0.0f)
							{
								st.shadow_o[id] = make_float4(next_o.x, next_o.y, next_o.z, rr * 0.9999f);
								st.shadow_d[id] = make_float4(wi.x, wi.y, wi.z, 0.0f);
								st.shadow_c[id] = make_float4(c.x, c.y, c.z, 0.0f);
								want_shadow = true;
							}
						}
					}
					if (RR && alive && depth >= PTB_RR_START_DEPTH) alive = russian_roulette(not_absorbed, seed, pixel_index, depth, cfg.sampler);
					// ray_o.w: the NEE flag, or (kernels.cuh: PTB_FROM_BITS) the lead of a path ahead of the loop (FUSED only) + the triangle this
					// segment leaves — the bounce-ray kernels start their search at that triangle's leaf (none for a sphere)
					st.ray_o[id] = make_float4(next_o.x, next_o.y, next_o.z, NEE ? nee_flag : __int_as_float(lead | PTB_FROM_BITS(prim)));
					st.throughput[id] = make_float4(not_absorbed.x, not_absorbed.y, not_absorbed.z, medium_bits);
				}
				else
				{
					float3 bg = background_color(sc.sky, ray_d);
					float4 r4 = st.radiance[id];
					float3 add = not_absorbed * bg;
					st.radiance[id] = make_float4(r4.x + add.x, r4.y + add.y, r4.z + add.z, 0.0f);
					alive = false;
				}
			}
		}
		// a path ahead of the loop runs out of bounces before the loop does (the reference's depth loop ends at MaxDepth for every path)
		if (FUSED && alive && loop_depth + path_lead + 1 >= cfg.max_depth) alive = false;
		if (octant_order)
		{
			// block-level compaction grouped by the direction octant of the next ray: the 32 rays an extend warp
			// fetches together then descend the tree in the same child order
			if (threadIdx.x < 10) s_oct[threadIdx.x] = 0;
			__syncthreads();
			const int key = alive ? oct_key : 8;
			const unsigned peers = __match_any_sync(0xffffffffu, key);
			const int leader = __ffs(peers) - 1;
			int in_key = 0;
			if ((int)lane == leader && alive) in_key = atomicAdd(&s_oct[key], __popc(peers));
			in_key = __shfl_sync(0xffffffffu, in_key, leader) + __popc(peers & ((1u << lane) - 1u));
			__syncthreads();
			if (threadIdx.x == 0)
			{
				int total = 0;
				for (int k = 0; k < 8; k++) total += s_oct[k];
				s_oct[9] = total ? atomicAdd(count_out, total) : 0;
			}
			__syncthreads();
			if (alive)
			{
				int before = s_oct[9];
				for (int k = 0; k < key; k++) before += s_oct[k];
				queue_out[before + in_key] = id;
			}
			__syncthreads();
		}
		else
		{
		// stream compaction of survivors (replaces thrust::remove_if + host sync)
		unsigned mask = __ballot_sync(0xffffffffu, alive);
#if PTB_SHADE_BLOCK_COMPACT
		// one atomic per BLOCK and iteration: the warps' counts meet in shared memory
		{
			const int warp = threadIdx.x >> 5;
			if (lane == 0) s_cmp[warp] = __popc(mask);
			__syncthreads();
			if (threadIdx.x == 0)
			{
				int sum = 0;
				for (int k = 0; k < 4; k++) { const int c = s_cmp[k]; s_cmp[k] = sum; sum += c; }
				s_cmp[4] = sum ? atomicAdd(count_out, sum) : 0;
			}
			__syncthreads();
			if (alive) queue_out[s_cmp[4] + s_cmp[warp] + __popc(mask & ((1u << lane) - 1u))] = id;
			__syncthreads();
		}
#else
		// one atomic per warp
		if (mask)
		{
			int pos = 0;
			if (lane == 0) pos = atomicAdd(count_out, __popc(mask));
			pos = __shfl_sync(0xffffffffu, pos, 0);
			if (alive) queue_out[pos + __popc(mask & ((1u << lane) - 1u))] = id;
		}
#endif
		}
		if (NEE)
		{
			const unsigned smask = __ballot_sync(0xffffffffu, want_shadow);
			if (smask)
			{
				int pos = 0;
				if (lane == 0) pos = atomicAdd(shadow_count, __popc(smask));
				pos = __shfl_sync(0xffffffffu, pos, 0);
				if (want_shadow) st.shadow_queue[pos + __popc(smask & ((1u << lane) - 1u))] = id;
			}
		}
	}
}

// ------------------------------------------------------------------------------------------
// k_shadow — visibility of the NEE light samples (any hit strictly before the light point)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_shadow(DeviceScene sc, PathState st, const int* __restrict__ count_ptr)
{
	const int count = *count_ptr;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
	{
		const int id = st.shadow_queue[i];
		const float4 o4 = st.shadow_o[id], d4 = st.shadow_d[id];
		const float3 o = make_float3(o4.x, o4.y, o4.z), d = make_float3(d4.x, d4.y, d4.z);
		const float t_max = o4.w;
		bool blocked = false;
		for (int s = 0; s < sc.n_spheres && !blocked; s++)
		{
			const float4 sp = __ldg(&sc.spheres[s]);
			float t;
			if (intersect_sphere(make_float3(sp.x, sp.y, sp.z), sp.w, o, d, t) && t > 0.0f && t < t_max) blocked = true;
		}
		if (!blocked && sc.n_triangles > 0)
		{
			const float tiny = 1e-30f;
			const float3 ds = make_float3(fabsf(d.x) < tiny ? copysignf(tiny, d.x) : d.x, fabsf(d.y) < tiny ? copysignf(tiny, d.y) : d.y,
				fabsf(d.z) < tiny ? copysignf(tiny, d.z) : d.z);
			const float3 idir = make_float3(1.0f / ds.x, 1.0f / ds.y, 1.0f / ds.z);
			const float3 noidir = make_float3(-o.x * idir.x, -o.y * idir.y, -o.z * idir.z);
			const float margin2 = 4.8e-7f * fmaxf(fmaxf(fabsf(noidir.x), fabsf(noidir.y)), fabsf(noidir.z));
			int stack[PTB_STACK_SIZE];
			int sp = 0;
			int node = sc.root_ref;
			while (!blocked)
			{
				if (node >= 0)
				{
					const float4* np = sc.bvh_nodes + (size_t)node * 4;
					float4 n0, n1, n2;
					float2 n3;
					load_node(np, n0, n1, n2, n3);
					const float c0x0 = fmaf(n0.x, idir.x, noidir.x), c0x1 = fmaf(n0.y, idir.x, noidir.x);
					const float c0y0 = fmaf(n0.z, idir.y, noidir.y), c0y1 = fmaf(n0.w, idir.y, noidir.y);
					const float c0z0 = fmaf(n2.x, idir.z, noidir.z), c0z1 = fmaf(n2.y, idir.z, noidir.z);
					const float c1x0 = fmaf(n1.x, idir.x, noidir.x), c1x1 = fmaf(n1.y, idir.x, noidir.x);
					const float c1y0 = fmaf(n1.z, idir.y, noidir.y), c1y1 = fmaf(n1.w, idir.y, noidir.y);
					const float c1z0 = fmaf(n2.z, idir.z, noidir.z), c1z1 = fmaf(n2.w, idir.z, noidir.z);
					const float tmin0 = fmaxf(fmaxf(fminf(c0x0, c0x1), fminf(c0y0, c0y1)), fmaxf(fminf(c0z0, c0z1), 0.0f));
					const float tmax0 = fminf(fminf(fmaxf(c0x0, c0x1), fmaxf(c0y0, c0y1)), fminf(fmaxf(c0z0, c0z1), t_max));
					const float tmin1 = fmaxf(fmaxf(fminf(c1x0, c1x1), fminf(c1y0, c1y1)), fmaxf(fminf(c1z0, c1z1), 0.0f));
					const float tmax1 = fminf(fminf(fmaxf(c1x0, c1x1), fmaxf(c1y0, c1y1)), fminf(fmaxf(c1z0, c1z1), t_max));
					const bool h0 = fmaf(tmin0, PTB_SLACK_LO / PTB_SLACK_HI, -margin2) <= tmax0;
					const bool h1 = fmaf(tmin1, PTB_SLACK_LO / PTB_SLACK_HI, -margin2) <= tmax1;
					const int child0 = __float_as_int(n3.x), child1 = __float_as_int(n3.y);
					if (h0 && h1) { if (sp < PTB_STACK_SIZE) stack[sp++] = child1; node = child0; }
					else if (h0) node = child0;
					else if (h1) node = child1;
					else { if (sp == 0) break; node = stack[--sp]; }
				}
				else
				{
					const int ref = ~node;
					const int first = ref >> 3, cnt = (ref & 7) + 1;
					for (int k = 0; k < cnt && !blocked; k++)
					{
						const float4* tp = sc.tri_isect + (size_t)(first + k) * 3;
						const float4 a = __ldg(tp + 0), b = __ldg(tp + 1), c = __ldg(tp + 2);
						float t, t1, t2;
						if (intersect_triangle(make_float3(a.x, a.y, a.z), make_float3(b.x, b.y, b.z), make_float3(c.x, c.y, c.z), o, d, t, t1, t2) && t > 0.0f && t < t_max)
							blocked = true;
					}
					if (blocked || sp == 0) break;
					node = stack[--sp];
				}
			}
		}
		if (!blocked)
		{
			const float4 c = st.shadow_c[id];
			float4 r4 = st.radiance[id];
			st.radiance[id] = make_float4(r4.x + c.x, r4.y + c.y, r4.z + c.z, 0.0f);
		}
	}
}

// ------------------------------------------------------------------------------------------
// k_accumulate / k_tonemap — pixel_256_transform_gamma_corrected_kernel (:627-682) split in two
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_accumulate(const float4* __restrict__ radiance, float* __restrict__ image_sum, float* __restrict__ last_pass,
	const int* __restrict__ counts, unsigned long long* __restrict__ segment_totals, int n_counts,
	int pixel_count, int n_slots, float clamp_hi, int generated = 0)
{
	int p = blockIdx.x * blockDim.x + threadIdx.x;
	// generated > 0: k_generate<.., SKY> finished the camera rays of empty tiles itself (background colour) and queued only counts[0] of the
	// `generated` rays; the others are depth-0 segments like the queued ones
	if (p == 0 && generated > 0 && n_counts > 0 && counts != nullptr) atomicAdd(&segment_totals[0], (unsigned long long)(generated - counts[0]));
	// tally this batch's live-path counters (ray segments per depth) into the call totals; n_counts stops at the first loop depth whose
	// closest-hit kernel tallies its own searches (option inline_scatter)
	if (p < n_counts && counts != nullptr) atomicAdd(&segment_totals[p], (unsigned long long)counts[p]);
	if (p >= pixel_count) return;
	float sx = image_sum[p * 3 + 0], sy = image_sum[p * 3 + 1], sz = image_sum[p * 3 + 2];
	float4 r = make_float4(0, 0, 0, 0);
	for (int s = 0; s < n_slots; s++)
	{
		r = radiance[(size_t)s * pixel_count + p];
		sx += clampf(r.x, 0.0f, clamp_hi);
		sy += clampf(r.y, 0.0f, clamp_hi);
		sz += clampf(r.z, 0.0f, clamp_hi);
	}
	image_sum[p * 3 + 0] = sx; image_sum[p * 3 + 1] = sy; image_sum[p * 3 + 2] = sz;
	last_pass[p * 3 + 0] = r.x; last_pass[p * 3 + 1] = r.y; last_pass[p * 3 + 2] = r.z;
}

__global__ void __launch_bounds__(256) k_tonemap(const float* __restrict__ image_sum, uint8_t* __restrict__ image_u8, int pixel_count, int pass_counter, int gamma_correction)
{
	int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= pixel_count) return;
	float3 pixel = make_float3(image_sum[p * 3 + 0] / (float)pass_counter, image_sum[p * 3 + 1] / (float)pass_counter, image_sum[p * 3 + 2] / (float)pass_counter);
	float x, y, z;
	if (gamma_correction)
	{
		float inverse_gamma = 0.45454545f;
		float cx = __expf(inverse_gamma * __logf(pixel.x));
		float cy = __expf(inverse_gamma * __logf(pixel.y));
		float cz = __expf(inverse_gamma * __logf(pixel.z));
		x = clampf(cx * 255.0f, 0.0f, 255.0f);
		y = clampf(cy * 255.0f, 0.0f, 255.0f);
		z = clampf(cz * 255.0f, 0.0f, 255.0f);
	}
	else
	{
		x = clampf(pixel.x * 255.0f, 0.0f, 255.0f);
		y = clampf(pixel.y * 255.0f, 0.0f, 255.0f);
		z = clampf(pixel.z * 255.0f, 0.0f, 255.0f);
	}
	image_u8[p * 3 + 0] = (uint8_t)x;
	image_u8[p * 3 + 1] = (uint8_t)y;
	image_u8[p * 3 + 2] = (uint8_t)z;
}

__global__ void k_iota(int* q, int* count, int n)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i == 0) *count = n;
	if (i < n) q[i] = i;
}

} // namespace ptb
