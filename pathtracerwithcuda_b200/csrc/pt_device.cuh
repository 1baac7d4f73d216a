// Device-side arithmetic of the ptb200 integrator.
//
// Every function restates the arithmetic of the reference integrator
// (/root/reference/gpu_path_tracer/Kernel/path_tracer_kernel.cu and the headers it includes) in
// the SAME expression order, because the stochastic branch decisions (u < F, d < t_hit, energy
// kill) and the closest-hit winner (t < min_t) are discontinuous in those values
// (SURVEY.md Appendix G).  Where the winner is decided — Moller-Trumbore and the sphere roots —
// the fused-multiply-add placement nvcc chose for the reference (read from its sm_100a SASS:
// a*b - c*d -> fma(a,b,-(c*d)); x*x' + y*y' + z*z' -> fma(z,z', fma(x,x', y*y'))) is pinned with
// explicit intrinsics so it cannot drift with inlining context.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace ptbdev
{

#define PTB_PI 3.1415926535897f
#define PTB_TWO_PI 6.2831853071795f
#define PTB_E 2.7182818284590f
#define PTB_SQRT_ONE_THIRD 0.5773502691896f

// ---- float3 helpers (same shapes as Math/cuda_math.hpp:450-453,679-682,903-921,1389,1438,1458,1581) ----
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return make_float3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float b) { return make_float3(a.x * b, a.y * b, a.z * b); }
__device__ __forceinline__ float3 operator*(float b, float3 a) { return make_float3(b * a.x, b * a.y, b * a.z); }
__device__ __forceinline__ float2 operator+(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 operator*(float2 a, float b) { return make_float2(a.x * b, a.y * b); }

// dot / cross with the reference kernel's FMA placement pinned
__device__ __forceinline__ float dot(float3 a, float3 b)
{
	return __fmaf_rn(a.z, b.z, __fmaf_rn(a.x, b.x, __fmul_rn(a.y, b.y)));
}

__device__ __forceinline__ float3 cross(float3 a, float3 b)
{
	return make_float3(
		__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)),
		__fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)),
		__fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)));
}

__device__ __forceinline__ float length(float3 v) { return sqrtf(dot(v, v)); }
__device__ __forceinline__ float3 normalize(float3 v) { float inv_len = rsqrtf(dot(v, v)); return v * inv_len; }
__device__ __forceinline__ float3 lerp(float3 a, float3 b, float t) { return a + t * (b - a); }
__device__ __forceinline__ float clampf(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }

// ---- RNG: hash (path_tracer_kernel.cu:35-44) + thrust::minstd_rand + uniform_real_distribution<float> ----
__device__ __forceinline__ int hash_ref(int a_in)
{
	uint32_t a = (uint32_t)a_in;
	a = (a + 0x7ed55d16u) + (a << 12);
	a = (a ^ 0xc761c23cu) ^ (uint32_t)(((int)a) >> 19);
	a = (a + 0x165667b1u) + (a << 5);
	a = (a + 0xd3a2646cu) ^ (a << 9);
	a = (a + 0xfd7046c5u) + (a << 3);
	a = (a ^ 0xb55a4f09u) ^ (uint32_t)(((int)a) >> 16);
	return (int)a;
}

// Sampler option "pcg" (estimator upgrade, SURVEY.md 8f rank 4 — NOT a parity mode): the reference keys its streams by the PRODUCT
// hash(seed) * hash(pixel) * hash(depth), which collides whenever two factors swap or one hash is 0 / even-heavy; the pcg sampler keys a
// 32-bit PCG (Jarzynski & Olano 2020 output permutation over an LCG) by a nested hash of the three coordinates and a stream salt.
__device__ __forceinline__ uint32_t pcg_permute(uint32_t state)
{
	const uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
	return (word >> 22u) ^ word;
}
__device__ __forceinline__ uint32_t pcg_hash(uint32_t v) { return pcg_permute(v * 747796405u + 2891336453u); }

struct Rng
{
	uint32_t x;
	float lo, span;
	bool pcg;
	// x mod (2^31 - 1) without a division: 2^31 = 1 (mod m), so fold the high bits onto the low 31 and subtract m at most once —
	// exactly thrust::minstd_rand's arithmetic (linear_congruential_engine.h), a third of the instructions of the 64-bit '%'
	__device__ __forceinline__ static uint32_t mod_m31(uint64_t v)
	{
		uint32_t t = (uint32_t)(v & 0x7fffffffull) + (uint32_t)(v >> 31);   // v < 2^47: the sum stays below 2^32
		t = (t & 0x7fffffffu) + (t >> 31);                                    // fold once more: now t <= 2^31 - 1 + 1
		return t >= 2147483647u ? t - 2147483647u : t;
	}
	__device__ __forceinline__ void seed(uint32_t s, float a, float b)
	{
		x = mod_m31((uint64_t)s);
		if (x == 0u) x = 1u;
		lo = a; span = b - a;
		pcg = false;
	}
	// stream keyed by three coordinates: sampler 0 = the reference's product of hashes (xor salt), sampler 1 = pcg
	__device__ __forceinline__ void seed3(int sampler, int c0, int c1, int c2, uint32_t salt, float a, float b)
	{
		if (sampler == 0) { seed((uint32_t)(hash_ref(c0) * hash_ref(c1) * hash_ref(c2)) ^ salt, a, b); return; }
		x = pcg_hash((uint32_t)c2 + pcg_hash((uint32_t)c1 + pcg_hash((uint32_t)c0 ^ salt ^ 0x9e3779b9u)));
		lo = a; span = b - a;
		pcg = true;
	}
	__device__ __forceinline__ float next()
	{
		if (pcg)
		{
			x = x * 747796405u + 2891336453u;
			return (float)(pcg_permute(x) >> 8) * (1.0f / 16777216.0f) * span + lo;
		}
		x = mod_m31((uint64_t)x * 48271ull);
		float r = (float)(x - 1u);
		r /= 2147483648.0f;
		return (r * span) + lo;
	}
};

// ---- camera (path_tracer_kernel.cu:299-379) ----
struct CameraParams
{
	float3 eye, view, up;
	float2 resolution, fov;
	float aperture_radius, focal_distance;
};

__device__ __forceinline__ void generate_camera_ray(const CameraParams& cam, int pixel_index, int seed, bool use_anti_alias,
	float3& origin, float3& direction, int sampler = 0)
{
	float2 resolution = cam.resolution;
	int image_y = (int)(pixel_index / resolution.x);
	int image_x = pixel_index - (image_y * resolution.x);

	Rng rng;
	rng.seed3(sampler, seed, seed, pixel_index, sampler ? 0x243f6a88u : 0u, -0.5f, 0.5f);

	float jitter_x = 0.0f;
	float jitter_y = 0.0f;
	if (use_anti_alias)
	{
		jitter_x = rng.next();
		jitter_y = rng.next();
	}

	float distance = length(cam.view);
	float3 horizontal = normalize(cross(cam.view, cam.up));
	float3 vertical = normalize(cross(horizontal, cam.view));

	float3 x_axis = horizontal * (distance * __tanf(cam.fov.x * 0.5f * (PTB_PI / 180.0f)));
	float3 y_axis = vertical * (distance * __tanf(-cam.fov.y * 0.5f * (PTB_PI / 180.0f)));

	float normalized_image_x = (((float)image_x + jitter_x) / (resolution.x - 1.0f)) * 2.0f - 1.0f;
	float normalized_image_y = (((float)image_y + jitter_y) / (resolution.y - 1.0f)) * 2.0f - 1.0f;

	float3 point_on_canvas_plane = cam.eye + cam.view + normalized_image_x * x_axis + normalized_image_y * y_axis;
	float3 point_on_image_plane = cam.eye + normalize(point_on_canvas_plane - cam.eye) * cam.focal_distance;

	float3 point_on_aperture;
	if (cam.aperture_radius > 0.00001f)
	{
		float rand1 = rng.next() + 0.5f;
		float rand2 = rng.next() + 0.5f;
		float angle = rand1 * PTB_TWO_PI;
		float dist = cam.aperture_radius * sqrtf(rand2);
		float aperture_x = __cosf(angle) * dist;
		float aperture_y = __sinf(angle) * dist;
		point_on_aperture = cam.eye + aperture_x * horizontal + aperture_y * vertical;
	}
	else
	{
		point_on_aperture = cam.eye;
	}
	direction = normalize(point_on_image_plane - point_on_aperture);
	origin = point_on_aperture;
}

// ---- primitives ----

// Core/triangle.h:27-62 with edges precomputed (same subtraction, same value).
__device__ __forceinline__ bool intersect_triangle(float3 v0, float3 e1, float3 e2, float3 o, float3 d, float& t, float& t1, float& t2)
{
	float3 p_vec = cross(d, e2);
	float det = dot(e1, p_vec);
	if (det == 0.0f) return false;
	float inverse_det = 1.0f / det;
	float3 t_vec = make_float3(__fsub_rn(o.x, v0.x), __fsub_rn(o.y, v0.y), __fsub_rn(o.z, v0.z));
	float3 q_vec = cross(t_vec, e1);
	float a = __fmul_rn(dot(t_vec, p_vec), inverse_det);
	float b = __fmul_rn(dot(d, q_vec), inverse_det);
	if (a >= 0.0f && b >= 0.0f && __fadd_rn(a, b) <= 1.0f)
	{
		t = __fmul_rn(dot(e2, q_vec), inverse_det);
		t1 = a;
		t2 = b;
		return true;
	}
	return false;
}

// Core/sphere.h:18-55 (distance only; point/normal are recomputed by the shade stage)
__device__ __forceinline__ bool intersect_sphere(float3 center, float radius, float3 o, float3 d, float& hit_t)
{
	float3 op = make_float3(__fsub_rn(center.x, o.x), __fsub_rn(center.y, o.y), __fsub_rn(center.z, o.z));
	float b = dot(op, d);
	float delta = __fmaf_rn(radius, radius, __fmaf_rn(b, b, -dot(op, op)));
	if (delta < 0) return false;
	float delta_root = sqrtf(delta);
	float t1 = __fsub_rn(b, delta_root);
	float t2 = __fadd_rn(b, delta_root);
	if (t1 < 0 && t2 < 0) return false;
	if (t1 > 0 && t2 > 0) hit_t = fminf(t1, t2);
	else hit_t = fmaxf(t1, t2);
	return true;
}

// ---- scattering helpers (path_tracer_kernel.cu:46-83,163-273) ----
__device__ __forceinline__ float3 reflection(float3 normal, float3 in_direction)
{
	return in_direction - 2.0f * dot(normal, in_direction) * normal;
}

__device__ __forceinline__ float3 refraction(float3 normal, float3 in_direction, float in_refraction_index, float out_refraction_index)
{
	float3 i = in_direction * -1.0f;
	float n_dot_i = dot(normal, i);
	float refraction_ratio = in_refraction_index / out_refraction_index;
	// `a` and `a -/+ sqrt(b)` stay an un-fused multiply and add, b = fma(-(r*r), 1 - n.i*n.i, 1): that is what ptxas emits for the reference
	// (and for this kernel without a register cap); under __launch_bounds__ it would contract them into one FMA,
	// which moved 4 % of the bounce rays by an ulp (tools/ray_bits.py)
	float a = __fmul_rn(refraction_ratio, n_dot_i);
	float b = __fmaf_rn(-__fmul_rn(refraction_ratio, refraction_ratio), __fsub_rn(1.0f, __fmul_rn(n_dot_i, n_dot_i)), 1.0f);
	if (b < 0.0f) return make_float3(0.0f, 0.0f, 0.0f);
	if (n_dot_i > 0) return normal * __fsub_rn(a, sqrtf(b)) - refraction_ratio * i;
	return normal * __fadd_rn(a, sqrtf(b)) - refraction_ratio * i;
}

__device__ __forceinline__ float3 tangent_axis(float3 normal)
{
	// first axis whose |component| < sqrt(1/3), in x, y, z order (:175-187). If none qualifies the
	// reference reads an uninitialised vector; we fall back to z.
	if (fabsf(normal.x) < PTB_SQRT_ONE_THIRD) return make_float3(1.0f, 0.0f, 0.0f);
	if (fabsf(normal.y) < PTB_SQRT_ONE_THIRD) return make_float3(0.0f, 1.0f, 0.0f);
	return make_float3(0.0f, 0.0f, 1.0f);
}

__device__ __forceinline__ float3 sample_on_hemisphere_cosine_weight(float3 normal, float rand1, float rand2)
{
	float cos_theta = sqrtf(rand1);
	float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
	float phi = rand2 * PTB_TWO_PI;
	float3 any_direction = tangent_axis(normal);
	float3 vec_i = normalize(cross(normal, any_direction));
	float3 vec_j = cross(normal, vec_i);
	return cos_theta * normal + __cosf(phi) * sin_theta * vec_i + __sinf(phi) * sin_theta * vec_j;
}

__device__ __forceinline__ float3 sample_on_hemisphere_ggx_weight(float3 normal, float roughness, float rand1, float rand2)
{
	float theta = atanf(roughness * sqrtf(rand1) / sqrtf(1.0f - rand1));
	float phi = rand2 * PTB_TWO_PI;
	float cos_theta = __cosf(theta);
	float sin_theta = __sinf(theta);
	float3 any_direction = tangent_axis(normal);
	float3 vec_i = normalize(cross(normal, any_direction));
	float3 vec_j = cross(normal, vec_i);
	return cos_theta * normal + __cosf(phi) * sin_theta * vec_i + __sinf(phi) * sin_theta * vec_j;
}

__device__ __forceinline__ float3 sample_on_sphere(float rand1, float rand2)
{
	float cos_theta = rand1 * 2.0f - 1.0f;
	float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
	float phi = rand2 * PTB_TWO_PI;
	return make_float3(cos_theta, __cosf(phi) * sin_theta, __sinf(phi) * sin_theta);
}

__device__ __forceinline__ float3 absorption_through_medium(float3 absorption_coefficient, float distance)
{
	return make_float3(
		__powf(PTB_E, -1.0f * absorption_coefficient.x * distance),
		__powf(PTB_E, -1.0f * absorption_coefficient.y * distance),
		__powf(PTB_E, -1.0f * absorption_coefficient.z * distance));
}

__device__ __forceinline__ float ggx_shadowing_masking(float roughness, float3 macro_normal, float3 micro_normal, float3 ray_direction)
{
	float3 v = -1.0f * ray_direction;
	float v_dot_n = dot(v, macro_normal);
	float v_dot_m = dot(v, micro_normal);
	float positive_value = (v_dot_m / v_dot_n) > 0.0f ? 1.0f : 0.0f;
	if (positive_value == 0.0f) return 0.0f;
	float roughness_square = roughness * roughness;
	float cos_v_square = v_dot_n * v_dot_n;
	float tan_v_square = (1.0f - cos_v_square) / cos_v_square;
	return 2.0f / (1.0f + sqrtf(1.0f + roughness_square * tan_v_square));
}

// ---- Fresnel (Core/fresnel.h:11-76) ----
__device__ __forceinline__ float fresnel_dielectric(float3 normal, float3 in_direction, float n_in, float n_out, float3 refraction_direction)
{
	float cos_theta_in = dot(normal, in_direction * -1.0f);
	float cos_theta_out = dot(-1.0f * normal, refraction_direction);
	if (n_in > n_out && acosf(cos_theta_in) >= asinf(n_out / n_in)) return 1.0f;
	if (length(refraction_direction) <= 0.000005f || cos_theta_out < 0) return 1.0f;
	float rs = powf((n_in * cos_theta_in - n_out * cos_theta_out) / (n_in * cos_theta_in + n_out * cos_theta_out), 2.0f);
	float rp = powf((n_in * cos_theta_out - n_out * cos_theta_in) / (n_in * cos_theta_out + n_out * cos_theta_in), 2.0f);
	return (rs + rp) / 2.0f;
}

__device__ __forceinline__ float fresnel_conductor(float3 normal, float3 in_direction, float refraction_index, float extinction_coefficient)
{
	float cos_theta_in = dot(normal, in_direction * -1.0f);
	float refraction_index_square = refraction_index * refraction_index;
	float extinction_coefficient_square = extinction_coefficient * extinction_coefficient;
	float refraction_extinction_square_add = refraction_index_square + extinction_coefficient_square;
	float cos_theta_in_square = cos_theta_in * cos_theta_in;
	float two_refraction_cos_theta_in = 2 * refraction_index * cos_theta_in;
	float rs = (refraction_extinction_square_add * cos_theta_in_square - two_refraction_cos_theta_in + 1.0f) /
		(refraction_extinction_square_add * cos_theta_in_square + two_refraction_cos_theta_in + 1.0f);
	float rp = (refraction_extinction_square_add - two_refraction_cos_theta_in + cos_theta_in_square) /
		(refraction_extinction_square_add + two_refraction_cos_theta_in + cos_theta_in_square);
	return (rs + rp) / 2.0f;
}

// ---- textures and sky (Core/texture.h:15-79, Core/cube_map.h:20-119, Math/cuda_math.hpp:56-126) ----

// v / 255.0f for an integer v in [0, 255], correctly rounded like the IEEE division the reference performs (Core/texture.h) but in three
// instructions instead of the division's ~10 + range check: one Newton correction of v * RN(1 / 255) is exact for all 256 inputs
// (checked exhaustively with correctly rounded fmaf; tests/test_oracle_kat.py holds the same check against the oracle's division).
__device__ __forceinline__ float byte_over_255(unsigned char v)
{
	const float r = 0x1.010102p-8f;        // RN(1 / 255)
	const float fv = (float)v;
	const float q = __fmul_rn(fv, r);
	return __fmaf_rn(__fmaf_rn(-255.0f, q, fv), r, q);
}
__device__ __forceinline__ float3 texel_rgb(const uint8_t* __restrict__ pixels, int width, int x, int y)
{
	uchar4 p = __ldg(reinterpret_cast<const uchar4*>(pixels) + ((size_t)y * width + x));
	return make_float3(byte_over_255(p.x), byte_over_255(p.y), byte_over_255(p.z));
}

__device__ __forceinline__ float3 sample_image(const uint8_t* __restrict__ pixels, int width, int height, float u, float v_flipped, bool use_bilinear)
{
	// u in [0,1] maps to x = u*(w-1); v_flipped is (1 - v) already
	if (use_bilinear)
	{
		float x_image_real = u * (float)(width - 1);
		float y_image_real = v_flipped * (float)(height - 1);
		int floor_x = (int)clampf(floorf(x_image_real), 0.0f, (float)(width - 1));
		int ceil_x = (int)clampf(ceilf(x_image_real), 0.0f, (float)(width - 1));
		int floor_y = (int)clampf(floorf(y_image_real), 0.0f, (float)(height - 1));
		int ceil_y = (int)clampf(ceilf(y_image_real), 0.0f, (float)(height - 1));
		float left_right_t = x_image_real - floorf(x_image_real);
		float bottom_top_t = y_image_real - floorf(y_image_real);
		float3 c0 = texel_rgb(pixels, width, floor_x, floor_y);
		float3 c1 = texel_rgb(pixels, width, ceil_x, floor_y);
		float3 c2 = texel_rgb(pixels, width, floor_x, ceil_y);
		float3 c3 = texel_rgb(pixels, width, ceil_x, ceil_y);
		return lerp(lerp(c0, c1, left_right_t), lerp(c2, c3, left_right_t), bottom_top_t);
	}
	int x_image = (int)clampf((u * (float)(width - 1)), 0.0f, (float)(width - 1));
	int y_image = (int)clampf((v_flipped * (float)(height - 1)), 0.0f, (float)(height - 1));
	return texel_rgb(pixels, width, x_image, y_image);
}

__device__ __forceinline__ float3 sample_image_hw(cudaTextureObject_t tex, int width, int height, float u, float v_flipped)
{
	const float4 c = tex2D<float4>(tex, u * (float)(width - 1) + 0.5f, v_flipped * (float)(height - 1) + 0.5f);
	return make_float3(c.x, c.y, c.z);
}

__device__ __forceinline__ void cube_uv(float x, float y, float z, int& index, float& u, float& v)
{
	float ax = fabsf(x), ay = fabsf(y), az = fabsf(z);
	bool xp = x > 0, yp = y > 0, zp = z > 0;
	float max_axis = 0.0f, uc = 0.0f, vc = 0.0f;
	index = 0;
	// later matches override earlier ones, exactly like the chain of independent ifs
	if (xp && ax >= ay && ax >= az) { max_axis = ax; uc = -z; vc = y; index = 0; }
	if (!xp && ax >= ay && ax >= az) { max_axis = ax; uc = z; vc = y; index = 1; }
	if (yp && ay >= ax && ay >= az) { max_axis = ay; uc = x; vc = -z; index = 2; }
	if (!yp && ay >= ax && ay >= az) { max_axis = ay; uc = x; vc = z; index = 3; }
	if (zp && az >= ax && az >= ay) { max_axis = az; uc = x; vc = y; index = 4; }
	if (!zp && az >= ax && az >= ay) { max_axis = az; uc = -x; vc = y; index = 5; }
	u = 0.5f * (uc / max_axis + 1.0f);
	v = 0.5f * (vc / max_axis + 1.0f);
}

struct SkyParams
{
	const uint8_t* faces[6];
	int length;
	int use_sky_box, use_sky, use_bilinear;
	cudaTextureObject_t face_tex[6];   // option texture_filter=hardware, else 0
};

__device__ __forceinline__ float3 background_color(const SkyParams& sky, float3 direction)
{
	if (sky.use_sky_box)
	{
		float u, v;
		int index;
		cube_uv(direction.x, direction.y, direction.z, index, u, v);
		if (sky.use_bilinear && sky.face_tex[index]) return sample_image_hw(sky.face_tex[index], sky.length, sky.length, u, 1.0f - v);
		return sample_image(sky.faces[index], sky.length, sky.length, u, 1.0f - v, sky.use_bilinear != 0);
	}
	if (sky.use_sky)
	{
		float t = (dot(direction, make_float3(-0.41f, 0.41f, -0.82f)) + 1.0f) / 2.0f;
		float3 a = make_float3(0.15f, 0.3f, 0.5f);
		float3 b = make_float3(1.0f, 1.0f, 1.0f);
		return ((1.0f - t) * a + t * b) * 1.0f;
	}
	return make_float3(0.0f, 0.0f, 0.0f);
}

} // namespace ptbdev
