// Kernel-side data layout of the ptb200 wavefront integrator (device buffers, all explicit
// cudaMalloc — no managed memory, unlike the reference's cudaMallocManaged scene).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pt_device.cuh"

namespace ptb
{

// Material padded from the reference's 84 bytes to 6 x 16-byte loads.
//  a = diffuse.xyz, roughness      b = emission.xyz, refraction_index   c = specular.xyz, extinction
//  d = sigma_a.xyz, sigma_s'.x     e = sigma_s'.y, sigma_s'.z, transparent(bits), diffuse_tex(bits)
//  f = specular_tex(bits), 0, 0, 0
struct DeviceMaterial
{
	float4 a, b, c, d, e, f;
};

struct DeviceTexture
{
	const uint8_t* pixels;
	int width, height;
	cudaTextureObject_t tex;      // option texture_filter=hardware: the same RGBA8 image as a cudaArray texture (bilinear by the texture unit), else 0
};

struct DeviceScene
{
	const float4* bvh_nodes;      // layout 2: 4 x float4 per binary node; layout 8: 5 x float4 per compressed wide node (bvh.h)
	const float4* tri_isect;      // 3 x float4 per triangle, leaf order
	int bvh_layout;               // 2 or 8
	const float4* bvh8_nodes;     // hybrid mode: a second, compressed 8-wide tree over the same triangles (bounce rays), else nullptr
	const float4* tri_isect8;     // its triangles in ITS leaf order
	const float4* tri_shade;      // 4 x float4 per triangle, by global triangle id:
	                              //   n0.xyz n1.x | n1.yz n2.xy | n2.z uv0.xy uv1.x | uv1.y uv2.xy material(bits)
	const DeviceMaterial* materials; // mesh materials, then one per sphere
	const float4* spheres;        // center.xyz, radius
	const DeviceTexture* textures;
	int n_triangles;
	int n_spheres;
	int sphere_material_base;
	int n_textures;
	int root_ref;                 // node index (>= 0) or leaf ref (< 0)
	// bounce rays that leave a triangle start their search AT that triangle's leaf (kernels_entry.cuh: k_up_level; layout 2 only, else nullptr):
	const float4* up_records;     // per child slot (node * 2 + side) of the binary tree, 4 x float4: the sibling's box + reference, the parent's sibling's box + reference, the slot two levels up (kernels_entry.cuh: k_up_pair)
	const int* tri_slot;          // child slot holding the leaf of triangle (global id), -1 = unknown
	// next-event estimation (estimator "nee", off by default): emissive, non-transparent triangles
	const float* tris24;          // raw triangles by global id (v0 v1 v2 ...), 24 floats each
	const int* light_tri;         // global triangle id of light k
	const float* light_cdf;       // cumulative area of lights 0..k, normalised to 1
	int n_lights;
	float light_area;             // total emissive area
	ptbdev::SkyParams sky;
};

// Config values the kernels read (Core/configuration.h), passed by value in kernel params.
struct DeviceConfig
{
	int max_depth;
	float bias_length;
	float energy_threshold;
	float sss_threshold;
	int use_bilinear;
	int gamma_correction;
	int use_anti_alias;
	float air_n;
	float3 air_sigma_a;
	float3 air_sigma_s;
	// estimator options (0 = the reference's behaviour)
	int sampler;        // 1: pcg streams instead of hash-product + minstd (pt_device.cuh)
	int sss_mode;       // 1: per-channel subsurface scattering (kernels_shade.cuh: spectral MIS over sigma_s'.xyz instead of sigma_s'.x only)
};

// hit.w of a path that ended inside the closest-hit kernel (option inline_scatter): neither a triangle (>= 0), a miss (-1) nor a sphere (<= -2 small)
#define PTB_PRIM_DEAD ((int)0x80000001)

// ray_o.w (every estimator but "nee", which keeps a flag there): bits 0-7 = bounces the path is ahead of the loop depth (its lead; FUSED
// instantiations, option inline_scatter, else 0), bits 8-31 = 1 + the triangle the segment leaves / the current walk entered through
// (0 = none): the bounce-ray kernels start their searches at that triangle's leaf (k_extend_upwalk, k_extend_persistent_fused<.., UPWALK>).
// Needs MaxDepth <= 255 for FUSED; triangles beyond 2^23 - 2 read "none".
#define PTB_LEAD_OF(w) ((w) & 0xff)
#define PTB_FROM_BITS_OF(w) ((w) & ~0xff)
#define PTB_FROM_BITS(prim) (((prim) >= 0 && (prim) < (1 << 23) - 2 ? (prim) + 1 : 0) << 8)

// terminator of a tile's entry cut (kernels_entry.cuh) == PTB_DONE of kernels_extend.cuh
#define PTB_ENTRY_END ((int)0x80000000)

// SoA path state, indexed by path id = slot * pixel_count + pixel.
struct PathState
{
	float4* ray_o;      // origin.xyz, 1.0f when this bounce estimated direct light by NEE (the next hit must not add emission)
	float4* ray_d;      // direction.xyz, (unused)
	float4* throughput; // not-absorbed colour .xyz, medium material index (bits; -1 = air)
	float4* radiance;   // accumulated colour .xyz of this pass
	float4* hit;        // t, t1, t2, primitive (bits): >= 0 triangle, -(s+2) sphere s, -1 miss
	// estimator "nee" only (nullptr otherwise): one pending shadow ray per path
	float4* shadow_o;   // origin.xyz, t_max
	float4* shadow_d;   // direction.xyz
	float4* shadow_c;   // radiance to add when the segment is unoccluded
	int* shadow_queue;  // path ids with a pending shadow ray at the current depth
};

} // namespace ptb
