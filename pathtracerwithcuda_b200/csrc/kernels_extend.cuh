// ptb200 wavefront integrator, stage 2: closest hit.  One-ray-per-thread kernels over both tree layouts (test / comparison),
// the production persistent warp-voting kernels (binary tree: k_extend_persistent, compressed 8-wide tree: k_extend_persistent8)
// and the exhaustive-scan test hook.  Included by render.cu only.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include "kernels.cuh"

namespace ptb
{

using namespace ptbdev;

// ------------------------------------------------------------------------------------------
// k_extend — closest hit (path_tracer_kernel.cu:418-454 + intersect_triangle_mesh_bvh :85-161)
// One conservative traversal over a single tree for all meshes; the accepted hit is decided by
// the reference's own Moller-Trumbore / sphere arithmetic (pt_device.cuh).
// ------------------------------------------------------------------------------------------
#define PTB_STACK_SIZE 64
#define PTB_STACK_SIZE8 32
#define PTB_SLACK_LO 0.9999995f
#define PTB_SLACK_HI 1.0000005f

struct HitRecord
{
	float t, t1, t2;
	int prim;
};

template <bool COUNT>
__device__ __forceinline__ HitRecord closest_hit(const DeviceScene& sc, float3 o, float3 d, float t_bound, unsigned& n_nodes, unsigned& n_tris)
{
	HitRecord best;
	best.t = t_bound; best.t1 = CUDART_INF_F; best.t2 = CUDART_INF_F; best.prim = -1;

	// spheres first, in index order, strict '<' (path_tracer_kernel.cu:431-441)
	for (int s = 0; s < sc.n_spheres; s++)
	{
		float4 sp = __ldg(&sc.spheres[s]);
		float t;
		if (intersect_sphere(make_float3(sp.x, sp.y, sp.z), sp.w, o, d, t) && t < best.t && t > 0.0f)
		{
			best.t = t;
			best.prim = -(s + 2);
		}
	}
	if (sc.n_triangles == 0) return best;

	const float3 idir = make_float3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
	int stack[PTB_STACK_SIZE];
	int sp = 0;
	int node = sc.root_ref;
	int best_tri = 0x7fffffff;

	while (true)
	{
		while (node >= 0)
		{
			if (COUNT) n_nodes++;
			const float4* np = sc.bvh_nodes + (size_t)node * 4;
			float4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
			float c0x0 = (n0.x - o.x) * idir.x, c0x1 = (n0.y - o.x) * idir.x;
			float c0y0 = (n0.z - o.y) * idir.y, c0y1 = (n0.w - o.y) * idir.y;
			float c0z0 = (n2.x - o.z) * idir.z, c0z1 = (n2.y - o.z) * idir.z;
			float c1x0 = (n1.x - o.x) * idir.x, c1x1 = (n1.y - o.x) * idir.x;
			float c1y0 = (n1.z - o.y) * idir.y, c1y1 = (n1.w - o.y) * idir.y;
			float c1z0 = (n2.z - o.z) * idir.z, c1z1 = (n2.w - o.z) * idir.z;
			float tmin0 = fmaxf(fmaxf(fminf(c0x0, c0x1), fminf(c0y0, c0y1)), fmaxf(fminf(c0z0, c0z1), 0.0f));
			float tmax0 = fminf(fminf(fmaxf(c0x0, c0x1), fmaxf(c0y0, c0y1)), fminf(fmaxf(c0z0, c0z1), best.t));
			float tmin1 = fmaxf(fmaxf(fminf(c1x0, c1x1), fminf(c1y0, c1y1)), fmaxf(fminf(c1z0, c1z1), 0.0f));
			float tmax1 = fminf(fminf(fmaxf(c1x0, c1x1), fmaxf(c1y0, c1y1)), fminf(fmaxf(c1z0, c1z1), best.t));
			bool h0 = tmin0 * PTB_SLACK_LO <= tmax0 * PTB_SLACK_HI;
			bool h1 = tmin1 * PTB_SLACK_LO <= tmax1 * PTB_SLACK_HI;
			int child0 = __float_as_int(n3.x), child1 = __float_as_int(n3.y);
			if (h0 && h1)
			{
				bool swap = tmin1 < tmin0;
				int near_c = swap ? child1 : child0;
				int far_c = swap ? child0 : child1;
				if (sp < PTB_STACK_SIZE) stack[sp++] = far_c;
				node = near_c;
			}
			else if (h0) node = child0;
			else if (h1) node = child1;
			else
			{
				if (sp == 0) return best;
				node = stack[--sp];
			}
		}
		// leaf: node = ~((first << 3) | (count - 1))
		{
			int ref = ~node;
			int first = ref >> 3;
			int count = (ref & 7) + 1;
			for (int k = 0; k < count; k++)
			{
				if (COUNT) n_tris++;
				const float4* tp = sc.tri_isect + (size_t)(first + k) * 3;
				float4 a = __ldg(tp + 0), b = __ldg(tp + 1), c = __ldg(tp + 2);
				float t, t1, t2;
				if (intersect_triangle(make_float3(a.x, a.y, a.z), make_float3(b.x, b.y, b.z), make_float3(c.x, c.y, c.z), o, d, t, t1, t2) && t > 0.0f)
				{
					int id = __float_as_int(a.w);
					// strict '<' like the reference; exact-t ties between triangles go to the lower
					// global index (the reference's tie winner depends on its tree layout)
					if (t < best.t || (t == best.t && best.prim >= 0 && id < best_tri))
					{
						best.t = t; best.t1 = t1; best.t2 = t2; best.prim = id; best_tri = id;
					}
				}
			}
			if (sp == 0) return best;
			node = stack[--sp];
		}
	}
}

// ---- compressed 8-wide traversal (layout in bvh.h; after Ylitie, Karras & Laine 2017) ----
__device__ __forceinline__ unsigned sign_extend_s8x4(unsigned x)
{
	unsigned r;
	asm("prmt.b32 %0, %1, 0x0, 0x0000BA98;" : "=r"(r) : "r"(x));
	return r;
}

__device__ __forceinline__ unsigned extract_byte(unsigned x, unsigned i) { return (x >> (i * 8)) & 0xffu; }

// (float)byte j of x.  The int -> float conversion issues on the quarter-rate XU pipe, the busiest pipe of the wide-tree kernels under ncu
// (63 %, 48 conversions per node); replacing it by PRMT into the mantissa of 2^23 + FADD (build flag PTB_BYTE_NO_I2F, exact) was measured
// SLOWER — c2 d2 1.47 -> 1.56 ms, c4 807 -> 761 Msamples/s (profiles/r02_experiments.md): the kernel is bound by issue slots, not by that pipe.
#ifdef PTB_BYTE_NO_I2F
__device__ __forceinline__ float byte_to_float(unsigned x, int j) { return __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7650u + (unsigned)j)) - 8388608.0f; }
#else
__device__ __forceinline__ float byte_to_float(unsigned x, int j) { return (float)extract_byte(x, j); }
#endif

template <bool COUNT>
__device__ __forceinline__ HitRecord closest_hit_bvh8(const DeviceScene& sc, float3 o, float3 d, float t_bound, unsigned& n_nodes, unsigned& n_tris)
{
	HitRecord best;
	best.t = t_bound; best.t1 = CUDART_INF_F; best.t2 = CUDART_INF_F; best.prim = -1;
	for (int s = 0; s < sc.n_spheres; s++)
	{
		float4 sp = __ldg(&sc.spheres[s]);
		float t;
		if (intersect_sphere(make_float3(sp.x, sp.y, sp.z), sp.w, o, d, t) && t < best.t && t > 0.0f)
		{
			best.t = t;
			best.prim = -(s + 2);
		}
	}
	if (sc.n_triangles == 0) return best;

	// box culling only: keep the reciprocal finite so 0 * inf never appears (the deciding
	// triangle test below still sees the exact direction)
	const float tiny = 1e-30f;
	const float3 ds = make_float3(fabsf(d.x) < tiny ? copysignf(tiny, d.x) : d.x, fabsf(d.y) < tiny ? copysignf(tiny, d.y) : d.y,
		fabsf(d.z) < tiny ? copysignf(tiny, d.z) : d.z);
	const float3 idir = make_float3(1.0f / ds.x, 1.0f / ds.y, 1.0f / ds.z);
	const unsigned oct_inv4 = (d.x < 0.0f ? 0u : 0x04040404u) | (d.y < 0.0f ? 0u : 0x02020202u) | (d.z < 0.0f ? 0u : 0x01010101u);

	uint2 stack[PTB_STACK_SIZE8];
	int sp = 0;
	uint2 current = make_uint2(0u, 0x80000000u);
	int best_tri = 0x7fffffff;

	while (true)
	{
		uint2 tri_group;
		if (current.y & 0xff000000u)
		{
			const unsigned hits_imask = current.y;
			const unsigned child_index_offset = 31u - __clz(hits_imask);
			const unsigned child_index_base = current.x;
			current.y &= ~(1u << child_index_offset);
			if (current.y & 0xff000000u) { if (sp < PTB_STACK_SIZE8) stack[sp++] = current; }
			const unsigned slot_index = (child_index_offset - 24u) ^ (oct_inv4 & 0xffu);
			const unsigned relative_index = __popc(hits_imask & ~(0xffffffffu << slot_index));
			const unsigned node_index = child_index_base + relative_index;
			if (COUNT) n_nodes++;

			const float4* np = sc.bvh_nodes + (size_t)node_index * 5;
			const float4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
			const unsigned e_imask = __float_as_uint(n0.w);
			const float3 adir = make_float3(__uint_as_float(extract_byte(e_imask, 0) << 23) * idir.x, __uint_as_float(extract_byte(e_imask, 1) << 23) * idir.y,
				__uint_as_float(extract_byte(e_imask, 2) << 23) * idir.z);
			const float3 aorg = make_float3((n0.x - o.x) * idir.x, (n0.y - o.y) * idir.y, (n0.z - o.z) * idir.z);

			unsigned hit_mask = 0;
#pragma unroll
			for (int half = 0; half < 2; half++)
			{
				const unsigned meta4 = __float_as_uint(half == 0 ? n1.z : n1.w);
				const unsigned is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
				const unsigned inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
				const unsigned bit_index4 = (meta4 ^ (oct_inv4 & inner_mask4)) & 0x1f1f1f1fu;
				const unsigned child_bits4 = (meta4 >> 5) & 0x07070707u;
				const unsigned qlox = __float_as_uint(half == 0 ? n2.x : n2.y), qhix = __float_as_uint(half == 0 ? n2.z : n2.w);
				const unsigned qloy = __float_as_uint(half == 0 ? n3.x : n3.y), qhiy = __float_as_uint(half == 0 ? n3.z : n3.w);
				const unsigned qloz = __float_as_uint(half == 0 ? n4.x : n4.y), qhiz = __float_as_uint(half == 0 ? n4.z : n4.w);
				const unsigned x_min = d.x < 0.0f ? qhix : qlox, x_max = d.x < 0.0f ? qlox : qhix;
				const unsigned y_min = d.y < 0.0f ? qhiy : qloy, y_max = d.y < 0.0f ? qloy : qhiy;
				const unsigned z_min = d.z < 0.0f ? qhiz : qloz, z_max = d.z < 0.0f ? qloz : qhiz;
#pragma unroll
				for (int j = 0; j < 4; j++)
				{
					const float tx0 = fmaf(byte_to_float(x_min, j), adir.x, aorg.x), tx1 = fmaf(byte_to_float(x_max, j), adir.x, aorg.x);
					const float ty0 = fmaf(byte_to_float(y_min, j), adir.y, aorg.y), ty1 = fmaf(byte_to_float(y_max, j), adir.y, aorg.y);
					const float tz0 = fmaf(byte_to_float(z_min, j), adir.z, aorg.z), tz1 = fmaf(byte_to_float(z_max, j), adir.z, aorg.z);
					const float tmin = fmaxf(fmaxf(tx0, ty0), fmaxf(tz0, 0.0f));
					const float tmax = fminf(fminf(tx1, ty1), fminf(tz1, best.t));
					if (tmin * PTB_SLACK_LO <= tmax * PTB_SLACK_HI)
						hit_mask |= extract_byte(child_bits4, j) << extract_byte(bit_index4, j);
				}
			}
			current.x = __float_as_uint(n1.x);
			tri_group.x = __float_as_uint(n1.y);
			current.y = (hit_mask & 0xff000000u) | (e_imask >> 24);
			tri_group.y = hit_mask & 0x00ffffffu;
		}
		else
		{
			tri_group = current;
			current = make_uint2(0u, 0u);
		}

		while (tri_group.y)
		{
			const unsigned k = 31u - __clz(tri_group.y);
			tri_group.y &= ~(1u << k);
			if (COUNT) n_tris++;
			const float4* tp = sc.tri_isect + (size_t)(tri_group.x + k) * 3;
			const float4 a = __ldg(tp + 0), b = __ldg(tp + 1), c = __ldg(tp + 2);
			float t, t1, t2;
			if (intersect_triangle(make_float3(a.x, a.y, a.z), make_float3(b.x, b.y, b.z), make_float3(c.x, c.y, c.z), o, d, t, t1, t2) && t > 0.0f)
			{
				const int id = __float_as_int(a.w);
				if (t < best.t || (t == best.t && best.prim >= 0 && id < best_tri))
				{
					best.t = t; best.t1 = t1; best.t2 = t2; best.prim = id; best_tri = id;
				}
			}
		}

		if ((current.y & 0xff000000u) == 0u)
		{
			if (sp == 0) return best;
			current = stack[--sp];
		}
	}
}

template <bool COUNT, bool WIDE>
__global__ void __launch_bounds__(128) k_extend(DeviceScene sc, PathState st, const int* __restrict__ queue, const int* __restrict__ count_ptr,
	unsigned long long* __restrict__ counters)
{
	const int count = *count_ptr;
	unsigned n_nodes = 0, n_tris = 0;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
	{
		int id = queue[i];
		float4 o4 = st.ray_o[id], d4 = st.ray_d[id];
		HitRecord h = WIDE ? closest_hit_bvh8<COUNT>(sc, make_float3(o4.x, o4.y, o4.z), make_float3(d4.x, d4.y, d4.z), d4.w, n_nodes, n_tris)
		                   : closest_hit<COUNT>(sc, make_float3(o4.x, o4.y, o4.z), make_float3(d4.x, d4.y, d4.z), d4.w, n_nodes, n_tris);
		st.hit[id] = make_float4(h.prim == -1 ? CUDART_INF_F : h.t, h.t1, h.t2, __int_as_float(h.prim));
	}
	if (COUNT)
	{
		for (int off = 16; off > 0; off >>= 1)
		{
			n_nodes += __shfl_down_sync(0xffffffffu, n_nodes, off);
			n_tris += __shfl_down_sync(0xffffffffu, n_tris, off);
		}
		if ((threadIdx.x & 31) == 0)
		{
			atomicAdd(&counters[0], (unsigned long long)n_nodes);
			atomicAdd(&counters[1], (unsigned long long)n_tris);
		}
	}
}

// ------------------------------------------------------------------------------------------
// k_extend_persistent — the production closest-hit kernel.
// ncu on the one-ray-per-thread kernels above showed them ISSUE-bound (65-70% issue slots busy)
// with only 8-18 of 32 lanes active per instruction: rays of one warp need very different numbers
// of node visits and the finished lanes idle.  Here warps are persistent: a lane whose ray
// terminates fetches the next queue entry (one atomicAdd per warp per refill) as soon as fewer than
// PTB_REFILL_THRESHOLD lanes are still traversing, so the 32 lanes stay populated.
// Binary tree, both child boxes per 64-byte node; slab distances as one FMA per plane
// (plane * (1/d) - o/d) with an absolute + relative safety margin so culling stays conservative.
// ------------------------------------------------------------------------------------------
#define PTB_DONE ((int)0x80000000)

// one 64-byte binary node: two 256-bit loads (sm_100 LDG.E.ENL2.256) instead of 3 x 128-bit + 1 x 64-bit.
// (Routing the node fetches through the TEX data pipe was measured 2-7 % slower: profiles/r01_experiments.md.)
__device__ __forceinline__ void load_node(const float4* np, float4& n0, float4& n1, float4& n2, float2& n3)
{
#ifndef PTB_NODE_LDG128
	float z, w;
	asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"(n0.x), "=f"(n0.y), "=f"(n0.z), "=f"(n0.w), "=f"(n1.x), "=f"(n1.y), "=f"(n1.z), "=f"(n1.w) : "l"(np));
	asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"(n2.x), "=f"(n2.y), "=f"(n2.z), "=f"(n2.w), "=f"(n3.x), "=f"(n3.y), "=f"(z), "=f"(w) : "l"(np + 2));
#else
	n0 = __ldg(np + 0); n1 = __ldg(np + 1); n2 = __ldg(np + 2);
	n3 = __ldg(reinterpret_cast<const float2*>(np + 3));
#endif
}
// resident blocks per SM the binary-tree kernels are compiled for.  8 (<= 64 registers: 54-56 used, 9 blocks fit) was the optimum while the
// searches started at the root and the kernels were issue-bound; with entry cuts and leaf starts they wait on dependent fetches, and 10
// blocks (48 registers, 8-12 bytes of spills) read c2 4443 -> 4529, c5 3647 -> 3739 Msamples/s; 12 blocks (40 registers, 60-70 bytes of
// spills) 4325 / 3663 (profiles/r02_experiments.md).  The wide-tree kernels stay at 8 (9 / 10: c4 830 -> 818 / 815).
#ifndef PTB_PERSISTENT_MIN_BLOCKS
#define PTB_PERSISTENT_MIN_BLOCKS 10
#endif
#ifndef PTB_PERSISTENT_MIN_BLOCKS8
#define PTB_PERSISTENT_MIN_BLOCKS8 8
#endif

// Leaf phase: 1 = every vote tests ONE triangle per waiting lane (a lane with more triangles stays in the leaf state), so the lanes
// of a leaf phase never idle behind the longest leaf; 0 = the whole leaf in one phase (round 1).
#ifndef PTB_LEAF_SINGLE
#define PTB_LEAF_SINGLE 1
#endif
#ifndef PTB_LEAF_FALLTHROUGH
#define PTB_LEAF_FALLTHROUGH 0
#endif

// FUSED (option inline_scatter, default on in parity mode): in a scattering medium most bounces are scatter events — the free flight ends
// before any surface (k_shade: d < t_hit), the path gets an isotropic direction, its throughput is attenuated, nothing else happens
// (path_tracer_kernel.cu:460-486).  The bounded search already knows it (no hit below the free-flight bound), so the lane performs the
// event here with the reference's arithmetic and random stream of that depth and goes on tracing the SAME path, instead of a round trip
// through hit record, queue, k_shade and the next launch per event.  The path runs ahead of the wavefront's loop depth by `lead`
// bounces (kept in ray_o.w for k_shade<.., FUSED>); scatter events are executed as a voted phase like node and triangle steps.
struct FusedArgs
{
	DeviceConfig cfg;
	int loop_depth, pixel_count, first_pass, pass_stride, scatter_min;
	// ray segments per ACTUAL depth (the call totals ptb_get_depth_profile reports): with paths ahead of the loop depth the queue sizes no longer
	// say at which depth a search ran, so this kernel tallies its own searches (block histogram in shared memory, flushed once per block)
	unsigned long long* depth_segments;
	int n_depth_slots;
};
#define PTB_FUSED_HIST 66


// One scatter event of the lane's path at bounce `depth` (= loop depth + lead): k_shade's medium branch (kernels_shade.cuh) with the
// reference's arithmetic and random stream.  Returns false when the path ends (energy cut or bounce limit; its hit record then carries
// PTB_PRIM_DEAD); otherwise o / d / bound describe the next segment, the throughput is in s_thr, and the caller restarts its search.
__device__ __forceinline__ bool fused_scatter_event(const DeviceScene& sc, const PathState& st, const FusedArgs& fa, int id, int depth,
	bool& thr_cached, float4* s_thr_slot, int* s_hist, float3& o, float3& d, float& bound)
{
	const DeviceConfig& cfg = fa.cfg;
	const int slot = id / fa.pixel_count;
	const int pixel_index = id - slot * fa.pixel_count;
	const int seed = fa.first_pass + slot * fa.pass_stride;
	const float4 t4 = thr_cached ? *s_thr_slot : (depth == 0 ? make_float4(1.0f, 1.0f, 1.0f, __int_as_float(-1)) : st.throughput[id]);
	float3 not_absorbed = make_float3(t4.x, t4.y, t4.z);
	const int medium_index = __float_as_int(t4.w);
	float3 sigma_a = cfg.air_sigma_a, sigma_s = cfg.air_sigma_s;
	if (medium_index >= 0)
	{
		const float4 md = __ldg(&sc.materials[medium_index].d), me = __ldg(&sc.materials[medium_index].e);
		sigma_a = make_float3(md.x, md.y, md.z);
		sigma_s = make_float3(md.w, me.x, me.y);
	}
	// streams of this bounce and the next: (hash(seed) * hash(pixel)) * hash(depth) — the product the reference seeds with, the
	// first two factors shared (integer multiplication is associative mod 2^32)
	const int hsp = hash_ref(seed) * hash_ref(pixel_index);
	Rng rng;
	rng.seed((uint32_t)(hsp * hash_ref(depth)), 0.0f, 1.0f);
	const float rand = rng.next();
	const float scattering_distance = -__logf(rand) / sigma_s.x;
	const float rand1 = rng.next();
	const float rand2 = rng.next();
	const float3 next_o = o + d * scattering_distance;
	const float3 next_d = sample_on_sphere(rand1, rand2);
	not_absorbed = not_absorbed * absorption_through_medium(sigma_a, scattering_distance);
	if (length(not_absorbed) <= cfg.energy_threshold || depth + 1 >= cfg.max_depth)
	{
		// the path ends here: energy cut (:480-483) or the bounce limit of the reference's depth loop
		__stcs(&st.hit[id], make_float4(CUDART_INF_F, 0.0f, 0.0f, __int_as_float(PTB_PRIM_DEAD)));
		return false;
	}
	*s_thr_slot = make_float4(not_absorbed.x, not_absorbed.y, not_absorbed.z, t4.w);
	thr_cached = true;
	atomicAdd(&s_hist[min(depth + 1, PTB_FUSED_HIST - 1)], 1);     // the search of the next bounce starts here
	o = next_o; d = next_d;
	bound = next_bounce_bound_hashed(cfg, sigma_a, sigma_s, hsp, depth + 1);
	return true;
}

// TREELET (k_extend_treelet below): the first `n_top` nodes of the array — the top levels of the tree, which the device builder
// numbers level by level — are copied into shared memory by the block and fetched from there: every ray walks through them, and a
// scattered 64-byte gather costs the L1 data stage ~1.45 cycles per lane (tools/roofs: 12.8 TB/s) where shared memory delivers it at
// bank rate.  Blocks are then as large as the launch allows (one 1024-thread block per SM shares one copy).
// ENTRY (k_extend_entry below; camera rays only): a ray starts at its 8x4 pixel tile's entry cut (kernels_entry.cuh: the sub-trees its
// shaft touches, ascending in a lower bound on the hit distance) instead of at the root, and stops walking the list at the first
// sub-tree that begins beyond its current hit.
struct EntryArgs
{
	const int2* cuts;
	int pixel_count, width, tiles_x, tile_w_shift, tile_h_shift, stride_shift;
};

// The leaf start of a bounce ray (UPWALK): `slot` = child slot (node * 2 + side) of the leaf the search starts at.  Sets `node` to that
// leaf and leaves on the stack the siblings of the leaf's ancestors the ray's box test accepts, deepest (= nearest) on top.  Every
// sub-tree of the scene is the start leaf or one of those siblings, so the search is exhaustive whatever leaf it starts at; a leaf near
// the ray's origin makes it cheap (32 bytes and one box test per level instead of a 64-byte node and two; records hold two levels).
template <bool COUNT>
__device__ __forceinline__ void upwalk_start(const DeviceScene& sc, int slot, float3 idir, float3 noidir, float margin2, float t_max,
	int* stack, int& sp, int& node, unsigned& n_nodes, unsigned& ray_nodes)
{
	const float2 ch = __ldg(reinterpret_cast<const float2*>(sc.bvh_nodes + (size_t)(slot >> 1) * 4 + 3));
	node = __float_as_int((slot & 1) ? ch.y : ch.x);
	int n = 0;
	do
	{
		// two levels per 64-byte record (kernels_entry.cuh: k_up_pair): the sibling of `slot`, then the sibling of its parent
		const float4* rp = sc.up_records + (size_t)slot * 4;
		float4 a_lo, a_hi, b_lo;
		float2 b_hi_xy; float b_hi_z;
		{
			float pad;
			asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
				: "=f"(a_lo.x), "=f"(a_lo.y), "=f"(a_lo.z), "=f"(a_lo.w), "=f"(a_hi.x), "=f"(a_hi.y), "=f"(a_hi.z), "=f"(a_hi.w) : "l"(rp));
			asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
				: "=f"(b_lo.x), "=f"(b_lo.y), "=f"(b_lo.z), "=f"(b_lo.w), "=f"(b_hi_xy.x), "=f"(b_hi_xy.y), "=f"(b_hi_z), "=f"(pad) : "l"(rp + 2));
		}
		if (COUNT) { n_nodes++; ray_nodes++; }
		{
			const float x0 = fmaf(a_lo.x, idir.x, noidir.x), x1 = fmaf(a_hi.x, idir.x, noidir.x);
			const float y0 = fmaf(a_lo.y, idir.y, noidir.y), y1 = fmaf(a_hi.y, idir.y, noidir.y);
			const float z0 = fmaf(a_lo.z, idir.z, noidir.z), z1 = fmaf(a_hi.z, idir.z, noidir.z);
			const float tmin = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
			const float tmax = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), t_max));
			if (fmaf(tmin, PTB_SLACK_LO / PTB_SLACK_HI, -margin2) <= tmax) { if (n < PTB_STACK_SIZE) stack[n] = __float_as_int(a_lo.w); n++; }
		}
		const int b_ref = __float_as_int(a_hi.w);
		if (b_ref != PTB_DONE)
		{
			if (COUNT) { n_nodes++; ray_nodes++; }
			const float x0 = fmaf(b_lo.x, idir.x, noidir.x), x1 = fmaf(b_hi_xy.x, idir.x, noidir.x);
			const float y0 = fmaf(b_lo.y, idir.y, noidir.y), y1 = fmaf(b_hi_xy.y, idir.y, noidir.y);
			const float z0 = fmaf(b_lo.z, idir.z, noidir.z), z1 = fmaf(b_hi_z, idir.z, noidir.z);
			const float tmin = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
			const float tmax = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), t_max));
			if (fmaf(tmin, PTB_SLACK_LO / PTB_SLACK_HI, -margin2) <= tmax) { if (n < PTB_STACK_SIZE) stack[n] = b_ref; n++; }
		}
		slot = __float_as_int(b_lo.w);
	} while (slot >= 0);
	n = min(n, PTB_STACK_SIZE);
	for (int a = 0, b = n - 1; a < b; a++, b--) { const int t = stack[a]; stack[a] = stack[b]; stack[b] = t; }
	sp = n;
}

// UPWALK (k_extend_upwalk below; bounce rays): a ray that leaves a triangle (k_shade left its id in ray_o.w) starts at that triangle's
// leaf and collects the siblings of the leaf's ancestors it hits by walking UP the tree (kernels_entry.cuh: k_up_level) instead of
// descending from the root.
template <bool COUNT, int REPS, bool TREELET, bool STAGED, bool FUSED = false, bool ENTRY = false, bool UPWALK = false>
__device__ __forceinline__ void extend_persistent_body(const DeviceScene& sc, const PathState& st, const int* __restrict__ queue, const int* __restrict__ count_ptr,
	int* __restrict__ work_counter, unsigned long long* __restrict__ counters, int refill_min, int leaf_min, int node_reps, const float4* s_top, int n_top, float4* s_stage, const FusedArgs* fa = nullptr, float4* s_thr = nullptr, int* s_hist = nullptr,
	const EntryArgs* ea = nullptr, int* s_from = nullptr)
{
	const int count = *count_ptr;
	const unsigned lane = threadIdx.x & 31u;
	const unsigned lane_lt = (1u << lane) - 1u;
	const unsigned FULL = 0xffffffffu;
	unsigned n_nodes = 0, n_tris = 0;

	int id = -1;                 // path id this lane is tracing; -1 = idle
	unsigned ray_nodes = 0;      // COUNT only: node visits of the current ray
	bool exhausted = false;      // warp-uniform: the queue has been handed out completely
	int staged = 0;              // STAGED: set-up rays waiting in this warp's shared-memory buffer (warp-uniform)
	float3 o = make_float3(0, 0, 0), d = o, idir = o, noidir = o;
	float margin2 = 0.0f;
	HitRecord best;
	best.t = CUDART_INF_F; best.t1 = 0.0f; best.t2 = 0.0f; best.prim = -1;
	int best_tri = 0x7fffffff;
	// per-thread traversal stack (local memory).  A shared-memory stack ([entry][thread], conflict-free at any depth mix) was measured
	// 3-4 % SLOWER at every size tried (12 / 16 / 24 / 32 entries, profiles/r02_experiments.md) and is not kept.  The push is bounded:
	// the builders keep every tree shallower than PTB_STACK_SIZE (render.cu rejects deeper ones), the guard costs nothing.
	int stack[PTB_STACK_SIZE];
#define PTB_PUSH(v) do { if (sp < PTB_STACK_SIZE) stack[sp] = (v); sp++; } while (0)
#define PTB_POP(dst) do { --sp; dst = stack[min(sp, PTB_STACK_SIZE - 1)]; } while (0)
	int sp = 0;
	int node = PTB_DONE;         // >= 0 inner node, PTB_DONE = nothing left, other negative = leaf reference
	// FUSED (see FusedArgs above): scatter events of the medium performed here, the path runs `lead` bounces ahead of the loop depth
	int lead = 0;
	bool at_scatter = false, thr_cached = false;
	int cut = 0;                 // ENTRY: index of the next entry of the lane's tile list
	int start_slot = -1;         // UPWALK: child slot of the leaf the lane's searches start at (FUSED: kept across the path's scatter events)

	// Every iteration starts with full-mask votes, so all 32 lanes are converged when a phase
	// begins; a phase is executed by the lanes in that state, the others are predicated off.
	while (true)
	{
		if (ENTRY && id >= 0 && node == PTB_DONE)
		{
			// the sub-tree is finished (or the ray is new): next entry of the tile's list, unless it begins beyond the current hit
			const int2 e = __ldg(ea->cuts + cut);
			if (e.x != PTB_DONE && __int_as_float(e.y) <= best.t) { node = e.x; cut++; }
		}
		// retire finished rays (no vote needed: a plain predicated store)
		if (FUSED && id >= 0 && node == PTB_DONE && !at_scatter && best.prim == -1 && best.t < CUDART_INF_F) at_scatter = true;
		if (id >= 0 && node == PTB_DONE && !at_scatter)
		{
			if (FUSED && thr_cached)
			{
				st.ray_o[id] = make_float4(o.x, o.y, o.z, __int_as_float(lead | s_from[threadIdx.x]));
				st.ray_d[id] = make_float4(d.x, d.y, d.z, 0.0f);
				st.throughput[id] = s_thr[threadIdx.x];
			}
#ifndef PTB_NO_STREAMING_HINTS
			__stcs(&st.hit[id], make_float4(best.prim == -1 ? CUDART_INF_F : best.t, best.t1, best.t2, __int_as_float(best.prim)));
#else
			st.hit[id] = make_float4(best.prim == -1 ? CUDART_INF_F : best.t, best.t1, best.t2, __int_as_float(best.prim));
#endif
			id = -1;
			if (COUNT)
			{
				// histogram of node visits per ray: counters[4 + floor(log2(n + 1))], max in counters[2]
				atomicMax(&counters[2], (unsigned long long)ray_nodes);
				atomicAdd(&counters[4 + min(27, 31 - __clz(ray_nodes + 1u))], 1ull);
				ray_nodes = 0;
			}
		}
		const bool has_ray = id >= 0;
		const unsigned m_idle = __ballot_sync(FULL, !has_ray);
		const unsigned m_scat = FUSED ? __ballot_sync(FULL, has_ray && at_scatter) : 0u;
		const unsigned m_node = __ballot_sync(FULL, has_ray && node >= 0) | m_scat;   // "work in flight" for the refill / exit tests
		const unsigned m_leaf = __ballot_sync(FULL, has_ray && node < 0 && !at_scatter);

		if (STAGED)
		{
			// ---- STAGED refill: rays are fetched and set up 32 at a time by the WHOLE warp (every lane busy) into a shared-memory buffer;
			// an idle lane then takes a ready ray for a dozen instructions, so idle lanes no longer wait until 20 of them justify the
			// ~150-instruction divergent set-up — fewer lanes idle per issued instruction.
			if (m_idle != 0u && (staged > 0 || !exhausted) && (__popc(m_idle) >= refill_min || (m_node | m_leaf) == 0u))
			{
				if (staged == 0)
				{
					int base = 0;
					if (lane == 0) base = atomicAdd(work_counter, 32);
					base = __shfl_sync(FULL, base, 0);
					if (base + 32 >= count) exhausted = true;
					staged = max(0, min(32, count - base));
					const int i = base + (int)lane;
					if (i < count)
					{
						const int rid = __ldcs(&queue[i]);
						const float4 o4 = __ldcs(&st.ray_o[rid]), d4 = __ldcs(&st.ray_d[rid]);
						const float3 ro = make_float3(o4.x, o4.y, o4.z), rd = make_float3(d4.x, d4.y, d4.z);
						float bt = d4.w;
						int bprim = -1;
						for (int sidx = 0; sidx < sc.n_spheres; sidx++)
						{
							const float4 sph = __ldg(&sc.spheres[sidx]);
							float t;
							if (intersect_sphere(make_float3(sph.x, sph.y, sph.z), sph.w, ro, rd, t) && t < bt && t > 0.0f) { bt = t; bprim = -(sidx + 2); }
						}
						const float tiny = 1e-30f;
						const float3 ds = make_float3(fabsf(rd.x) < tiny ? copysignf(tiny, rd.x) : rd.x, fabsf(rd.y) < tiny ? copysignf(tiny, rd.y) : rd.y,
							fabsf(rd.z) < tiny ? copysignf(tiny, rd.z) : rd.z);
						s_stage[lane] = make_float4(ro.x, ro.y, ro.z, bt);
						s_stage[32 + lane] = make_float4(rd.x, rd.y, rd.z, __int_as_float(rid));
						s_stage[64 + lane] = make_float4(1.0f / ds.x, 1.0f / ds.y, 1.0f / ds.z, __int_as_float(bprim));
					}
					__syncwarp();
				}
				const int rank = __popc(m_idle & lane_lt);
				if (!has_ray && rank < staged)
				{
					const int slot = staged - 1 - rank;
					const float4 a = s_stage[slot], b = s_stage[32 + slot], c = s_stage[64 + slot];
					o = make_float3(a.x, a.y, a.z);
					d = make_float3(b.x, b.y, b.z);
					id = __float_as_int(b.w);
					idir = make_float3(c.x, c.y, c.z);
					best.t = a.w; best.t1 = CUDART_INF_F; best.t2 = CUDART_INF_F; best.prim = __float_as_int(c.w);
					best_tri = 0x7fffffff;
					noidir = make_float3(-o.x * idir.x, -o.y * idir.y, -o.z * idir.z);
					margin2 = 4.8e-7f * fmaxf(fmaxf(fabsf(noidir.x), fabsf(noidir.y)), fabsf(noidir.z));
					sp = 0;
					node = sc.n_triangles > 0 ? sc.root_ref : PTB_DONE;
				}
				staged -= min(staged, __popc(m_idle));
				__syncwarp();
				continue;
			}
		}
		else
		if (m_idle != 0u && !exhausted && (__popc(m_idle) >= refill_min || (m_node | m_leaf) == 0u))
		{
			// ---- refill idle lanes from the queue: one atomic per warp.  (Fetching a chunk of 32-1024 entries per atomic and handing it out
			// over the warp's next refills was measured SLOWER at every size, c2 d0 3.78 -> 4.2-5.0 ms: the same-address atomics overlap
			// with the other warps' work, the extra warp state cost a resident block per SM; profiles/r02_experiments.md.)
			const int n = __popc(m_idle);
			int base = 0;
			if (lane == 0) base = atomicAdd(work_counter, n);
			base = __shfl_sync(FULL, base, 0);
			if (base + n >= count) exhausted = true;
			if (!has_ray)
			{
				const int i = base + __popc(m_idle & lane_lt);
				if (i < count)
				{
#ifndef PTB_NO_STREAMING_HINTS   // ray records are read once: keep them from displacing tree nodes in L1 (+0.5 %)
					id = __ldcs(&queue[i]);
					const float4 o4 = __ldcs(&st.ray_o[id]), d4 = __ldcs(&st.ray_d[id]);
#else
					id = queue[i];
					const float4 o4 = st.ray_o[id], d4 = st.ray_d[id];
#endif
					o = make_float3(o4.x, o4.y, o4.z);
					d = make_float3(d4.x, d4.y, d4.z);
					best.t = d4.w; best.t1 = CUDART_INF_F; best.t2 = CUDART_INF_F; best.prim = -1;   // d4.w: free-flight bound (next_bounce_bound)
					best_tri = 0x7fffffff;
					if (FUSED) { lead = PTB_LEAD_OF(__float_as_int(o4.w)); s_from[threadIdx.x] = PTB_FROM_BITS_OF(__float_as_int(o4.w)); at_scatter = false; thr_cached = false; atomicAdd(&s_hist[min(fa->loop_depth + lead, PTB_FUSED_HIST - 1)], 1); }
					for (int s = 0; s < sc.n_spheres; s++)
					{
						const float4 sph = __ldg(&sc.spheres[s]);
						float t;
						if (intersect_sphere(make_float3(sph.x, sph.y, sph.z), sph.w, o, d, t) && t < best.t && t > 0.0f)
						{
							best.t = t;
							best.prim = -(s + 2);
						}
					}
					// box culling only: finite reciprocal so 0 * inf never appears
					const float tiny = 1e-30f;
					const float3 ds = make_float3(fabsf(d.x) < tiny ? copysignf(tiny, d.x) : d.x, fabsf(d.y) < tiny ? copysignf(tiny, d.y) : d.y,
						fabsf(d.z) < tiny ? copysignf(tiny, d.z) : d.z);
					idir = make_float3(1.0f / ds.x, 1.0f / ds.y, 1.0f / ds.z);
					noidir = make_float3(-o.x * idir.x, -o.y * idir.y, -o.z * idir.z);
					// |error| of fma(plane, idir, -o*idir) <= 2^-23 * (|o*idir| + |t|): absolute part here, relative part in the slack factors
					margin2 = 4.8e-7f * fmaxf(fmaxf(fabsf(noidir.x), fabsf(noidir.y)), fabsf(noidir.z));
					sp = 0;
					node = sc.n_triangles > 0 ? sc.root_ref : PTB_DONE;
					if (UPWALK)
					{
						// the triangle this segment leaves (k_shade: ray_o.w, kernels.cuh: PTB_FROM_BITS)
						const int from = (PTB_FROM_BITS_OF(__float_as_int(o4.w)) >> 8) - 1;
						start_slot = (from >= 0 && from < sc.n_triangles) ? __ldg(&sc.tri_slot[from]) : -1;
						if (start_slot >= 0) upwalk_start<COUNT>(sc, start_slot, idir, noidir, margin2, best.t, stack, sp, node, n_nodes, ray_nodes);
					}
					if (ENTRY)
					{
						const int pixel = id % ea->pixel_count;
						const int py = pixel / ea->width, px = pixel - py * ea->width;
						cut = ((py >> ea->tile_h_shift) * ea->tiles_x + (px >> ea->tile_w_shift)) << ea->stride_shift;
						const int2 e = __ldg(ea->cuts + cut);
						node = PTB_DONE;
						if (e.x != PTB_DONE && __int_as_float(e.y) <= best.t) { node = e.x; cut++; }
					}
				}
			}
			continue;
		}
		if ((m_node | m_leaf) == 0u) break;   // nothing in flight and nothing left to fetch

		if (FUSED && m_scat != 0u && (__popc(m_scat) >= fa->scatter_min || ((m_node & ~m_scat) | m_leaf) == 0u))
		{
			// ---- scatter phase (fused_scatter_event): the lanes whose search ended below the free-flight bound without a hit
			if (has_ray && at_scatter)
			{
				at_scatter = false;
				if (!fused_scatter_event(sc, st, *fa, id, fa->loop_depth + lead, thr_cached, &s_thr[threadIdx.x], s_hist, o, d, best.t)) id = -1;
				else
				{
					lead++;
					best.t1 = CUDART_INF_F; best.t2 = CUDART_INF_F; best.prim = -1;
					best_tri = 0x7fffffff;
					for (int s = 0; s < sc.n_spheres; s++)
					{
						const float4 sph = __ldg(&sc.spheres[s]);
						float t;
						if (intersect_sphere(make_float3(sph.x, sph.y, sph.z), sph.w, o, d, t) && t < best.t && t > 0.0f)
						{
							best.t = t;
							best.prim = -(s + 2);
						}
					}
					const float tiny = 1e-30f;
					const float3 ds = make_float3(fabsf(d.x) < tiny ? copysignf(tiny, d.x) : d.x, fabsf(d.y) < tiny ? copysignf(tiny, d.y) : d.y,
						fabsf(d.z) < tiny ? copysignf(tiny, d.z) : d.z);
					idir = make_float3(1.0f / ds.x, 1.0f / ds.y, 1.0f / ds.z);
					noidir = make_float3(-o.x * idir.x, -o.y * idir.y, -o.z * idir.z);
					margin2 = 4.8e-7f * fmaxf(fmaxf(fabsf(noidir.x), fabsf(noidir.y)), fabsf(noidir.z));
					sp = 0;
					node = sc.n_triangles > 0 ? sc.root_ref : PTB_DONE;
					// the walk stays near the surface it entered through: the next search starts at the same leaf
					if (UPWALK && start_slot >= 0) upwalk_start<COUNT>(sc, start_slot, idir, noidir, margin2, best.t, stack, sp, node, n_nodes, ray_nodes);
				}
			}
			continue;
		}

		if (m_leaf != 0u && (__popc(m_leaf) >= leaf_min || (m_node & ~m_scat) == 0u))
		{
			// ---- leaf phase: node = ~((first << 3) | (count - 1))
			if (has_ray && node < 0 && !(FUSED && at_scatter))   // a lane waiting for its scatter event sits at PTB_DONE, which is negative too
			{
				const int ref = ~node;
				const int first = ref >> 3;
#if PTB_LEAF_SINGLE
				{
					if (COUNT) n_tris++;
					const float4* tp = sc.tri_isect + (size_t)first * 3;
#else
				const int cnt = (ref & 7) + 1;
				for (int k = 0; k < cnt; k++)
				{
					if (COUNT) n_tris++;
					const float4* tp = sc.tri_isect + (size_t)(first + k) * 3;
#endif
					const float4 a = __ldg(tp + 0), b = __ldg(tp + 1), c = __ldg(tp + 2);
					float t, t1, t2;
					if (intersect_triangle(make_float3(a.x, a.y, a.z), make_float3(b.x, b.y, b.z), make_float3(c.x, c.y, c.z), o, d, t, t1, t2) && t > 0.0f)
					{
						const int tid = __float_as_int(a.w);
						if (t < best.t || (t == best.t && best.prim >= 0 && tid < best_tri))
						{
							best.t = t; best.t1 = t1; best.t2 = t2; best.prim = tid; best_tri = tid;
						}
					}
				}
#if PTB_LEAF_SINGLE
				// (first + 1, count - 1) is the same reference plus 7: ((first + 1) << 3 | (count - 2)) - (first << 3 | (count - 1))
				if (ref & 7) node = ~(ref + 7);
				else
#endif
				if (sp > 0) PTB_POP(node); else node = PTB_DONE;
			}
#if PTB_LEAF_FALLTHROUGH
			// lanes at an inner node take their node steps in the SAME iteration (one round of votes for both phases); a lane whose leaf
			// just ended joins them with the node it popped
			if ((m_node & ~m_scat) == 0u) continue;
#else
			continue;
#endif
		}

		// ---- node phase (node_reps steps for every lane sitting at an inner node)
		// REPS > 0: the shipped count, unrolled at compile time; REPS == 0: runtime count (tuning sweeps)
#pragma unroll
		for (int rep = 0; rep < (REPS > 0 ? REPS : node_reps); rep++)
		if (id >= 0 && node >= 0)
		{
			if (COUNT) { n_nodes++; ray_nodes++; }
			const float4* np = sc.bvh_nodes + (size_t)node * 4;
			float4 n0, n1, n2;
			float2 n3;
			if (TREELET && node < n_top)
			{
				const float4* sp4 = s_top + node * 4;
				n0 = sp4[0]; n1 = sp4[1]; n2 = sp4[2];
				n3 = *reinterpret_cast<const float2*>(sp4 + 3);
			}
			else load_node(np, n0, n1, n2, n3);
			const float c0x0 = fmaf(n0.x, idir.x, noidir.x), c0x1 = fmaf(n0.y, idir.x, noidir.x);
			const float c0y0 = fmaf(n0.z, idir.y, noidir.y), c0y1 = fmaf(n0.w, idir.y, noidir.y);
			const float c0z0 = fmaf(n2.x, idir.z, noidir.z), c0z1 = fmaf(n2.y, idir.z, noidir.z);
			const float c1x0 = fmaf(n1.x, idir.x, noidir.x), c1x1 = fmaf(n1.y, idir.x, noidir.x);
			const float c1y0 = fmaf(n1.z, idir.y, noidir.y), c1y1 = fmaf(n1.w, idir.y, noidir.y);
			const float c1z0 = fmaf(n2.z, idir.z, noidir.z), c1z1 = fmaf(n2.w, idir.z, noidir.z);
			const float tmin0 = fmaxf(fmaxf(fminf(c0x0, c0x1), fminf(c0y0, c0y1)), fmaxf(fminf(c0z0, c0z1), 0.0f));
			const float tmax0 = fminf(fminf(fmaxf(c0x0, c0x1), fmaxf(c0y0, c0y1)), fminf(fmaxf(c0z0, c0z1), best.t));
			const float tmin1 = fmaxf(fmaxf(fminf(c1x0, c1x1), fminf(c1y0, c1y1)), fmaxf(fminf(c1z0, c1z1), 0.0f));
			const float tmax1 = fminf(fminf(fmaxf(c1x0, c1x1), fmaxf(c1y0, c1y1)), fminf(fmaxf(c1z0, c1z1), best.t));
			// conservative overlap test tmin * LO - margin <= tmax * HI, divided through by HI (one FMA per box)
			const bool h0 = fmaf(tmin0, PTB_SLACK_LO / PTB_SLACK_HI, -margin2) <= tmax0;
			const bool h1 = fmaf(tmin1, PTB_SLACK_LO / PTB_SLACK_HI, -margin2) <= tmax1;
			const int child0 = __float_as_int(n3.x), child1 = __float_as_int(n3.y);
			const bool both = h0 && h1;
			const bool swap = tmin1 < tmin0;
			const int near_c = swap ? child1 : child0;
			const int far_c = swap ? child0 : child1;
			int next = both ? near_c : (h0 ? child0 : (h1 ? child1 : PTB_DONE));
			if (both) PTB_PUSH(far_c);
			else if (!(h0 || h1) && sp > 0) PTB_POP(next);
			if (ENTRY && next == PTB_DONE)
			{
				// sub-tree finished inside the phase: go on with the tile's next entry instead of idling until the next vote
				const int2 e = __ldg(ea->cuts + cut);
				if (e.x != PTB_DONE && __int_as_float(e.y) <= best.t) { next = e.x; cut++; }
			}
			node = next;
		}
	}
	if (COUNT)
	{
		for (int off = 16; off > 0; off >>= 1)
		{
			n_nodes += __shfl_down_sync(FULL, n_nodes, off);
			n_tris += __shfl_down_sync(FULL, n_tris, off);
		}
		if (lane == 0)
		{
			atomicAdd(&counters[0], (unsigned long long)n_nodes);
			atomicAdd(&counters[1], (unsigned long long)n_tris);
		}
	}
}

#undef PTB_PUSH
#undef PTB_POP

template <bool COUNT, int REPS>
__global__ void __launch_bounds__(128, PTB_PERSISTENT_MIN_BLOCKS) k_extend_persistent(DeviceScene sc, PathState st, const int* __restrict__ queue, const int* __restrict__ count_ptr,
	int* __restrict__ work_counter, unsigned long long* __restrict__ counters, int refill_min, int leaf_min, int node_reps)
{
	extend_persistent_body<COUNT, REPS, false, false>(sc, st, queue, count_ptr, work_counter, counters, refill_min, leaf_min, node_reps, nullptr, 0, nullptr);
}

// bounce rays on the binary tree started at the leaf of the triangle they leave (kernels_entry.cuh: k_up_level)
template <bool COUNT, int REPS>
__global__ void __launch_bounds__(128, PTB_PERSISTENT_MIN_BLOCKS) k_extend_upwalk(DeviceScene sc, PathState st, const int* __restrict__ queue, const int* __restrict__ count_ptr,
	int* __restrict__ work_counter, unsigned long long* __restrict__ counters, int refill_min, int leaf_min, int node_reps)
{
	extend_persistent_body<COUNT, REPS, false, false, false, false, true>(sc, st, queue, count_ptr, work_counter, counters, refill_min, leaf_min, node_reps, nullptr, 0, nullptr);
}

// camera rays (depth 0) started at their tile's entry cut (kernels_entry.cuh); REPS as in k_extend_persistent
template <bool COUNT, int REPS>
__global__ void __launch_bounds__(128, PTB_PERSISTENT_MIN_BLOCKS) k_extend_entry(DeviceScene sc, PathState st, const int* __restrict__ queue, const int* __restrict__ count_ptr,
	int* __restrict__ work_counter, unsigned long long* __restrict__ counters, int refill_min, int leaf_min, int node_reps, EntryArgs ea)
{
	extend_persistent_body<COUNT, REPS, false, false, false, true>(sc, st, queue, count_ptr, work_counter, counters, refill_min, leaf_min, node_reps, nullptr, 0, nullptr, nullptr, nullptr, nullptr, &ea);
}

// the binary-tree kernel with inline scatter events (option inline_scatter with fused_tree=2: the whole subsurface walk over the binary tree)
template <bool COUNT, bool UPWALK = false>
__global__ void __launch_bounds__(128, PTB_PERSISTENT_MIN_BLOCKS) k_extend_persistent_fused(DeviceScene sc, PathState st, const int* __restrict__ queue, const int* __restrict__ count_ptr,
	int* __restrict__ work_counter, unsigned long long* __restrict__ counters, int refill_min, int leaf_min, FusedArgs fa)
{
	__shared__ int s_hist[PTB_FUSED_HIST];
	__shared__ float4 s_thr[128];
	__shared__ int s_from[128];
	if (threadIdx.x < PTB_FUSED_HIST) s_hist[threadIdx.x] = 0;
	__syncthreads();
	extend_persistent_body<COUNT, 6, false, false, true, false, UPWALK>(sc, st, queue, count_ptr, work_counter, counters, refill_min, leaf_min, 6, nullptr, 0, nullptr, &fa, s_thr, s_hist, nullptr, s_from);
	__syncthreads();
	if (threadIdx.x < PTB_FUSED_HIST && s_hist[threadIdx.x])
		atomicAdd(&fa.depth_segments[min((int)threadIdx.x, fa.n_depth_slots - 1)], (unsigned long long)s_hist[threadIdx.x]);
}

// STAGED refill (extend_variant 4): 96 float4 (1.5 KB) of shared memory per warp hold 32 set-up rays
template <bool COUNT>
__global__ void __launch_bounds__(128, PTB_PERSISTENT_MIN_BLOCKS) k_extend_staged(DeviceScene sc, PathState st, const int* __restrict__ queue, const int* __restrict__ count_ptr,
	int* __restrict__ work_counter, unsigned long long* __restrict__ counters, int refill_min, int leaf_min)
{
	__shared__ float4 s_stage_all[4 * 96];
	extend_persistent_body<COUNT, 6, false, true>(sc, st, queue, count_ptr, work_counter, counters, refill_min, leaf_min, 6, nullptr, 0, s_stage_all + (threadIdx.x >> 5) * 96);
}

// up to 1024 threads per block (1 x 1024 or 2 x 512 per SM: 64 registers per thread either way), dynamic shared memory = n_top x 64 bytes
template <bool COUNT>
__global__ void __launch_bounds__(1024, 1) k_extend_treelet(DeviceScene sc, PathState st, const int* __restrict__ queue, const int* __restrict__ count_ptr,
	int* __restrict__ work_counter, unsigned long long* __restrict__ counters, int refill_min, int leaf_min, int n_top)
{
	extern __shared__ float4 s_top_nodes[];
	for (int i = threadIdx.x; i < n_top * 4; i += blockDim.x) s_top_nodes[i] = __ldg(sc.bvh_nodes + i);
	__syncthreads();
	extend_persistent_body<COUNT, 6, true, false>(sc, st, queue, count_ptr, work_counter, counters, refill_min, leaf_min, 6, s_top_nodes, n_top, nullptr);
}


// ------------------------------------------------------------------------------------------
// k_extend_speculative — k_extend_persistent with POSTPONED LEAVES (after Aila & Laine 2009, "speculative traversal"):
// a lane that reaches a leaf parks it in `leaf` and keeps traversing (pops the next node) for the rest of the node phase instead of
// idling until enough lanes wait at a leaf; it only stalls when it meets a second leaf.  The parked leaf's first triangle can be
// prefetched into L1 meanwhile (PREFETCH).  Culling uses whatever best.t the lane has at that moment, so postponing a leaf can only
// visit MORE nodes, never miss one: hits are identical (tests/test_gpu_parity.py runs every variant against the exhaustive scan).
// ------------------------------------------------------------------------------------------
template <bool COUNT, bool PREFETCH>
__global__ void __launch_bounds__(128, PTB_PERSISTENT_MIN_BLOCKS) k_extend_speculative(DeviceScene sc, PathState st, const int* __restrict__ queue, const int* __restrict__ count_ptr,
	int* __restrict__ work_counter, unsigned long long* __restrict__ counters, int refill_min, int leaf_min)
{
	const int count = *count_ptr;
	const unsigned lane = threadIdx.x & 31u;
	const unsigned lane_lt = (1u << lane) - 1u;
	const unsigned FULL = 0xffffffffu;
	unsigned n_nodes = 0, n_tris = 0;

	int id = -1;
	bool exhausted = false;
	float3 o = make_float3(0, 0, 0), d = o, idir = o, noidir = o;
	float margin2 = 0.0f;
	HitRecord best;
	best.t = CUDART_INF_F; best.t1 = 0.0f; best.t2 = 0.0f; best.prim = -1;
	int best_tri = 0x7fffffff;
	int stack[PTB_STACK_SIZE];
	int sp = 0;
	int node = PTB_DONE;         // >= 0 inner node, PTB_DONE = nothing left, other negative = a leaf this lane is stuck at (slot taken)
	int leaf = 0;                // < 0: parked leaf reference still to be tested, 0: none

	while (true)
	{
		if (id >= 0 && node == PTB_DONE && leaf == 0)
		{
			__stcs(&st.hit[id], make_float4(best.prim == -1 ? CUDART_INF_F : best.t, best.t1, best.t2, __int_as_float(best.prim)));
			id = -1;
		}
		const bool has_ray = id >= 0;
		const unsigned m_idle = __ballot_sync(FULL, !has_ray);
		const unsigned m_node = __ballot_sync(FULL, has_ray && node >= 0);
		const unsigned m_leaf = __ballot_sync(FULL, has_ray && leaf < 0);

		if (m_idle != 0u && !exhausted && (__popc(m_idle) >= refill_min || (m_node | m_leaf) == 0u))
		{
			const int n = __popc(m_idle);
			int base = 0;
			if (lane == 0) base = atomicAdd(work_counter, n);
			base = __shfl_sync(FULL, base, 0);
			if (base + n >= count) exhausted = true;
			if (!has_ray)
			{
				const int i = base + __popc(m_idle & lane_lt);
				if (i < count)
				{
					id = __ldcs(&queue[i]);
					const float4 o4 = __ldcs(&st.ray_o[id]), d4 = __ldcs(&st.ray_d[id]);
					o = make_float3(o4.x, o4.y, o4.z);
					d = make_float3(d4.x, d4.y, d4.z);
					best.t = d4.w; best.t1 = CUDART_INF_F; best.t2 = CUDART_INF_F; best.prim = -1;
					best_tri = 0x7fffffff;
					for (int s = 0; s < sc.n_spheres; s++)
					{
						const float4 sph = __ldg(&sc.spheres[s]);
						float t;
						if (intersect_sphere(make_float3(sph.x, sph.y, sph.z), sph.w, o, d, t) && t < best.t && t > 0.0f)
						{
							best.t = t;
							best.prim = -(s + 2);
						}
					}
					const float tiny = 1e-30f;
					const float3 ds = make_float3(fabsf(d.x) < tiny ? copysignf(tiny, d.x) : d.x, fabsf(d.y) < tiny ? copysignf(tiny, d.y) : d.y,
						fabsf(d.z) < tiny ? copysignf(tiny, d.z) : d.z);
					idir = make_float3(1.0f / ds.x, 1.0f / ds.y, 1.0f / ds.z);
					noidir = make_float3(-o.x * idir.x, -o.y * idir.y, -o.z * idir.z);
					margin2 = 4.8e-7f * fmaxf(fmaxf(fabsf(noidir.x), fabsf(noidir.y)), fabsf(noidir.z));
					sp = 0;
					leaf = 0;
					node = sc.n_triangles > 0 ? sc.root_ref : PTB_DONE;
				}
			}
			continue;
		}
		if ((m_node | m_leaf) == 0u) break;

		if (m_leaf != 0u && (__popc(m_leaf) >= leaf_min || m_node == 0u))
		{
			// ---- leaf phase: ONE triangle of the parked leaf per vote; leaf = ~((first << 3) | (count - 1))
			if (has_ray && leaf < 0)
			{
				const int ref = ~leaf;
				if (COUNT) n_tris++;
				const float4* tp = sc.tri_isect + (size_t)(ref >> 3) * 3;
				const float4 a = __ldg(tp + 0), b = __ldg(tp + 1), c = __ldg(tp + 2);
				float t, t1, t2;
				if (intersect_triangle(make_float3(a.x, a.y, a.z), make_float3(b.x, b.y, b.z), make_float3(c.x, c.y, c.z), o, d, t, t1, t2) && t > 0.0f)
				{
					const int tid = __float_as_int(a.w);
					if (t < best.t || (t == best.t && best.prim >= 0 && tid < best_tri))
					{
						best.t = t; best.t1 = t1; best.t2 = t2; best.prim = tid; best_tri = tid;
					}
				}
				if (ref & 7) leaf = ~(ref + 7);          // (first + 1, count - 1)
				else if (node < 0 && node != PTB_DONE)
				{
					// the leaf this lane was stuck at moves into the slot; traversal resumes from the stack
					leaf = node;
					if (PREFETCH) asm volatile("prefetch.global.L1 [%0];" :: "l"(sc.tri_isect + (size_t)((~leaf) >> 3) * 3));
					node = sp > 0 ? stack[--sp] : PTB_DONE;
				}
				else leaf = 0;
			}
			continue;
		}

		// ---- node phase
#pragma unroll
		for (int rep = 0; rep < 6; rep++)
		if (id >= 0 && node >= 0)
		{
			if (COUNT) n_nodes++;
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
			const float tmax0 = fminf(fminf(fmaxf(c0x0, c0x1), fmaxf(c0y0, c0y1)), fminf(fmaxf(c0z0, c0z1), best.t));
			const float tmin1 = fmaxf(fmaxf(fminf(c1x0, c1x1), fminf(c1y0, c1y1)), fmaxf(fminf(c1z0, c1z1), 0.0f));
			const float tmax1 = fminf(fminf(fmaxf(c1x0, c1x1), fmaxf(c1y0, c1y1)), fminf(fmaxf(c1z0, c1z1), best.t));
			const bool h0 = fmaf(tmin0, PTB_SLACK_LO / PTB_SLACK_HI, -margin2) <= tmax0;
			const bool h1 = fmaf(tmin1, PTB_SLACK_LO / PTB_SLACK_HI, -margin2) <= tmax1;
			const int child0 = __float_as_int(n3.x), child1 = __float_as_int(n3.y);
			const bool both = h0 && h1;
			const bool swap = tmin1 < tmin0;
			const int near_c = swap ? child1 : child0;
			const int far_c = swap ? child0 : child1;
			int next = both ? near_c : (h0 ? child0 : (h1 ? child1 : PTB_DONE));
			if (both) { if (sp < PTB_STACK_SIZE) stack[sp] = far_c; sp++; }
			else if (!(h0 || h1) && sp > 0) next = stack[min(--sp, PTB_STACK_SIZE - 1)];
			if (next < 0 && next != PTB_DONE && leaf == 0)
			{
				// park the leaf, keep traversing
				leaf = next;
				if (PREFETCH) asm volatile("prefetch.global.L1 [%0];" :: "l"(sc.tri_isect + (size_t)((~leaf) >> 3) * 3));
				next = sp > 0 ? stack[min(--sp, PTB_STACK_SIZE - 1)] : PTB_DONE;
			}
			node = next;
		}
	}
	if (COUNT)
	{
		for (int off = 16; off > 0; off >>= 1)
		{
			n_nodes += __shfl_down_sync(FULL, n_nodes, off);
			n_tris += __shfl_down_sync(FULL, n_tris, off);
		}
		if (lane == 0)
		{
			atomicAdd(&counters[0], (unsigned long long)n_nodes);
			atomicAdd(&counters[1], (unsigned long long)n_tris);
		}
	}
}

// ------------------------------------------------------------------------------------------
// k_extend_persistent8 — the same persistent, warp-voting scheme over the compressed 8-wide layout
// (bvh.h layout #2).  Lane state: `current` = a group of not-yet-visited inner children of one wide node
// (child base + hit bits), `tri_group` = leaf triangles still to test; stack entries are node groups.
// Phases: refill | wide-node step (one child popped, 8 quantised boxes decoded and tested) | triangle step.
// ------------------------------------------------------------------------------------------
template <bool COUNT, bool FUSED = false>
__global__ void __launch_bounds__(128, PTB_PERSISTENT_MIN_BLOCKS8) k_extend_persistent8(DeviceScene sc, PathState st, const int* __restrict__ queue, const int* __restrict__ count_ptr,
	int* __restrict__ work_counter, unsigned long long* __restrict__ counters, int refill_min, int leaf_min, FusedArgs fa = FusedArgs())
{
	const int count = *count_ptr;
	const unsigned lane = threadIdx.x & 31u;
	const unsigned lane_lt = (1u << lane) - 1u;
	const unsigned FULL = 0xffffffffu;
	unsigned n_nodes = 0, n_tris = 0;

	int id = -1;
	bool exhausted = false;
	float3 o = make_float3(0, 0, 0), d = o, idir = o;
	unsigned oct_inv4 = 0;
	HitRecord best;
	best.t = CUDART_INF_F; best.t1 = 0.0f; best.t2 = 0.0f; best.prim = -1;
	int best_tri = 0x7fffffff;
	uint2 stack[PTB_STACK_SIZE8];
#define PTB_PUSH8(v) do { if (sp < PTB_STACK_SIZE8) stack[sp] = (v); sp++; } while (0)
#define PTB_POP8(dst) do { --sp; dst = stack[min(sp, PTB_STACK_SIZE8 - 1)]; } while (0)
	int sp = 0;
	uint2 current = make_uint2(0u, 0u), tri_group = make_uint2(0u, 0u);
	int lead = 0;                // FUSED: bounces this path is ahead of the loop depth
	bool at_scatter = false;     // FUSED: the search ended below the free-flight bound without a hit: a scatter event is due
	__shared__ int s_hist[FUSED ? PTB_FUSED_HIST : 1];
	// FUSED: throughput + medium index of the lane's path while it scatters (loaded at its first event here, written back when the path
	// leaves the kernel alive): every later event costs shared-memory latency instead of a scattered 16-byte global load and store
	__shared__ float4 s_thr[FUSED ? 128 : 1];
	__shared__ int s_from[FUSED ? 128 : 1];     // the "segment leaves triangle" bits of ray_o.w (kernels.cuh: PTB_FROM_BITS_OF), carried through unchanged
	bool thr_cached = false;
	if (FUSED)
	{
		if (threadIdx.x < PTB_FUSED_HIST) s_hist[threadIdx.x] = 0;
		__syncthreads();
	}

	while (true)
	{
		// bookkeeping without votes: pop a node group when the lane ran dry, retire when nothing is left
		if (id >= 0 && !at_scatter && (current.y & 0xff000000u) == 0u && tri_group.y == 0u)
		{
			if (sp > 0) PTB_POP8(current);
			else if (FUSED && best.prim == -1 && best.t < CUDART_INF_F) at_scatter = true;
			else
			{
				__stcs(&st.hit[id], make_float4(best.prim == -1 ? CUDART_INF_F : best.t, best.t1, best.t2, __int_as_float(best.prim)));
				if (FUSED && thr_cached)
				{
					// the path scattered here: the ray in memory is the one it had when it entered (k_shade needs the segment that found the
					// surface) and its throughput lives in shared memory
					st.ray_o[id] = make_float4(o.x, o.y, o.z, __int_as_float(lead | s_from[threadIdx.x]));
					st.ray_d[id] = make_float4(d.x, d.y, d.z, 0.0f);
					st.throughput[id] = s_thr[threadIdx.x];
				}
				id = -1;
			}
		}
		const bool has_ray = id >= 0;
		const bool at_tri = has_ray && tri_group.y != 0u;
		const bool at_node = has_ray && !at_tri && !at_scatter && (current.y & 0xff000000u) != 0u;
		const unsigned m_idle = __ballot_sync(FULL, !has_ray);
		const unsigned m_scat = FUSED ? __ballot_sync(FULL, has_ray && at_scatter) : 0u;
		const unsigned m_node = __ballot_sync(FULL, at_node) | m_scat;   // "work in flight" for the refill / exit tests below
		const unsigned m_tri = __ballot_sync(FULL, at_tri);

		if (m_idle != 0u && !exhausted && (__popc(m_idle) >= refill_min || (m_node | m_tri) == 0u))
		{
			const int n = __popc(m_idle);
			int base = 0;
			if (lane == 0) base = atomicAdd(work_counter, n);
			base = __shfl_sync(FULL, base, 0);
			if (base + n >= count) exhausted = true;
			if (!has_ray)
			{
				const int i = base + __popc(m_idle & lane_lt);
				if (i < count)
				{
					id = __ldcs(&queue[i]);
					const float4 o4 = __ldcs(&st.ray_o[id]), d4 = __ldcs(&st.ray_d[id]);
					o = make_float3(o4.x, o4.y, o4.z);
					d = make_float3(d4.x, d4.y, d4.z);
					best.t = d4.w; best.t1 = CUDART_INF_F; best.t2 = CUDART_INF_F; best.prim = -1;
					best_tri = 0x7fffffff;
					if (FUSED) { lead = PTB_LEAD_OF(__float_as_int(o4.w)); s_from[threadIdx.x] = PTB_FROM_BITS_OF(__float_as_int(o4.w)); at_scatter = false; thr_cached = false; atomicAdd(&s_hist[min(fa.loop_depth + lead, PTB_FUSED_HIST - 1)], 1); }
					for (int s = 0; s < sc.n_spheres; s++)
					{
						const float4 sph = __ldg(&sc.spheres[s]);
						float t;
						if (intersect_sphere(make_float3(sph.x, sph.y, sph.z), sph.w, o, d, t) && t < best.t && t > 0.0f)
						{
							best.t = t;
							best.prim = -(s + 2);
						}
					}
					const float tiny = 1e-30f;
					const float3 ds = make_float3(fabsf(d.x) < tiny ? copysignf(tiny, d.x) : d.x, fabsf(d.y) < tiny ? copysignf(tiny, d.y) : d.y,
						fabsf(d.z) < tiny ? copysignf(tiny, d.z) : d.z);
					idir = make_float3(1.0f / ds.x, 1.0f / ds.y, 1.0f / ds.z);
					oct_inv4 = (d.x < 0.0f ? 0u : 0x04040404u) | (d.y < 0.0f ? 0u : 0x02020202u) | (d.z < 0.0f ? 0u : 0x01010101u);
					sp = 0;
					tri_group = make_uint2(0u, 0u);
					current = sc.n_triangles > 0 ? make_uint2(0u, 0x80000000u) : make_uint2(0u, 0u);
				}
			}
			continue;
		}
		if ((m_node | m_tri) == 0u) break;

		if (FUSED && m_scat != 0u && (__popc(m_scat) >= fa.scatter_min || ((m_node & ~m_scat) | m_tri) == 0u))
		{
			// ---- scatter phase: the medium event of k_shade (kernels_shade.cuh) for the lanes whose search ended below the free-flight bound
			if (has_ray && at_scatter)
			{
				at_scatter = false;
				if (!fused_scatter_event(sc, st, fa, id, fa.loop_depth + lead, thr_cached, &s_thr[threadIdx.x], s_hist, o, d, best.t)) id = -1;
				else
				{
					lead++;
					best.t1 = CUDART_INF_F; best.t2 = CUDART_INF_F; best.prim = -1;
					best_tri = 0x7fffffff;
					for (int sidx = 0; sidx < sc.n_spheres; sidx++)
					{
						const float4 sph = __ldg(&sc.spheres[sidx]);
						float t;
						if (intersect_sphere(make_float3(sph.x, sph.y, sph.z), sph.w, o, d, t) && t < best.t && t > 0.0f)
						{
							best.t = t;
							best.prim = -(sidx + 2);
						}
					}
					const float tiny = 1e-30f;
					const float3 ds = make_float3(fabsf(d.x) < tiny ? copysignf(tiny, d.x) : d.x, fabsf(d.y) < tiny ? copysignf(tiny, d.y) : d.y,
						fabsf(d.z) < tiny ? copysignf(tiny, d.z) : d.z);
					idir = make_float3(1.0f / ds.x, 1.0f / ds.y, 1.0f / ds.z);
					oct_inv4 = (d.x < 0.0f ? 0u : 0x04040404u) | (d.y < 0.0f ? 0u : 0x02020202u) | (d.z < 0.0f ? 0u : 0x01010101u);
					sp = 0;
					tri_group = make_uint2(0u, 0u);
					current = sc.n_triangles > 0 ? make_uint2(0u, 0x80000000u) : make_uint2(0u, 0u);
				}
			}
			continue;
		}
		if (m_tri != 0u && (__popc(m_tri) >= leaf_min || (m_node & ~m_scat) == 0u))
		{
			// ---- triangle phase
			if (at_tri)
			{
#if PTB_LEAF_SINGLE
				{
#else
				while (tri_group.y)
				{
#endif
					const unsigned k = 31u - __clz(tri_group.y);
					tri_group.y &= ~(1u << k);
					if (COUNT) n_tris++;
					const float4* tp = sc.tri_isect + (size_t)(tri_group.x + k) * 3;
					const float4 a = __ldg(tp + 0), b = __ldg(tp + 1), c = __ldg(tp + 2);
					float t, t1, t2;
					if (intersect_triangle(make_float3(a.x, a.y, a.z), make_float3(b.x, b.y, b.z), make_float3(c.x, c.y, c.z), o, d, t, t1, t2) && t > 0.0f)
					{
						const int tid = __float_as_int(a.w);
						if (t < best.t || (t == best.t && best.prim >= 0 && tid < best_tri))
						{
							best.t = t; best.t1 = t1; best.t2 = t2; best.prim = tid; best_tri = tid;
						}
					}
				}
			}
			continue;
		}

		// ---- wide-node phase
		if (at_node)
		{
			const unsigned hits_imask = current.y;
			const unsigned child_index_offset = 31u - __clz(hits_imask);
			const unsigned child_index_base = current.x;
			current.y &= ~(1u << child_index_offset);
			if (current.y & 0xff000000u) PTB_PUSH8(current);
			const unsigned slot_index = (child_index_offset - 24u) ^ (oct_inv4 & 0xffu);
			const unsigned relative_index = __popc(hits_imask & ~(0xffffffffu << slot_index));
			const unsigned node_index = child_index_base + relative_index;
			if (COUNT) n_nodes++;

			const float4* np = sc.bvh_nodes + (size_t)node_index * 5;
			const float4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
			const unsigned e_imask = __float_as_uint(n0.w);
			const float3 adir = make_float3(__uint_as_float(extract_byte(e_imask, 0) << 23) * idir.x, __uint_as_float(extract_byte(e_imask, 1) << 23) * idir.y,
				__uint_as_float(extract_byte(e_imask, 2) << 23) * idir.z);
			const float3 aorg = make_float3((n0.x - o.x) * idir.x, (n0.y - o.y) * idir.y, (n0.z - o.z) * idir.z);
			unsigned hit_mask = 0;
#pragma unroll
			for (int half = 0; half < 2; half++)
			{
				const unsigned meta4 = __float_as_uint(half == 0 ? n1.z : n1.w);
				const unsigned is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
				const unsigned inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
				const unsigned bit_index4 = (meta4 ^ (oct_inv4 & inner_mask4)) & 0x1f1f1f1fu;
				const unsigned child_bits4 = (meta4 >> 5) & 0x07070707u;
				const unsigned qlox = __float_as_uint(half == 0 ? n2.x : n2.y), qhix = __float_as_uint(half == 0 ? n2.z : n2.w);
				const unsigned qloy = __float_as_uint(half == 0 ? n3.x : n3.y), qhiy = __float_as_uint(half == 0 ? n3.z : n3.w);
				const unsigned qloz = __float_as_uint(half == 0 ? n4.x : n4.y), qhiz = __float_as_uint(half == 0 ? n4.z : n4.w);
				const unsigned x_min = d.x < 0.0f ? qhix : qlox, x_max = d.x < 0.0f ? qlox : qhix;
				const unsigned y_min = d.y < 0.0f ? qhiy : qloy, y_max = d.y < 0.0f ? qloy : qhiy;
				const unsigned z_min = d.z < 0.0f ? qhiz : qloz, z_max = d.z < 0.0f ? qloz : qhiz;
#pragma unroll
				for (int j = 0; j < 4; j++)
				{
					const float tx0 = fmaf(byte_to_float(x_min, j), adir.x, aorg.x), tx1 = fmaf(byte_to_float(x_max, j), adir.x, aorg.x);
					const float ty0 = fmaf(byte_to_float(y_min, j), adir.y, aorg.y), ty1 = fmaf(byte_to_float(y_max, j), adir.y, aorg.y);
					const float tz0 = fmaf(byte_to_float(z_min, j), adir.z, aorg.z), tz1 = fmaf(byte_to_float(z_max, j), adir.z, aorg.z);
					const float tmin = fmaxf(fmaxf(tx0, ty0), fmaxf(tz0, 0.0f));
					const float tmax = fminf(fminf(tx1, ty1), fminf(tz1, best.t));
					if (tmin * PTB_SLACK_LO <= tmax * PTB_SLACK_HI)
						hit_mask |= extract_byte(child_bits4, j) << extract_byte(bit_index4, j);
				}
			}
			current.x = __float_as_uint(n1.x);
			tri_group.x = __float_as_uint(n1.y);
			current.y = (hit_mask & 0xff000000u) | (e_imask >> 24);
			tri_group.y = hit_mask & 0x00ffffffu;
		}
	}
	if (FUSED)
	{
		__syncthreads();
		if (threadIdx.x < PTB_FUSED_HIST && s_hist[threadIdx.x])
			atomicAdd(&fa.depth_segments[min((int)threadIdx.x, fa.n_depth_slots - 1)], (unsigned long long)s_hist[threadIdx.x]);
	}
	if (COUNT)
	{
		for (int off = 16; off > 0; off >>= 1)
		{
			n_nodes += __shfl_down_sync(FULL, n_nodes, off);
			n_tris += __shfl_down_sync(FULL, n_tris, off);
		}
		if (lane == 0)
		{
			atomicAdd(&counters[3], (unsigned long long)n_nodes);   // wide-node visits are tallied apart from binary-node visits
			atomicAdd(&counters[1], (unsigned long long)n_tris);
		}
	}
}

#undef PTB_PUSH8
#undef PTB_POP8

// brute-force closest hit over every primitive (test hook; same acceptance arithmetic)
__global__ void k_bruteforce(DeviceScene sc, const float4* __restrict__ ray_o, const float4* __restrict__ ray_d, float4* __restrict__ hit, int n)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float3 o = make_float3(ray_o[i].x, ray_o[i].y, ray_o[i].z), d = make_float3(ray_d[i].x, ray_d[i].y, ray_d[i].z);
	float best_t = CUDART_INF_F, bt1 = CUDART_INF_F, bt2 = CUDART_INF_F;
	int prim = -1;
	for (int s = 0; s < sc.n_spheres; s++)
	{
		float4 sp = sc.spheres[s];
		float t;
		if (intersect_sphere(make_float3(sp.x, sp.y, sp.z), sp.w, o, d, t) && t < best_t && t > 0.0f) { best_t = t; prim = -(s + 2); }
	}
	for (int k = 0; k < sc.n_triangles; k++)
	{
		const float4* tp = sc.tri_isect + (size_t)k * 3;
		float4 a = tp[0], b = tp[1], c = tp[2];
		float t, t1, t2;
		if (intersect_triangle(make_float3(a.x, a.y, a.z), make_float3(b.x, b.y, b.z), make_float3(c.x, c.y, c.z), o, d, t, t1, t2) && t > 0.0f)
		{
			int id = __float_as_int(a.w);
			if (t < best_t || (t == best_t && prim >= 0 && id < prim)) { best_t = t; bt1 = t1; bt2 = t2; prim = id; }
		}
	}
	hit[i] = make_float4(best_t, bt1, bt2, __int_as_float(prim));
}

} // namespace ptb
