// ptb200 wavefront integrator: entry cuts for camera rays (round 2).
//
// Every camera ray of an 8x4 pixel tile lies inside one thin shaft: the pyramid from the eye through the tile's rectangle on the canvas
// plane (generate_camera_ray, pt_device.cuh; path_tracer_kernel.cu:299-379), widened by the lens disc when the camera has an aperture.
// The closest-hit search of such a ray spends most of its node visits walking from the root down to the few leaves its shaft touches —
// the same walk for all 32 pixels of the tile and for every pass, because the shaft only depends on the camera.  k_entry_cut does that
// walk ONCE per tile: starting at the root it replaces a node by those of its children whose boxes overlap the shaft (one child
// overlapping = a free descent; both = the list grows by one) until the list holds `k_max` sub-trees, and stores them sorted by a lower
// bound on the hit distance.  k_extend_entry (kernels_extend.cuh, ENTRY) then starts each camera ray at its tile's list instead of at
// the root and stops walking the list at the first sub-tree that begins beyond its current hit.
//
// Correctness: every state of the list covers all triangles whose boxes overlap the shaft, the overlap test and the distance bound are
// conservative (margins below), so the rays find exactly the hits a search from the root finds (closest hit with the same acceptance
// arithmetic; tests/test_gpu_entry.py compares the accumulated images and the per-depth segment counts bit for bit over hand-picked and
// random cameras; tests/test_entry_shaft_math.py pins the lens derivation).  render.cu only builds the lists for cameras whose generator
// arithmetic stays within the shafts' slack (entry_cuts_usable) and for batches that amortise the build (ensure_entry_cuts).
// The same file holds k_tile_rank (queue positions of the non-empty tiles, for k_generate's sky fast path) and the sibling records of
// the bounce rays' leaf starts (k_up_level, k_up_pair).  Included by render.cu only.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include "kernels.cuh"

namespace ptb
{

using namespace ptbdev;

#define PTB_ENTRY_STRIDE 32          // most int2 slots per tile: <= 31 sub-trees + terminator (the lists are stored with a stride of
                                     // the next power of two above k_max)

// distances to the shaft's planes / eye of one box, relative to the eye
struct EntryShaft
{
	float3 eye, w;          // eye, unit view direction
	float3 n[4];            // inward unit normals of the four side planes of the pinhole pyramid
	float lens;             // aperture radius (0 = pinhole)
	float focal_lo;         // lower bound on dot(F - eye, w) over the tile's focal points F
	float focal_hi;         // upper bound (= focal distance)
};

// Conservative: false only when no camera ray of the tile can touch the box.  A ray from lens point eye + delta (|delta| <= lens,
// delta perpendicular to w) through focal point F is P(u) = [eye + u (F - eye)] + (1 - u) delta, u >= 0: the pinhole ray's point at
// depth fraction u displaced by at most |1 - u| * lens.  So P is at most |1 - u| * lens outside any side plane of the pinhole pyramid,
// with u = dot(P - eye, w) / dot(F - eye, w).
__device__ __forceinline__ bool entry_box_overlaps(const EntryShaft& s, float3 lo, float3 hi, float& t_lower)
{
	const float3 a = lo - s.eye, b = hi - s.eye;
	// depth range of the box along the view direction
	const float z_hi = fmaxf(a.x * s.w.x, b.x * s.w.x) + fmaxf(a.y * s.w.y, b.y * s.w.y) + fmaxf(a.z * s.w.z, b.z * s.w.z);
	const float z_lo = fminf(a.x * s.w.x, b.x * s.w.x) + fminf(a.y * s.w.y, b.y * s.w.y) + fminf(a.z * s.w.z, b.z * s.w.z);
	const float scale = fabsf(a.x) + fabsf(a.y) + fabsf(a.z) + fabsf(b.x) + fabsf(b.y) + fabsf(b.z);
	const float eps = 4e-6f * scale + 1e-30f;
	if (z_hi < -eps) return false;                           // behind the lens plane
	float widen = 0.0f;
	if (s.lens > 0.0f)
	{
		const float u_lo = fmaxf(z_lo, 0.0f) / s.focal_hi, u_hi = fmaxf(z_hi, 0.0f) / s.focal_lo;
		widen = s.lens * fmaxf(fabsf(1.0f - u_lo), fabsf(1.0f - u_hi)) * 1.0001f;
	}
#pragma unroll
	for (int k = 0; k < 4; k++)
	{
		const float3 n = s.n[k];
		const float m = fmaxf(a.x * n.x, b.x * n.x) + fmaxf(a.y * n.y, b.y * n.y) + fmaxf(a.z * n.z, b.z * n.z);   // most-inside corner
		if (m < -(widen + eps)) return false;
	}
	// lower bound on the ray parameter of any point in the box: |P - o| >= |P - eye| - lens (unit directions)
	const float dx = fmaxf(fmaxf(a.x, -b.x), 0.0f), dy = fmaxf(fmaxf(a.y, -b.y), 0.0f), dz = fmaxf(fmaxf(a.z, -b.z), 0.0f);
	t_lower = fmaxf(0.0f, (sqrtf(dx * dx + dy * dy + dz * dz) - s.lens) * 0.99999f - eps);
	return true;
}

// One thread per tile.  cuts[tile * stride + i] = (node reference, lower bound on t as float bits), ascending in t,
// terminated by (PTB_ENTRY_END, +inf).
__global__ void __launch_bounds__(128) k_entry_cut(DeviceScene sc, CameraParams cam, int width, int height, int tiles_x, int n_tiles, int k_max,
	int tile_w, int tile_h, int stride, int2* __restrict__ cuts)
{
	const int tile = blockIdx.x * blockDim.x + threadIdx.x;
	if (tile >= n_tiles) return;
	const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
	int2* out = cuts + (size_t)tile * stride;

	// the camera frame of generate_camera_ray, same expressions
	const float distance = length(cam.view);
	const float3 horizontal = normalize(cross(cam.view, cam.up));
	const float3 vertical = normalize(cross(horizontal, cam.view));
	const float3 x_axis = horizontal * (distance * __tanf(cam.fov.x * 0.5f * (PTB_PI / 180.0f)));
	const float3 y_axis = vertical * (distance * __tanf(-cam.fov.y * 0.5f * (PTB_PI / 180.0f)));
	// the tile's rectangle on the canvas plane: pixel centres +- 0.5 of jitter, + 1/16 pixel of slack for the rounding of the
	// generator's own arithmetic
	const float slack = 0.5f + 0.0625f;
	const float px0 = (float)(tx * tile_w) - slack, px1 = (float)min(tx * tile_w + tile_w - 1, width - 1) + slack;
	const float py0 = (float)(ty * tile_h) - slack, py1 = (float)min(ty * tile_h + tile_h - 1, height - 1) + slack;
	const float nx0 = (px0 / (cam.resolution.x - 1.0f)) * 2.0f - 1.0f, nx1 = (px1 / (cam.resolution.x - 1.0f)) * 2.0f - 1.0f;
	const float ny0 = (py0 / (cam.resolution.y - 1.0f)) * 2.0f - 1.0f, ny1 = (py1 / (cam.resolution.y - 1.0f)) * 2.0f - 1.0f;
	float3 c[4];
	c[0] = cam.view + nx0 * x_axis + ny0 * y_axis;
	c[1] = cam.view + nx1 * x_axis + ny0 * y_axis;
	c[2] = cam.view + nx1 * x_axis + ny1 * y_axis;
	c[3] = cam.view + nx0 * x_axis + ny1 * y_axis;
	const float3 centre = cam.view + (0.5f * (nx0 + nx1)) * x_axis + (0.5f * (ny0 + ny1)) * y_axis;

	EntryShaft s;
	s.eye = cam.eye;
	s.w = cam.view * (1.0f / distance);
	s.lens = cam.aperture_radius > 0.00001f ? cam.aperture_radius : 0.0f;
	float cos_min = 1.0f;
#pragma unroll
	for (int k = 0; k < 4; k++)
	{
		float3 n = normalize(cross(c[k], c[(k + 1) & 3]));
		if (dot(n, centre) < 0.0f) n = n * -1.0f;
		s.n[k] = n;
		cos_min = fminf(cos_min, dot(normalize(c[k]), s.w));
	}
	s.focal_hi = fmaxf(cam.focal_distance, 1e-20f) * 1.0001f;
	s.focal_lo = fmaxf(cam.focal_distance * cos_min * 0.9999f, 1e-20f);

	int ref[PTB_ENTRY_STRIDE];
	float tn[PTB_ENTRY_STRIDE];
	unsigned frozen = 0u;      // bit i: entry i is final
	int n = 0;
	k_max = max(1, min(k_max, min(stride, PTB_ENTRY_STRIDE) - 1));
	// a degenerate frame (zero view / up parallel to view) makes every test pass: the list stays at the root, which is always valid
	if (sc.n_triangles > 0)
	{
		ref[0] = sc.root_ref; tn[0] = 0.0f; n = 1;
		if (sc.root_ref < 0) frozen = 1u;
	}
	for (int iter = 0; iter < 256; iter++)
	{
		// the front-most sub-tree that may still be opened
		int pick = -1;
		float front = CUDART_INF_F;
		for (int i = 0; i < n; i++)
			if (!((frozen >> i) & 1u) && tn[i] < front) { front = tn[i]; pick = i; }
		if (pick < 0) break;
		const float4* np = sc.bvh_nodes + (size_t)ref[pick] * 4;
		const float4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
		const int child0 = __float_as_int(n3.x), child1 = __float_as_int(n3.y);
		float t0 = 0.0f, t1 = 0.0f;
		const bool ok0 = entry_box_overlaps(s, make_float3(n0.x, n0.z, n2.x), make_float3(n0.y, n0.w, n2.y), t0);
		const bool ok1 = entry_box_overlaps(s, make_float3(n1.x, n1.z, n2.z), make_float3(n1.y, n1.w, n2.w), t1);
		if (!ok0 && !ok1)
		{
			// nothing of this sub-tree is inside the shaft: drop it (last entry moves into the hole)
			n--;
			ref[pick] = ref[n]; tn[pick] = tn[n];
			frozen = (frozen & ~(1u << pick)) | (((frozen >> n) & 1u) << pick);
			frozen &= ~(1u << n);
			continue;
		}
		// a leaf never becomes an entry: the box test that lets a ray skip it sits in its parent
		if ((ok0 && child0 < 0) || (ok1 && child1 < 0)) { frozen |= 1u << pick; continue; }
		if (ok0 != ok1)
		{
			ref[pick] = ok0 ? child0 : child1; tn[pick] = fmaxf(tn[pick], ok0 ? t0 : t1);   // free descent
			continue;
		}
		if (n >= k_max) { frozen |= 1u << pick; continue; }
		const float parent_t = tn[pick];
		ref[pick] = child0; tn[pick] = fmaxf(parent_t, t0);
		ref[n] = child1; tn[n] = fmaxf(parent_t, t1);
		n++;
	}
	// ascending lower bounds: a ray stops at the first entry beyond its current hit
	for (int i = 1; i < n; i++)
	{
		const int r_i = ref[i]; const float t_i = tn[i];
		int j = i - 1;
		while (j >= 0 && tn[j] > t_i) { ref[j + 1] = ref[j]; tn[j + 1] = tn[j]; j--; }
		ref[j + 1] = r_i; tn[j + 1] = t_i;
	}
	for (int i = 0; i < n; i++) out[i] = make_int2(ref[i], __float_as_int(tn[i]));
	out[n] = make_int2(PTB_ENTRY_END, __float_as_int(CUDART_INF_F));
}

// Numbers the tiles whose entry cut is not empty, in tile (= queue) order: rank[t] = index among them, -1 for an empty cut; *n_nonempty =
// how many.  One block; k_generate<.., SKY> (kernels_generate.cuh) queues only their rays and keeps the queue in tile order without atomics.
__global__ void __launch_bounds__(1024) k_tile_rank(const int2* __restrict__ cuts, int stride, int n_tiles, int* __restrict__ rank, int* __restrict__ n_nonempty)
{
	__shared__ int s_sum[1024];
	const int per = (n_tiles + 1023) / 1024;
	const int begin = min(n_tiles, (int)threadIdx.x * per), end = min(n_tiles, begin + per);
	int mine = 0;
	for (int t = begin; t < end; t++) mine += cuts[(size_t)t * stride].x != PTB_ENTRY_END ? 1 : 0;
	s_sum[threadIdx.x] = mine;
	__syncthreads();
	// inclusive scan over the 1024 partial sums
	for (int off = 1; off < 1024; off <<= 1)
	{
		const int v = threadIdx.x >= off ? s_sum[threadIdx.x - off] : 0;
		__syncthreads();
		s_sum[threadIdx.x] += v;
		__syncthreads();
	}
	int next = s_sum[threadIdx.x] - mine;
	for (int t = begin; t < end; t++) rank[t] = cuts[(size_t)t * stride].x != PTB_ENTRY_END ? next++ : -1;
	if (threadIdx.x == 1023) *n_nonempty = s_sum[1023];
}

// ------------------------------------------------------------------------------------------
// Leaf starts for bounce rays (k_extend_persistent<.., UPWALK>, kernels_extend.cuh).
// A ray that leaves a triangle spends ~18 of its ~22 node visits walking from the root down to the leaf it starts in: at every level
// the child that contains the origin is "hit" trivially and the only question is whether the SIBLING is hit too.  With one record per
// child slot — the sibling's box, the sibling's reference and the slot of the node in its own parent — the ray walks UP from its leaf
// instead: one 32-byte record and one box test per level (a top-down step loads 64 bytes and tests two boxes), the siblings it hits
// go on the traversal stack (deepest = nearest first) and the search proceeds as usual from the leaf's own triangles.  Every sub-tree
// of the scene is either the start leaf or the sibling of one of its ancestors, so the search is as exhaustive as one from the root.
// k_up_level builds the records top-down, one launch per tree level (unused pool slots of the node array are never touched).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_up_level(const float4* __restrict__ nodes, const float4* __restrict__ tri_isect, const int2* __restrict__ frontier_in,
	int2* __restrict__ frontier_out, int* __restrict__ level_counts, int level, int capacity, float4* __restrict__ up, int* __restrict__ tri_slot, int n_tris)
{
	const int count = min(level_counts[level], capacity);
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
	{
		const int2 f = frontier_in[i];         // node index, the node's slot in its parent
		const int n = f.x;
		const float4* np = nodes + (size_t)n * 4;
		const float4 n0 = np[0], n1 = np[1], n2 = np[2], n3 = np[3];
		const int child[2] = { __float_as_int(n3.x), __float_as_int(n3.y) };
		// slot of child 0: its sibling is child 1, and vice versa
		up[((size_t)n * 2 + 0) * 2 + 0] = make_float4(n1.x, n1.z, n2.z, __int_as_float(child[1]));
		up[((size_t)n * 2 + 0) * 2 + 1] = make_float4(n1.y, n1.w, n2.w, __int_as_float(f.y));
		up[((size_t)n * 2 + 1) * 2 + 0] = make_float4(n0.x, n0.z, n2.x, __int_as_float(child[0]));
		up[((size_t)n * 2 + 1) * 2 + 1] = make_float4(n0.y, n0.w, n2.y, __int_as_float(f.y));
		for (int s = 0; s < 2; s++)
		{
			const int c = child[s];
			if (c >= 0)
			{
				const int pos = atomicAdd(&level_counts[level + 1], 1);
				if (pos < capacity) frontier_out[pos] = make_int2(c, n * 2 + s);
			}
			else if (c != PTB_ENTRY_END)
			{
				const int ref = ~c, first = ref >> 3, cnt = (ref & 7) + 1;
				for (int k = 0; k < cnt; k++)
				{
					const int gid = __float_as_int(tri_isect[(size_t)(first + k) * 3].w);
					if (gid >= 0 && gid < n_tris) tri_slot[gid] = n * 2 + s;
				}
			}
		}
	}
}

// Two levels per record: up2[slot] = { A.lo.xyz, A.ref | A.hi.xyz, B.ref | B.lo.xyz, slot two levels up | B.hi.xyz, - } with A the sibling of
// `slot` and B the sibling of its parent (B.ref = PTB_ENTRY_END when `slot` hangs off the root).  Halves the chain of dependent loads of
// the walk.  `up` must have been filled with 0xff before k_up_level ran, so unused pool slots read "no parent".
__global__ void __launch_bounds__(128) k_up_pair(const float4* __restrict__ up, float4* __restrict__ up2, int n_slots)
{
	const int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n_slots) return;
	const float4 a_lo = up[(size_t)s * 2], a_hi = up[(size_t)s * 2 + 1];
	const int p = __float_as_int(a_hi.w);
	float4 b_lo = make_float4(1.0f, 1.0f, 1.0f, __int_as_float(-1)), b_hi = make_float4(-1.0f, -1.0f, -1.0f, 0.0f);
	int b_ref = PTB_ENTRY_END;
	if (p >= 0 && p < n_slots)
	{
		const float4 q_lo = up[(size_t)p * 2], q_hi = up[(size_t)p * 2 + 1];
		b_ref = __float_as_int(q_lo.w);
		b_lo = make_float4(q_lo.x, q_lo.y, q_lo.z, q_hi.w);      // .w: the slot two levels up
		b_hi = make_float4(q_hi.x, q_hi.y, q_hi.z, 0.0f);
	}
	up2[(size_t)s * 4 + 0] = a_lo;
	up2[(size_t)s * 4 + 1] = make_float4(a_hi.x, a_hi.y, a_hi.z, __int_as_float(b_ref));
	up2[(size_t)s * 4 + 2] = b_lo;
	up2[(size_t)s * 4 + 3] = b_hi;
}

} // namespace ptb
