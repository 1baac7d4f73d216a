// Acceleration-structure build for the ptb200 extend stage.
//
// Replaces the reference's per-mesh binary trees flattened to a skip-pointer array
// (Bvh/bvh.cpp:185-330,667-780,862-1047; Kernel/bvh_morton_code_kernel.cu:298-346) with ONE tree
// over all meshes' world-space triangles (global triangle indices, same numbering as the
// reference's concatenated triangle array, triangle_mesh.cpp:498-540).
#pragma once
#include <cstdint>
#include <vector>
#include "scene.h"

namespace ptb
{

struct Aabb
{
	float lo[3], hi[3];
};

// Binary tree in host memory; node 0 is the root. Leaves reference a range of `prim_order`.
struct Bvh2Node
{
	Aabb box;
	int left = -1, right = -1;   // inner node
	int first = 0, count = 0;    // leaf if count > 0
};

struct Bvh2
{
	std::vector<Bvh2Node> nodes;
	std::vector<int> prim_order; // permutation of triangle indices
	float sah_cost = 0.0f;
	int max_depth = 0;           // depth of the deepest node (root = 0); the builders keep it below the traversal stack size
};

// Top-down binned surface-area-heuristic build (host, multi-threaded).
void build_bvh2_sah(const TriangleArray& tris, int max_leaf_size, Bvh2& out, float intersect_cost = 1.5f);

// GPU layout #1: binary nodes holding BOTH children's boxes (64 bytes = 4 x 16-byte loads).
//   n[0] = c0.lo.x c0.hi.x c0.lo.y c0.hi.y
//   n[1] = c1.lo.x c1.hi.x c1.lo.y c1.hi.y
//   n[2] = c0.lo.z c0.hi.z c1.lo.z c1.hi.z
//   n[3] = child0, child1 (int bits), 0, 0    child >= 0: node index; child < 0: ~((first<<3)|(count-1))
// Child boxes are padded outwards by a few ulps so a conservative traversal never culls a
// triangle whose exact Moller-Trumbore distance rounds across its box face.
struct GpuBvh2
{
	std::vector<float> nodes;    // 16 floats per node
	std::vector<float> tris;     // 12 floats per triangle in leaf order: v0.xyz,id | e1.xyz,0 | e2.xyz,0
	int root_is_leaf = 0;
	int root_ref = 0;
};

void flatten_bvh2(const Bvh2& bvh, const TriangleArray& tris, GpuBvh2& out);

// GPU layout #2 (default): compressed 8-wide BVH, 80-byte nodes = 5 x 16-byte loads, after
// Ylitie, Karras & Laine 2017.  Child boxes are quantised to 8 bits per plane on a per-node grid
// (origin p, per-axis power-of-two cell 2^e), rounded OUTWARDS, so the node array is ~1/4 the size
// of the binary layout and far more of it stays in L1/L2.
//   n[0] = p.x, p.y, p.z, {e.x, e.y, e.z, imask}                       imask bit s: slot s is an inner node
//   n[1] = child_base (u32), tri_base (u32), meta[0..3], meta[4..7]     meta: inner 0b001xxxxx (x = 24 + slot),
//                                                                              leaf  (unary count) << 5 | first tri offset
//   n[2] = qlo.x[0..3], qlo.x[4..7], qhi.x[0..3], qhi.x[4..7]
//   n[3] = same for y        n[4] = same for z
// Inner children of a node are stored contiguously from child_base in slot order; the triangles of
// its leaf children contiguously from tri_base (<= 24 per node, <= 3 per leaf slot).  Slots are
// assigned so that visiting set bits of (hit mask) from the top, after XOR with the ray octant,
// walks the children front to back.
struct GpuBvh8
{
	std::vector<uint32_t> nodes; // 20 words per node
	std::vector<float> tris;     // 12 floats per triangle: v0.xyz,id | e1.xyz,0 | e2.xyz,0
	int max_depth = 0;
};

void build_bvh8(const Bvh2& bvh, const TriangleArray& tris, GpuBvh8& out);

} // namespace ptb
