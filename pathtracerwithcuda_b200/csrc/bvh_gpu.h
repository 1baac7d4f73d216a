// GPU binned-SAH builder of the ptb200 acceleration structure (csrc/bvh_build.cu).
//
// Replaces the reference's BVH producers (Bvh/bvh.cpp:185-219 naive, :667-780 Morton CPU,
// :862-1047 "Morton CUDA" — host-heavy even in CUDA mode: AABB loop, box propagation, pointer-tree
// rebuild and two DFS passes are single-threaded; Kernel/bvh_morton_code_kernel.cu:298-346 only
// sorts the codes) with a top-down binned surface-area-heuristic build that runs on the device and
// emits the traversal layout (bvh.h, layout #1: 64-byte binary nodes holding both child boxes,
// triangles in leaf order as v0|e1|e2) directly into device buffers.  The split arithmetic is the
// host builder's (bvh_host.cpp) operation for operation, so both produce the same leaf partition.
#pragma once
#include <cuda_runtime.h>
#include <string>

namespace ptb
{

struct GpuBuildOutput
{
	float4* nodes = nullptr;      // cudaMalloc'ed, 4 x float4 per node (caller frees)
	float4* tri_isect = nullptr;  // cudaMalloc'ed, 3 x float4 per triangle in leaf order (caller frees)
	int* prim_order = nullptr;    // cudaMalloc'ed, triangle id per leaf slot (caller frees)
	int n_nodes = 0;              // node records allocated (the array may contain a few unused pool slots)
	int n_prims = 0;
	int max_depth = 0;            // deepest task level reached
	int levels = 0;               // level-synchronous rounds of the large-node phase
	int small_tasks = 0;          // sub-trees finished by one block in shared memory
	float build_ms = 0.0f;        // CUDA-event time of the whole build on `stream`
};

// d_tris24: n triangles x 24 floats on the DEVICE (v0 v1 v2 n0 n1 n2 uv0 uv1 uv2 — scene.h Triangle).
// Returns 0 on success; on failure `err` says why (capacity overflow on adversarial input, CUDA error)
// and the caller falls back to / reports the host builder.
int build_bvh2_gpu(const float* d_tris24, int n, int max_leaf_size, float intersect_cost, cudaStream_t stream, GpuBuildOutput& out, std::string& err);

struct GpuWideOutput
{
	float4* nodes = nullptr;      // cudaMalloc'ed, 5 x float4 per compressed 8-wide node (bvh.h layout #2; caller frees)
	float4* tris = nullptr;       // cudaMalloc'ed, 3 x float4 per triangle in the wide tree's leaf order (caller frees)
	int n_nodes = 0, n_tris = 0;
	int max_depth = 0;            // depth of the wide tree (traversal stack bound)
	int levels = 0;
	float collapse_ms = 0.0f;
};

// Collapses a binary tree built by build_bvh2_gpu (max_leaf_size <= 3) into the compressed 8-wide layout ON THE DEVICE
// (same rules as the host build_bvh8 in bvh_host.cpp).
int collapse_bvh8_gpu(const GpuBuildOutput& tree, const float* d_tris24, cudaStream_t stream, GpuWideOutput& out, std::string& err);

// packs the shading attributes (DeviceScene::tri_shade: 4 x float4 per triangle by global id) on the device
void pack_tri_shade_gpu(const float* d_tris24, const int* d_material, int n, float4* d_out, cudaStream_t stream);

} // namespace ptb
