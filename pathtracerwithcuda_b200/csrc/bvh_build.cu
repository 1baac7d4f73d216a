// GPU binned-SAH BVH builder (see bvh_gpu.h).
//
// Top-down, 16 bins per axis, same split rule and float arithmetic as bvh_host.cpp:
//   phase 1 (large nodes, > kSmall = 128 triangles): level-synchronous.  Every level runs
//       k_plan (chunk list) -> k_init_bins -> k_bin (shared-memory bins per 2048-triangle chunk, flushed
//       with global atomics) -> k_split (one warp per node: SAH sweep over 45 candidate planes, writes
//       the 64-byte traversal node, emits child tasks, scans the chunks' left counts) -> k_partition
//       (stable out-of-place partition of the triangle-id array, ping-pong buffers);
//       kSmall = 128 / 64 threads measured best: c2 2.3 ms, c4 7.8 ms (1024 / 128: 9.1 and 16.4 ms — too few, too long blocks)
//   phase 2 (sub-trees of <= kSmall triangles): ONE thread block finishes the whole sub-tree in shared
//       memory (ids, boxes, bins, an explicit stack; smaller child first so the stack stays <= log2 n),
//       blocks pull sub-trees from a queue with an atomic cursor.
// The node array is written in its final traversal layout as the tree is built: a node learns its
// children's boxes from its own bins, so no bottom-up pass is needed, and children patch their
// reference (inner index or leaf range) into the parent record when they are decided.
#include "bvh_gpu.h"
#include <math_constants.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

namespace ptb
{

namespace
{

constexpr int kBins = 16;
constexpr int kBinWords = 13;                           // count | box lo xyz | box hi xyz | centroid lo xyz | centroid hi xyz
constexpr int kTaskBinWords = 3 * kBins * kBinWords;    // 624 words per node
#ifndef PTB_BUILD_SMALL
#define PTB_BUILD_SMALL 128
#endif
#ifndef PTB_BUILD_SMALL_THREADS
#define PTB_BUILD_SMALL_THREADS 64
#endif
constexpr int kSmall = PTB_BUILD_SMALL;                  // sub-trees of <= kSmall triangles are finished by one block
constexpr int kChunk = 2048;                            // triangles per block-chunk in phase 1
constexpr int kThreads = 256;
constexpr int kItems = kChunk / kThreads;
constexpr int kSmallThreads = PTB_BUILD_SMALL_THREADS;
constexpr int kDepthLimit = 40;                         // below this depth splits fall back to halving by index (stack bound)
constexpr int kPool = 32;                               // node indices a phase-2 block reserves per atomic

enum { KIND_LEAF = 0, KIND_SAH = 1, KIND_MEDIAN = 2 };

struct BuildTask
{
	int begin, end;     // slot range in the triangle-id array
	int parent;         // node index of the parent (-1: root)
	int slot;           // which child of the parent (0/1)
	int depth;
	int buf;            // which ping-pong id buffer holds the range
	float blo[3], bhi[3], clo[3], chi[3];   // bounds, centroid bounds
};

struct SplitInfo
{
	int axis;           // -1: halve by index
	int bin;
	int nl;
	int pad;
	float lo, scale;
};

struct Counters
{
	int n_next, n_small, n_nodes, n_chunks, max_depth, overflow, small_cursor, root_ref;
	int root_bounds[12];
};

struct Choice
{
	int kind, axis, bin, nl;
	float l[12], r[12];   // per child: box lo, box hi, centroid lo, centroid hi
};

struct BuildArgs
{
	const float* tris24;
	float4* plo; float4* phi;      // per triangle id: box
	int* idx[2];                   // ping-pong triangle ids by slot
	int* idx_final;
	BuildTask* tasks[2];
	BuildTask* small_tasks;
	SplitInfo* splits;
	int* bins;
	int* task_chunk_begin;
	int* chunk_task;
	int* chunk_counts;             // [chunk][3][16]
	int* chunk_left;               // left elements of the task before this chunk
	float4* nodes;
	Counters* counters;
	int n, max_leaf;
	float intersect_cost;          // SAH: cost of one triangle test relative to one node visit (bvh_host.cpp: 1.5)
	int cap_tasks, cap_small, cap_chunks, cap_nodes;
};

__device__ __forceinline__ int enc(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float dec(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
#define PTB_ENC_POS_INF 0x7f800000
#define PTB_ENC_NEG_INF ((int)0x807fffff)

__device__ __forceinline__ float area_half(const float* lo, const float* hi)
{
	const float dx = __fsub_rn(hi[0], lo[0]), dy = __fsub_rn(hi[1], lo[1]), dz = __fsub_rn(hi[2], lo[2]);
	if (dx < 0.0f || dy < 0.0f || dz < 0.0f) return 0.0f;
	return __fadd_rn(__fadd_rn(__fmul_rn(dx, dy), __fmul_rn(dy, dz)), __fmul_rn(dz, dx));
}

__device__ __forceinline__ int bin_of(float c, float lo, float scale)
{
	int b = (int)__fmul_rn(__fsub_rn(c, lo), scale);
	return b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
}

__device__ __forceinline__ float centroid_of(float lo, float hi) { return __fmul_rn(0.5f, __fadd_rn(lo, hi)); }

__device__ __forceinline__ float pad_down(float v) { return __fsub_rn(v, __fadd_rn(__fmul_rn(fabsf(v), 4.76837158e-7f), 1e-30f)); }
__device__ __forceinline__ float pad_up(float v) { return __fadd_rn(v, __fadd_rn(__fmul_rn(fabsf(v), 4.76837158e-7f), 1e-30f)); }

__device__ __forceinline__ void bins_init(int* bins, int i)
{
	const int w = i % kBinWords;
	bins[i] = w == 0 ? 0 : ((w <= 3 || (w >= 7 && w <= 9)) ? PTB_ENC_POS_INF : PTB_ENC_NEG_INF);
}

// accumulate one triangle into the three axes' bins (shared or global memory)
__device__ __forceinline__ void bins_add(int* bins, const float* lo, const float* hi, const float* clo, const float* chi)
{
	float c[3];
#pragma unroll
	for (int a = 0; a < 3; a++) c[a] = centroid_of(lo[a], hi[a]);
#pragma unroll
	for (int a = 0; a < 3; a++)
	{
		const float extent = __fsub_rn(chi[a], clo[a]);
		if (!(extent > 0.0f)) continue;
		const int b = bin_of(c[a], clo[a], __fdiv_rn((float)kBins, extent));
		int* p = bins + (a * kBins + b) * kBinWords;
		atomicAdd(p, 1);
#pragma unroll
		for (int k = 0; k < 3; k++)
		{
			atomicMin(p + 1 + k, enc(lo[k]));
			atomicMax(p + 4 + k, enc(hi[k]));
			atomicMin(p + 7 + k, enc(c[k]));
			atomicMax(p + 10 + k, enc(c[k]));
		}
	}
}

// One warp evaluates the 3 x 15 candidate planes of a node from its bins and decides leaf / split.
// Mirrors Builder::build in bvh_host.cpp (first minimum in axis-major, bin-minor order wins).
__device__ Choice warp_choose(const int* bins, int count, const float* blo, const float* bhi, const float* clo, const float* chi, int depth, int max_leaf, float kIntersectCost)
{
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	Choice ch;
	ch.kind = KIND_LEAF; ch.axis = -1; ch.bin = 0; ch.nl = 0;
#pragma unroll
	for (int i = 0; i < 12; i++) { ch.l[i] = 0.0f; ch.r[i] = 0.0f; }
	if (count <= 1) return ch;

	// 45 candidate planes over 32 lanes: lane evaluates candidates `lane` and `lane + 32`, keeping its better one
	float L[12], R[12];
	int cl = 0, cr = 0, cand_best = lane;
	float cost = CUDART_INF_F;
	for (int cand = lane; cand < 45; cand += 32)
	{
		const int a = cand / 15, b = cand - a * 15;
		if (!(__fsub_rn(chi[a], clo[a]) > 0.0f)) continue;
		float l[12], r[12];
#pragma unroll
		for (int i = 0; i < 3; i++)
		{
			l[i] = r[i] = l[6 + i] = r[6 + i] = CUDART_INF_F;
			l[3 + i] = r[3 + i] = l[9 + i] = r[9 + i] = -CUDART_INF_F;
		}
		int nl = 0, nr = 0;
		for (int k = 0; k < kBins; k++)
		{
			const int* p = bins + (a * kBins + k) * kBinWords;
			const int c = p[0];
			if (c == 0) continue;
			if (k <= b)
			{
				nl += c;
#pragma unroll
				for (int i = 0; i < 3; i++)
				{
					l[i] = fminf(l[i], dec(p[1 + i])); l[3 + i] = fmaxf(l[3 + i], dec(p[4 + i]));
					l[6 + i] = fminf(l[6 + i], dec(p[7 + i])); l[9 + i] = fmaxf(l[9 + i], dec(p[10 + i]));
				}
			}
			else
			{
				nr += c;
#pragma unroll
				for (int i = 0; i < 3; i++)
				{
					r[i] = fminf(r[i], dec(p[1 + i])); r[3 + i] = fmaxf(r[3 + i], dec(p[4 + i]));
					r[6 + i] = fminf(r[6 + i], dec(p[7 + i])); r[9 + i] = fmaxf(r[9 + i], dec(p[10 + i]));
				}
			}
		}
		if (nl > 0 && nr > 0)
		{
			const float c = __fadd_rn(__fmul_rn(area_half(l, l + 3), (float)nl), __fmul_rn(area_half(r, r + 3), (float)nr));
			if (c < cost)   // strict: the lower candidate index wins ties, like the host's sweep order
			{
				cost = c; cand_best = cand; cl = nl; cr = nr;
#pragma unroll
				for (int i = 0; i < 12; i++) { L[i] = l[i]; R[i] = r[i]; }
			}
		}
	}
	(void)cr;
	float best = cost;
	int who = cand_best;
	for (int off = 16; off > 0; off >>= 1)
	{
		const float oc = __shfl_down_sync(FULL, best, off);
		const int ow = __shfl_down_sync(FULL, who, off);
		if (oc < best || (oc == best && ow < who)) { best = oc; who = ow; }
	}
	best = __shfl_sync(FULL, best, 0);
	who = __shfl_sync(FULL, who, 0);   // winning candidate index (axis * 15 + bin); it lives in lane who & 31
	const bool have = best < CUDART_INF_F;

	const float parent_area = area_half(blo, bhi);
	const float leaf_cost = __fmul_rn(kIntersectCost, (float)count);
	const float split_cost = parent_area > 0.0f ? __fadd_rn(1.0f, __fdiv_rn(__fmul_rn(kIntersectCost, best), parent_area)) : 0.0f;
	int kind;
	if (!have) kind = count <= max_leaf ? KIND_LEAF : KIND_MEDIAN;
	else if (count <= max_leaf && leaf_cost <= split_cost) kind = KIND_LEAF;
	else kind = KIND_SAH;
	if (kind == KIND_SAH && depth >= kDepthLimit) kind = count <= max_leaf ? KIND_LEAF : KIND_MEDIAN;
	ch.kind = kind;
	if (kind == KIND_SAH)
	{
		ch.axis = who / 15; ch.bin = who - (who / 15) * 15;
		ch.nl = __shfl_sync(FULL, cl, who & 31);
#pragma unroll
		for (int i = 0; i < 12; i++) { ch.l[i] = __shfl_sync(FULL, L[i], who & 31); ch.r[i] = __shfl_sync(FULL, R[i], who & 31); }
	}
	else if (kind == KIND_MEDIAN)
	{
		ch.axis = -1; ch.nl = count / 2;
#pragma unroll
		for (int i = 0; i < 3; i++)
		{
			ch.l[i] = ch.r[i] = blo[i]; ch.l[3 + i] = ch.r[3 + i] = bhi[i];
			ch.l[6 + i] = ch.r[6 + i] = clo[i]; ch.l[9 + i] = ch.r[9 + i] = chi[i];
		}
	}
	return ch;
}

__device__ __forceinline__ void write_inner_node(float4* nodes, int index, const float* l, const float* r)
{
	float4* n = nodes + (size_t)index * 4;
	n[0] = make_float4(pad_down(l[0]), pad_up(l[3]), pad_down(l[1]), pad_up(l[4]));
	n[1] = make_float4(pad_down(r[0]), pad_up(r[3]), pad_down(r[1]), pad_up(r[4]));
	n[2] = make_float4(pad_down(l[2]), pad_up(l[5]), pad_down(r[2]), pad_up(r[5]));
	n[3] = make_float4(__int_as_float(~0), __int_as_float(~0), 0.0f, 0.0f);
}

__device__ __forceinline__ void patch_parent(const BuildArgs& A, int parent, int slot, int ref)
{
	if (parent < 0) A.counters->root_ref = ref;
	else reinterpret_cast<int*>(A.nodes)[(size_t)parent * 16 + 12 + slot] = ref;
}

// ------------------------------------------------------------------------------------------
// setup: per-triangle boxes, identity order, scene bounds
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_setup_prims(BuildArgs A)
{
	float lo[3] = { CUDART_INF_F, CUDART_INF_F, CUDART_INF_F }, hi[3] = { -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F };
	float clo[3] = { CUDART_INF_F, CUDART_INF_F, CUDART_INF_F }, chi[3] = { -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F };
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += gridDim.x * blockDim.x)
	{
		const float* t = A.tris24 + (size_t)i * 24;
		float l[3], h[3];
#pragma unroll
		for (int a = 0; a < 3; a++)
		{
			const float v0 = t[a], v1 = t[3 + a], v2 = t[6 + a];
			l[a] = fminf(v0, fminf(v1, v2));
			h[a] = fmaxf(v0, fmaxf(v1, v2));
			const float c = centroid_of(l[a], h[a]);
			lo[a] = fminf(lo[a], l[a]); hi[a] = fmaxf(hi[a], h[a]);
			clo[a] = fminf(clo[a], c); chi[a] = fmaxf(chi[a], c);
		}
		A.plo[i] = make_float4(l[0], l[1], l[2], 0.0f);
		A.phi[i] = make_float4(h[0], h[1], h[2], 0.0f);
		A.idx[0][i] = i;
	}
	int* rb = A.counters->root_bounds;
#pragma unroll
	for (int a = 0; a < 3; a++)
	{
		for (int off = 16; off > 0; off >>= 1)
		{
			lo[a] = fminf(lo[a], __shfl_down_sync(0xffffffffu, lo[a], off)); hi[a] = fmaxf(hi[a], __shfl_down_sync(0xffffffffu, hi[a], off));
			clo[a] = fminf(clo[a], __shfl_down_sync(0xffffffffu, clo[a], off)); chi[a] = fmaxf(chi[a], __shfl_down_sync(0xffffffffu, chi[a], off));
		}
		if ((threadIdx.x & 31) == 0)
		{
			atomicMin(rb + a, enc(lo[a])); atomicMax(rb + 3 + a, enc(hi[a]));
			atomicMin(rb + 6 + a, enc(clo[a])); atomicMax(rb + 9 + a, enc(chi[a]));
		}
	}
}

__global__ void k_init_counters(Counters* c)
{
	c->n_next = 0; c->n_small = 0; c->n_nodes = 0; c->n_chunks = 0; c->max_depth = 0; c->overflow = 0; c->small_cursor = 0; c->root_ref = 0;
	for (int i = 0; i < 12; i++) c->root_bounds[i] = (i < 3 || (i >= 6 && i < 9)) ? PTB_ENC_POS_INF : PTB_ENC_NEG_INF;
}

__global__ void k_make_root(BuildArgs A, int to_small)
{
	BuildTask t;
	t.begin = 0; t.end = A.n; t.parent = -1; t.slot = 0; t.depth = 0; t.buf = 0;
	const int* rb = A.counters->root_bounds;
	for (int a = 0; a < 3; a++) { t.blo[a] = dec(rb[a]); t.bhi[a] = dec(rb[3 + a]); t.clo[a] = dec(rb[6 + a]); t.chi[a] = dec(rb[9 + a]); }
	if (to_small) { A.small_tasks[0] = t; A.counters->n_small = 1; }
	else A.tasks[0][0] = t;
}

// ------------------------------------------------------------------------------------------
// phase 1
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_plan(BuildArgs A, const BuildTask* __restrict__ tasks, int n_tasks)
{
	__shared__ int s_warp[32];
	__shared__ int s_total;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int per = (n_tasks + 1023) / 1024;
	const int t0 = min(n_tasks, tid * per), t1 = min(n_tasks, t0 + per);
	int local = 0;
	for (int t = t0; t < t1; t++) local += (tasks[t].end - tasks[t].begin + kChunk - 1) / kChunk;
	int incl = local;
	for (int off = 1; off < 32; off <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += v; }
	if (lane == 31) s_warp[warp] = incl;
	__syncthreads();
	if (warp == 0)
	{
		int w = s_warp[lane], wi = w;
		for (int off = 1; off < 32; off <<= 1) { int v = __shfl_up_sync(0xffffffffu, wi, off); if (lane >= off) wi += v; }
		s_warp[lane] = wi - w;
		if (lane == 31) s_total = wi;
	}
	__syncthreads();
	int running = s_warp[warp] + incl - local;
	for (int t = t0; t < t1; t++)
	{
		const int nc = (tasks[t].end - tasks[t].begin + kChunk - 1) / kChunk;
		A.task_chunk_begin[t] = running;
		for (int c = 0; c < nc; c++) if (running + c < A.cap_chunks) A.chunk_task[running + c] = t;
		running += nc;
	}
	if (tid == 0)
	{
		A.task_chunk_begin[n_tasks] = s_total;
		A.counters->n_chunks = min(s_total, A.cap_chunks);
		if (s_total > A.cap_chunks) A.counters->overflow = 1;
		A.counters->n_next = 0;
	}
}

__global__ void __launch_bounds__(kThreads) k_init_bins(int* bins, int n_words)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += gridDim.x * blockDim.x) bins_init(bins, i);
}

__global__ void __launch_bounds__(kThreads) k_bin(BuildArgs A, const BuildTask* __restrict__ tasks)
{
	__shared__ int s_bins[kTaskBinWords];
	__shared__ float s_cb[6];
	const int n_chunks = A.counters->n_chunks;
	for (int c = blockIdx.x; c < n_chunks; c += gridDim.x)
	{
		const int ti = A.chunk_task[c];
		const BuildTask* t = tasks + ti;
		const int cb = t->begin + (c - A.task_chunk_begin[ti]) * kChunk;
		const int ce = min(t->end, cb + kChunk);
		for (int i = threadIdx.x; i < kTaskBinWords; i += kThreads) bins_init(s_bins, i);
		if (threadIdx.x < 3) { s_cb[threadIdx.x] = t->clo[threadIdx.x]; s_cb[3 + threadIdx.x] = t->chi[threadIdx.x]; }
		__syncthreads();
		const int* idx = A.idx[t->buf];
		for (int s = cb + threadIdx.x; s < ce; s += kThreads)
		{
			const int id = idx[s];
			const float4 l4 = A.plo[id], h4 = A.phi[id];
			const float lo[3] = { l4.x, l4.y, l4.z }, hi[3] = { h4.x, h4.y, h4.z };
			bins_add(s_bins, lo, hi, s_cb, s_cb + 3);
		}
		__syncthreads();
		int* g = A.bins + (size_t)ti * kTaskBinWords;
		for (int i = threadIdx.x; i < kTaskBinWords; i += kThreads)
		{
			const int w = i % kBinWords, v = s_bins[i];
			if (w == 0)
			{
				A.chunk_counts[(size_t)c * 48 + i / kBinWords] = v;
				if (v) atomicAdd(g + i, v);
			}
			else if (w <= 3 || (w >= 7 && w <= 9)) { if (v != PTB_ENC_POS_INF) atomicMin(g + i, v); }
			else { if (v != PTB_ENC_NEG_INF) atomicMax(g + i, v); }
		}
		__syncthreads();
	}
}

__global__ void __launch_bounds__(128) k_split(BuildArgs A, const BuildTask* __restrict__ tasks, BuildTask* __restrict__ next, int n_tasks)
{
	const int lane = threadIdx.x & 31;
	const int ti = blockIdx.x * 4 + (threadIdx.x >> 5);
	if (ti >= n_tasks) return;
	const BuildTask* t = tasks + ti;
	const int count = t->end - t->begin;
	const Choice ch = warp_choose(A.bins + (size_t)ti * kTaskBinWords, count, t->blo, t->bhi, t->clo, t->chi, t->depth, A.max_leaf, A.intersect_cost);
	// count > kSmall > max_leaf: never a leaf here
	int node = 0;
	if (lane == 0)
	{
		node = atomicAdd(&A.counters->n_nodes, 1);
		if (node >= A.cap_nodes) { A.counters->overflow = 1; node = A.cap_nodes - 1; }
		write_inner_node(A.nodes, node, ch.l, ch.r);
		patch_parent(A, t->parent, t->slot, node);
		atomicMax(&A.counters->max_depth, t->depth + 1);
		SplitInfo si;
		si.axis = ch.axis; si.bin = ch.bin; si.nl = ch.nl; si.pad = 0;
		si.lo = ch.axis >= 0 ? t->clo[ch.axis] : 0.0f;
		si.scale = ch.axis >= 0 ? __fdiv_rn((float)kBins, __fsub_rn(t->chi[ch.axis], t->clo[ch.axis])) : 0.0f;
		A.splits[ti] = si;
		for (int k = 0; k < 2; k++)
		{
			BuildTask ct;
			ct.begin = k ? t->begin + ch.nl : t->begin;
			ct.end = k ? t->end : t->begin + ch.nl;
			ct.parent = node; ct.slot = k; ct.depth = t->depth + 1; ct.buf = t->buf ^ 1;
			const float* src = k ? ch.r : ch.l;
			for (int i = 0; i < 3; i++) { ct.blo[i] = src[i]; ct.bhi[i] = src[3 + i]; ct.clo[i] = src[6 + i]; ct.chi[i] = src[9 + i]; }
			if (ct.end - ct.begin > kSmall)
			{
				const int pos = atomicAdd(&A.counters->n_next, 1);
				if (pos < A.cap_tasks) next[pos] = ct; else A.counters->overflow = 1;
			}
			else
			{
				const int pos = atomicAdd(&A.counters->n_small, 1);
				if (pos < A.cap_small) A.small_tasks[pos] = ct; else A.counters->overflow = 1;
			}
		}
	}
	// left elements of the task that precede each of its chunks (for the stable partition)
	const int c0 = A.task_chunk_begin[ti], c1 = min(A.task_chunk_begin[ti + 1], A.cap_chunks);
	int carry = 0;
	for (int base = c0; base < c1; base += 32)
	{
		const int c = base + lane;
		int v = 0;
		if (c < c1)
		{
			if (ch.axis >= 0)
			{
				const int* cc = A.chunk_counts + (size_t)c * 48 + ch.axis * kBins;
				for (int k = 0; k <= ch.bin; k++) v += cc[k];
			}
			else
			{
				const int start = (c - c0) * kChunk;
				v = max(0, min(kChunk, ch.nl - start));
				v = min(v, count - start);
			}
		}
		int incl = v;
		for (int off = 1; off < 32; off <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += u; }
		if (c < c1) A.chunk_left[c] = carry + incl - v;
		carry += __shfl_sync(0xffffffffu, incl, 31);
	}
}

__global__ void __launch_bounds__(kThreads) k_partition(BuildArgs A, const BuildTask* __restrict__ tasks)
{
	__shared__ int s_warp[kThreads / 32];
	const int n_chunks = A.counters->n_chunks;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int c = blockIdx.x; c < n_chunks; c += gridDim.x)
	{
		const int ti = A.chunk_task[c];
		const BuildTask* t = tasks + ti;
		const SplitInfo si = A.splits[ti];
		const int cb = t->begin + (c - A.task_chunk_begin[ti]) * kChunk;
		const int ce = min(t->end, cb + kChunk);
		const int* src = A.idx[t->buf];
		int* dst = A.idx[t->buf ^ 1];
		// thread handles kItems consecutive slots so the partition is stable
		const int s0 = cb + threadIdx.x * kItems;
		int ids[kItems];
		unsigned left_mask = 0;
		int n_left = 0, n_mine = 0;
#pragma unroll
		for (int k = 0; k < kItems; k++)
		{
			const int s = s0 + k;
			if (s < ce)
			{
				const int id = src[s];
				ids[k] = id;
				bool left;
				if (si.axis >= 0)
				{
					const float4 l4 = A.plo[id], h4 = A.phi[id];
					const float lo = si.axis == 0 ? l4.x : (si.axis == 1 ? l4.y : l4.z), hi = si.axis == 0 ? h4.x : (si.axis == 1 ? h4.y : h4.z);
					left = bin_of(centroid_of(lo, hi), si.lo, si.scale) <= si.bin;
				}
				else left = s < t->begin + si.nl;
				if (left) { left_mask |= 1u << k; n_left++; }
				n_mine++;
			}
		}
		int incl = n_left;
		for (int off = 1; off < 32; off <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += v; }
		if (lane == 31) s_warp[warp] = incl;
		__syncthreads();
		int before = incl - n_left;
		for (int w = 0; w < warp; w++) before += s_warp[w];
		__syncthreads();
		const int chunk_left = A.chunk_left[c];
		int lpos = t->begin + chunk_left + before;
		int rpos = t->begin + si.nl + (cb - t->begin - chunk_left) + (threadIdx.x * kItems - before);
		// (threadIdx.x * kItems) counts this thread's predecessors in the chunk: all of them hold kItems valid slots
#pragma unroll
		for (int k = 0; k < kItems; k++)
		{
			if (k < n_mine)
			{
				if (left_mask & (1u << k)) dst[lpos++] = ids[k];
				else dst[rpos++] = ids[k];
			}
		}
	}
}

// ------------------------------------------------------------------------------------------
// phase 2: one block per sub-tree of <= kSmall triangles
// ------------------------------------------------------------------------------------------
struct StackEntry
{
	int begin, end, parent, slot, depth;
	float blo[3], bhi[3], clo[3], chi[3];
};

__global__ void __launch_bounds__(kSmallThreads) k_build_small(BuildArgs A, int n_small)
{
	__shared__ int s_gid[kSmall];
	__shared__ float s_box[6][kSmall];
	__shared__ int s_perm[kSmall];
	__shared__ int s_tmp[kSmall];
	__shared__ int s_bins[kTaskBinWords];
	__shared__ StackEntry s_stack[24];
	__shared__ StackEntry s_cur, s_child[2];
	__shared__ SplitInfo s_split;
	__shared__ int s_sp, s_has_cur, s_kind, s_task, s_pool_next, s_pool_end, s_max_depth;
	__shared__ int s_warp[kSmallThreads / 32];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) { s_pool_next = 0; s_pool_end = 0; s_max_depth = 0; }
	while (true)
	{
		__syncthreads();
		if (tid == 0) s_task = atomicAdd(&A.counters->small_cursor, 1);
		__syncthreads();
		const int ti = s_task;
		if (ti >= n_small) break;
		const BuildTask* task = A.small_tasks + ti;
		const int base = task->begin;
		const int n = task->end - task->begin;
		const int* src = A.idx[task->buf];
		for (int i = tid; i < n; i += kSmallThreads)
		{
			const int gid = src[base + i];
			s_gid[i] = gid;
			const float4 l4 = A.plo[gid], h4 = A.phi[gid];
			s_box[0][i] = l4.x; s_box[1][i] = l4.y; s_box[2][i] = l4.z;
			s_box[3][i] = h4.x; s_box[4][i] = h4.y; s_box[5][i] = h4.z;
			s_perm[i] = i;
		}
		if (tid == 0)
		{
			StackEntry e;
			e.begin = 0; e.end = n; e.parent = task->parent; e.slot = task->slot; e.depth = task->depth;
			for (int a = 0; a < 3; a++) { e.blo[a] = task->blo[a]; e.bhi[a] = task->bhi[a]; e.clo[a] = task->clo[a]; e.chi[a] = task->chi[a]; }
			s_cur = e;
			s_has_cur = 1;
			s_sp = 0;
		}
		while (true)
		{
			__syncthreads();
			if (tid == 0 && !s_has_cur && s_sp > 0) { s_cur = s_stack[--s_sp]; s_has_cur = 1; }
			__syncthreads();
			if (!s_has_cur) break;
			const int cb = s_cur.begin, ce = s_cur.end;
			const int count = ce - cb;
			for (int i = tid; i < kTaskBinWords; i += kSmallThreads) bins_init(s_bins, i);
			__syncthreads();
			if (count > 1)
			{
				for (int i = cb + tid; i < ce; i += kSmallThreads)
				{
					const int p = s_perm[i];
					const float lo[3] = { s_box[0][p], s_box[1][p], s_box[2][p] }, hi[3] = { s_box[3][p], s_box[4][p], s_box[5][p] };
					bins_add(s_bins, lo, hi, s_cur.clo, s_cur.chi);
				}
			}
			__syncthreads();
			if (warp == 0)
			{
				const Choice ch = warp_choose(s_bins, count, s_cur.blo, s_cur.bhi, s_cur.clo, s_cur.chi, s_cur.depth, A.max_leaf, A.intersect_cost);
				if (lane == 0)
				{
					s_kind = ch.kind;
					if (ch.kind == KIND_LEAF)
					{
						// triangles of a leaf in ascending id order: the tree is then independent of scheduling
						for (int i = cb + 1; i < ce; i++)
						{
							const int p = s_perm[i];
							int j = i - 1;
							while (j >= cb && s_gid[s_perm[j]] > s_gid[p]) { s_perm[j + 1] = s_perm[j]; j--; }
							s_perm[j + 1] = p;
						}
						const int ref = ~(((base + cb) << 3) | (count - 1));
						if (s_cur.parent < 0)
						{
							// single-leaf tree: one inner node whose second child is an empty box
							if (s_pool_next >= s_pool_end) { s_pool_next = atomicAdd(&A.counters->n_nodes, kPool); s_pool_end = s_pool_next + kPool; }
							const int node = min(s_pool_next++, A.cap_nodes - 1);
							const float empty[6] = { 1.0f, 1.0f, 1.0f, -1.0f, -1.0f, -1.0f };
							float own[6];
							for (int a = 0; a < 3; a++) { own[a] = s_cur.blo[a]; own[3 + a] = s_cur.bhi[a]; }
							write_inner_node(A.nodes, node, own, empty);
							float4* nd = A.nodes + (size_t)node * 4;
							// the empty child must stay inverted after padding: overwrite it exactly
							nd[1] = make_float4(1.0f, -1.0f, 1.0f, -1.0f);
							nd[2].z = 1.0f; nd[2].w = -1.0f;
							reinterpret_cast<int*>(A.nodes)[(size_t)node * 16 + 12] = ref;
							A.counters->root_ref = node;
						}
						else patch_parent(A, s_cur.parent, s_cur.slot, ref);
						s_has_cur = 0;
					}
					else
					{
						if (s_pool_next >= s_pool_end) { s_pool_next = atomicAdd(&A.counters->n_nodes, kPool); s_pool_end = s_pool_next + kPool; }
						int node = s_pool_next++;
						if (node >= A.cap_nodes) { A.counters->overflow = 1; node = A.cap_nodes - 1; }
						write_inner_node(A.nodes, node, ch.l, ch.r);
						patch_parent(A, s_cur.parent, s_cur.slot, node);
						s_max_depth = max(s_max_depth, s_cur.depth + 1);
						SplitInfo si;
						si.axis = ch.axis; si.bin = ch.bin; si.nl = ch.nl; si.pad = 0;
						si.lo = ch.axis >= 0 ? s_cur.clo[ch.axis] : 0.0f;
						si.scale = ch.axis >= 0 ? __fdiv_rn((float)kBins, __fsub_rn(s_cur.chi[ch.axis], s_cur.clo[ch.axis])) : 0.0f;
						s_split = si;
						for (int k = 0; k < 2; k++)
						{
							StackEntry e;
							e.begin = k ? cb + ch.nl : cb; e.end = k ? ce : cb + ch.nl;
							e.parent = node; e.slot = k; e.depth = s_cur.depth + 1;
							const float* b = k ? ch.r : ch.l;
							for (int a = 0; a < 3; a++) { e.blo[a] = b[a]; e.bhi[a] = b[3 + a]; e.clo[a] = b[6 + a]; e.chi[a] = b[9 + a]; }
							s_child[k] = e;
						}
					}
				}
			}
			__syncthreads();
			if (s_kind == KIND_LEAF) continue;

			// stable partition of s_perm[cb, ce): each thread owns a consecutive run
			const SplitInfo si = s_split;
			const int per = (count + kSmallThreads - 1) / kSmallThreads;
			const int i0 = min(ce, cb + tid * per), i1 = min(ce, i0 + per);
			int n_left = 0;
			for (int i = i0; i < i1; i++)
			{
				bool left;
				if (si.axis >= 0)
				{
					const int p = s_perm[i];
					left = bin_of(centroid_of(s_box[si.axis][p], s_box[3 + si.axis][p]), si.lo, si.scale) <= si.bin;
				}
				else left = i < cb + si.nl;
				n_left += left ? 1 : 0;
			}
			int incl = n_left;
			for (int off = 1; off < 32; off <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += v; }
			if (lane == 31) s_warp[warp] = incl;
			__syncthreads();
			int before = incl - n_left;
			for (int w = 0; w < warp; w++) before += s_warp[w];
			int lpos = cb + before;
			int rpos = cb + si.nl + (i0 - cb - before);
			for (int i = i0; i < i1; i++)
			{
				const int p = s_perm[i];
				bool left;
				if (si.axis >= 0) left = bin_of(centroid_of(s_box[si.axis][p], s_box[3 + si.axis][p]), si.lo, si.scale) <= si.bin;
				else left = i < cb + si.nl;
				if (left) s_tmp[lpos++] = p; else s_tmp[rpos++] = p;
			}
			__syncthreads();
			for (int i = cb + tid; i < ce; i += kSmallThreads) s_perm[i] = s_tmp[i];
			if (tid == 0)
			{
				// smaller child next, larger child on the stack: depth of the stack <= log2(n)
				const int small = (si.nl <= count - si.nl) ? 0 : 1;
				s_stack[s_sp++] = s_child[small ^ 1];
				s_cur = s_child[small];
				s_has_cur = 1;
			}
		}
		// final order of this sub-tree's triangles
		__syncthreads();
		for (int i = tid; i < n; i += kSmallThreads) A.idx_final[base + i] = s_gid[s_perm[i]];
	}
	if (tid == 0 && s_max_depth > 0) atomicMax(&A.counters->max_depth, s_max_depth);
}

// triangles in leaf order: v0.xyz,id | e1.xyz,0 | e2.xyz,0 (e1/e2 are the subtractions Core/triangle.h:33-34 performs)
__global__ void __launch_bounds__(kThreads) k_emit_tris(const float* __restrict__ tris24, const int* __restrict__ order, float4* __restrict__ out, int n)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
	{
		const int id = order[i];
		const float* t = tris24 + (size_t)id * 24;
		const float v0x = t[0], v0y = t[1], v0z = t[2];
		out[(size_t)i * 3 + 0] = make_float4(v0x, v0y, v0z, __int_as_float(id));
		out[(size_t)i * 3 + 1] = make_float4(__fsub_rn(t[3], v0x), __fsub_rn(t[4], v0y), __fsub_rn(t[5], v0z), 0.0f);
		out[(size_t)i * 3 + 2] = make_float4(__fsub_rn(t[6], v0x), __fsub_rn(t[7], v0y), __fsub_rn(t[8], v0z), 0.0f);
	}
}

__global__ void __launch_bounds__(kThreads) k_pack_shade(const float* __restrict__ tris24, const int* __restrict__ material, float4* __restrict__ out, int n)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
	{
		const float* t = tris24 + (size_t)i * 24;
		out[(size_t)i * 4 + 0] = make_float4(t[9], t[10], t[11], t[12]);
		out[(size_t)i * 4 + 1] = make_float4(t[13], t[14], t[15], t[16]);
		out[(size_t)i * 4 + 2] = make_float4(t[17], t[18], t[19], t[20]);
		out[(size_t)i * 4 + 3] = make_float4(t[21], t[22], t[23], __int_as_float(material[i]));
	}
}

// ------------------------------------------------------------------------------------------
// Collapse of the binary tree into the compressed 8-wide layout (bvh.h layout #2) on the device.
// Same rules as the host WideBuilder (bvh_host.cpp): open the inner child with the largest surface area until 8
// children, greedy octant-slot assignment, per-node power-of-two quantisation grid rounded outwards, inner
// children contiguous from child_base, leaf triangles (<= 3 per slot) contiguous from tri_base.
// One thread per wide node, one launch per level of the wide tree (breadth first).
// ------------------------------------------------------------------------------------------
struct WideItem
{
	int ref2;        // binary node this wide node is made from
	int wide;        // index of the wide node to write
	float lo[3], hi[3];
};

struct WideCounters
{
	int n_out, n_nodes, n_tris, max_depth, overflow;
};

__device__ __forceinline__ void read_binary_children(const float4* nodes2, int ref2, float (&blo)[2][3], float (&bhi)[2][3], int (&refs)[2])
{
	const float4* n = nodes2 + (size_t)ref2 * 4;
	const float4 n0 = n[0], n1 = n[1], n2 = n[2], n3 = n[3];
	blo[0][0] = n0.x; bhi[0][0] = n0.y; blo[0][1] = n0.z; bhi[0][1] = n0.w; blo[0][2] = n2.x; bhi[0][2] = n2.y;
	blo[1][0] = n1.x; bhi[1][0] = n1.y; blo[1][1] = n1.z; bhi[1][1] = n1.w; blo[1][2] = n2.z; bhi[1][2] = n2.w;
	refs[0] = __float_as_int(n3.x); refs[1] = __float_as_int(n3.y);
}

__global__ void __launch_bounds__(128) k_collapse8(const float4* __restrict__ nodes2, const int* __restrict__ prim_order, const float* __restrict__ tris24,
	const WideItem* __restrict__ items_in, int n_in, WideItem* __restrict__ items_out, WideCounters* __restrict__ wc,
	uint32_t* __restrict__ nodes8, float* __restrict__ tris8, int cap_nodes, int cap_tris, int depth)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_in) return;
	const WideItem it = items_in[i];
	int ref[8];
	float clo[8][3], chi[8][3];
	int n = 0;
	{
		float blo[2][3], bhi[2][3];
		int refs[2];
		read_binary_children(nodes2, it.ref2, blo, bhi, refs);
		for (int k = 0; k < 2; k++)
		{
			if (blo[k][0] > bhi[k][0]) continue;    // the empty second child of a single-leaf tree
			ref[n] = refs[k];
			for (int a = 0; a < 3; a++) { clo[n][a] = blo[k][a]; chi[n][a] = bhi[k][a]; }
			n++;
		}
	}
	while (n < 8)
	{
		int best = -1;
		float best_area = -1.0f;
		for (int c = 0; c < n; c++)
		{
			if (ref[c] < 0) continue;
			const float a = area_half(clo[c], chi[c]);
			if (a > best_area) { best_area = a; best = c; }
		}
		if (best < 0) break;
		float blo[2][3], bhi[2][3];
		int refs[2];
		read_binary_children(nodes2, ref[best], blo, bhi, refs);
		ref[best] = refs[0];
		ref[n] = refs[1];
		for (int a = 0; a < 3; a++) { clo[best][a] = blo[0][a]; chi[best][a] = bhi[0][a]; clo[n][a] = blo[1][a]; chi[n][a] = bhi[1][a]; }
		n++;
	}

	// greedy octant slots: slot s prefers the child lying furthest along (s&4 ? +x : -x, s&2 ? +y : -y, s&1 ? +z : -z)
	int slot_child[8];
	{
		float dcen[8][3];
		for (int c = 0; c < n; c++)
			for (int a = 0; a < 3; a++) dcen[c][a] = 0.5f * (clo[c][a] + chi[c][a]) - 0.5f * (it.lo[a] + it.hi[a]);
		unsigned child_done = 0, slot_done = 0;
		for (int s = 0; s < 8; s++) slot_child[s] = -1;
		for (int k = 0; k < n; k++)
		{
			int bc = -1, bs = -1;
			float bestv = -CUDART_INF_F;
			for (int c = 0; c < n; c++)
			{
				if (child_done & (1u << c)) continue;
				for (int s = 0; s < 8; s++)
				{
					if (slot_done & (1u << s)) continue;
					const float v = ((s & 4) ? dcen[c][0] : -dcen[c][0]) + ((s & 2) ? dcen[c][1] : -dcen[c][1]) + ((s & 1) ? dcen[c][2] : -dcen[c][2]);
					if (v > bestv || bc < 0) { bestv = v; bc = c; bs = s; }
				}
			}
			child_done |= 1u << bc; slot_done |= 1u << bs;
			slot_child[bs] = bc;
		}
	}

	// quantisation grid of this node
	float p[3], scale[3];
	uint32_t e[3];
	for (int a = 0; a < 3; a++)
	{
		p[a] = pad_down(it.lo[a]);
		const float extent = __fsub_rn(pad_up(it.hi[a]), p[a]);
		int e2;
		const float m = frexpf(fmaxf(extent, 1e-30f) / 255.0f, &e2);   // m in [0.5, 1)
		int ex = m > 0.5f ? e2 : e2 - 1;
		while (ldexpf(255.0f, ex) < extent) ex++;
		ex = max(-126, min(127, ex));
		e[a] = (uint32_t)(ex + 127);
		scale[a] = ldexpf(1.0f, -ex);
	}

	uint32_t imask = 0;
	int n_inner = 0, n_leaf_tris = 0;
	for (int s = 0; s < 8; s++)
	{
		const int c = slot_child[s];
		if (c < 0) continue;
		if (ref[c] >= 0) { imask |= 1u << s; n_inner++; }
		else n_leaf_tris += min(((~ref[c]) & 7) + 1, 3);
	}
	int child_base = n_inner ? atomicAdd(&wc->n_nodes, n_inner) : 0;
	int tri_base = n_leaf_tris ? atomicAdd(&wc->n_tris, n_leaf_tris) : 0;
	int out_base = n_inner ? atomicAdd(&wc->n_out, n_inner) : 0;
	if (child_base + n_inner > cap_nodes || tri_base + n_leaf_tris > cap_tris) { wc->overflow = 1; return; }
	if (n_inner) atomicMax(&wc->max_depth, depth + 1);

	uint32_t meta[8], qlo[3][8], qhi[3][8];
	int inner_rank = 0, tri_offset = 0;
	for (int s = 0; s < 8; s++)
	{
		meta[s] = 0;
		for (int a = 0; a < 3; a++) { qlo[a][s] = 0; qhi[a][s] = 0; }
		const int c = slot_child[s];
		if (c < 0) continue;
		for (int a = 0; a < 3; a++)
		{
			const float lo = floorf(__fmul_rn(__fsub_rn(pad_down(clo[c][a]), p[a]), scale[a]));
			const float hi = ceilf(__fmul_rn(__fsub_rn(pad_up(chi[c][a]), p[a]), scale[a]));
			qlo[a][s] = (uint32_t)fmaxf(0.0f, fminf(255.0f, lo));
			qhi[a][s] = (uint32_t)fmaxf(0.0f, fminf(255.0f, hi));
		}
		if (ref[c] >= 0)
		{
			meta[s] = (1u << 5) | (24u + (uint32_t)s);
			WideItem o;
			o.ref2 = ref[c]; o.wide = child_base + inner_rank;
			for (int a = 0; a < 3; a++) { o.lo[a] = clo[c][a]; o.hi[a] = chi[c][a]; }
			items_out[out_base + inner_rank] = o;
			inner_rank++;
		}
		else
		{
			const int lr = ~ref[c], first = lr >> 3, count = min((lr & 7) + 1, 3);
			const uint32_t unary = count == 1 ? 1u : (count == 2 ? 3u : 7u);
			meta[s] = (unary << 5) | (uint32_t)tri_offset;
			for (int k = 0; k < count; k++)
			{
				const int id = prim_order[first + k];
				const float* t = tris24 + (size_t)id * 24;
				float* d = tris8 + (size_t)(tri_base + tri_offset + k) * 12;
				const float v0x = t[0], v0y = t[1], v0z = t[2];
				d[0] = v0x; d[1] = v0y; d[2] = v0z; d[3] = __int_as_float(id);
				d[4] = __fsub_rn(t[3], v0x); d[5] = __fsub_rn(t[4], v0y); d[6] = __fsub_rn(t[5], v0z); d[7] = 0.0f;
				d[8] = __fsub_rn(t[6], v0x); d[9] = __fsub_rn(t[7], v0y); d[10] = __fsub_rn(t[8], v0z); d[11] = 0.0f;
			}
			tri_offset += count;
		}
	}
	uint32_t* d = nodes8 + (size_t)it.wide * 20;
	d[0] = __float_as_uint(p[0]); d[1] = __float_as_uint(p[1]); d[2] = __float_as_uint(p[2]);
	d[3] = e[0] | (e[1] << 8) | (e[2] << 16) | (imask << 24);
	d[4] = (uint32_t)child_base;
	d[5] = (uint32_t)tri_base;
	d[6] = meta[0] | (meta[1] << 8) | (meta[2] << 16) | (meta[3] << 24);
	d[7] = meta[4] | (meta[5] << 8) | (meta[6] << 16) | (meta[7] << 24);
	for (int a = 0; a < 3; a++)
	{
		d[8 + a * 4 + 0] = qlo[a][0] | (qlo[a][1] << 8) | (qlo[a][2] << 16) | (qlo[a][3] << 24);
		d[8 + a * 4 + 1] = qlo[a][4] | (qlo[a][5] << 8) | (qlo[a][6] << 16) | (qlo[a][7] << 24);
		d[8 + a * 4 + 2] = qhi[a][0] | (qhi[a][1] << 8) | (qhi[a][2] << 16) | (qhi[a][3] << 24);
		d[8 + a * 4 + 3] = qhi[a][4] | (qhi[a][5] << 8) | (qhi[a][6] << 16) | (qhi[a][7] << 24);
	}
}

__global__ void k_collapse8_root(const float4* __restrict__ nodes2, WideItem* items, WideCounters* wc)
{
	float blo[2][3], bhi[2][3];
	int refs[2];
	read_binary_children(nodes2, 0, blo, bhi, refs);
	WideItem it;
	it.ref2 = 0; it.wide = 0;
	const bool second_empty = blo[1][0] > bhi[1][0];
	for (int a = 0; a < 3; a++)
	{
		it.lo[a] = second_empty ? blo[0][a] : fminf(blo[0][a], blo[1][a]);
		it.hi[a] = second_empty ? bhi[0][a] : fmaxf(bhi[0][a], bhi[1][a]);
	}
	items[0] = it;
	wc->n_out = 0; wc->n_nodes = 1; wc->n_tris = 0; wc->max_depth = 1; wc->overflow = 0;
}

struct DeviceBuffers
{
	std::vector<void*> ptrs;
	~DeviceBuffers() { for (void* p : ptrs) cudaFree(p); }
	template <class T>
	bool alloc(T** out, size_t count)
	{
		void* p = nullptr;
		if (cudaMalloc(&p, std::max<size_t>(count * sizeof(T), 16)) != cudaSuccess) return false;
		ptrs.push_back(p);
		*out = (T*)p;
		return true;
	}
	void release(void* keep) { ptrs.erase(std::remove(ptrs.begin(), ptrs.end(), keep), ptrs.end()); }
};

} // namespace

void pack_tri_shade_gpu(const float* d_tris24, const int* d_material, int n, float4* d_out, cudaStream_t stream)
{
	if (n <= 0) return;
	k_pack_shade<<<std::min((n + kThreads - 1) / kThreads, 148 * 8), kThreads, 0, stream>>>(d_tris24, d_material, d_out, n);
}

int build_bvh2_gpu(const float* d_tris24, int n, int max_leaf_size, float intersect_cost, cudaStream_t stream, GpuBuildOutput& out, std::string& err)
{
	out = GpuBuildOutput();
	if (n <= 0) { err = "build_bvh2_gpu: no triangles"; return 1; }
	if (n > (1 << 27)) { err = "build_bvh2_gpu: more than 2^27 triangles"; return 1; }
	auto fail = [&](const char* what, cudaError_t e) { err = std::string("[Cuda]build_bvh2_gpu: ") + what + ": " + cudaGetErrorString(e); return 1; };

	BuildArgs A;
	A.tris24 = d_tris24;
	A.n = n;
	A.max_leaf = std::max(1, std::min(max_leaf_size, 8));
	A.intersect_cost = intersect_cost;
	A.cap_tasks = n / kSmall + 4;
	A.cap_small = 4096 + n / (kSmall >= 512 ? 64 : 8);
	A.cap_chunks = n / kChunk + A.cap_tasks + 4;
	A.cap_nodes = n + 65536;   // inner nodes <= n - 1, plus the unused tail of each block's last pool (<= 31 x blocks)

	DeviceBuffers scratch, results;
	bool ok = scratch.alloc(&A.plo, n) && scratch.alloc(&A.phi, n) && scratch.alloc(&A.idx[0], n) && scratch.alloc(&A.idx[1], n) &&
		scratch.alloc(&A.tasks[0], A.cap_tasks) && scratch.alloc(&A.tasks[1], A.cap_tasks) && scratch.alloc(&A.small_tasks, A.cap_small) &&
		scratch.alloc(&A.splits, A.cap_tasks) && scratch.alloc(&A.bins, (size_t)A.cap_tasks * kTaskBinWords) &&
		scratch.alloc(&A.task_chunk_begin, A.cap_tasks + 1) && scratch.alloc(&A.chunk_task, A.cap_chunks) &&
		scratch.alloc(&A.chunk_counts, (size_t)A.cap_chunks * 48) && scratch.alloc(&A.chunk_left, A.cap_chunks) && scratch.alloc(&A.counters, 1) &&
		results.alloc(&A.idx_final, n) && results.alloc(&A.nodes, (size_t)A.cap_nodes * 4);
	float4* tri_isect = nullptr;
	ok = ok && results.alloc(&tri_isect, (size_t)n * 3);
	if (!ok) { cudaGetLastError(); err = "[Cuda]build_bvh2_gpu: out of device memory"; return 1; }

	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0, stream);
	const int wide_grid = 148 * 8;
	k_init_counters<<<1, 1, 0, stream>>>(A.counters);
	k_setup_prims<<<std::min((n + kThreads - 1) / kThreads, wide_grid), kThreads, 0, stream>>>(A);
	const bool root_small = n <= kSmall;
	k_make_root<<<1, 1, 0, stream>>>(A, root_small ? 1 : 0);

	Counters hc;
	int n_cur = root_small ? 0 : 1, level = 0;
	while (n_cur > 0)
	{
		const BuildTask* cur = A.tasks[level & 1];
		BuildTask* next = A.tasks[(level + 1) & 1];
		k_plan<<<1, 1024, 0, stream>>>(A, cur, n_cur);
		k_init_bins<<<std::min((n_cur * kTaskBinWords + kThreads - 1) / kThreads, wide_grid), kThreads, 0, stream>>>(A.bins, n_cur * kTaskBinWords);
		k_bin<<<wide_grid, kThreads, 0, stream>>>(A, cur);
		k_split<<<(n_cur + 3) / 4, 128, 0, stream>>>(A, cur, next, n_cur);
		k_partition<<<wide_grid, kThreads, 0, stream>>>(A, cur);
		cudaError_t e = cudaMemcpyAsync(&hc, A.counters, sizeof(Counters), cudaMemcpyDeviceToHost, stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
		if (e != cudaSuccess) { cudaEventDestroy(e0); cudaEventDestroy(e1); return fail("large-node level", e); }
		if (hc.overflow) { cudaEventDestroy(e0); cudaEventDestroy(e1); err = "build_bvh2_gpu: task/chunk/node capacity exceeded (adversarial input)"; return 1; }
		n_cur = hc.n_next;
		level++;
		if (level > 96) { cudaEventDestroy(e0); cudaEventDestroy(e1); err = "build_bvh2_gpu: too many levels"; return 1; }
	}
	if (root_small) hc.n_small = 1;
	k_build_small<<<std::max(1, std::min(hc.n_small, 148 * (kSmall >= 512 ? 5 : 16))), kSmallThreads, 0, stream>>>(A, hc.n_small);
	k_emit_tris<<<std::min((n + kThreads - 1) / kThreads, wide_grid), kThreads, 0, stream>>>(d_tris24, A.idx_final, tri_isect, n);
	cudaEventRecord(e1, stream);
	cudaError_t e = cudaMemcpyAsync(&hc, A.counters, sizeof(Counters), cudaMemcpyDeviceToHost, stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
	if (e == cudaSuccess) e = cudaGetLastError();
	float ms = 0.0f;
	if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	if (e != cudaSuccess) return fail("small-sub-tree phase", e);
	if (hc.overflow || hc.n_nodes > A.cap_nodes) { err = "build_bvh2_gpu: node capacity exceeded (adversarial input)"; return 1; }

	out.nodes = A.nodes; out.tri_isect = tri_isect; out.prim_order = A.idx_final;
	results.release(A.nodes); results.release(tri_isect); results.release(A.idx_final);
	out.n_nodes = hc.n_nodes; out.n_prims = n; out.max_depth = hc.max_depth; out.levels = level; out.small_tasks = hc.n_small; out.build_ms = ms;
	if (hc.root_ref != 0) { err = "build_bvh2_gpu: internal error (root is not node 0)"; cudaFree(out.nodes); cudaFree(out.tri_isect); cudaFree(out.prim_order); out = GpuBuildOutput(); return 1; }
	return 0;
}

int collapse_bvh8_gpu(const GpuBuildOutput& tree, const float* d_tris24, cudaStream_t stream, GpuWideOutput& out, std::string& err)
{
	out = GpuWideOutput();
	const int n = tree.n_prims;
	if (n <= 0 || tree.n_nodes <= 0) { err = "collapse_bvh8_gpu: empty tree"; return 1; }
	const int cap_nodes = tree.n_nodes + 1, cap_tris = n;
	DeviceBuffers scratch, results;
	WideItem* items[2] = { nullptr, nullptr };
	WideCounters* wc = nullptr;
	uint32_t* nodes8 = nullptr;
	float* tris8 = nullptr;
	if (!scratch.alloc(&items[0], cap_nodes) || !scratch.alloc(&items[1], cap_nodes) || !scratch.alloc(&wc, 1) ||
		!results.alloc(&nodes8, (size_t)cap_nodes * 20) || !results.alloc(&tris8, (size_t)cap_tris * 12))
	{
		cudaGetLastError();
		err = "[Cuda]collapse_bvh8_gpu: out of device memory";
		return 1;
	}
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0, stream);
	k_collapse8_root<<<1, 1, 0, stream>>>(tree.nodes, items[0], wc);
	WideCounters hc;
	int n_cur = 1, level = 0;
	cudaError_t e = cudaSuccess;
	while (n_cur > 0 && e == cudaSuccess)
	{
		k_collapse8<<<(n_cur + 127) / 128, 128, 0, stream>>>(tree.nodes, tree.prim_order, d_tris24, items[level & 1], n_cur, items[(level + 1) & 1], wc,
			nodes8, tris8, cap_nodes, cap_tris, level + 1);
		e = cudaMemcpyAsync(&hc, wc, sizeof(hc), cudaMemcpyDeviceToHost, stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
		if (e != cudaSuccess || hc.overflow) break;
		n_cur = hc.n_out;
		level++;
		if (n_cur > 0) e = cudaMemsetAsync(&wc->n_out, 0, sizeof(int), stream);
		if (level > 128) break;
	}
	cudaEventRecord(e1, stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
	if (e == cudaSuccess) e = cudaGetLastError();
	float ms = 0.0f;
	if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	if (e != cudaSuccess) { err = std::string("[Cuda]collapse_bvh8_gpu: ") + cudaGetErrorString(e); return 1; }
	if (hc.overflow || level > 128) { err = "collapse_bvh8_gpu: capacity exceeded"; return 1; }
	out.nodes = reinterpret_cast<float4*>(nodes8); out.tris = reinterpret_cast<float4*>(tris8);
	results.release(nodes8); results.release(tris8);
	out.n_nodes = hc.n_nodes; out.n_tris = hc.n_tris; out.max_depth = hc.max_depth; out.levels = level; out.collapse_ms = ms;
	return 0;
}

} // namespace ptb
