// Host binned-SAH builder + flattening to the GPU node layout (see bvh.h).
#include "bvh.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace ptb
{

namespace
{

const int kBins = 16;
const float kTraversalCost = 1.0f;
const int kDepthLimit = 40;   // as in bvh_build.cu

inline void box_reset(Aabb& b)
{
	for (int a = 0; a < 3; a++) { b.lo[a] = std::numeric_limits<float>::infinity(); b.hi[a] = -std::numeric_limits<float>::infinity(); }
}

inline void box_grow(Aabb& b, const Aabb& o)
{
	for (int a = 0; a < 3; a++) { b.lo[a] = std::min(b.lo[a], o.lo[a]); b.hi[a] = std::max(b.hi[a], o.hi[a]); }
}

inline float half_area(const Aabb& b)
{
	float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
	if (dx < 0.0f || dy < 0.0f || dz < 0.0f) return 0.0f;
	return dx * dy + dy * dz + dz * dx;
}

struct PrimRef
{
	Aabb box;
	float c[3];
	int index;
};

struct Builder
{
	std::vector<PrimRef> prims;
	std::vector<Bvh2Node>* nodes;
	int max_leaf;
	int max_depth = 0;
	float kIntersectCost = 1.5f;

	// builds the subtree for prims[first, first+count) into node `node_index`
	void build(int node_index, int first, int count)
	{
		// explicit stack keeps deep, degenerate inputs from overflowing the call stack
		// `depth`: at kDepthLimit and below SAH splits give way to halving by index (the same rule as the device builder,
		// bvh_build.cu kDepthLimit), so the tree is never deeper than kDepthLimit + log2(n) < the 64-entry traversal stack
		struct Item { int node, first, count, depth; };
		std::vector<Item> stack;
		stack.push_back({ node_index, first, count, 0 });
		while (!stack.empty())
		{
			Item it = stack.back();
			stack.pop_back();
			Aabb box, cbox;
			box_reset(box); box_reset(cbox);
			for (int i = it.first; i < it.first + it.count; i++)
			{
				box_grow(box, prims[i].box);
				for (int a = 0; a < 3; a++) { cbox.lo[a] = std::min(cbox.lo[a], prims[i].c[a]); cbox.hi[a] = std::max(cbox.hi[a], prims[i].c[a]); }
			}
			(*nodes)[it.node].box = box;
			max_depth = std::max(max_depth, it.depth);
			if (it.count <= 1) { make_leaf(it.node, it.first, it.count); continue; }
			if (it.depth >= kDepthLimit)
			{
				if (it.count <= max_leaf) { make_leaf(it.node, it.first, it.count); continue; }
				push_children(stack, it.node, it.first, it.first + it.count / 2, it.first + it.count, it.depth + 1);
				continue;
			}

			int best_axis = -1, best_split = -1;
			float best_cost = std::numeric_limits<float>::infinity();
			for (int a = 0; a < 3; a++)
			{
				float extent = cbox.hi[a] - cbox.lo[a];
				if (!(extent > 0.0f)) continue;
				Aabb bin_box[kBins];
				int bin_count[kBins];
				for (int b = 0; b < kBins; b++) { box_reset(bin_box[b]); bin_count[b] = 0; }
				float scale = (float)kBins / extent;
				for (int i = it.first; i < it.first + it.count; i++)
				{
					int b = (int)((prims[i].c[a] - cbox.lo[a]) * scale);
					b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
					bin_count[b]++;
					box_grow(bin_box[b], prims[i].box);
				}
				float right_area[kBins];
				int right_count[kBins];
				Aabb acc; box_reset(acc);
				int cnt = 0;
				for (int b = kBins - 1; b > 0; b--)
				{
					box_grow(acc, bin_box[b]); cnt += bin_count[b];
					right_area[b] = half_area(acc); right_count[b] = cnt;
				}
				box_reset(acc); cnt = 0;
				for (int b = 0; b < kBins - 1; b++)
				{
					box_grow(acc, bin_box[b]); cnt += bin_count[b];
					if (cnt == 0 || right_count[b + 1] == 0) continue;
					float cost = half_area(acc) * (float)cnt + right_area[b + 1] * (float)right_count[b + 1];
					if (cost < best_cost) { best_cost = cost; best_axis = a; best_split = b; }
				}
			}

			float parent_area = half_area(box);
			float leaf_cost = kIntersectCost * (float)it.count;
			float split_cost = parent_area > 0.0f ? kTraversalCost + kIntersectCost * best_cost / parent_area : 0.0f;
			if (best_axis < 0)
			{
				// all centroids coincide: leaf if it fits, else split by index
				if (it.count <= max_leaf) { make_leaf(it.node, it.first, it.count); continue; }
				int mid = it.first + it.count / 2;
				push_children(stack, it.node, it.first, mid, it.first + it.count, it.depth + 1);
				continue;
			}
			if (it.count <= max_leaf && leaf_cost <= split_cost) { make_leaf(it.node, it.first, it.count); continue; }

			float extent = cbox.hi[best_axis] - cbox.lo[best_axis];
			float scale = (float)kBins / extent;
			float lo = cbox.lo[best_axis];
			int axis = best_axis, split = best_split;
			auto mid_it = std::partition(prims.begin() + it.first, prims.begin() + it.first + it.count, [&](const PrimRef& p) {
				int b = (int)((p.c[axis] - lo) * scale);
				b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
				return b <= split;
			});
			int mid = (int)(mid_it - prims.begin());
			if (mid == it.first || mid == it.first + it.count) mid = it.first + it.count / 2;
			push_children(stack, it.node, it.first, mid, it.first + it.count, it.depth + 1);
		}
	}

	template <class Stack>
	void push_children(Stack& stack, int node, int first, int mid, int end, int depth)
	{
		int l = (int)nodes->size();
		nodes->emplace_back();
		nodes->emplace_back();
		(*nodes)[node].left = l;
		(*nodes)[node].right = l + 1;
		(*nodes)[node].count = 0;
		stack.push_back({ l + 1, mid, end - mid, depth });
		stack.push_back({ l, first, mid - first, depth });
	}

	void make_leaf(int node, int first, int count)
	{
		(*nodes)[node].first = first;
		(*nodes)[node].count = count;
		(*nodes)[node].left = (*nodes)[node].right = -1;
	}
};

} // namespace

void build_bvh2_sah(const TriangleArray& tris, int max_leaf_size, Bvh2& out, float intersect_cost)
{
	const float kIntersectCost = intersect_cost;
	out.nodes.clear();
	out.prim_order.clear();
	const int n = (int)tris.size();
	if (n == 0) return;
	Builder b;
	b.nodes = &out.nodes;
	b.max_leaf = std::max(1, std::min(max_leaf_size, 8));
	b.kIntersectCost = intersect_cost;
	b.prims.resize(n);
	for (int i = 0; i < n; i++)
	{
		const Triangle& t = tris[i];
		PrimRef& p = b.prims[i];
		const float* v[3] = { &t.v0.x, &t.v1.x, &t.v2.x };
		for (int a = 0; a < 3; a++)
		{
			p.box.lo[a] = std::min(v[0][a], std::min(v[1][a], v[2][a]));
			p.box.hi[a] = std::max(v[0][a], std::max(v[1][a], v[2][a]));
			p.c[a] = 0.5f * (p.box.lo[a] + p.box.hi[a]);
		}
		p.index = i;
	}
	out.nodes.reserve((size_t)2 * n);
	out.nodes.emplace_back();
	b.build(0, 0, n);
	out.max_depth = b.max_depth;
	out.prim_order.resize(n);
	for (int i = 0; i < n; i++) out.prim_order[i] = b.prims[i].index;

	// SAH cost of the finished tree (reported by the stats / DESIGN.md numbers)
	double cost = 0.0;
	double root_area = half_area(out.nodes[0].box);
	if (root_area > 0.0)
	{
		for (auto& nd : out.nodes)
			cost += (double)half_area(nd.box) / root_area * (nd.count > 0 ? kIntersectCost * nd.count : kTraversalCost);
	}
	out.sah_cost = (float)cost;
}

// Outward padding: a few ulps of the coordinate magnitude, plus a tiny absolute floor.
static inline float pad_down(float v)
{
	float m = std::fabs(v) * 4.76837158e-7f + 1e-30f; // 4 * 2^-23
	return v - m;
}

static inline float pad_up(float v)
{
	float m = std::fabs(v) * 4.76837158e-7f + 1e-30f;
	return v + m;
}

void flatten_bvh2(const Bvh2& bvh, const TriangleArray& tris, GpuBvh2& out)
{
	out.nodes.clear();
	out.tris.clear();
	out.root_is_leaf = 0;
	out.root_ref = 0;
	const int n = (int)bvh.prim_order.size();
	out.tris.resize((size_t)n * 12);
	for (int i = 0; i < n; i++)
	{
		int id = bvh.prim_order[i];
		const Triangle& t = tris[id];
		float* d = &out.tris[(size_t)i * 12];
		d[0] = t.v0.x; d[1] = t.v0.y; d[2] = t.v0.z;
		memcpy(&d[3], &id, 4);
		// same subtraction the reference performs per intersection (Core/triangle.h:33-34)
		d[4] = t.v1.x - t.v0.x; d[5] = t.v1.y - t.v0.y; d[6] = t.v1.z - t.v0.z; d[7] = 0.0f;
		d[8] = t.v2.x - t.v0.x; d[9] = t.v2.y - t.v0.y; d[10] = t.v2.z - t.v0.z; d[11] = 0.0f;
	}
	if (bvh.nodes.empty()) return;

	auto leaf_ref = [](const Bvh2Node& nd) { return ~((nd.first << 3) | (nd.count - 1)); };
	if (bvh.nodes[0].count > 0)
	{
		// single-leaf tree: wrap it in one inner node whose second child is an empty box
		out.nodes.assign(16, 0.0f);
		const Aabb& b = bvh.nodes[0].box;
		float* d = out.nodes.data();
		d[0] = pad_down(b.lo[0]); d[1] = pad_up(b.hi[0]); d[2] = pad_down(b.lo[1]); d[3] = pad_up(b.hi[1]);
		d[4] = 1.0f; d[5] = -1.0f; d[6] = 1.0f; d[7] = -1.0f;
		d[8] = pad_down(b.lo[2]); d[9] = pad_up(b.hi[2]); d[10] = 1.0f; d[11] = -1.0f;
		int c0 = leaf_ref(bvh.nodes[0]), c1 = ~0;
		memcpy(&d[12], &c0, 4); memcpy(&d[13], &c1, 4);
		// c1 = ~0 encodes (first 0, count 1); its inverted box can never be hit
		return;
	}

	// inner nodes get consecutive indices in DFS order so parents and near children share cache lines
	std::vector<int> inner_index(bvh.nodes.size(), -1);
	std::vector<int> order;
	order.reserve(bvh.nodes.size() / 2 + 1);
	std::vector<int> stack;
	stack.push_back(0);
	while (!stack.empty())
	{
		int k = stack.back();
		stack.pop_back();
		if (bvh.nodes[k].count > 0) continue;
		inner_index[k] = (int)order.size();
		order.push_back(k);
		stack.push_back(bvh.nodes[k].right);
		stack.push_back(bvh.nodes[k].left);
	}
	out.nodes.resize(order.size() * 16);
	for (size_t i = 0; i < order.size(); i++)
	{
		const Bvh2Node& nd = bvh.nodes[order[i]];
		const Bvh2Node& c0 = bvh.nodes[nd.left];
		const Bvh2Node& c1 = bvh.nodes[nd.right];
		float* d = &out.nodes[i * 16];
		d[0] = pad_down(c0.box.lo[0]); d[1] = pad_up(c0.box.hi[0]); d[2] = pad_down(c0.box.lo[1]); d[3] = pad_up(c0.box.hi[1]);
		d[4] = pad_down(c1.box.lo[0]); d[5] = pad_up(c1.box.hi[0]); d[6] = pad_down(c1.box.lo[1]); d[7] = pad_up(c1.box.hi[1]);
		d[8] = pad_down(c0.box.lo[2]); d[9] = pad_up(c0.box.hi[2]); d[10] = pad_down(c1.box.lo[2]); d[11] = pad_up(c1.box.hi[2]);
		int r0 = c0.count > 0 ? leaf_ref(c0) : inner_index[nd.left];
		int r1 = c1.count > 0 ? leaf_ref(c1) : inner_index[nd.right];
		memcpy(&d[12], &r0, 4); memcpy(&d[13], &r1, 4);
		d[14] = 0.0f; d[15] = 0.0f;
	}
}

} // namespace ptb

// ==========================================================================================
// 8-wide compressed BVH
// ==========================================================================================
namespace ptb
{

namespace
{

struct WideChild
{
	int node2;      // index into Bvh2::nodes
};

struct WideBuilder
{
	const Bvh2& bvh;
	const TriangleArray& tris;
	GpuBvh8& out;
	int max_depth = 0;

	WideBuilder(const Bvh2& b, const TriangleArray& t, GpuBvh8& o) : bvh(b), tris(t), out(o) {}

	// open the inner child with the largest surface area until 8 children (or only leaves remain)
	void gather_children(int node2, std::vector<int>& children) const
	{
		children.clear();
		const Bvh2Node& n = bvh.nodes[node2];
		children.push_back(n.left);
		children.push_back(n.right);
		while (children.size() < 8)
		{
			int best = -1;
			float best_area = -1.0f;
			for (size_t i = 0; i < children.size(); i++)
			{
				const Bvh2Node& c = bvh.nodes[children[i]];
				if (c.count > 0) continue;
				float a = half_area(c.box);
				if (a > best_area) { best_area = a; best = (int)i; }
			}
			if (best < 0) break;
			int open = children[best];
			children[best] = bvh.nodes[open].left;
			children.push_back(bvh.nodes[open].right);
		}
	}

	// greedy assignment of children to octant slots: slot s prefers the child lying furthest along
	// (s&4 ? +x : -x, s&2 ? +y : -y, s&1 ? +z : -z) from the node centre
	void assign_slots(const Aabb& parent, const std::vector<int>& children, int slot_child[8]) const
	{
		float pc[3];
		for (int a = 0; a < 3; a++) pc[a] = 0.5f * (parent.lo[a] + parent.hi[a]);
		float cost[8][8];
		const int n = (int)children.size();
		for (int c = 0; c < n; c++)
		{
			const Aabb& b = bvh.nodes[children[c]].box;
			float d[3];
			for (int a = 0; a < 3; a++) d[a] = 0.5f * (b.lo[a] + b.hi[a]) - pc[a];
			for (int s = 0; s < 8; s++)
				cost[c][s] = ((s & 4) ? d[0] : -d[0]) + ((s & 2) ? d[1] : -d[1]) + ((s & 1) ? d[2] : -d[2]);
		}
		bool child_done[8] = { false }, slot_done[8] = { false };
		for (int s = 0; s < 8; s++) slot_child[s] = -1;
		for (int k = 0; k < n; k++)
		{
			int bc = -1, bs = -1;
			float best = -std::numeric_limits<float>::infinity();
			for (int c = 0; c < n; c++)
			{
				if (child_done[c]) continue;
				for (int s = 0; s < 8; s++)
				{
					if (slot_done[s]) continue;
					if (cost[c][s] > best || bc < 0) { best = cost[c][s]; bc = c; bs = s; }     // bc < 0: a pair is always chosen, whatever the values (same rule as k_collapse8)
				}
			}
			child_done[bc] = true; slot_done[bs] = true;
			slot_child[bs] = children[bc];
		}
	}

	void emit_triangle(int prim)
	{
		const Triangle& t = tris[prim];
		size_t o = out.tris.size();
		out.tris.resize(o + 12);
		float* d = &out.tris[o];
		d[0] = t.v0.x; d[1] = t.v0.y; d[2] = t.v0.z;
		memcpy(&d[3], &prim, 4);
		d[4] = t.v1.x - t.v0.x; d[5] = t.v1.y - t.v0.y; d[6] = t.v1.z - t.v0.z; d[7] = 0.0f;
		d[8] = t.v2.x - t.v0.x; d[9] = t.v2.y - t.v0.y; d[10] = t.v2.z - t.v0.z; d[11] = 0.0f;
	}

	void build()
	{
		out.nodes.clear();
		out.tris.clear();
		out.tris.reserve(bvh.prim_order.size() * 12);
		if (bvh.nodes.empty()) return;

		struct Item { int node2; int index; int depth; bool root_leaf; };
		std::vector<Item> queue;
		out.nodes.resize(20, 0u);
		queue.push_back({ 0, 0, 1, bvh.nodes[0].count > 0 });
		std::vector<int> children;
		size_t head = 0;
		while (head < queue.size())
		{
			Item it = queue[head++];
			max_depth = std::max(max_depth, it.depth);
			const Aabb& box = bvh.nodes[it.node2].box;
			int slot_child[8];
			if (it.root_leaf)
			{
				for (int s = 0; s < 8; s++) slot_child[s] = -1;
				slot_child[0] = it.node2;
			}
			else
			{
				gather_children(it.node2, children);
				assign_slots(box, children, slot_child);
			}

			// quantisation grid of this node (boxes are padded outwards before quantising)
			float p[3], scale[3];
			uint32_t e[3];
			for (int a = 0; a < 3; a++)
			{
				p[a] = pad_down(box.lo[a]);
				float extent = pad_up(box.hi[a]) - p[a];
				int ex = (int)std::ceil(std::log2(std::max(extent, 1e-30f) / 255.0f));
				// make sure 255 cells really cover the extent after rounding
				while (std::ldexp(255.0f, ex) < extent) ex++;
				ex = std::max(-126, std::min(127, ex));
				e[a] = (uint32_t)(ex + 127);
				scale[a] = std::ldexp(1.0f, -ex);
			}

			uint32_t imask = 0;
			int n_inner = 0;
			for (int s = 0; s < 8; s++)
				if (slot_child[s] >= 0 && bvh.nodes[slot_child[s]].count == 0) { imask |= 1u << s; n_inner++; }
			uint32_t child_base = (uint32_t)(out.nodes.size() / 20);
			out.nodes.resize(out.nodes.size() + (size_t)n_inner * 20, 0u);
			uint32_t tri_base = (uint32_t)(out.tris.size() / 12);

			uint8_t meta[8], qlo[3][8], qhi[3][8];
			int inner_rank = 0, tri_offset = 0;
			for (int s = 0; s < 8; s++)
			{
				meta[s] = 0;
				for (int a = 0; a < 3; a++) { qlo[a][s] = 0; qhi[a][s] = 0; }
				int c = slot_child[s];
				if (c < 0) continue;
				const Bvh2Node& cn = bvh.nodes[c];
				for (int a = 0; a < 3; a++)
				{
					float lo = std::floor((pad_down(cn.box.lo[a]) - p[a]) * scale[a]);
					float hi = std::ceil((pad_up(cn.box.hi[a]) - p[a]) * scale[a]);
					qlo[a][s] = (uint8_t)std::max(0.0f, std::min(255.0f, lo));
					qhi[a][s] = (uint8_t)std::max(0.0f, std::min(255.0f, hi));
				}
				if (cn.count == 0)
				{
					meta[s] = (uint8_t)((1u << 5) | (24u + (uint32_t)s));
					queue.push_back({ c, (int)child_base + inner_rank, it.depth + 1, false });
					inner_rank++;
				}
				else
				{
					int count = std::min(cn.count, 3);
					uint32_t unary = count == 1 ? 1u : (count == 2 ? 3u : 7u);
					meta[s] = (uint8_t)((unary << 5) | (uint32_t)tri_offset);
					for (int k = 0; k < count; k++) emit_triangle(bvh.prim_order[cn.first + k]);
					tri_offset += count;
				}
			}

			uint32_t* d = &out.nodes[(size_t)it.index * 20];
			memcpy(&d[0], &p[0], 4); memcpy(&d[1], &p[1], 4); memcpy(&d[2], &p[2], 4);
			d[3] = e[0] | (e[1] << 8) | (e[2] << 16) | (imask << 24);
			d[4] = child_base;
			d[5] = tri_base;
			d[6] = meta[0] | (meta[1] << 8) | (meta[2] << 16) | ((uint32_t)meta[3] << 24);
			d[7] = meta[4] | (meta[5] << 8) | (meta[6] << 16) | ((uint32_t)meta[7] << 24);
			for (int a = 0; a < 3; a++)
			{
				d[8 + a * 4 + 0] = qlo[a][0] | (qlo[a][1] << 8) | (qlo[a][2] << 16) | ((uint32_t)qlo[a][3] << 24);
				d[8 + a * 4 + 1] = qlo[a][4] | (qlo[a][5] << 8) | (qlo[a][6] << 16) | ((uint32_t)qlo[a][7] << 24);
				d[8 + a * 4 + 2] = qhi[a][0] | (qhi[a][1] << 8) | (qhi[a][2] << 16) | ((uint32_t)qhi[a][3] << 24);
				d[8 + a * 4 + 3] = qhi[a][4] | (qhi[a][5] << 8) | (qhi[a][6] << 16) | ((uint32_t)qhi[a][7] << 24);
			}
		}
		out.max_depth = max_depth;
	}
};

} // namespace

void build_bvh8(const Bvh2& bvh, const TriangleArray& tris, GpuBvh8& out)
{
	WideBuilder b(bvh, tris, out);
	b.build();
}

} // namespace ptb
