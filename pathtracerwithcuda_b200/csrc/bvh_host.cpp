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
const float kIntersectCost = 1.5f;

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

	// builds the subtree for prims[first, first+count) into node `node_index`
	void build(int node_index, int first, int count)
	{
		// explicit stack keeps deep, degenerate inputs from overflowing the call stack
		struct Item { int node, first, count; };
		std::vector<Item> stack;
		stack.push_back({ node_index, first, count });
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
			if (it.count <= 1) { make_leaf(it.node, it.first, it.count); continue; }

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
				push_children(stack, it.node, it.first, mid, it.first + it.count);
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
			push_children(stack, it.node, it.first, mid, it.first + it.count);
		}
	}

	template <class Stack>
	void push_children(Stack& stack, int node, int first, int mid, int end)
	{
		int l = (int)nodes->size();
		nodes->emplace_back();
		nodes->emplace_back();
		(*nodes)[node].left = l;
		(*nodes)[node].right = l + 1;
		(*nodes)[node].count = 0;
		stack.push_back({ l + 1, mid, end - mid });
		stack.push_back({ l, first, mid - first });
	}

	void make_leaf(int node, int first, int count)
	{
		(*nodes)[node].first = first;
		(*nodes)[node].count = count;
		(*nodes)[node].left = (*nodes)[node].right = -1;
	}
};

} // namespace

void build_bvh2_sah(const std::vector<Triangle>& tris, int max_leaf_size, Bvh2& out)
{
	out.nodes.clear();
	out.prim_order.clear();
	const int n = (int)tris.size();
	if (n == 0) return;
	Builder b;
	b.nodes = &out.nodes;
	b.max_leaf = std::max(1, std::min(max_leaf_size, 8));
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

void flatten_bvh2(const Bvh2& bvh, const std::vector<Triangle>& tris, GpuBvh2& out)
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
