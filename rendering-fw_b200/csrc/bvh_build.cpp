// bvh_build.cpp — see bvh_build.h
#include "bvh_build.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <limits>
#include <mutex>
#include <thread>

namespace rfwb200
{
namespace
{

constexpr int BINS = 16;
constexpr int MAX_LEAF = 4;
constexpr int MAX_DEPTH2 = 31;		   // BVH2 depth bound => BVH4 depth <= 31 => stack <= 94 < TRAVERSAL_STACK
constexpr size_t PAR_THRESHOLD = 16384; // subtrees above this size become pool tasks
constexpr float BOX_PAD = 1e-5f;		   // reference pads primitive and node boxes by 1e-5 (bvh_tree.cpp:446, bvh_node.h:221)

struct Box
{
	float lo[3], hi[3];
	void reset()
	{
		for (int a = 0; a < 3; a++)
			lo[a] = 3.0e38f, hi[a] = -3.0e38f;
	}
	void grow(const Box &b)
	{
		for (int a = 0; a < 3; a++)
			lo[a] = std::min(lo[a], b.lo[a]), hi[a] = std::max(hi[a], b.hi[a]);
	}
	void grow(const float *p)
	{
		for (int a = 0; a < 3; a++)
			lo[a] = std::min(lo[a], p[a]), hi[a] = std::max(hi[a], p[a]);
	}
	float area() const
	{
		const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
		if (ex < 0 || ey < 0 || ez < 0)
			return 0;
		return ex * ey + ey * ez + ez * ex;
	}
};

struct Node2
{
	Box box;
	int32_t left = -1, right = -1; // children (inner)
	uint32_t first = 0, count = 0; // leaf range in idx[] when count > 0
	int32_t depth = 0;
};

struct Task
{
	int32_t node;
	uint32_t first, count;
	int32_t depth;
};

struct Builder
{
	const Box *boxes;
	const float *cent; // 3 per prim
	std::vector<uint32_t> idx;
	std::vector<Node2> nodes;
	std::atomic<int32_t> next_node{1};

	std::mutex mtx;
	std::condition_variable cv;
	std::deque<Task> queue;
	std::atomic<int> outstanding{0};
	bool parallel = false;

	static int ceil_log2(uint32_t v)
	{
		int r = 0;
		while ((1u << r) < v)
			r++;
		return r;
	}

	// returns false when the range became a leaf
	bool split(const Task &t, Task &lt, Task &rt)
	{
		Node2 &node = nodes[t.node];
		node.depth = t.depth;
		const uint32_t first = t.first, count = t.count;
		if (count <= 1)
		{
			node.first = first, node.count = count;
			return false;
		}
		Box cb;
		cb.reset();
		for (uint32_t i = first; i < first + count; i++)
			cb.grow(cent + 3 * size_t(idx[i]));

		const int levels_needed = ceil_log2((count + MAX_LEAF - 1) / MAX_LEAF);
		const bool force_median = t.depth + levels_needed >= MAX_DEPTH2;

		int best_axis = -1, best_bin = -1;
		float best_cost = 3.0e38f;
		if (!force_median)
		{
			for (int axis = 0; axis < 3; axis++)
			{
				const float ext = cb.hi[axis] - cb.lo[axis];
				if (!(ext > 0))
					continue;
				Box bb[BINS];
				uint32_t bc[BINS];
				for (int b = 0; b < BINS; b++)
					bb[b].reset(), bc[b] = 0;
				const float scale = float(BINS) * (1.0f - 1e-6f) / ext;
				for (uint32_t i = first; i < first + count; i++)
				{
					const uint32_t p = idx[i];
					int b = int((cent[3 * size_t(p) + axis] - cb.lo[axis]) * scale);
					b = b < 0 ? 0 : (b >= BINS ? BINS - 1 : b);
					bb[b].grow(boxes[p]);
					bc[b]++;
				}
				float right_area[BINS];
				uint32_t right_count[BINS];
				Box acc;
				acc.reset();
				uint32_t c = 0;
				for (int b = BINS - 1; b > 0; b--)
				{
					acc.grow(bb[b]);
					c += bc[b];
					right_area[b] = acc.area(), right_count[b] = c;
				}
				acc.reset();
				c = 0;
				for (int b = 0; b < BINS - 1; b++)
				{
					acc.grow(bb[b]);
					c += bc[b];
					if (c == 0 || right_count[b + 1] == 0)
						continue;
					const float cost = acc.area() * float(c) + right_area[b + 1] * float(right_count[b + 1]);
					if (cost < best_cost)
						best_cost = cost, best_axis = axis, best_bin = b;
				}
			}
		}
		const float node_area = node.box.area();
		const float leaf_cost = float(count) * node_area;
		const float split_cost = node_area * 1.0f + best_cost; // C_trav = C_isect = 1
		if (count <= MAX_LEAF && (best_axis < 0 || split_cost >= leaf_cost))
		{
			node.first = first, node.count = count;
			return false;
		}
		uint32_t mid;
		if (best_axis >= 0)
		{
			const float ext = cb.hi[best_axis] - cb.lo[best_axis];
			const float scale = float(BINS) * (1.0f - 1e-6f) / ext;
			const float lo = cb.lo[best_axis];
			const int axis = best_axis, bin = best_bin;
			auto it = std::partition(idx.begin() + first, idx.begin() + first + count, [&](uint32_t p) {
				int b = int((cent[3 * size_t(p) + axis] - lo) * scale);
				b = b < 0 ? 0 : (b >= BINS ? BINS - 1 : b);
				return b <= bin;
			});
			mid = uint32_t(it - idx.begin());
		}
		else
		{
			// all centroids coincide, or the depth budget forces balanced splits: object median on
			// the widest centroid axis
			int axis = 0;
			float w = cb.hi[0] - cb.lo[0];
			for (int a = 1; a < 3; a++)
				if (cb.hi[a] - cb.lo[a] > w)
					w = cb.hi[a] - cb.lo[a], axis = a;
			mid = first + count / 2;
			std::nth_element(idx.begin() + first, idx.begin() + mid, idx.begin() + first + count, [&](uint32_t a, uint32_t b) {
				return cent[3 * size_t(a) + axis] < cent[3 * size_t(b) + axis];
			});
		}
		if (mid == first || mid == first + count)
			mid = first + count / 2;
		const int32_t l = next_node.fetch_add(2), r = l + 1;
		node.left = l, node.right = r, node.count = 0;
		Box lb, rb;
		lb.reset(), rb.reset();
		for (uint32_t i = first; i < mid; i++)
			lb.grow(boxes[idx[i]]);
		for (uint32_t i = mid; i < first + count; i++)
			rb.grow(boxes[idx[i]]);
		nodes[l].box = lb, nodes[r].box = rb;
		lt = {l, first, mid - first, t.depth + 1};
		rt = {r, mid, first + count - mid, t.depth + 1};
		return true;
	}

	void build_serial(const Task &t)
	{
		// explicit stack: depth-first, large right halves may be handed to the pool
		std::vector<Task> stack;
		stack.push_back(t);
		while (!stack.empty())
		{
			const Task cur = stack.back();
			stack.pop_back();
			Task l, r;
			if (!split(cur, l, r))
				continue;
			if (parallel && r.count > PAR_THRESHOLD)
				push_task(r);
			else
				stack.push_back(r);
			stack.push_back(l);
		}
	}

	void push_task(const Task &t)
	{
		outstanding.fetch_add(1);
		{
			std::lock_guard<std::mutex> lk(mtx);
			queue.push_back(t);
		}
		cv.notify_one();
	}

	void worker()
	{
		for (;;)
		{
			Task t;
			{
				std::unique_lock<std::mutex> lk(mtx);
				cv.wait(lk, [&] { return !queue.empty() || outstanding.load() == 0; });
				if (queue.empty())
					return;
				t = queue.front();
				queue.pop_front();
			}
			build_serial(t);
			if (outstanding.fetch_sub(1) == 1)
			{
				std::lock_guard<std::mutex> lk(mtx);
				cv.notify_all();
			}
		}
	}
};

inline Box tri_box(const BuildTriangle &t)
{
	Box b;
	b.reset();
	b.grow(t.v0), b.grow(t.v1), b.grow(t.v2);
	for (int a = 0; a < 3; a++)
	{
		// pad relative to magnitude as well: 1e-5 absolute vanishes next to coordinates ~1e3
		const float m = std::max(std::fabs(b.lo[a]), std::fabs(b.hi[a]));
		const float pad = std::max(BOX_PAD, m * 2.4e-7f);
		b.lo[a] -= pad, b.hi[a] += pad;
	}
	return b;
}

inline void set_child_box(BvhNode4 &n, int slot, const Box &b)
{
	n.minx[slot] = b.lo[0], n.miny[slot] = b.lo[1], n.minz[slot] = b.lo[2];
	n.maxx[slot] = b.hi[0], n.maxy[slot] = b.hi[1], n.maxz[slot] = b.hi[2];
}
// An unused slot must never be entered: min/max slab arithmetic turns an inverted box into a valid
// one (the reference relies on count == 0 for that, mbvh_node.cpp:207-212), so the box is NaN — every
// comparison in the hit test is false — and the child word is a harmless 1-triangle leaf reference.
// pad[0] of the node holds the number of used slots.
inline void set_child_empty(BvhNode4 &n, int slot)
{
	const float q = std::numeric_limits<float>::quiet_NaN();
	n.minx[slot] = n.miny[slot] = n.minz[slot] = q;
	n.maxx[slot] = n.maxy[slot] = n.maxz[slot] = q;
	n.child[slot] = -1;
}
inline Box child_box(const BvhNode4 &n, int slot)
{
	Box b;
	b.lo[0] = n.minx[slot], b.lo[1] = n.miny[slot], b.lo[2] = n.minz[slot];
	b.hi[0] = n.maxx[slot], b.hi[1] = n.maxy[slot], b.hi[2] = n.maxz[slot];
	return b;
}

} // namespace

void build_bvh4(const BuildTriangle *tris, size_t count, int threads, BvhBuildResult &out)
{
	const auto t0 = std::chrono::steady_clock::now();
	out = BvhBuildResult();
	if (count == 0)
	{
		BvhNode4 root;
		memset(&root, 0, sizeof(root));
		for (int s = 0; s < 4; s++)
			set_child_empty(root, s);
		out.nodes.push_back(root);
		out.node_parent.push_back(0xffffffffu);
		return;
	}
	std::vector<Box> boxes(count);
	std::vector<float> cent(count * 3);
	Box root_box;
	root_box.reset();
	for (size_t i = 0; i < count; i++)
	{
		boxes[i] = tri_box(tris[i]);
		for (int a = 0; a < 3; a++)
			cent[3 * i + a] = (boxes[i].lo[a] + boxes[i].hi[a]) * 0.5f;
		root_box.grow(boxes[i]);
	}
	Builder b;
	b.boxes = boxes.data();
	b.cent = cent.data();
	b.idx.resize(count);
	for (size_t i = 0; i < count; i++)
		b.idx[i] = uint32_t(i);
	b.nodes.resize(2 * count + 2);
	b.nodes[0].box = root_box;
	threads = std::max(1, threads);
	b.parallel = threads > 1 && count > 4 * PAR_THRESHOLD;
	if (b.parallel)
	{
		b.push_task({0, 0, uint32_t(count), 0});
		std::vector<std::thread> pool;
		for (int t = 1; t < threads; t++)
			pool.emplace_back([&] { b.worker(); });
		b.worker();
		for (auto &t : pool)
			t.join();
	}
	else
		b.build_serial({0, 0, uint32_t(count), 0});

	// ---- collapse to 4-wide, breadth-first layout ----
	const std::vector<Node2> &n2 = b.nodes;
	struct Pending
	{
		int32_t n2; // BVH2 inner node (or -1: synthetic root around a leaf root)
		uint32_t parent;
	};
	std::vector<Pending> fifo;
	fifo.reserve(count);
	out.nodes.reserve(count / 2 + 1);
	fifo.push_back({0, 0xffffffffu});
	float cost = 0;
	const float inv_root_area = root_box.area() > 0 ? 1.0f / root_box.area() : 0.0f;
	std::vector<int> depth4;
	depth4.push_back(1);
	int max_depth = 1;
	for (size_t head = 0; head < fifo.size(); head++)
	{
		const Pending p = fifo[head];
		BvhNode4 node;
		memset(&node, 0, sizeof(node));
		int32_t kids[4];
		int nk = 0;
		const Node2 &src = n2[p.n2];
		if (src.count > 0 || src.left < 0)
			kids[nk++] = p.n2; // root is a leaf
		else
		{
			kids[nk++] = src.left, kids[nk++] = src.right;
			while (nk < 4)
			{
				int best = -1;
				float best_area = -1;
				for (int k = 0; k < nk; k++)
				{
					const Node2 &c = n2[kids[k]];
					if (c.count == 0 && c.left >= 0 && c.box.area() > best_area)
						best_area = c.box.area(), best = k;
				}
				if (best < 0)
					break;
				const Node2 &c = n2[kids[best]];
				kids[best] = c.left;
				kids[nk++] = c.right;
			}
		}
		cost += src.box.area() * inv_root_area;
		const uint32_t self = uint32_t(head);
		for (int k = 0; k < 4; k++)
		{
			if (k >= nk)
			{
				set_child_empty(node, k);
				continue;
			}
			const Node2 &c = n2[kids[k]];
			set_child_box(node, k, c.box);
			if (c.count > 0 || c.left < 0)
			{
				const uint32_t cnt = std::max<uint32_t>(c.count, 1u);
				node.child[k] = ~int32_t((c.first << 2) | (cnt - 1));
				cost += c.box.area() * inv_root_area * float(c.count);
			}
			else
			{
				node.child[k] = int32_t(fifo.size());
				fifo.push_back({kids[k], self});
				depth4.push_back(depth4[head] + 1);
				max_depth = std::max(max_depth, depth4[head] + 1);
			}
		}
		node.pad[0] = nk;
		out.nodes.push_back(node);
		out.node_parent.push_back(p.parent);
	}
	out.tri_order = std::move(b.idx);
	out.sah_cost = cost;
	out.depth = max_depth;
	out.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

void refit_bvh4(const BuildTriangle *tris, size_t count, BvhBuildResult &bvh)
{
	(void)count;
	for (size_t ni = bvh.nodes.size(); ni-- > 0;)
	{
		BvhNode4 &n = bvh.nodes[ni];
		for (int s = 0; s < n.pad[0]; s++)
		{
			const int32_t c = n.child[s];
			Box b;
			b.reset();
			if (c < 0)
			{
				const uint32_t v = uint32_t(~c), first = v >> 2, cnt = (v & 3u) + 1u;
				for (uint32_t i = first; i < first + cnt; i++)
					b.grow(tri_box(tris[bvh.tri_order[i]]));
			}
			else
			{
				const BvhNode4 &ch = bvh.nodes[c];
				for (int k = 0; k < ch.pad[0]; k++)
					b.grow(child_box(ch, k));
			}
			set_child_box(n, s, b);
		}
	}
}

} // namespace rfwb200
