// bvh_build.cpp — see bvh_build.h
//
// Builder: top-down SAH over triangle *references* (primitive id + clipped box) with spatial splits
// (Stich, Friedrich, Dietrich: "Spatial Splits in Bounding Volume Hierarchies", HPG 2009) — the reference
// asks its Rust crate for the same kind of tree (BVHTree::SpatialSAH, RFW/system/bvh/src/top_level_bvh.cpp:41).
// Object splits are binned (16 bins on reference-box centroids), spatial splits use chopped binning with
// exact triangle clipping; a spatial split is only tried when the object-split children overlap, and a global
// budget bounds the number of duplicated references.  Large subtrees are built by a pool of host threads.
// The BVH2 is then collapsed to 4-wide (open the child with the largest area) and laid out breadth-first.
#include "bvh_build.h"
#include "cwbvh.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <limits>
#include <memory>
#include <mutex>
#include <thread>

namespace rfwb200
{
namespace
{

constexpr int BINS = 16;
constexpr int SPATIAL_BINS = 32;
constexpr int MAX_LEAF = 4; // 4-wide layout; the compressed 8-wide layout takes 3 (Builder::max_leaf)
constexpr int MAX_DEPTH2 = 31;			// BVH2 depth bound => BVH4 depth <= 31 => stack <= 94 + sentinel <= TRAVERSAL_STACK
constexpr size_t PAR_THRESHOLD = 16384; // subtrees above this many references become pool tasks
constexpr float BOX_PAD = 1e-5f;		// reference pads primitive and node boxes by 1e-5 (bvh_tree.cpp:446, bvh_node.h:221)
#ifndef RFW_SPATIAL_ALPHA
#define RFW_SPATIAL_ALPHA 1e-5f
#endif
constexpr float SPATIAL_ALPHA = RFW_SPATIAL_ALPHA;	// overlap / root area above which a spatial split is considered
#ifndef RFW_REF_BUDGET
#define RFW_REF_BUDGET 0.6f
#endif
#ifndef RFW_SAH_CTRAV
#define RFW_SAH_CTRAV 1.0f
#endif
#ifndef RFW_COLLAPSE_CTRAV
#define RFW_COLLAPSE_CTRAV 1.0f // node visit : triangle test cost in the collapse (measured: 1 is best, 2.4 = the kernels' instruction ratio is not)
#endif
constexpr float REF_BUDGET = RFW_REF_BUDGET;		// at most this fraction of extra (duplicated) references

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
	void clip_to(const Box &b)
	{
		for (int a = 0; a < 3; a++)
			lo[a] = std::max(lo[a], b.lo[a]), hi[a] = std::min(hi[a], b.hi[a]);
	}
	bool valid() const { return lo[0] <= hi[0] && lo[1] <= hi[1] && lo[2] <= hi[2]; }
	float area() const
	{
		const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
		if (ex < 0 || ey < 0 || ez < 0)
			return 0;
		return ex * ey + ey * ez + ez * ex;
	}
};

inline Box intersection(const Box &a, const Box &b)
{
	Box r = a;
	r.clip_to(b);
	return r;
}

struct Ref
{
	Box box;
	uint32_t prim;
};

struct Node2
{
	Box box;
	int32_t left = -1, right = -1; // children (inner)
	std::vector<uint32_t> prims;   // leaf content
	std::vector<Box> prim_boxes;   // the (possibly clipped) box of each reference
	bool leaf = false;
};

struct Task
{
	int32_t node;
	std::vector<Ref> refs;
	int32_t depth;
	int64_t budget = 0; // references spatial splits below this node may still duplicate.  The budget travels with the task
						// (a split hands what is left to its children in proportion to their reference counts) instead of
						// living in one shared counter, so the tree does not depend on which worker thread splits first:
						// every build of the same triangles — on every rank of a sharded frame — gives the same tree.
};

inline int ceil_log2(uint32_t v)
{
	int r = 0;
	while ((1u << r) < v)
		r++;
	return r;
}

// clip the triangle of `r` against the plane x[axis] = pos, bound both halves (SBVH reference splitting)
inline void split_ref(const Ref &r, const BuildTriangle &t, int axis, float pos, Ref &l, Ref &rt)
{
	l.prim = rt.prim = r.prim;
	l.box.reset(), rt.box.reset();
	const float *v[3] = {t.v0, t.v1, t.v2};
	for (int i = 0; i < 3; i++)
	{
		const float *a = v[i], *b = v[(i + 1) % 3];
		const float pa = a[axis], pb = b[axis];
		if (pa <= pos)
			l.box.grow(a);
		if (pa >= pos)
			rt.box.grow(a);
		if ((pa < pos && pb > pos) || (pa > pos && pb < pos))
		{
			const float s = std::min(std::max((pos - pa) / (pb - pa), 0.0f), 1.0f);
			float p[3];
			for (int k = 0; k < 3; k++)
				p[k] = a[k] + (b[k] - a[k]) * s;
			p[axis] = pos;
			l.box.grow(p), rt.box.grow(p);
		}
	}
	l.box.hi[axis] = std::min(l.box.hi[axis], pos);
	rt.box.lo[axis] = std::max(rt.box.lo[axis], pos);
	l.box.clip_to(r.box), rt.box.clip_to(r.box);
}

struct Builder
{
	const BuildTriangle *tris = nullptr;
	std::vector<Node2> nodes;
	std::atomic<int32_t> next_node{1};
	float root_area = 1.0f;
	bool spatial = true;
	int max_leaf = MAX_LEAF;
	int max_depth2 = MAX_DEPTH2;

	std::mutex mtx;
	std::condition_variable cv;
	std::deque<std::unique_ptr<Task>> queue;
	std::atomic<int> outstanding{0};
	bool parallel = false;

	void make_leaf(Node2 &node, const std::vector<Ref> &refs)
	{
		node.leaf = true;
		node.prims.resize(refs.size());
		node.prim_boxes.resize(refs.size());
		for (size_t i = 0; i < refs.size(); i++)
			node.prims[i] = refs[i].prim, node.prim_boxes[i] = refs[i].box;
	}

	// returns false when the node became a leaf; otherwise fills the two child tasks
	bool split(Task &t, Task &lt, Task &rt)
	{
		Node2 &node = nodes[t.node];
		std::vector<Ref> &refs = t.refs;
		const uint32_t count = uint32_t(refs.size());
		if (count <= 1)
		{
			make_leaf(node, refs);
			return false;
		}
		Box cb;
		cb.reset();
		for (const Ref &r : refs)
		{
			float c[3];
			for (int a = 0; a < 3; a++)
				c[a] = (r.box.lo[a] + r.box.hi[a]) * 0.5f;
			cb.grow(c);
		}
		const int levels_needed = ceil_log2((count + max_leaf - 1) / max_leaf);
		const bool force_median = t.depth + levels_needed >= max_depth2;

		// ---- object split: binned SAH on reference centroids -------------------------------------------
		int obj_axis = -1, obj_bin = -1;
		float obj_cost = 3.0e38f, obj_overlap = 0.0f;
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
				for (const Ref &r : refs)
				{
					int b = int(((r.box.lo[axis] + r.box.hi[axis]) * 0.5f - cb.lo[axis]) * scale);
					b = b < 0 ? 0 : (b >= BINS ? BINS - 1 : b);
					bb[b].grow(r.box);
					bc[b]++;
				}
				Box right_box[BINS];
				uint32_t right_count[BINS];
				Box acc;
				acc.reset();
				uint32_t c = 0;
				for (int b = BINS - 1; b > 0; b--)
				{
					acc.grow(bb[b]);
					c += bc[b];
					right_box[b] = acc, right_count[b] = c;
				}
				acc.reset();
				c = 0;
				for (int b = 0; b < BINS - 1; b++)
				{
					acc.grow(bb[b]);
					c += bc[b];
					if (c == 0 || right_count[b + 1] == 0)
						continue;
					const float cost = acc.area() * float(c) + right_box[b + 1].area() * float(right_count[b + 1]);
					if (cost < obj_cost)
					{
						obj_cost = cost, obj_axis = axis, obj_bin = b;
						const Box ov = intersection(acc, right_box[b + 1]);
						obj_overlap = ov.valid() ? ov.area() : 0.0f;
					}
				}
			}
		}

		// ---- spatial split: chopped binning over the node box ---------------------------------------------
		int sp_axis = -1;
		float sp_pos = 0.0f, sp_cost = 3.0e38f;
		int64_t budget_left = t.budget;
		if (spatial && !force_median && count > uint32_t(max_leaf) && t.budget > 0 &&
			(obj_axis < 0 || obj_overlap > SPATIAL_ALPHA * root_area))
		{
			for (int axis = 0; axis < 3; axis++)
			{
				const float lo = node.box.lo[axis], ext = node.box.hi[axis] - lo;
				if (!(ext > 1e-12f))
					continue;
				struct SBin
				{
					Box box;
					uint32_t enter, exit;
				} bins[SPATIAL_BINS];
				for (auto &b : bins)
					b.box.reset(), b.enter = b.exit = 0;
				const float scale = float(SPATIAL_BINS) / ext, binw = ext / float(SPATIAL_BINS);
				for (const Ref &r : refs)
				{
					int first = int((r.box.lo[axis] - lo) * scale), last = int((r.box.hi[axis] - lo) * scale);
					first = std::min(std::max(first, 0), SPATIAL_BINS - 1);
					last = std::min(std::max(last, first), SPATIAL_BINS - 1);
					Ref cur = r;
					for (int b = first; b < last; b++)
					{
						Ref l, rr;
						split_ref(cur, tris[r.prim], axis, lo + binw * float(b + 1), l, rr);
						if (l.box.valid())
							bins[b].box.grow(l.box);
						cur = rr;
						if (!cur.box.valid())
							break;
					}
					if (cur.box.valid())
						bins[last].box.grow(cur.box);
					bins[first].enter++, bins[last].exit++;
				}
				Box right_box[SPATIAL_BINS];
				Box acc;
				acc.reset();
				for (int b = SPATIAL_BINS - 1; b > 0; b--)
				{
					acc.grow(bins[b].box);
					right_box[b] = acc;
				}
				acc.reset();
				uint32_t ln = 0, rn = count;
				for (int b = 1; b < SPATIAL_BINS; b++)
				{
					acc.grow(bins[b - 1].box);
					ln += bins[b - 1].enter;
					rn -= bins[b - 1].exit;
					if (ln == 0 || rn == 0)
						continue;
					const float cost = acc.area() * float(ln) + right_box[b].area() * float(rn);
					if (cost < sp_cost)
						sp_cost = cost, sp_axis = axis, sp_pos = lo + binw * float(b);
				}
			}
		}

		const float node_area = node.box.area();
		const float leaf_cost = float(count) * node_area;
		const float best_cost = std::min(obj_cost, sp_cost);
		const float split_cost = node_area * RFW_SAH_CTRAV + best_cost; // C_trav (default 1) relative to C_isect = 1
		if (count <= uint32_t(max_leaf) && (best_cost >= 3.0e38f || split_cost >= leaf_cost))
		{
			make_leaf(node, refs);
			return false;
		}

		std::vector<Ref> left, right;
		bool done = false;
		if (sp_axis >= 0 && sp_cost < obj_cost)
		{
			left.reserve(count), right.reserve(count);
			for (const Ref &r : refs)
			{
				if (r.box.hi[sp_axis] <= sp_pos)
					left.push_back(r);
				else if (r.box.lo[sp_axis] >= sp_pos)
					right.push_back(r);
				else
				{
					Ref l, rr;
					split_ref(r, tris[r.prim], sp_axis, sp_pos, l, rr);
					const bool lv = l.box.valid(), rv = rr.box.valid();
					if (lv && rv)
						left.push_back(l), right.push_back(rr);
					else if (lv)
						left.push_back(r);
					else
						right.push_back(r);
				}
			}
			const int64_t extra = int64_t(left.size() + right.size()) - int64_t(count);
			if (!left.empty() && !right.empty() && left.size() < count + count / 2 && right.size() < count + count / 2 &&
				(left.size() < count || right.size() < count) && extra <= t.budget)
			{
				budget_left = t.budget - extra;
				done = true;
			}
			else
				left.clear(), right.clear();
		}
		if (!done && obj_axis >= 0)
		{
			const float ext = cb.hi[obj_axis] - cb.lo[obj_axis];
			const float scale = float(BINS) * (1.0f - 1e-6f) / ext, lo = cb.lo[obj_axis];
			left.reserve(count), right.reserve(count);
			for (const Ref &r : refs)
			{
				int b = int(((r.box.lo[obj_axis] + r.box.hi[obj_axis]) * 0.5f - lo) * scale);
				b = b < 0 ? 0 : (b >= BINS ? BINS - 1 : b);
				(b <= obj_bin ? left : right).push_back(r);
			}
			done = !left.empty() && !right.empty();
			if (!done)
				left.clear(), right.clear();
		}
		if (!done)
		{
			// all centroids coincide, or the depth budget forces balanced splits: object median on the widest axis
			int axis = 0;
			float w = cb.hi[0] - cb.lo[0];
			for (int a = 1; a < 3; a++)
				if (cb.hi[a] - cb.lo[a] > w)
					w = cb.hi[a] - cb.lo[a], axis = a;
			const size_t mid = count / 2;
			std::nth_element(refs.begin(), refs.begin() + mid, refs.end(), [&](const Ref &a, const Ref &b) {
				return a.box.lo[axis] + a.box.hi[axis] < b.box.lo[axis] + b.box.hi[axis];
			});
			left.assign(refs.begin(), refs.begin() + mid);
			right.assign(refs.begin() + mid, refs.end());
		}
		std::vector<Ref>().swap(refs);

		const int32_t l = next_node.fetch_add(2), r = l + 1;
		if (size_t(r) >= nodes.size()) // cannot happen while the budget holds (nodes are sized for it); never write past the array
		{
			t.refs.swap(left);
			t.refs.insert(t.refs.end(), right.begin(), right.end());
			make_leaf(node, t.refs);
			return false;
		}
		node.left = l, node.right = r;
		lt.budget = int64_t(double(budget_left) * double(left.size()) / double(left.size() + right.size()));
		rt.budget = budget_left - lt.budget;
		Box lb, rb;
		lb.reset(), rb.reset();
		for (const Ref &x : left)
			lb.grow(x.box);
		for (const Ref &x : right)
			rb.grow(x.box);
		nodes[l].box = lb, nodes[r].box = rb;
		lt.node = l, lt.refs = std::move(left), lt.depth = t.depth + 1;
		rt.node = r, rt.refs = std::move(right), rt.depth = t.depth + 1;
		return true;
	}

	void build_serial(std::unique_ptr<Task> root)
	{
		std::vector<std::unique_ptr<Task>> stack;
		stack.push_back(std::move(root));
		while (!stack.empty())
		{
			std::unique_ptr<Task> cur = std::move(stack.back());
			stack.pop_back();
			std::unique_ptr<Task> l(new Task()), r(new Task());
			if (!split(*cur, *l, *r))
				continue;
			if (parallel && r->refs.size() > PAR_THRESHOLD)
				push_task(std::move(r));
			else
				stack.push_back(std::move(r));
			stack.push_back(std::move(l));
		}
	}

	void push_task(std::unique_ptr<Task> t)
	{
		outstanding.fetch_add(1);
		{
			std::lock_guard<std::mutex> lk(mtx);
			queue.push_back(std::move(t));
		}
		cv.notify_one();
	}

	void worker()
	{
		for (;;)
		{
			std::unique_ptr<Task> t;
			{
				std::unique_lock<std::mutex> lk(mtx);
				cv.wait(lk, [&] { return !queue.empty() || outstanding.load() == 0; });
				if (queue.empty())
					return;
				t = std::move(queue.front());
				queue.pop_front();
			}
			build_serial(std::move(t));
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
	return b;
}

// pad relative to magnitude as well: 1e-5 absolute vanishes next to coordinates ~1e3
inline Box padded(const Box &in)
{
	Box b = in;
	for (int a = 0; a < 3; a++)
	{
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

// top-down SBVH over the triangles into b.nodes (BVH2); returns the number of references the tree may hold
static size_t build_tree2(Builder &b, const BuildTriangle *tris, size_t count, int threads, bool spatial_splits)
{
	std::unique_ptr<Task> root(new Task());
	root->node = 0, root->depth = 0;
	root->refs.resize(count);
	Box root_box;
	root_box.reset();
	for (size_t i = 0; i < count; i++)
	{
		root->refs[i].box = tri_box(tris[i]);
		root->refs[i].prim = uint32_t(i);
		root_box.grow(root->refs[i].box);
	}
	b.tris = tris;
	b.spatial = spatial_splits;
	b.root_area = std::max(root_box.area(), 1e-30f);
	root->budget = int64_t(double(count) * REF_BUDGET);
	const size_t max_refs = count + size_t(double(count) * REF_BUDGET) + 64;
	b.nodes.resize(2 * max_refs + 2);
	b.nodes[0].box = root_box;
	threads = std::max(1, threads);
	b.parallel = threads > 1 && count > 4 * PAR_THRESHOLD;
	if (b.parallel)
	{
		b.push_task(std::move(root));
		std::vector<std::thread> pool;
		for (int t = 1; t < threads; t++)
			pool.emplace_back([&] { b.worker(); });
		b.worker();
		for (auto &t : pool)
			t.join();
	}
	else
		b.build_serial(std::move(root));
	return max_refs;
}

void build_bvh4(const BuildTriangle *tris, size_t count, int threads, BvhBuildResult &out, bool spatial_splits)
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
	Builder b;
	const size_t max_refs = build_tree2(b, tris, count, threads, spatial_splits);

	// ---- collapse to 4-wide, breadth-first layout; leaves get their slice of the triangle order here ----
	const std::vector<Node2> &n2 = b.nodes;
	// Optional SAH-optimal collapse (-DRFW_OPTIMAL_COLLAPSE): G(i, k) = cheapest cover of BVH2 subtree i by at most k roots,
	// a root being a leaf or a 4-wide node (area * C_trav + cheapest cover of its two subtrees by 4 roots); since every
	// BVH2 leaf ends up as a leaf slot either way, this minimises the summed area of the 4-wide nodes.  Against the greedy
	// rule "open the child with the largest area" on Sponza: 72.7 k instead of 86.0 k nodes, 4.8 % fewer node visits per
	// bounce ray for 2.7 % more triangle tests — measured on the B200: frame 13.81 vs 13.88 ms on Sponza, but 10.95 vs
	// 10.57 ms on the 10 M-triangle scene (profiles/r01b/collapse_sweep.jsonl), so greedy stays the default.
	struct Dp
	{
		float g[5];
		uint8_t left_share[5]; // roots given to the left subtree when the node is opened with k roots
		bool as_one[5];		   // with k roots available the node is still cheapest as ONE 4-wide node
		bool as_leaf;		   // an inner BVH2 node whose <= 4 references are cheapest as one leaf slot of its parent
	};
#ifdef RFW_OPTIMAL_COLLAPSE
	const bool optimal_collapse = true;
#else
	const bool optimal_collapse = false;
#endif
	std::vector<Dp> dp;
	if (optimal_collapse)
	{
		const int32_t nn2 = b.next_node.load();
		dp.resize(size_t(nn2));
		std::vector<uint32_t> sub(size_t(nn2), 0);
		for (int32_t i = nn2; i-- > 0;) // children are allocated after their parents
		{
			Dp &d = dp[i];
			const Node2 &nd = n2[i];
			d.as_leaf = false;
			sub[i] = nd.leaf ? uint32_t(std::max<size_t>(nd.prims.size(), 1)) : (sub[nd.left] + sub[nd.right]);
			if (nd.leaf)
			{
				for (int k = 1; k <= 4; k++)
					d.g[k] = nd.box.area() * float(std::max<size_t>(nd.prims.size(), 1)), d.left_share[k] = 0, d.as_one[k] = true;
				continue;
			}
			const Dp &l = dp[nd.left], &r = dp[nd.right];
			float f[5] = {3e38f, 3e38f, 3e38f, 3e38f, 3e38f};
			for (int k = 2; k <= 4; k++)
				for (int a = 1; a < k; a++)
					if (l.g[a] + r.g[k - a] < f[k])
						f[k] = l.g[a] + r.g[k - a], d.left_share[k] = uint8_t(a);
			const float one = nd.box.area() * RFW_COLLAPSE_CTRAV + f[4];
			d.g[1] = one, d.as_one[1] = true, d.left_share[1] = d.left_share[4];
			for (int k = 2; k <= 4; k++)
				d.as_one[k] = one <= f[k], d.g[k] = std::min(one, f[k]);
#ifndef RFW_NO_LEAF_MERGE
			// the wide node pays per visit, not per triangle: a small subtree can be cheaper as one leaf slot
			const float as_leaf_cost = nd.box.area() * float(sub[i]);
			if (sub[i] <= uint32_t(MAX_LEAF) && as_leaf_cost <= d.g[4])
			{
				d.as_leaf = true;
				for (int k = 1; k <= 4; k++)
					d.g[k] = as_leaf_cost, d.as_one[k] = true;
			}
#endif
		}
	}
	// references of a (merged) leaf: the two halves of a spatially split triangle collapse into one reference
	std::vector<uint32_t> leaf_prims;
	std::vector<Box> leaf_boxes;
	auto gather_leaf = [&](int32_t root) {
		leaf_prims.clear(), leaf_boxes.clear();
		int32_t st[16];
		int sp = 0;
		st[sp++] = root;
		while (sp)
		{
			const Node2 &c = n2[st[--sp]];
			if (!c.leaf)
			{
				st[sp++] = c.right, st[sp++] = c.left;
				continue;
			}
			for (size_t pi = 0; pi < c.prims.size(); pi++)
			{
				const auto it = std::find(leaf_prims.begin(), leaf_prims.end(), c.prims[pi]);
				if (it == leaf_prims.end())
					leaf_prims.push_back(c.prims[pi]), leaf_boxes.push_back(c.prim_boxes[pi]);
				else
					leaf_boxes[size_t(it - leaf_prims.begin())].grow(c.prim_boxes[pi]);
			}
		}
	};
	struct Pending
	{
		int32_t n2;
		uint32_t parent;
	};
	std::vector<Pending> fifo;
	fifo.reserve(count);
	out.nodes.reserve(count / 2 + 1);
	out.tri_order.reserve(max_refs);
	fifo.push_back({0, 0xffffffffu});
	float cost = 0;
	const float inv_root_area = 1.0f / b.root_area;
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
		if (src.leaf)
			kids[nk++] = p.n2; // root is a leaf
		else if (optimal_collapse)
		{
			// children chosen by the dynamic programme above: the cheapest way to cover the two subtrees with <= 4 roots
			struct Item
			{
				int32_t n;
				int k;
			};
			Item st[8];
			int sp = 0;
			const int a = dp[p.n2].left_share[4];
			st[sp++] = {src.right, 4 - a}, st[sp++] = {src.left, a};
			while (sp)
			{
				const Item it = st[--sp];
				const Node2 &c = n2[it.n];
				if (c.leaf || dp[it.n].as_leaf || it.k == 1 || dp[it.n].as_one[it.k])
					kids[nk++] = it.n;
				else
				{
					const int la = dp[it.n].left_share[it.k];
					st[sp++] = {c.right, it.k - la}, st[sp++] = {c.left, la};
				}
			}
		}
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
					if (!c.leaf && c.box.area() > best_area)
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
			set_child_box(node, k, padded(c.box));
			if (c.leaf || (optimal_collapse && dp[kids[k]].as_leaf))
			{
				gather_leaf(kids[k]);
				const uint32_t first = uint32_t(out.tri_order.size());
				const uint32_t cnt = uint32_t(std::min<size_t>(std::max<size_t>(leaf_prims.size(), 1), MAX_LEAF));
				if (leaf_prims.empty())
				{
					out.tri_order.push_back(0);
					out.ref_boxes.insert(out.ref_boxes.end(), {0.f, 0.f, 0.f, 0.f, 0.f, 0.f});
				}
				for (size_t i = 0; i < leaf_prims.size(); i++)
				{
					if (i >= size_t(MAX_LEAF))
						break; // cannot happen: leaves are only made at <= MAX_LEAF references
					out.tri_order.push_back(leaf_prims[i]);
					const Box &rb = leaf_boxes[i];
					out.ref_boxes.insert(out.ref_boxes.end(), {rb.lo[0], rb.lo[1], rb.lo[2], rb.hi[0], rb.hi[1], rb.hi[2]});
				}
				node.child[k] = ~int32_t((first << 2) | (cnt - 1));
				cost += c.box.area() * inv_root_area * float(leaf_prims.size());
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
	out.sah_cost = cost;
	out.depth = max_depth;
	out.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

// box of leaf reference i for a refit: whole (moved) triangle, or the builder's clipped box when the triangle has not
// moved since the build
static inline Box refit_ref_box(const BuildTriangle *tris, const BvhBuildResult &bvh, const uint8_t *tri_moved, uint32_t i)
{
	const uint32_t src = bvh.tri_order[i];
	if (tri_moved && !tri_moved[src] && bvh.ref_boxes.size() >= 6 * size_t(i + 1))
	{
		Box b;
		for (int a = 0; a < 3; a++)
			b.lo[a] = bvh.ref_boxes[6 * size_t(i) + a], b.hi[a] = bvh.ref_boxes[6 * size_t(i) + 3 + a];
		return padded(b);
	}
	return padded(tri_box(tris[src]));
}

void refit_bvh4(const BuildTriangle *tris, size_t count, BvhBuildResult &bvh, const uint8_t *tri_moved)
{
	(void)count;
	// Only triangles that moved since the build get whole-triangle boxes; the others keep the builder's reference boxes,
	// so the spatial splits of the static part of a scene survive any number of refits (refitting every reference from
	// its whole triangle made Sponza 7x slower to trace after the first animated frame, profiles/r01b).
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
					b.grow(refit_ref_box(tris, bvh, tri_moved, i));
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


// ------------------------------------------------------------------------------------------------
// compressed 8-wide layout (cwbvh.h)
// ------------------------------------------------------------------------------------------------
void build_cwbvh(const BuildTriangle *tris, size_t count, int threads, BvhBuildResult &out, bool spatial_splits)
{
	const auto t0 = std::chrono::steady_clock::now();
	out = BvhBuildResult();
	out.wide8 = true;
	if (count == 0)
	{
		CwNode root;
		memset(&root, 0, sizeof(root));
		CwAux aux;
		memset(&aux, 0, sizeof(aux));
		cw_quantize(root, aux);
		out.cw_nodes.push_back(root), out.cw_aux.push_back(aux), out.cw_parent_slot.push_back(0xffffffffu);
		return;
	}
	Builder b;
	b.max_leaf = 3;				  // a leaf slot holds 1-3 triangles (unary count in three meta bits)
	b.max_depth2 = CW_MAX_DEPTH; // one stack entry per level at most
	const size_t max_refs = build_tree2(b, tris, count, threads, spatial_splits);
	const std::vector<Node2> &n2 = b.nodes;
	// The SAH builder (C_trav = C_isect) splits down to 1-2 triangles per leaf, which leaves wide nodes half empty.
	// A BVH2 subtree holding <= 3 distinct triangles becomes ONE leaf slot here (the wide node pays per node visit,
	// not per triangle: Ylitie et al. section 3.1 fold the same decision into their collapse cost).
	const int32_t n_nodes2 = b.next_node.load();
	std::vector<uint32_t> sub(size_t(n_nodes2), 0);
	for (int32_t i = n_nodes2; i-- > 0;) // children are allocated after their parents
		sub[i] = n2[i].leaf ? uint32_t(n2[i].prims.size()) : (sub[n2[i].left] + sub[n2[i].right]);
	std::vector<Box> leaf_boxes;
	auto gather = [&](int32_t root, std::vector<uint32_t> &prims) {
		prims.clear();
		leaf_boxes.clear();
		int32_t st[16];
		int sp = 0;
		st[sp++] = root;
		while (sp)
		{
			const Node2 &c = n2[st[--sp]];
			if (c.leaf)
			{
				for (size_t pi = 0; pi < c.prims.size(); pi++)
				{
					const auto it = std::find(prims.begin(), prims.end(), c.prims[pi]);
					if (it == prims.end())
						prims.push_back(c.prims[pi]), leaf_boxes.push_back(c.prim_boxes[pi]);
					else
						leaf_boxes[size_t(it - prims.begin())].grow(c.prim_boxes[pi]); // two halves of one split triangle
				}
			}
			else
				st[sp++] = c.left, st[sp++] = c.right;
		}
	};
	auto is_leaf = [&](int32_t i) { return n2[i].leaf || sub[i] <= 3u; };
	std::vector<uint32_t> leaf_prims;
	struct Pending
	{
		int32_t n2;
		uint32_t parent_slot;
		int depth;
	};
	std::vector<Pending> fifo;
	fifo.reserve(count / 2 + 1);
	out.tri_order.reserve(max_refs);
	fifo.push_back({0, 0xffffffffu, 1});
	float cost = 0;
	const float inv_root_area = 1.0f / b.root_area;
	int max_depth = 1;
	for (size_t head = 0; head < fifo.size(); head++)
	{
		const Pending p = fifo[head];
		const Node2 &src = n2[p.n2];
		int32_t kids[8];
		int nk = 0;
		if (is_leaf(p.n2))
			kids[nk++] = p.n2; // the root is a leaf
		else
		{
			kids[nk++] = src.left, kids[nk++] = src.right;
			while (nk < 8)
			{
				int best = -1;
				float best_area = -1;
				for (int k = 0; k < nk; k++)
				{
					const Node2 &c = n2[kids[k]];
					if (!is_leaf(kids[k]) && c.box.area() > best_area)
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
		// slot assignment: child with the largest projection of (centroid - node centroid) on a slot's octant
		// direction takes that slot, greedily (Ylitie et al. section 3.2 use an auction; greedy is within a few %)
		float nc[3];
		for (int a = 0; a < 3; a++)
			nc[a] = 0.5f * (src.box.lo[a] + src.box.hi[a]);
		int slot_of[8], kid_in[8];
		for (int k = 0; k < 8; k++)
			slot_of[k] = -1, kid_in[k] = -1;
		for (int round = 0; round < nk; round++)
		{
			float best = -3.0e38f;
			int bk = -1, bs = -1;
			for (int k = 0; k < nk; k++)
			{
				if (slot_of[k] >= 0)
					continue;
				const Box &cb = n2[kids[k]].box;
				const float d[3] = {0.5f * (cb.lo[0] + cb.hi[0]) - nc[0], 0.5f * (cb.lo[1] + cb.hi[1]) - nc[1],
									0.5f * (cb.lo[2] + cb.hi[2]) - nc[2]};
				for (int sl = 0; sl < 8; sl++)
				{
					if (kid_in[sl] >= 0)
						continue;
					const float v = ((sl & 4) ? d[0] : -d[0]) + ((sl & 2) ? d[1] : -d[1]) + ((sl & 1) ? d[2] : -d[2]);
					if (v > best)
						best = v, bk = k, bs = sl;
				}
			}
			slot_of[bk] = bs, kid_in[bs] = bk;
		}
		CwNode node;
		memset(&node, 0, sizeof(node));
		CwAux aux;
		memset(&aux, 0, sizeof(aux));
		node.child_base = uint32_t(fifo.size());
		node.tri_base = uint32_t(out.tri_order.size());
		const uint32_t self = uint32_t(head);
		for (int sl = 0; sl < 8; sl++)
		{
			if (kid_in[sl] < 0)
				continue;
			const Node2 &c = n2[kids[kid_in[sl]]];
			const Box pb = padded(c.box);
			for (int a = 0; a < 3; a++)
				aux.lo[a][sl] = pb.lo[a], aux.hi[a][sl] = pb.hi[a];
			if (is_leaf(kids[kid_in[sl]]))
			{
				gather(kids[kid_in[sl]], leaf_prims);
				const uint32_t off = uint32_t(out.tri_order.size()) - node.tri_base;
				const uint32_t cnt = uint32_t(std::min<size_t>(leaf_prims.size(), 3));
				if (cnt == 0)
					continue; // cannot happen: the builder never makes empty leaves
				for (uint32_t i = 0; i < cnt; i++)
				{
					out.tri_order.push_back(leaf_prims[i]);
					const Box &rb = leaf_boxes[i];
					out.ref_boxes.insert(out.ref_boxes.end(), {rb.lo[0], rb.lo[1], rb.lo[2], rb.hi[0], rb.hi[1], rb.hi[2]});
				}
				node.meta[sl] = uint8_t((((1u << cnt) - 1u) << 5) | off);
				cost += c.box.area() * inv_root_area * float(cnt);
			}
			else
			{
				node.meta[sl] = uint8_t(0x20u | (24u + uint32_t(sl)));
				node.imask |= uint8_t(1u << sl);
				fifo.push_back({kids[kid_in[sl]], (self << 3) | uint32_t(sl), p.depth + 1});
				max_depth = std::max(max_depth, p.depth + 1);
			}
		}
		cw_quantize(node, aux);
		out.cw_nodes.push_back(node), out.cw_aux.push_back(aux), out.cw_parent_slot.push_back(p.parent_slot);
	}
	out.sah_cost = cost;
	out.depth = max_depth;
	out.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

void refit_cwbvh(const BuildTriangle *tris, size_t count, BvhBuildResult &bvh, const uint8_t *tri_moved)
{
	(void)count;
	// children are stored after their parents (breadth-first): walk backwards so a child's boxes are final before its
	// parent reads them; boxes come from whole triangles, as in refit_bvh4
	for (size_t ni = bvh.cw_nodes.size(); ni-- > 0;)
	{
		CwNode &n = bvh.cw_nodes[ni];
		CwAux &a = bvh.cw_aux[ni];
		uint32_t inner_rank = 0;
		for (int s = 0; s < 8; s++)
		{
			const uint32_t meta = n.meta[s];
			if (!meta)
				continue;
			Box b;
			b.reset();
			if ((meta & 0x18u) == 0x18u)
			{
				const size_t ci = size_t(n.child_base) + inner_rank++;
				const CwNode &cn = bvh.cw_nodes[ci];
				const CwAux &ca = bvh.cw_aux[ci];
				for (int k = 0; k < 8; k++)
					if (cn.meta[k])
						for (int ax = 0; ax < 3; ax++)
							b.lo[ax] = std::min(b.lo[ax], ca.lo[ax][k]), b.hi[ax] = std::max(b.hi[ax], ca.hi[ax][k]);
			}
			else
			{
				const uint32_t first = n.tri_base + (meta & 31u), cnt = uint32_t(__builtin_popcount(meta >> 5));
				for (uint32_t i = first; i < first + cnt; i++)
					b.grow(refit_ref_box(tris, bvh, tri_moved, i));
			}
			for (int ax = 0; ax < 3; ax++)
				a.lo[ax][s] = b.lo[ax], a.hi[ax][s] = b.hi[ax];
		}
		cw_quantize(n, a);
	}
}

// ---- top level of a two-level scene ---------------------------------------------------------------------------------------
// A 4-wide tree over item boxes (instances) whose leaf slots name ONE item each (leaf word ~(item << 2)): binned SAH on the
// box centres, two binary splits per node.  The reference builds its top level the same way over instance boxes
// (RFW/system/bvh/src/top_level_bvh.cpp:17-102), with one instance per leaf as well.
namespace
{
struct TlasBuild
{
	const float *boxes;
	std::vector<BvhNode4> *nodes;
	int depth = 0;

	Box item_box(uint32_t i) const
	{
		Box b;
		for (int a = 0; a < 3; a++)
			b.lo[a] = boxes[6 * i + a], b.hi[a] = boxes[6 * i + 3 + a];
		return b;
	}
	// splits items[lo, hi) in two non-empty halves, returns the split position
	size_t split(std::vector<uint32_t> &items, size_t lo, size_t hi) const
	{
		const size_t n = hi - lo;
		Box cb;
		cb.reset();
		for (size_t i = lo; i < hi; i++)
		{
			const Box b = item_box(items[i]);
			const float c[3] = {0.5f * (b.lo[0] + b.hi[0]), 0.5f * (b.lo[1] + b.hi[1]), 0.5f * (b.lo[2] + b.hi[2])};
			cb.grow(c);
		}
		constexpr int BINS = 16;
		float best = 3.0e38f;
		int best_axis = -1, best_bin = 0;
		for (int a = 0; a < 3; a++)
		{
			const float ext = cb.hi[a] - cb.lo[a];
			if (!(ext > 0.0f))
				continue;
			Box bb[BINS];
			size_t bc[BINS] = {};
			for (int k = 0; k < BINS; k++)
				bb[k].reset();
			const float sc = float(BINS) / ext;
			for (size_t i = lo; i < hi; i++)
			{
				const Box b = item_box(items[i]);
				const int k = std::min(BINS - 1, std::max(0, int((0.5f * (b.lo[a] + b.hi[a]) - cb.lo[a]) * sc)));
				bb[k].grow(b), bc[k]++;
			}
			float right_area[BINS];
			Box acc;
			acc.reset();
			size_t cnt = 0;
			for (int k = BINS - 1; k > 0; k--)
				acc.grow(bb[k]), right_area[k] = acc.area();
			acc.reset();
			size_t right = n;
			for (int k = 0; k < BINS - 1; k++)
			{
				acc.grow(bb[k]), cnt += bc[k], right = n - cnt;
				if (cnt == 0 || right == 0)
					continue;
				const float cost = acc.area() * float(cnt) + right_area[k + 1] * float(right);
				if (cost < best)
					best = cost, best_axis = a, best_bin = k;
			}
		}
		if (best_axis < 0) // all centres coincide: split the run in the middle
			return lo + n / 2;
		const int a = best_axis;
		const float sc = float(BINS) / (cb.hi[a] - cb.lo[a]);
		const auto mid = std::partition(items.begin() + lo, items.begin() + hi, [&](uint32_t it) {
			const Box b = item_box(it);
			return std::min(BINS - 1, std::max(0, int((0.5f * (b.lo[a] + b.hi[a]) - cb.lo[a]) * sc))) <= best_bin;
		});
		const size_t m = size_t(mid - items.begin());
		return (m == lo || m == hi) ? lo + n / 2 : m;
	}
	uint32_t build(std::vector<uint32_t> &items, size_t lo, size_t hi, int level)
	{
		depth = std::max(depth, level);
		const uint32_t me = uint32_t(nodes->size());
		nodes->emplace_back();
		// up to four groups: split once, then the larger-area halves again
		size_t cut[5] = {lo, hi, hi, hi, hi};
		int groups = 1;
		while (groups < 4)
		{
			int pick = -1;
			float pick_area = -1.0f;
			for (int g = 0; g < groups; g++)
			{
				if (cut[g + 1] - cut[g] < 2)
					continue;
				Box b;
				b.reset();
				for (size_t i = cut[g]; i < cut[g + 1]; i++)
					b.grow(item_box(items[i]));
				if (b.area() > pick_area)
					pick_area = b.area(), pick = g;
			}
			if (pick < 0)
				break;
			const size_t m = split(items, cut[pick], cut[pick + 1]);
			for (int g = groups; g > pick; g--)
				cut[g + 1] = cut[g];
			cut[pick + 1] = m;
			groups++;
		}
		BvhNode4 n;
		memset(&n, 0, sizeof(n));
		for (int s = 0; s < 4; s++)
			set_child_empty(n, s);
		for (int g = 0; g < groups; g++)
		{
			Box b;
			b.reset();
			for (size_t i = cut[g]; i < cut[g + 1]; i++)
				b.grow(item_box(items[i]));
			n.minx[g] = b.lo[0], n.miny[g] = b.lo[1], n.minz[g] = b.lo[2];
			n.maxx[g] = b.hi[0], n.maxy[g] = b.hi[1], n.maxz[g] = b.hi[2];
			if (cut[g + 1] - cut[g] == 1)
				n.child[g] = ~int32_t(items[cut[g]] << 2);
			else
				n.child[g] = int32_t(build(items, cut[g], cut[g + 1], level + 1));
		}
		n.pad[0] = groups;
		(*nodes)[me] = n;
		return me;
	}
};
} // namespace

int build_tlas4(const float *boxes, size_t count, std::vector<BvhNode4> &nodes)
{
	nodes.clear();
	TlasBuild b;
	b.boxes = boxes, b.nodes = &nodes;
	if (count == 0)
	{
		BvhNode4 root;
		memset(&root, 0, sizeof(root));
		for (int s = 0; s < 4; s++)
			set_child_empty(root, s);
		nodes.push_back(root);
		return 1;
	}
	std::vector<uint32_t> items(count);
	for (size_t i = 0; i < count; i++)
		items[i] = uint32_t(i);
	b.build(items, 0, count, 1);
	return b.depth;
}

} // namespace rfwb200
