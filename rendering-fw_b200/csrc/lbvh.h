// lbvh.h — per-element steps of the GPU BVH builder (setting "builder" = lbvh), written once as host/device functions:
// the kernels in lbvh.cu run them one element per thread, the host self check (rfwb200_host_lbvh_check, no GPU) runs the
// very same functions in a loop.
//
// Replaces, for scenes whose topology changes (new meshes / instances, or first load of very large scenes), the host-side
// builder the reference delegates to its Rust crate (RFW/system/bvh/src/bvh_tree.cpp:48-102, top_level_bvh.cpp:54-102;
// CUDART rebuilds its TLAS on the CPU every update, CUDART/src/Context.cpp:394-456):
//   1. Morton code of every triangle centroid (30 bits) | triangle index  -> radix sort
//   2. binary radix tree over the sorted keys (Karras, "Maximizing Parallelism in the Construction of BVHs, Octrees, and
//      k-d Trees", HPG 2012): every inner node finds its range and split independently
//   3. boxes bottom-up with one arrival counter per inner node
//      (setting "builder" = ploc replaces 2-3 by parallel locally-ordered clustering over the same sorted references:
//      a bottom-up build that pairs clusters by the surface area of their union, see ploc_* below)
//   4. collapse to the 4-wide BvhNode4 layout level by level (a subtree of <= 4 triangles becomes a leaf: its triangles
//      are contiguous in the sorted order), children allocated contiguously -> the same node format, leaf encoding and
//      parent links as the host builder's output, so the refit / pack / trace kernels do not care who built the tree.
// Quality is that of an LBVH (no SAH, no spatial splits): traversal is slower than on the host-built SBVH; it is the
// fast path for dynamic topology, not the default for static scenes.
#pragma once
#include "device_types.h"

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LB_HD __host__ __device__ __forceinline__
#else
#define LB_HD inline
#endif

namespace rfwb200
{

struct LbvhBox
{
	float lo[3], pad0;
	float hi[3], pad1;
};

// BVH2 produced by steps 2-3: inner nodes 0..n-2 (root 0), leaves 0..n-1 in sorted order; a child word >= 0 is an inner
// node, < 0 is leaf ~child
struct Lbvh2View
{
	const uint64_t *keys; // sorted (morton30 << 32 | triangle index): unique
	int32_t n;			  // triangles
	int32_t *left, *right;		   // [n-1]
	int32_t *first, *last;		   // [n-1] leaf range covered by an inner node
	int32_t *parent_inner;		   // [n-1] parent of an inner node (root: -1)
	int32_t *parent_leaf;		   // [n]
	LbvhBox *inner_box, *leaf_box; // [n-1], [n] (leaf boxes in sorted order, padded like the host builder's)
	uint32_t *arrivals;			   // [n-1]
};

LB_HD uint32_t lb_expand10(uint32_t v)
{
	v = (v * 0x00010001u) & 0xFF0000FFu;
	v = (v * 0x00000101u) & 0x0F00F00Fu;
	v = (v * 0x00000011u) & 0xC30C30C3u;
	v = (v * 0x00000005u) & 0x49249249u;
	return v;
}

// centroid relative to the scene box -> 30-bit Morton code
LB_HD uint32_t lb_morton30(float cx, float cy, float cz, const float *scene_lo, const float *scene_inv_ext)
{
	float f[3] = {(cx - scene_lo[0]) * scene_inv_ext[0], (cy - scene_lo[1]) * scene_inv_ext[1], (cz - scene_lo[2]) * scene_inv_ext[2]};
	uint32_t q[3];
	for (int a = 0; a < 3; a++)
	{
		float v = f[a] * 1024.0f;
		v = v < 0.0f ? 0.0f : (v > 1023.0f ? 1023.0f : v);
		q[a] = uint32_t(v);
	}
	return (lb_expand10(q[0]) << 2) | (lb_expand10(q[1]) << 1) | lb_expand10(q[2]);
}

LB_HD int lb_clz64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
	return __clzll((long long)x);
#else
	return x ? __builtin_clzll(x) : 64;
#endif
}

// length of the common prefix of keys i and j, -1 outside the array
LB_HD int lb_delta(const uint64_t *keys, int n, int i, int j)
{
	if (j < 0 || j >= n)
		return -1;
	return lb_clz64(keys[i] ^ keys[j]);
}

// step 2: inner node i of the radix tree
LB_HD void lb_build_inner(const Lbvh2View &t, int i)
{
	const uint64_t *k = t.keys;
	const int n = t.n;
	const int d = (lb_delta(k, n, i, i + 1) - lb_delta(k, n, i, i - 1)) >= 0 ? 1 : -1;
	const int dmin = lb_delta(k, n, i, i - d);
	int lmax = 2;
	while (lb_delta(k, n, i, i + lmax * d) > dmin)
		lmax *= 2;
	int l = 0;
	for (int s = lmax / 2; s >= 1; s /= 2)
		if (lb_delta(k, n, i, i + (l + s) * d) > dmin)
			l += s;
	const int j = i + l * d;
	const int dnode = lb_delta(k, n, i, j);
	int s = 0;
	int step = l;
	do
	{
		step = (step + 1) >> 1;
		if (lb_delta(k, n, i, i + (s + step) * d) > dnode)
			s += step;
	} while (step > 1);
	const int gamma = i + s * d + (d < 0 ? d : 0);
	const int lo = i < j ? i : j, hi = i < j ? j : i;
	const int32_t lc = (lo == gamma) ? ~gamma : gamma;
	const int32_t rc = (hi == gamma + 1) ? ~(gamma + 1) : (gamma + 1);
	t.left[i] = lc, t.right[i] = rc;
	t.first[i] = lo, t.last[i] = hi;
	if (lc >= 0)
		t.parent_inner[lc] = i;
	else
		t.parent_leaf[~lc] = i;
	if (rc >= 0)
		t.parent_inner[rc] = i;
	else
		t.parent_leaf[~rc] = i;
	if (i == 0)
		t.parent_inner[0] = -1;
}

LB_HD LbvhBox lb_union(const LbvhBox &a, const LbvhBox &b)
{
	LbvhBox r;
	for (int k = 0; k < 3; k++)
	{
		r.lo[k] = fminf(a.lo[k], b.lo[k]);
		r.hi[k] = fmaxf(a.hi[k], b.hi[k]);
	}
	r.pad0 = r.pad1 = 0.0f;
	return r;
}

LB_HD float lb_area(const LbvhBox &b)
{
	const float ex = b.hi[0] - b.lo[0], ey = b.hi[1] - b.lo[1], ez = b.hi[2] - b.lo[2];
	return ex * ey + ey * ez + ez * ex;
}

// ---- steps 2-3, SAH-aware variant (setting "builder" = ploc): parallel locally-ordered clustering ------------------------------
// Meister & Bittner, "Parallel Locally-Ordered Clustering for Bounding Volume Hierarchy Construction" (TVCG 2018).  The
// Morton-sorted references start as one cluster each.  Per round every cluster looks `radius` neighbours to either side in
// the (Morton-ordered) cluster array for the partner with which it forms the smallest box (surface area — the SAH's
// measure); two clusters that chose each other merge into a new inner node; the array is compacted and the round repeats
// until one cluster — the root — is left.  Unlike the radix tree this is a bottom-up agglomerative build driven by box area,
// so a large triangle is paired late with a large cluster instead of wherever its centroid's Morton prefix puts it.
// The merged tree no longer keeps the references of a subtree contiguous in Morton order, which the leaf encoding
// (first, count) needs: afterwards every reference gets its position in a depth-first walk (ploc_leaf_position) and the
// per-reference arrays are permuted to it.
#ifndef RFW_PLOC_RADIUS
#define RFW_PLOC_RADIUS 16
#endif
constexpr int PLOC_RADIUS = RFW_PLOC_RADIUS;

struct PlocRound
{
	int32_t c;			   // clusters at the start of the round
	const int32_t *id;	   // [c] BVH2 node of the cluster: ~reference or inner node
	const LbvhBox *box;	   // [c]
	int32_t *nn;		   // [c] chosen partner
	int32_t *out_id;	   // [c] the round's result before compaction
	LbvhBox *out_box;	   // [c]
	uint32_t *keep;		   // [c] 1: the entry survives the round (merged pairs survive once, at the lower index)
};

LB_HD void ploc_nearest(const PlocRound &p, int i, int radius)
{
	const LbvhBox me = p.box[i];
	const int lo = i - radius < 0 ? 0 : i - radius, hi = i + radius > p.c - 1 ? p.c - 1 : i + radius;
	float best = 3.0e38f;
	int best_j = -1;
	for (int j = lo; j <= hi; j++)
	{
		if (j == i)
			continue;
		const float a = lb_area(lb_union(me, p.box[j]));
		if (a < best) // ties go to the lower index: both sides of a pair see the same order
			best = a, best_j = j;
	}
	p.nn[i] = best_j;
}

// counts[k]: references below inner node k.  `alloc()` returns how many inner nodes were made before this one (atomicAdd on
// the device); node indices run downwards from n - 2 so that the last merge, the root, is node 0 like in the radix tree.
template <typename Alloc>
LB_HD void ploc_merge(const PlocRound &p, const Lbvh2View &t, int32_t *counts, int i, Alloc alloc)
{
	const int j = p.nn[i];
	if (j >= 0 && p.nn[j] == i)
	{
		if (i > j)
		{
			p.keep[i] = 0u; // the pair lives on at index j
			return;
		}
		const int32_t k = int32_t(t.n - 2) - int32_t(alloc());
		const int32_t l = p.id[i], r = p.id[j];
		t.left[k] = l, t.right[k] = r;
		if (l >= 0)
			t.parent_inner[l] = k;
		else
			t.parent_leaf[~l] = k;
		if (r >= 0)
			t.parent_inner[r] = k;
		else
			t.parent_leaf[~r] = k;
		const LbvhBox u = lb_union(p.box[i], p.box[j]);
		t.inner_box[k] = u;
		counts[k] = (l < 0 ? 1 : counts[l]) + (r < 0 ? 1 : counts[r]);
		if (k == 0)
			t.parent_inner[0] = -1;
		p.out_id[i] = k, p.out_box[i] = u, p.keep[i] = 1u;
		return;
	}
	p.out_id[i] = p.id[i], p.out_box[i] = p.box[i], p.keep[i] = 1u;
}

// position of reference `leaf` in a depth-first, left-to-right walk of the finished tree
LB_HD int32_t ploc_leaf_position(const Lbvh2View &t, const int32_t *counts, int leaf)
{
	int32_t pos = 0, child = ~leaf, cur = t.parent_leaf[leaf];
	while (cur >= 0)
	{
		if (t.right[cur] == child)
			pos += t.left[cur] < 0 ? 1 : counts[t.left[cur]];
		child = cur, cur = t.parent_inner[cur];
	}
	return pos;
}

// reference range of inner node k once the references are in depth-first order (children still carry the old labels)
LB_HD void ploc_node_range(const Lbvh2View &t, const int32_t *counts, const int32_t *position, int k)
{
	int32_t c = k;
	while (c >= 0)
		c = t.left[c];
	t.first[k] = position[~c], t.last[k] = position[~c] + counts[k] - 1;
}

// ---- step 0 (optional, setting lbvh_presplit): early split clipping --------------------------------------------------
// A Morton-ordered tree cannot separate a huge triangle from the small ones around it, and Sponza's floor and walls are
// exactly that.  Before the sort, a triangle whose box is longer than `cell` along its longest axis is cut into up to
// LBVH_MAX_PIECES slabs; every slab becomes its own reference (same triangle, box of the clipped polygon), the way the
// host builder's spatial splits do (bvh_build.cpp split_ref) but decided up front instead of by SAH.
constexpr int LBVH_MAX_PIECES = 16;

LB_HD int lb_longest_axis(const LbvhBox &b)
{
	const float ex = b.hi[0] - b.lo[0], ey = b.hi[1] - b.lo[1], ez = b.hi[2] - b.lo[2];
	return ex >= ey ? (ex >= ez ? 0 : 2) : (ey >= ez ? 1 : 2);
}

// number of pieces of a triangle with (unpadded) box b
LB_HD int lb_piece_count(const LbvhBox &b, float cell)
{
	const int a = lb_longest_axis(b);
	const float len = b.hi[a] - b.lo[a];
	if (!(cell > 0.0f) || !(len > cell))
		return 1;
	const float k = ceilf(len / cell);
	return k > float(LBVH_MAX_PIECES) ? LBVH_MAX_PIECES : int(k);
}

// box of the part of triangle (v0, v1, v2) inside the slab lo <= x[axis] <= hi, intersected with the triangle's own box
LB_HD LbvhBox lb_clip_to_slab(const float *v0, const float *v1, const float *v2, const LbvhBox &full, int axis, float lo, float hi)
{
	LbvhBox r;
	for (int k = 0; k < 3; k++)
		r.lo[k] = 3.0e38f, r.hi[k] = -3.0e38f;
	r.pad0 = r.pad1 = 0.0f;
	const float *v[3] = {v0, v1, v2};
	for (int i = 0; i < 3; i++)
	{
		const float *a = v[i], *b = v[(i + 1) % 3];
		const float pa = a[axis], pb = b[axis];
		if (pa >= lo && pa <= hi)
			for (int k = 0; k < 3; k++)
				r.lo[k] = fminf(r.lo[k], a[k]), r.hi[k] = fmaxf(r.hi[k], a[k]);
		for (int side = 0; side < 2; side++)
		{
			const float pl = side ? hi : lo;
			if ((pa < pl && pb > pl) || (pa > pl && pb < pl))
			{
				float t = (pl - pa) / (pb - pa);
				t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
				for (int k = 0; k < 3; k++)
				{
					const float q = k == axis ? pl : a[k] + (b[k] - a[k]) * t;
					r.lo[k] = fminf(r.lo[k], q), r.hi[k] = fmaxf(r.hi[k], q);
				}
			}
		}
	}
	// conservative against rounding of the intersection points: never smaller than the slab says, never larger than the triangle
	for (int k = 0; k < 3; k++)
	{
		if (k == axis)
			r.lo[k] = fminf(r.lo[k], fmaxf(lo, full.lo[k])), r.hi[k] = fmaxf(r.hi[k], fminf(hi, full.hi[k]));
		r.lo[k] = fmaxf(r.lo[k], full.lo[k]), r.hi[k] = fminf(r.hi[k], full.hi[k]);
		if (!(r.lo[k] <= r.hi[k]))
			r.lo[k] = full.lo[k], r.hi[k] = full.hi[k]; // degenerate piece: fall back to the whole box
	}
	return r;
}

// piece j of `pieces` of a triangle: the slab bounds along the longest axis of its box
LB_HD void lb_piece_slab(const LbvhBox &full, int pieces, int j, int &axis, float &lo, float &hi)
{
	axis = lb_longest_axis(full);
	const float w = (full.hi[axis] - full.lo[axis]) / float(pieces);
	lo = j == 0 ? full.lo[axis] : full.lo[axis] + w * float(j);
	hi = j == pieces - 1 ? full.hi[axis] : full.lo[axis] + w * float(j + 1);
}

// one entry of the level queue of step 4
struct LbvhPending
{
	int32_t node2;		  // BVH2 inner node this 4-wide node is made from (or ~leaf when the whole scene is one leaf)
	uint32_t parent_slot; // (parent << 2) | slot, root 0xffffffff
};

#ifndef RFW_LBVH_MAX_LEAF
#define RFW_LBVH_MAX_LEAF 4
#endif
constexpr int LBVH_MAX_LEAF = RFW_LBVH_MAX_LEAF;

LB_HD int lb_count(const Lbvh2View &t, int32_t c) { return c < 0 ? 1 : (t.last[c] - t.first[c] + 1); }
LB_HD int lb_first(const Lbvh2View &t, int32_t c) { return c < 0 ? ~c : t.first[c]; }
LB_HD const LbvhBox &lb_box(const Lbvh2View &t, int32_t c) { return c < 0 ? t.leaf_box[~c] : t.inner_box[c]; }

// step 4: 4-wide node `self` from BVH2 node p.node2.  Children that stay inner are appended to the next level:
// `alloc(count)` returns the position of the first of `count` consecutive entries (atomicAdd on the device).
template <typename Alloc>
LB_HD void lb_collapse_node(const Lbvh2View &t, const LbvhPending &p, uint32_t self, uint32_t next_level_base, BvhNode4 *nodes,
							uint32_t *parent_slot_out, LbvhPending *next_queue, Alloc alloc)
{
	int32_t kids[4];
	int nk = 0;
	if (p.node2 < 0 || lb_count(t, p.node2) <= LBVH_MAX_LEAF)
		kids[nk++] = p.node2; // the whole tree is one leaf
	else
	{
		kids[nk++] = t.left[p.node2], kids[nk++] = t.right[p.node2];
		while (nk < 4)
		{
			int best = -1;
			float best_area = -1.0f;
			for (int k = 0; k < nk; k++)
				if (kids[k] >= 0 && lb_count(t, kids[k]) > LBVH_MAX_LEAF)
				{
					const float a = lb_area(lb_box(t, kids[k]));
					if (a > best_area)
						best_area = a, best = k;
				}
			if (best < 0)
				break;
			const int32_t c = kids[best];
			kids[best] = t.left[c];
			kids[nk++] = t.right[c];
		}
	}
	int n_inner = 0;
	for (int k = 0; k < nk; k++)
		if (kids[k] >= 0 && lb_count(t, kids[k]) > LBVH_MAX_LEAF)
			n_inner++;
	const uint32_t pos = n_inner ? alloc(uint32_t(n_inner)) : 0u;
	BvhNode4 node;
	const float qnan = nanf("");
	int inner_rank = 0;
	for (int k = 0; k < 4; k++)
	{
		if (k >= nk)
		{
			node.minx[k] = node.miny[k] = node.minz[k] = node.maxx[k] = node.maxy[k] = node.maxz[k] = qnan;
			node.child[k] = -1;
			continue;
		}
		const LbvhBox &b = lb_box(t, kids[k]);
		node.minx[k] = b.lo[0], node.miny[k] = b.lo[1], node.minz[k] = b.lo[2];
		node.maxx[k] = b.hi[0], node.maxy[k] = b.hi[1], node.maxz[k] = b.hi[2];
		const int cnt = lb_count(t, kids[k]);
		if (cnt <= LBVH_MAX_LEAF)
			node.child[k] = ~int32_t((uint32_t(lb_first(t, kids[k])) << 2) | uint32_t(cnt - 1));
		else
		{
			const uint32_t q = pos + uint32_t(inner_rank++);
			node.child[k] = int32_t(next_level_base + q);
			next_queue[q].node2 = kids[k];
			next_queue[q].parent_slot = (self << 2) | uint32_t(k);
		}
	}
	node.pad[0] = nk, node.pad[1] = node.pad[2] = node.pad[3] = 0;
	nodes[self] = node;
	parent_slot_out[self] = p.parent_slot;
}

} // namespace rfwb200
