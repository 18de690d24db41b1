// bvh_build.h — host-side builder of the flattened 4-wide BVH the trace kernels walk.
// Replaces, for this path, what the reference delegates to the un-vendored Rust crate rtbvh
// (RFW/system/bvh/src/bvh_tree.cpp:48-102, mbvh_tree.cpp:26-49) and the in-tree binned-SAH spec
// (RFW/system/bvh/include/bvh/bvh_node.h:136-233) + MBVH collapse (src/mbvh_node.cpp:194-374).
#pragma once
#include "cwbvh.h"
#include "device_types.h"

#include <cstddef>
#include <vector>

namespace rfwb200
{

struct BuildTriangle
{
	float v0[3], v1[3], v2[3]; // world space
};

struct BvhBuildResult
{
	std::vector<BvhNode4> nodes;	   // breadth-first, root = 0
	std::vector<uint32_t> tri_order;   // leaf-ordered position -> input triangle index (spatial splits may
									   // reference a triangle from several leaves, so size() >= triangle count)
	std::vector<uint32_t> node_parent; // for refit: parent index per node (root: 0xffffffff)
	std::vector<float> ref_boxes;	   // 6 floats (lo, hi) per leaf-ordered reference: the builder's box of that reference
									   // (clipped by spatial splits), kept by refits for triangles that did not move
	// compressed 8-wide layout (cwbvh.h), filled by build_cwbvh instead of nodes / node_parent
	bool wide8 = false;
	std::vector<CwNode> cw_nodes;			// breadth-first, root = 0
	std::vector<CwAux> cw_aux;				// full-precision child boxes, for refits
	std::vector<uint32_t> cw_parent_slot;	// (parent << 3) | slot, root 0xffffffff
	float sah_cost = 0;
	int depth = 0;
	double build_ms = 0;
};

// SAH BVH2 with spatial splits (binned object splits + chopped-binning spatial splits, leaves <= 4
// triangles) built with a task pool over `threads` host threads, collapsed to 4-wide by repeatedly
// opening the child with the largest area, then laid out breadth-first.  Max stack need of a
// depth-first traversal is 3*depth+1.
void build_bvh4(const BuildTriangle *tris, size_t count, int threads, BvhBuildResult &out, bool spatial_splits = true);

// Recompute all boxes bottom-up for moved vertices with unchanged topology (the reference's refit,
// bvh_tree.cpp:104-114, top_level_bvh.cpp:46-52).
// `tri_moved` (one byte per input triangle, may be null = all moved): triangles that have not moved since the build keep
// their reference boxes, the others are bounded as whole triangles.
void refit_bvh4(const BuildTriangle *tris, size_t count, BvhBuildResult &bvh, const uint8_t *tri_moved = nullptr);

// The same SBVH collapsed to 8-wide nodes with quantised child boxes (cwbvh.h): leaves of <= 3 triangles, children
// placed in octant-ordered slots, depth <= CW_MAX_DEPTH.
void build_cwbvh(const BuildTriangle *tris, size_t count, int threads, BvhBuildResult &out, bool spatial_splits = true);
void refit_cwbvh(const BuildTriangle *tris, size_t count, BvhBuildResult &bvh, const uint8_t *tri_moved = nullptr);

// Top level of a two-level scene: a 4-wide tree over `count` boxes (6 floats each: lo, hi) whose leaf slots name one box each
// (child word ~(index << 2)); nodes[0] is the root.  Returns the depth of the tree.
int build_tlas4(const float *boxes, size_t count, std::vector<BvhNode4> &nodes);

constexpr int TRAVERSAL_STACK = 96; // ints per thread in the kernels; builder keeps 3*depth+1 below this

} // namespace rfwb200
