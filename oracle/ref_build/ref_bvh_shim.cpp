// ref_bvh_shim.cpp — TEST INFRASTRUCTURE.  The reference's in-tree BVH code compiled from where it lies under /root/reference:
// RFW/system/bvh/include/bvh/{aabb,bvh_node,mbvh_node}.h and RFW/system/bvh/src/{aabb,bvh_node,mbvh_node}.cpp — the binned-SAH
// BVHNode::subdivide / partition templates (bvh_node.h:56-81,136-233) and the 4-wide collapse MBVHNode::merge_nodes / merge_node
// (mbvh_node.cpp:194-374).  The tree CLASSES around them (bvh_tree.cpp, mbvh_tree.cpp) hand the build to the un-vendored Rust crate
// rtbvh today and cannot be built; the node-level code above is the only builder the reference tree itself holds, and it is what
// oracle/rfw_oracle.cpp restates (bvh_partition, bvh_subdivide, mbvh_merge_node(s)).  The few lines of driver below (identity
// primitive order, root = node 0 bounded by calculate_bounds, pool pointers starting at 2 and 1) are the same on both sides.
// Used by tests/test_ref_pin_bvh.py; never shipped.
#include <bvh/BVH.h>

#include <atomic>
#include <cstring>
#include <vector>

// the sources, in place (they include <bvh/BVH.h>, resolved to shim_inc/bvh_pin/bvh/BVH.h)
#include <aabb.cpp>
#include <bvh_node.cpp>
#include <mbvh_node.cpp>

#define REF_API extern "C" __attribute__((visibility("default")))

using namespace rfw::bvh;

static_assert(sizeof(BVHNode) == 32 && sizeof(MBVHNode) == 128 && sizeof(AABB) == 24, "node layouts");

// aabbs6: n x (min3, max3).  nodes_out: capacity max(2n, 2) BVHNode (32 B each), prims_out: n indices, mnodes_out: capacity
// max(2n, 2) MBVHNode (128 B each).  Returns 0, or 1 when the root stayed a leaf (merge_nodes refuses leaves: no MBVH written).
REF_API int rfwref_bvh_build(const float *aabbs6, int n, void *nodes_out, unsigned *prims_out, int *n_nodes, void *mnodes_out, int *n_mnodes)
{
	std::vector<AABB> aabbs(n);
	memcpy(static_cast<void *>(aabbs.data()), aabbs6, size_t(n) * sizeof(AABB));
	const int cap = n * 2 > 2 ? n * 2 : 2;
	std::vector<BVHNode> nodes(cap);
	std::vector<unsigned> prims(n);
	for (int i = 0; i < n; i++)
		prims[i] = unsigned(i);
	nodes[0].set_left_first(0);
	nodes[0].set_count(n);
	nodes[0].calculate_bounds(aabbs.data(), prims.data());
	std::atomic_int pool{2};
	nodes[0].subdivide<9, 32, 3>(aabbs.data(), nodes.data(), prims.data(), 0, pool);
	*n_nodes = pool.load();
	memcpy(nodes_out, static_cast<const void *>(nodes.data()), size_t(pool.load()) * sizeof(BVHNode));
	memcpy(prims_out, prims.data(), size_t(n) * sizeof(unsigned));
	*n_mnodes = 0;
	if (nodes[0].is_leaf())
		return 1;
	std::vector<MBVHNode> mnodes(cap);
	std::atomic_int mpool{1};
	mnodes[0].merge_nodes(nodes[0], rfw::utils::array_proxy<BVHNode>(uint32_t(nodes.size()), nodes.data()), mnodes.data(), mpool);
	*n_mnodes = mpool.load();
	memcpy(mnodes_out, static_cast<const void *>(mnodes.data()), size_t(mpool.load()) * sizeof(MBVHNode));
	return 0;
}
