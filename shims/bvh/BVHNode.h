// Layout-compatible stand-ins for rfw::bvh::{AABB,BVHNode,BVHTraversal} (RFW/system/bvh/include/bvh/aabb.h,
// bvh_node.h:13-52): only the members CUDART/src/CUDAIntersect.h touches.  OUR code (the real headers pull in the
// SIMD math library, TBB and glm proper).
#pragma once
#include <glm/glm.hpp>
namespace rfw
{
namespace bvh
{
struct AABB
{
	float bmin[3], bmax[3];
};
struct BVHTraversal
{
	int nodeIdx;
	BVHTraversal() : nodeIdx(0) {}
	BVHTraversal(int n) : nodeIdx(n) {}
};
struct BVHNode
{
	AABB bounds;
	int left_first;
	int count;
	int get_count() const { return count; }
	int get_left_first() const { return left_first; }
};
} // namespace bvh
} // namespace rfw
