// Layout-compatible stand-in for rfw::bvh::MBVHNode / MBVHTraversal (RFW/system/bvh/include/bvh/mbvh_node.h:20-106).
#pragma once
#include <glm/glm.hpp>
namespace rfw
{
namespace bvh
{
struct MBVHTraversal
{
	int leftFirst;
	int count;
};
struct MBVHNode
{
	glm::vec4 bminx4, bmaxx4, bminy4, bmaxy4, bminz4, bmaxz4;
	glm::ivec4 childs, counts;
};
} // namespace bvh
} // namespace rfw
