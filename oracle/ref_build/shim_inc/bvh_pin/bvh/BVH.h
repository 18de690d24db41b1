// Stand-in for the umbrella header the reference's bvh sources name (<bvh/BVH.h>; on disk it is bvh.h, and it pulls in the tree
// classes that need the un-vendored rtbvh crate): only the node-level headers the pin build compiles.  TEST INFRASTRUCTURE.
#pragma once
#include <rfw/math.h>
namespace glm
{
inline vec3 make_vec3(const float *p) { return vec3(p[0], p[1], p[2]); } // glm/gtc/type_ptr.hpp, absent from the stand-in
} // namespace glm
#include <bvh/aabb.h>
#include <bvh/bvh_node.h>
#include <bvh/mbvh_node.h>
