// ref_skin_shim.cpp — TEST INFRASTRUCTURE.  Compiles the reference's own SIMD math (RFW/system/math/src/rfw/math.h:
// simd::matrix4 operator*, operator+, inversed(), operator*(vector4, matrix4), vector4::length) from where it lies under
// /root/reference, against the small glm stand-in of shims/, and runs the reference's skinning loop body
// (RFW/system/src/rfw/geometry/gltf/mesh.cpp:30-45, indexed branch) on caller arrays.  mesh.cpp itself cannot be
// compiled here (TBB, tiny_gltf, the whole rfw:: object model); what it does per vertex is these six lines, restated
// verbatim around the reference's operators, which carry all the arithmetic (matrix blend, 4x4 inverse, row-vector
// product, the 4-component length).  Used by tests/test_ref_pin.py to pin oracle/skinning.py; never shipped.
#include <rfw/math.h>

#include <cstring>

#define REF_API extern "C" __attribute__((visibility("default")))

using namespace rfw;

// joint_matrices: n_joints column-major mat4 (glm memory layout); base_vertices / base_normals: vec4 per vertex (normal w = 0
// as simd::vector4(glm::vec3) constructs it, math.h:780); outputs: vec4 vertices, vec3 normals
REF_API void rfwref_set_pose(const float *joint_matrices, const float *base_vertices, const float *base_normals, const unsigned *joints4,
							 const float *weights4, int vertex_count, float *out_vertices4, float *out_normals3)
{
	const simd::vector4 normal_mask = simd::vector4(_mm_castsi128_ps(_mm_set_epi32(0, ~0, ~0, ~0)));
	(void)normal_mask;
	for (int vIndex = 0; vIndex < vertex_count; vIndex++)
	{
		const unsigned *j4 = joints4 + 4 * vIndex;
		const float *w4 = weights4 + 4 * vIndex;
		simd::matrix4 J[4];
		for (int k = 0; k < 4; k++)
			memcpy(&J[k], joint_matrices + 16 * j4[k], 64);
		// mesh.cpp:35-38
		simd::matrix4 skinMatrix = J[0] * w4[0];
		skinMatrix = skinMatrix + (J[1] * w4[1]);
		skinMatrix = skinMatrix + (J[2] * w4[2]);
		skinMatrix = skinMatrix + (J[3] * w4[3]);
		// mesh.cpp:39-40
		simd::vector4 bv(base_vertices[4 * vIndex], base_vertices[4 * vIndex + 1], base_vertices[4 * vIndex + 2], base_vertices[4 * vIndex + 3]);
		simd::vector4 result = skinMatrix * bv;
		float tmp[4];
		result.write_to(tmp);
		memcpy(out_vertices4 + 4 * vIndex, tmp, 16);
		// mesh.cpp:42-44
		simd::vector4 bn(base_normals[4 * vIndex], base_normals[4 * vIndex + 1], base_normals[4 * vIndex + 2], base_normals[4 * vIndex + 3]);
		result = bn * skinMatrix.inversed();
		result = result / result.length();
		result.write_to(tmp);
		memcpy(out_normals3 + 3 * vIndex, tmp, 12);
	}
}
