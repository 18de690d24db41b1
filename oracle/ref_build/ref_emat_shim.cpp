// ref_emat_shim.cpp — TEST INFRASTRUCTURE.  The material lookup of the reference's CPU renderer,
// Context::retrieve_material (RFW/backends/EmbreeRT/src/Context.cpp:417-476) with the createTangentSpace it calls (:10-17),
// compiled from the reference tree.  Context.cpp as a whole needs Embree, TBB and GL through its PCH, so the Makefile
// extracts exactly those two definitions (line ranges checked) into a temporary directory for the duration of the compile
// and this file supplies what they need: the reference's own structs.h (Triangle, Material, TextureData) and math.h
// (simd::matrix4 / vector4), and a stand-in for the class around the member function (the ShadingData struct and the
// texture list, as declared in EmbreeRT/src/Context.h:35-57).  Used by tests/test_ref_pin.py to pin the E-mode material
// step of the oracle; never shipped.
#include <cmath>
#include <cstring>
#include <vector>
#include <glm/glm.hpp>
#include <glm/ext.hpp>
using namespace glm;
using uint = unsigned int;
#include <rfw/math.h>
#include <rfw/context/structs.h>

namespace rfw
{
class Context // stand-in for EmbreeRT/src/Context.h: only what retrieve_material touches
{
  public:
	struct ShadingData
	{
		glm::vec3 color, N, iN, T, B;
	};
	ShadingData retrieve_material(const Triangle &tri, const Material &material, const glm::vec3 &p, const glm::vec3 bary,
								  const simd::matrix4 &normal_matrix) const;
	std::vector<TextureData> m_Textures;
};
} // namespace rfw

#include "emode_tangent_extract.inc" // temporary (Makefile), = Context.cpp:10-17
using namespace rfw;
#include "emode_material_extract.inc" // temporary (Makefile), = Context.cpp:417-476

#define REF_API extern "C" __attribute__((visibility("default")))

struct RefTexture
{
	int type; // 0 = FLOAT4, 1 = UINT (TextureData::DataType)
	unsigned width, height;
	const void *data;
};

// triangle160: rfw::Triangle; material192: rfw::Material with texaddr0 = index into `textures` (the unpatched id the CPU
// backend keeps); normal16: column-major mat4; out: color3, N3, iN3
REF_API void rfwref_emode_material(const void *triangle160, const void *material192, const RefTexture *textures, int n_textures,
								   const float *bary3, const float *normal16, float *color_out, float *N_out, float *iN_out)
{
	Context ctx;
	for (int i = 0; i < n_textures; i++)
	{
		TextureData t{};
		t.type = textures[i].type == 0 ? TextureData::FLOAT4 : TextureData::UINT;
		t.width = textures[i].width, t.height = textures[i].height;
		t.data = const_cast<void *>(textures[i].data);
		ctx.m_Textures.push_back(t);
	}
	simd::matrix4 nm;
	memcpy(&nm, normal16, 64);
	const auto sd = ctx.retrieve_material(*static_cast<const Triangle *>(triangle160), *static_cast<const Material *>(material192),
										  vec3(0.0f), vec3(bary3[0], bary3[1], bary3[2]), nm);
	memcpy(color_out, &sd.color, 12), memcpy(N_out, &sd.N, 12), memcpy(iN_out, &sd.iN, 12);
}
