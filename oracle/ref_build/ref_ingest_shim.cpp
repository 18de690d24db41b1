// ref_ingest_shim.cpp — TEST INFRASTRUCTURE.  Four rules of the reference's scene ingest (the producers in front of the
// plugin boundary), compiled from where they lie under /root/reference so that the Python bake tools
// (tools/bake_sponza.py, rendering-fw_b200/python/scenes.py) can be pinned on them:
//   texture::construct_mipmaps            RFW/system/src/rfw/texture.cpp:163-209   (5-level box mips, alpha = min)
//   the per-triangle LOD constant         RFW/system/src/rfw/geometry/assimp/object.cpp:728-731
//   the MTL / assimp material rule        RFW/system/src/rfw/material_list.cpp:54-78
//   system::update_area_lights, per light RFW/system/src/rfw/system.cpp:1003-1025  (+ Triangle::updateArea, context.cpp, whole file)
// The files as a whole need FreeImage / assimp / GLFW, so the Makefile extracts exactly those line ranges (first and last
// lines checked) into a temporary directory for the duration of the compile; this file supplies the names they use.
// Never shipped; used by tests/test_ref_pin_ingest.py and tests/golden/make_ref_ingest_golden.py.
#include <cassert>
#include <cmath>
#include <cstring>
#include <vector>
#include <glm/glm.hpp>
#include <glm/ext.hpp>
using namespace glm;
using uint = unsigned int;
#include <rfw/math.h>
#include <rfw/context/settings.h>
#include <rfw/context/structs.h>

#include "context_cpp_extract.inc" // temporary (Makefile) = RFW/system/context/rfw/context/context.cpp: Triangle::calculateArea / updateArea

namespace rfw
{
class texture // stand-in for RFW/system/src/rfw/texture.h: only what construct_mipmaps touches
{
  public:
	uint *udata = nullptr;
	uint width = 0, height = 0, texelCount = 0, mipLevels = 1;
	void construct_mipmaps();
};
} // namespace rfw
#include "mipmaps_extract.inc" // temporary = texture.cpp:163-209

#define REF_API extern "C" __attribute__((visibility("default")))
using namespace rfw;

// level0: width * height RGBA8 texels; out: all 5 levels (texelCount = sum of (w >> l) * (h >> l))
REF_API unsigned rfwref_construct_mipmaps(const unsigned *level0, unsigned width, unsigned height, unsigned *out, unsigned capacity)
{
	unsigned need = 0;
	for (unsigned l = 0, w = width, h = height; l < MIPLEVELCOUNT; l++, w >>= 1u, h >>= 1u)
		need += w * h;
	if (need > capacity)
		return need;
	memcpy(out, level0, size_t(width) * height * 4);
	rfw::texture t;
	t.udata = out, t.width = width, t.height = height, t.texelCount = need;
	t.construct_mipmaps();
	return need;
}

// tri160: rfw::Triangle with u/v/vertex fields set; returns the LOD the loader stores
REF_API float rfwref_triangle_lod(const void *tri160, unsigned tex_width, unsigned tex_height)
{
	Triangle tri;
	memcpy(&tri, tri160, sizeof(tri));
	struct
	{
		uint width, height;
	} texture{tex_width, tex_height};
	{
#include "lod_extract.inc" // temporary = object.cpp:728-731
	}
	return tri.LOD;
}

// in: emissive3, diffuse3, transparent3, opacity, shininess, shininessStrength, eta, reflectivity (what assimp hands over)
// out: color3, absorption3, metallic, subsurface, specular, roughness, eta, transmission (HostMaterial defaults where untouched)
REF_API void rfwref_material_rule(const float *in14, float *out12)
{
	struct C3
	{
		float r, g, b;
	} emissive{in14[0], in14[1], in14[2]}, diffuse{in14[3], in14[4], in14[5]}, transparent{in14[6], in14[7], in14[8]};
	float opacity = in14[9], shininess = in14[10], shininessStrength = in14[11], eta = in14[12], reflectivity = in14[13];
	struct
	{
		vec3 color = vec3(1.0f), absorption = vec3(0.0f);
		float metallic = 0.0f, subsurface = 0.0f, specular = 0.5f, roughness = 0.5f, eta = 1.0f, transmission = 0.0f; // material_list.h:50-62
		void setFlag(int) {}
	} mat;
#include "material_rule_extract.inc" // temporary = material_list.cpp:54-78
	memcpy(out12, &mat.color, 12), memcpy(out12 + 3, &mat.absorption, 12);
	out12[6] = mat.metallic, out12[7] = mat.subsurface, out12[8] = mat.specular, out12[9] = mat.roughness, out12[10] = mat.eta, out12[11] = mat.transmission;
}

// one light of system::update_area_lights: triangle (mesh-local rfw::Triangle), its material colour, the instance matrix
// (column-major mat4, glm layout).  out: the 96-byte rfw::AreaLight + the triangle's updated area
REF_API void rfwref_area_light(const void *tri160, const float *color3, const float *matrix16, int index, int i, void *light96, float *tri_area)
{
	Triangle triangle;
	memcpy(&triangle, tri160, sizeof(triangle));
	struct
	{
		vec3 color;
	} material{vec3(color3[0], color3[1], color3[2])};
	simd::matrix4 transform;
	memcpy(&transform, matrix16, 64);
	const auto normal_transform = transform.inversed().transposed(); // system.cpp:991
	std::vector<AreaLight> m_AreaLights;
	{
#include "area_light_extract.inc" // temporary = system.cpp:1003-1025
	}
	memcpy(light96, &m_AreaLights[0], sizeof(AreaLight));
	*tri_area = triangle.area;
}
