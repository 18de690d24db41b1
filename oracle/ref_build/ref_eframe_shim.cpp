// ref_eframe_shim.cpp — TEST INFRASTRUCTURE.  The per-pixel body of the frame loop of the reference's CPU renderer,
// Context::render_frame (RFW/backends/EmbreeRT/src/Context.cpp:179-282: sky lookup on a miss, probe, material lookup, the
// "colour > 1 is a light" early-out, the area-light and point-light loops around rtcOccluded1, 0.1 ambient, the pixel write),
// compiled from the reference tree together with the retrieve_material it calls (:417-476, :10-17).  Context.cpp as a whole
// needs Embree, TBB and GL through its PCH, so the Makefile extracts exactly those line ranges (first / last lines checked)
// into a temporary directory for the duration of the compile, and this file supplies what they name:
//   * the reference's own structs.h (Triangle, Material, TextureData, AreaLight, PointLight) and math.h (simd::matrix4);
//   * stand-ins with Embree's field names for the two things Embree OWNS on this path — the ray packet with its hit record
//     (RTCRayHit8: filled by the caller from ITS closest-hit query) and rtcOccluded1 (forwarded to a callback of the caller:
//     "is anything in [tnear, tfar]?", answered by setting tfar = -inf like Embree does);
//   * a stand-in for the class around the member function (the members the extract touches, EmbreeRT/src/Context.h:35-86).
// So everything the reference computes AROUND Embree's two intersection calls runs as the reference wrote it.  Used by
// tests/test_ref_pin.py to pin the E-mode frame of the oracle (SURVEY.md §8 row "E-mode shading"); never shipped.
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>
#include <glm/glm.hpp>
#include <glm/ext.hpp>
namespace glm
{
inline float atan(float y, float x) { return std::atan2(y, x); } // glm::atan(y, x)
} // namespace glm
using namespace glm;
using uint = unsigned int;
#include <rfw/math.h>
#include <rfw/context/structs.h>

#define PACKET_WIDTH 8
constexpr unsigned RTC_INVALID_GEOMETRY_ID = 0xffffffffu;

// Embree's names (rtcore_ray.h), only the fields the extract reads or writes
struct RTCRay8
{
	float org_x[8], org_y[8], org_z[8], tnear[8], dir_x[8], dir_y[8], dir_z[8], time[8], tfar[8];
	unsigned mask[8];
	int id[8]; // the reference keeps the pixel id here (Ray.cpp GenerateRay8)
	unsigned flags[8];
};
struct RTCHit8
{
	float Ng_x[8], Ng_y[8], Ng_z[8], u[8], v[8];
	unsigned primID[8], geomID[8], instID[1][8];
};
struct RTCRayHit8
{
	RTCRay8 ray;
	RTCHit8 hit;
};
struct RTCRay
{
	float org_x, org_y, org_z, tnear, dir_x, dir_y, dir_z, time, tfar;
	unsigned mask, id, flags;
};
struct RTCIntersectContext
{
	int flags;
};
typedef void *RTCScene;
typedef int (*OccludedFn)(void *user, const float *org3, const float *dir3, float tnear, float tfar); // != 0: occluded
static OccludedFn g_occluded = nullptr;
static void *g_user = nullptr;
static inline void rtcOccluded1(RTCScene, RTCIntersectContext *, RTCRay *ray)
{
	const float o[3] = {ray->org_x, ray->org_y, ray->org_z}, d[3] = {ray->dir_x, ray->dir_y, ray->dir_z};
	if (g_occluded(g_user, o, d, ray->tnear, ray->tfar))
		ray->tfar = -std::numeric_limits<float>::infinity(); // Embree's answer for "occluded"
}

namespace rfw
{
struct CPUMesh // EmbreeRT/src/Mesh.h: the one member the extract reads
{
	const rfw::Triangle *triangles = nullptr;
};
class Context // stand-in for EmbreeRT/src/Context.h: only what the extracts touch
{
  public:
	struct ShadingData
	{
		glm::vec3 color, N, iN, T, B;
	};
	ShadingData retrieve_material(const Triangle &tri, const Material &material, const glm::vec3 &p, const glm::vec3 bary,
								  const simd::matrix4 &normal_matrix) const;
	void shade_packet(RTCRayHit8 &packet, int maxPixelID, int probe_id);
	std::vector<PointLight> m_PointLights;
	std::vector<AreaLight> m_AreaLights;
	std::vector<Material> m_Materials;
	std::vector<TextureData> m_Textures;
	std::vector<CPUMesh> m_Meshes;
	RTCScene m_Scene = nullptr;
	std::vector<uint> m_InstanceMesh;
	std::vector<simd::matrix4> m_InverseMatrices;
	int m_SkyboxWidth = 0, m_SkyboxHeight = 0;
	std::vector<glm::vec3> m_Skybox = {glm::vec3(0)};
	glm::vec4 *m_Pixels = nullptr;
	unsigned int m_ProbedInstance = 0, m_ProbedTriangle = 0;
	float m_ProbedDist = -1.0f;
};
} // namespace rfw

#include "emode_tangent_extract.inc" // temporary (Makefile), = Context.cpp:10-17
using namespace rfw;
#include "emode_material_extract.inc" // temporary (Makefile), = Context.cpp:417-476

void Context::shade_packet(RTCRayHit8 &packet, int maxPixelID, int probe_id)
{
	RTCIntersectContext shadow_context{};
#include "emode_pixel_loop_extract.inc" // temporary (Makefile), = Context.cpp:179-282
}

#define REF_API extern "C" __attribute__((visibility("default")))

struct RefTexture
{
	int type; // 0 = FLOAT4, 1 = UINT (TextureData::DataType)
	unsigned width, height;
	const void *data;
};
struct RefEScene
{
	const void *materials192; // rfw::Material[n_materials], texaddr0 = index into textures (the id the CPU backend keeps)
	int n_materials;
	const RefTexture *textures;
	int n_textures;
	const void *const *mesh_triangles160; // per mesh: rfw::Triangle[]
	int n_meshes;
	const unsigned *instance_mesh;	 // per instance
	const float *instance_normal16; // per instance: the matrix the reference calls m_InverseMatrices[inst] (column-major mat4)
	int n_instances;
	const void *area_lights96; // rfw::AreaLight[n_area]
	int n_area;
	const void *point_lights32; // rfw::PointLight[n_point]
	int n_point;
	const float *sky3; // vec3[sky_w * sky_h]
	int sky_w, sky_h;
};

// One packet of eight camera rays with the caller's closest hits in Embree's record -> eight pixels (pixels_out is indexed by the
// packet's ray.id values, < max_pixel_id); probe_out3: (instance, triangle, distance) when one of the rays is the probe pixel.
REF_API void rfwref_emode_shade_packet(const RefEScene *s, RTCRayHit8 *packet, int max_pixel_id, int probe_id, OccludedFn occluded, void *user,
									   float *pixels_out4, float *probe_out3)
{
	Context ctx;
	const Material *mats = static_cast<const Material *>(s->materials192);
	ctx.m_Materials.assign(mats, mats + s->n_materials);
	for (int i = 0; i < s->n_textures; i++)
	{
		TextureData t{};
		t.type = s->textures[i].type == 0 ? TextureData::FLOAT4 : TextureData::UINT;
		t.width = s->textures[i].width, t.height = s->textures[i].height;
		t.data = const_cast<void *>(s->textures[i].data);
		ctx.m_Textures.push_back(t);
	}
	for (int i = 0; i < s->n_meshes; i++)
	{
		CPUMesh m;
		m.triangles = static_cast<const Triangle *>(s->mesh_triangles160[i]);
		ctx.m_Meshes.push_back(m);
	}
	ctx.m_InstanceMesh.assign(s->instance_mesh, s->instance_mesh + s->n_instances);
	ctx.m_InverseMatrices.resize(s->n_instances);
	for (int i = 0; i < s->n_instances; i++)
		memcpy(&ctx.m_InverseMatrices[i], s->instance_normal16 + 16 * i, 64);
	const AreaLight *al = static_cast<const AreaLight *>(s->area_lights96);
	ctx.m_AreaLights.assign(al, al + s->n_area);
	const PointLight *pl = static_cast<const PointLight *>(s->point_lights32);
	ctx.m_PointLights.assign(pl, pl + s->n_point);
	ctx.m_SkyboxWidth = s->sky_w, ctx.m_SkyboxHeight = s->sky_h;
	ctx.m_Skybox.resize(size_t(s->sky_w) * s->sky_h);
	memcpy(ctx.m_Skybox.data(), s->sky3, ctx.m_Skybox.size() * sizeof(glm::vec3));
	ctx.m_Pixels = reinterpret_cast<glm::vec4 *>(pixels_out4);
	g_occluded = occluded, g_user = user;
	ctx.shade_packet(*packet, max_pixel_id, probe_id);
	if (probe_out3 && ctx.m_ProbedDist >= 0.0f) // this packet held the probe pixel
		probe_out3[0] = float(ctx.m_ProbedInstance), probe_out3[1] = float(ctx.m_ProbedTriangle), probe_out3[2] = ctx.m_ProbedDist;
}
