// ref_shim.cpp — compiles the REFERENCE's own shading / intersection headers, where they lie under
// /root/reference, into oracle/_ref/librfwref.so so that tests can pin the oracle's restatement on the real
// reference arithmetic.  TEST INFRASTRUCTURE.  No reference source is copied: the headers are #included from
// their original location; what is ours is the glm / CUDA stand-in under /shims and the C exports below.
//
// Reference code exercised:
//   RFW/system/context/rfw/bsdf/{tools,compat,disney}.h       (EvaluateBSDF, SampleBSDF, PackNormal, WangHash, ...)
//   RFW/backends/CUDART/src/CUDAIntersect.h                    (intersect_triangle, intersect_quad_node, intersect_mbvh[_shadow])
//   RFW/backends/CUDART/src/getShadingData.h                   (FetchTexel[Trilinear], getShadingData)
//   RFW/backends/CUDART/src/lights.h                           (RandomBarycentrics, RandomPointOnLight, LightPickProb, ...)
// Kernels.cu itself cannot be built by a host compiler (CUDA launch syntax, CUDA-11 surface references), so
// shade_rays / generatePrimaryRay stay restatements in oracle/rfw_oracle.cpp.
#include <cuda_runtime.h> // shim
#include <glm/glm.hpp>	  // shim
using namespace glm;

#include <Structures.h>
#include <DeviceStructures.h>
using namespace rfw;

#include "bvh/BVHNode.h"
#include "bvh/MBVHNode.h"
using namespace rfw::bvh;

#include "CUDAIntersect.h"
#include "getShadingData.h"
#include "bsdf/bsdf.h"
#include "lights.h"

#include <cstdint>
#include <vector>

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API unsigned rfwref_wang_hash(unsigned s) { return WangHash(s); }
REF_API unsigned rfwref_random_int(unsigned *s) { return RandomInt(*s); }
REF_API float rfwref_random_float(unsigned *s) { return RandomFloat(*s); }
REF_API unsigned rfwref_pack_normal(const float *n) { return PackNormal(vec3(n[0], n[1], n[2])); }
REF_API void rfwref_unpack_normal(unsigned p, float *out)
{
	const vec3 n = UnpackNormal(p);
	out[0] = n.x, out[1] = n.y, out[2] = n.z;
}
REF_API float rfwref_blue_noise(const unsigned *table, int x, int y, int s, int d) { return blueNoiseSampler(table, x, y, s, d); }
REF_API float rfwref_survival_probability(const float *c) { return SurvivalProbability(vec3(c[0], c[1], c[2])); }
REF_API void rfwref_tangent_space(const float *n, float *T, float *B)
{
	vec3 t, b;
	createTangentSpace(vec3(n[0], n[1], n[2]), t, b);
	T[0] = t.x, T[1] = t.y, T[2] = t.z, B[0] = b.x, B[1] = b.y, B[2] = b.z;
}
REF_API void rfwref_clamp_intensity(float *v, float c)
{
	vec3 x(v[0], v[1], v[2]);
	clampIntensity(x, c);
	v[0] = x.x, v[1] = x.y, v[2] = x.z;
}

static ShadingData make_sd(const float *color, const float *absorption, const unsigned *params)
{
	ShadingData sd{};
	sd.color = vec3(color[0], color[1], color[2]);
	sd.absorption = absorption ? vec3(absorption[0], absorption[1], absorption[2]) : vec3(0.0f);
	sd.parameters = uvec4(params[0], params[1], params[2], params[3]);
	return sd;
}
// EvaluateBSDF (disney.h:266-272)
REF_API void rfwref_bsdf_eval(const float *color, const unsigned *params, const float *N, const float *wo, const float *wi,
							  float *bsdf_out, float *pdf_out)
{
	const ShadingData sd = make_sd(color, nullptr, params);
	vec3 T, B;
	const vec3 n(N[0], N[1], N[2]);
	createTangentSpace(n, T, B);
	float pdf = 0;
	unsigned seed = 1;
	const vec3 r = EvaluateBSDF(sd, n, T, B, vec3(wo[0], wo[1], wo[2]), vec3(wi[0], wi[1], wi[2]), pdf, seed);
	bsdf_out[0] = r.x, bsdf_out[1] = r.y, bsdf_out[2] = r.z;
	*pdf_out = pdf;
}
// BSDFSample + BSDFEval with explicit r3, r4 (disney.h:188-262,274-280; the argument-evaluation order of
// SampleBSDF's two RandomFloat(seed) calls is compiler-defined, so the randoms are passed in)
REF_API void rfwref_bsdf_sample(const float *color, const float *absorption, const unsigned *params, const float *N,
								const float *wo, float t, int backfacing, float r3, float r4, float *wi_out, float *bsdf_out,
								float *pdf_out)
{
	const ShadingData sd = make_sd(color, absorption, params);
	vec3 T, B, wi(0.0f);
	const vec3 n(N[0], N[1], N[2]);
	createTangentSpace(n, T, B);
	float pdf = 0;
	int type = 0;
	const vec3 w(wo[0], wo[1], wo[2]);
	BSDFSample(sd, T, B, n, w, wi, pdf, type, t, backfacing != 0, r3, r4);
	const vec3 r = BSDFEval(sd, n, w, wi, t, backfacing != 0);
	wi_out[0] = wi.x, wi_out[1] = wi.y, wi_out[2] = wi.z;
	bsdf_out[0] = r.x, bsdf_out[1] = r.y, bsdf_out[2] = r.z;
	*pdf_out = pdf;
}

// intersect_triangle with barycentrics (CUDAIntersect.h:48-94); returns hit flag, t and the reference's area-ratio
// barycentrics (weights of vertex0, vertex1)
REF_API int rfwref_intersect_triangle(const float *org, const float *dir, float tmin, float tmax, const float *p0,
									  const float *p1, const float *p2, float eps, float *t_out, float *bary_out)
{
	float t = tmax;
	vec2 bary(0.0f);
	const bool hit = intersect_triangle(vec3(org[0], org[1], org[2]), vec3(dir[0], dir[1], dir[2]), tmin, &t,
										vec4(p0[0], p0[1], p0[2], 1.0f), vec4(p1[0], p1[1], p1[2], 1.0f),
										vec4(p2[0], p2[1], p2[2], 1.0f), &bary, eps);
	*t_out = t, bary_out[0] = bary.x, bary_out[1] = bary.y;
	return hit ? 1 : 0;
}

// intersect_mbvh / intersect_mbvh_shadow (CUDAIntersect.h:270-322,391-439) over a caller-supplied MBVH (128-byte nodes
// in the reference layout), prim-index array and vertex/index arrays of ONE mesh
REF_API int rfwref_traverse_mbvh(const void *nodes, const unsigned *prim_indices, const float *vertices4,
								 const unsigned *indices3, const float *org, const float *dir, float tmin, float *t_inout,
								 int *prim_out, float *bary_out)
{
	const MBVHNode *n = static_cast<const MBVHNode *>(nodes);
	const vec4 *verts = reinterpret_cast<const vec4 *>(vertices4);
	const vec3 o(org[0], org[1], org[2]), d(dir[0], dir[1], dir[2]);
	vec2 bary(0.0f);
	int prim = -1;
	float t = *t_inout;
	const bool hit = intersect_mbvh(o, d, tmin, &t, &prim, n, prim_indices, [&](unsigned triangleID) {
		const unsigned i0 = indices3 ? indices3[triangleID * 3 + 0] : triangleID * 3 + 0;
		const unsigned i1 = indices3 ? indices3[triangleID * 3 + 1] : triangleID * 3 + 1;
		const unsigned i2 = indices3 ? indices3[triangleID * 3 + 2] : triangleID * 3 + 2;
		return intersect_triangle(o, d, tmin, &t, verts[i0], verts[i1], verts[i2], &bary, 1e-6f);
	});
	*t_inout = t, *prim_out = prim, bary_out[0] = bary.x, bary_out[1] = bary.y;
	return hit ? 1 : 0;
}
REF_API int rfwref_occluded_mbvh(const void *nodes, const unsigned *prim_indices, const float *vertices4,
								 const unsigned *indices3, const float *org, const float *dir, float tmin, float tmax)
{
	const MBVHNode *n = static_cast<const MBVHNode *>(nodes);
	const vec4 *verts = reinterpret_cast<const vec4 *>(vertices4);
	const vec3 o(org[0], org[1], org[2]), d(dir[0], dir[1], dir[2]);
	return intersect_mbvh_shadow(o, d, tmin, tmax, n, prim_indices, [&](unsigned triangleID) {
		const unsigned i0 = indices3 ? indices3[triangleID * 3 + 0] : triangleID * 3 + 0;
		const unsigned i1 = indices3 ? indices3[triangleID * 3 + 1] : triangleID * 3 + 1;
		const unsigned i2 = indices3 ? indices3[triangleID * 3 + 2] : triangleID * 3 + 2;
		float tm = tmax;
		return intersect_triangle(o, d, tmin, &tm, verts[i0], verts[i1], verts[i2], 1e-6f);
	}) ? 1 : 0;
}

// getShadingData (getShadingData.h:100-217) on caller-supplied material / texture pools.  u, v follow the
// REFERENCE (CUDART) convention: weights of vertex0, vertex1.
REF_API void rfwref_get_shading_data(const void *materials192, const unsigned *uint_texels, const void *triangle160,
									 const float *D, float u, float v, float cone_width, const float *normal_matrix9,
									 float *color_out, unsigned *flags_out, float *N_out, float *iN_out)
{
	materials = const_cast<DeviceMaterial *>(static_cast<const DeviceMaterial *>(materials192));
	uintTextures = const_cast<uint *>(uint_texels);
	mat3 m;
	for (int c = 0; c < 3; c++)
		m[c] = vec3(normal_matrix9[3 * c], normal_matrix9[3 * c + 1], normal_matrix9[3 * c + 2]);
	vec3 N, iN, T, B;
	const ShadingData sd = getShadingData(vec3(D[0], D[1], D[2]), u, v, cone_width, *static_cast<const DeviceTriangle *>(triangle160),
										  0, N, iN, T, B, m);
	color_out[0] = sd.color.x, color_out[1] = sd.color.y, color_out[2] = sd.color.z;
	*flags_out = sd.flags;
	N_out[0] = N.x, N_out[1] = N.y, N_out[2] = N.z;
	iN_out[0] = iN.x, iN_out[1] = iN.y, iN_out[2] = iN.z;
}

// lights.h
REF_API void rfwref_set_lights(unsigned na, const void *area, unsigned np, const void *point, unsigned ns, const void *spot,
							   unsigned nd, const void *dir)
{
	areaLights = const_cast<DeviceAreaLight *>(static_cast<const DeviceAreaLight *>(area));
	pointLights = const_cast<DevicePointLight *>(static_cast<const DevicePointLight *>(point));
	spotLights = const_cast<DeviceSpotLight *>(static_cast<const DeviceSpotLight *>(spot));
	directionalLights = const_cast<DeviceDirectionalLight *>(static_cast<const DeviceDirectionalLight *>(dir));
	lightCounts.areaLightCount = na, lightCounts.pointLightCount = np, lightCounts.spotLightCount = ns,
	lightCounts.directionalLightCount = nd;
}
REF_API void rfwref_random_barycentrics(float r0, float *out)
{
	const vec3 b = RandomBarycentrics(r0);
	out[0] = b.x, out[1] = b.y, out[2] = b.z;
}
REF_API void rfwref_random_point_on_light(float r0, float r1, const float *I, const float *N, float *P_out, float *pick_out,
										  float *pdf_out, float *color_out)
{
	float pick = 0, pdf = 0;
	vec3 color(0.0f);
	const vec3 P = RandomPointOnLight(r0, r1, vec3(I[0], I[1], I[2]), vec3(N[0], N[1], N[2]), pick, pdf, color);
	P_out[0] = P.x, P_out[1] = P.y, P_out[2] = P.z;
	*pick_out = pick, *pdf_out = pdf;
	color_out[0] = color.x, color_out[1] = color.y, color_out[2] = color.z;
}
REF_API float rfwref_light_pick_prob(int idx, const float *O, const float *N, const float *I)
{
	return LightPickProb(idx, vec3(O[0], O[1], O[2]), vec3(N[0], N[1], N[2]), vec3(I[0], I[1], I[2]));
}
REF_API float rfwref_light_pdf(const float *D, float t, float area, const float *LN)
{
	return CalculateLightPDF(vec3(D[0], D[1], D[2]), t, area, vec3(LN[0], LN[1], LN[2]));
}
