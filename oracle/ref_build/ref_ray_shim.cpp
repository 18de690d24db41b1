// ref_ray_shim.cpp — TEST INFRASTRUCTURE.  The reference's scalar specification of the generate stage,
// Ray::CameraParams::CameraParams and Ray::generateFromView (RFW/backends/EmbreeRT/src/Ray.cpp:3-47), compiled from the
// reference tree.  Ray.cpp as a whole needs Embree, TBB and the GL window headers through its PCH, so the Makefile
// extracts exactly those two definitions (lines 3-47, checked) into a temporary directory for the duration of the compile —
// nothing of the reference's source stays in the tree — and this file supplies the declarations they need: the reference's own Ray.h,
// device_structs.h and rng.h, plus field-compatible stand-ins for the Embree packet structs Ray.h names.
#include <cassert>
#include <cmath>
#include <cstring>
#include <glm/glm.hpp>
#include <glm/ext.hpp>
using namespace glm;
using uint = unsigned int;
#include <rfw/context/device_structs.h>
#include <rfw/utils/rng.h>

template <int N> struct RefRayN
{
	float org_x[N], org_y[N], org_z[N], tnear[N], dir_x[N], dir_y[N], dir_z[N], time[N], tfar[N];
	unsigned mask[N], id[N], flags[N];
};
template <int N> struct RefHitN
{
	float Ng_x[N], Ng_y[N], Ng_z[N], u[N], v[N];
	unsigned primID[N], geomID[N], instID[1][N];
};
template <int N> struct RTCRayHitNt
{
	RefRayN<N> ray;
	RefHitN<N> hit;
};
using RTCRayHit4 = RTCRayHitNt<4>;
using RTCRayHit8 = RTCRayHitNt<8>;
using RTCRayHit16 = RTCRayHitNt<16>;

#include "Ray.h" // RFW/backends/EmbreeRT/src
#include "ray_generate_extract.inc" // temporary (Makefile), = Ray.cpp:3-47

#define REF_API extern "C" __attribute__((visibility("default")))

// view14: pos, p1, p2, p3, aperture, spread (rfw::CameraView); out6: origin, direction
REF_API void rfwref_generate_from_view(const float *view14, int width, int height, int x, int y, float r0, float r1, float r2, float r3,
									   float *out6)
{
	rfw::CameraView view;
	view.pos = vec3(view14[0], view14[1], view14[2]), view.p1 = vec3(view14[3], view14[4], view14[5]);
	view.p2 = vec3(view14[6], view14[7], view14[8]), view.p3 = vec3(view14[9], view14[10], view14[11]);
	view.aperture = view14[12], view.spreadAngle = view14[13];
	const Ray::CameraParams params(view, 0, 1e-5f, uint(width), uint(height));
	const Ray r = Ray::generateFromView(params, x, y, r0, r1, r2, r3);
	out6[0] = r.origin.x, out6[1] = r.origin.y, out6[2] = r.origin.z;
	out6[3] = r.direction.x, out6[4] = r.direction.y, out6[5] = r.direction.z;
}
