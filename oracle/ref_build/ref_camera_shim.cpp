// ref_camera_shim.cpp — TEST INFRASTRUCTURE.  Compiles the reference's own RFW/system/context/rfw/context/Camera.cpp
// (get_view :74-88, calculate_matrix :109-115) from where it lies under /root/reference, against the glm stand-in of
// shims/ and two stand-ins for file / serialisation helpers it names but the pin never runs.  tests/test_ref_pin.py pins
// rfwb200.Camera.get_view (python) and the adapter's restatement on it; never shipped.
#include <cmath>
namespace glm
{
using std::tan;
} // namespace glm
#include <glm/glm.hpp>
#include <glm/ext.hpp>
namespace glm
{
// named by Camera::get_matrix (rasteriser path, not exercised by the pin)
inline float radians(float d) { return d * 0.01745329251994329576923690768489f; }
inline mat4 scale(const mat4 &m, const vec3 &) { return m; }
inline mat4 perspective(float, float, float, float) { return mat4(1.0f); }
inline mat4 lookAt(const vec3 &, const vec3 &, const vec3 &) { return mat4(1.0f); }
inline mat4 operator*(const mat4 &a, const mat4 &) { return a; }
} // namespace glm
#include "Camera.cpp" // found through -I$(REF)/RFW/system/context/rfw/context

#define REF_API extern "C" __attribute__((visibility("default")))

// in: position[3], direction[3] (normalised), fov, focal distance, aperture, width, height; out: pos, p1, p2, p3, aperture, spread (14 floats)
REF_API void rfwref_camera_get_view(const float *position, const float *direction, float fov, float focal_distance, float aperture, int width,
									int height, float *out14)
{
	rfw::Camera cam;
	cam.position = glm::vec3(position[0], position[1], position[2]);
	cam.direction = glm::vec3(direction[0], direction[1], direction[2]);
	cam.FOV = fov, cam.focalDistance = focal_distance, cam.aperture = aperture;
	cam.resize(width, height);
	const rfw::CameraView v = cam.get_view();
	const float r[14] = {v.pos.x, v.pos.y, v.pos.z, v.p1.x, v.p1.y, v.p1.z, v.p2.x, v.p2.y, v.p2.z, v.p3.x, v.p3.y, v.p3.z, v.aperture, v.spreadAngle};
	for (int i = 0; i < 14; i++)
		out14[i] = r[i];
}
