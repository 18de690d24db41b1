#pragma once
#include "glm.hpp"
// <glm/gtc/type_ptr.hpp> / <glm/gtc/constants.hpp> pieces the reference's rfw/math.h uses
namespace glm
{
template <typename T> inline T *value_ptr(tvec2<T> &v) { return &v.x; }
template <typename T> inline T *value_ptr(tvec3<T> &v) { return &v.x; }
template <typename T> inline T *value_ptr(tvec4<T> &v) { return &v.x; }
template <typename T> inline const T *value_ptr(const tvec2<T> &v) { return &v.x; }
template <typename T> inline const T *value_ptr(const tvec3<T> &v) { return &v.x; }
template <typename T> inline const T *value_ptr(const tvec4<T> &v) { return &v.x; }
inline float *value_ptr(mat4 &m) { return &m.c[0].x; }
inline const float *value_ptr(const mat4 &m) { return &m.c[0].x; }
template <typename T> constexpr T half_pi() { return T(1.57079632679489661923132169163975144); }
} // namespace glm
