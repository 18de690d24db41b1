// Minimal stand-in for the parts of GLM the reference's shading / intersection headers use.
// OUR code (not GLM, not reference code): it exists only so that oracle/ref_build can compile the reference's
// own headers (bsdf/disney.h, bsdf/tools.h, CUDART/src/{lights,getShadingData,CUDAIntersect}.h) where they lie
// under /root/reference, to pin the oracle's arithmetic on the real reference source.  GLM itself is absent
// from this image.  Semantics follow GLM: column-major matrices, component-wise operators,
// normalize(v) = v * inversesqrt(dot(v, v)).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

namespace glm
{
typedef unsigned int uint;

template <typename T> struct tvec2
{
	T x, y;
	tvec2() = default;
	tvec2(T a) : x(a), y(a) {}
	tvec2(T a, T b) : x(a), y(b) {}
	template <typename U> tvec2(const tvec2<U> &o) : x(T(o.x)), y(T(o.y)) {}
	T &operator[](int i) { return (&x)[i]; }
	const T &operator[](int i) const { return (&x)[i]; }
};
template <typename T> struct tvec4;
template <typename T> struct tvec3
{
	union
	{
		struct
		{
			T x, y, z;
		};
		struct
		{
			T r, g, b;
		};
	};
	tvec3() = default;
	tvec3(T a) : x(a), y(a), z(a) {}
	tvec3(T a, T b, T c) : x(a), y(b), z(c) {}
	template <typename A, typename B, typename C> tvec3(A a, B b, C c) : x(T(a)), y(T(b)), z(T(c)) {}
	tvec3(const tvec4<T> &v);
	template <typename U> explicit tvec3(const tvec3<U> &o) : x(T(o.x)), y(T(o.y)), z(T(o.z)) {}
	T &operator[](int i) { return (&x)[i]; }
	const T &operator[](int i) const { return (&x)[i]; }
};
template <typename T> struct tvec4
{
	union
	{
		struct
		{
			T x, y, z, w;
		};
		struct
		{
			T r, g, b, a;
		};
	};
	tvec4() = default;
	tvec4(T s) : x(s), y(s), z(s), w(s) {}
	tvec4(T a_, T b_, T c_, T d_) : x(a_), y(b_), z(c_), w(d_) {}
	template <typename A, typename B, typename C, typename D> tvec4(A a_, B b_, C c_, D d_) : x(T(a_)), y(T(b_)), z(T(c_)), w(T(d_)) {}
	template <typename D> tvec4(const tvec3<T> &v, D d_) : x(v.x), y(v.y), z(v.z), w(T(d_)) {}
	T &operator[](int i) { return (&x)[i]; }
	const T &operator[](int i) const { return (&x)[i]; }
};
template <typename T> tvec3<T>::tvec3(const tvec4<T> &v) : x(v.x), y(v.y), z(v.z) {}

typedef tvec2<float> vec2;
typedef tvec3<float> vec3;
typedef tvec4<float> vec4;
typedef tvec2<uint> uvec2;
typedef tvec3<uint> uvec3;
typedef tvec4<uint> uvec4;
typedef tvec2<int> ivec2;
typedef tvec3<int> ivec3;
typedef tvec4<int> ivec4;
typedef tvec3<bool> bvec3;
typedef tvec4<bool> bvec4;

template <typename T> inline tvec2<T> operator+(const tvec2<T> &a, const tvec2<T> &b) { return {a.x + b.x, a.y + b.y}; }
template <typename T> inline tvec2<T> operator-(const tvec2<T> &a, const tvec2<T> &b) { return {a.x - b.x, a.y - b.y}; }
template <typename T> inline tvec2<T> operator*(const tvec2<T> &a, const tvec2<T> &b) { return {a.x * b.x, a.y * b.y}; }
template <typename T> inline tvec2<T> operator*(const tvec2<T> &a, T s) { return {a.x * s, a.y * s}; }
template <typename T> inline tvec2<T> operator*(T s, const tvec2<T> &a) { return {a.x * s, a.y * s}; }

template <typename T> inline tvec3<T> operator+(const tvec3<T> &a, const tvec3<T> &b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename T> inline tvec3<T> operator-(const tvec3<T> &a, const tvec3<T> &b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename T> inline tvec3<T> operator*(const tvec3<T> &a, const tvec3<T> &b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
template <typename T> inline tvec3<T> operator/(const tvec3<T> &a, const tvec3<T> &b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline vec3 operator*(const vec3 &a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, const vec3 &a) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(const vec3 &a, double s) { return {float(a.x * s), float(a.y * s), float(a.z * s)}; }
inline vec3 operator/(const vec3 &a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline vec3 operator/(float s, const vec3 &a) { return {s / a.x, s / a.y, s / a.z}; }
inline vec3 operator+(const vec3 &a, float s) { return {a.x + s, a.y + s, a.z + s}; }
inline vec3 operator-(const vec3 &a, float s) { return {a.x - s, a.y - s, a.z - s}; }
inline vec3 operator-(const vec3 &a) { return {-a.x, -a.y, -a.z}; }
inline vec3 &operator+=(vec3 &a, const vec3 &b) { return a = a + b; }
inline vec3 &operator-=(vec3 &a, const vec3 &b) { return a = a - b; }
inline vec3 &operator*=(vec3 &a, float s) { return a = a * s; }
inline vec3 &operator*=(vec3 &a, const vec3 &b) { return a = a * b; }
inline uvec3 operator+(const uvec3 &a, const uvec3 &b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }

template <typename T> inline tvec4<T> operator+(const tvec4<T> &a, const tvec4<T> &b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
template <typename T> inline tvec4<T> operator-(const tvec4<T> &a, const tvec4<T> &b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
template <typename T> inline tvec4<T> operator*(const tvec4<T> &a, const tvec4<T> &b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
inline vec4 operator*(const vec4 &a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline vec4 operator*(float s, const vec4 &a) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline vec4 operator-(const vec4 &a, float s) { return {a.x - s, a.y - s, a.z - s, a.w - s}; }
inline vec4 operator+(const vec4 &a, float s) { return {a.x + s, a.y + s, a.z + s, a.w + s}; }
inline vec4 &operator+=(vec4 &a, const vec4 &b) { return a = a + b; }
inline bvec4 operator&&(const bvec4 &a, const bvec4 &b) { return {a.x && b.x, a.y && b.y, a.z && b.z, a.w && b.w}; }

inline float dot(const vec2 &a, const vec2 &b) { return a.x * b.x + a.y * b.y; }
inline float dot(const vec3 &a, const vec3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(const vec4 &a, const vec4 &b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline vec3 cross(const vec3 &a, const vec3 &b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline float length(const vec3 &a) { return std::sqrt(dot(a, a)); }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline vec3 normalize(const vec3 &a) { return a * inversesqrt(dot(a, a)); }
inline vec2 normalize(const vec2 &a) { return a * inversesqrt(dot(a, a)); }
inline vec4 normalize(const vec4 &a) { return a * inversesqrt(dot(a, a)); }
inline vec3 reflect(const vec3 &I, const vec3 &N) { return I - N * dot(N, I) * 2.0f; }

using std::abs;
using std::acos;
using std::atan2;
using std::cos;
using std::exp;
using std::floor;
using std::log;
using std::log2;
using std::sin;
using std::sqrt;
inline float min(float a, float b) { return b < a ? b : a; }
inline float max(float a, float b) { return a < b ? b : a; }
inline int min(int a, int b) { return b < a ? b : a; }
inline int max(int a, int b) { return a < b ? b : a; }
inline uint min(uint a, uint b) { return b < a ? b : a; }
inline uint max(uint a, uint b) { return a < b ? b : a; }
inline double min(double a, double b) { return b < a ? b : a; }
inline double max(double a, double b) { return a < b ? b : a; }
inline float max(float a, double b) { return a < float(b) ? float(b) : a; }
inline float max(double a, float b) { return float(a) < b ? b : float(a); }
inline float min(float a, double b) { return float(b) < a ? float(b) : a; }
inline vec3 min(const vec3 &a, const vec3 &b) { return {min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)}; }
inline vec3 max(const vec3 &a, const vec3 &b) { return {max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)}; }
inline vec4 min(const vec4 &a, const vec4 &b) { return {min(a.x, b.x), min(a.y, b.y), min(a.z, b.z), min(a.w, b.w)}; }
inline vec4 max(const vec4 &a, const vec4 &b) { return {max(a.x, b.x), max(a.y, b.y), max(a.z, b.z), max(a.w, b.w)}; }
template <typename T> inline T clamp(T v, T lo, T hi) { return v < lo ? lo : (hi < v ? hi : v); }
inline float clamp(float v, float lo, float hi) { return min(max(v, lo), hi); }
inline float sign(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }
inline bool isnan(float v) { return std::isnan(v); }
inline bvec3 isnan(const vec3 &v) { return {std::isnan(v.x), std::isnan(v.y), std::isnan(v.z)}; }
inline bool any(const bvec3 &b) { return b.x || b.y || b.z; }
inline bool any(const bvec4 &b) { return b.x || b.y || b.z || b.w; }
inline bvec3 greaterThan(const vec3 &a, const vec3 &b) { return {a.x > b.x, a.y > b.y, a.z > b.z}; }
inline bvec3 lessThan(const vec3 &a, const vec3 &b) { return {a.x < b.x, a.y < b.y, a.z < b.z}; }
inline bvec3 notEqual(const vec3 &a, const vec3 &b) { return {a.x != b.x, a.y != b.y, a.z != b.z}; }
inline bvec4 lessThan(const vec4 &a, const vec4 &b) { return {a.x < b.x, a.y < b.y, a.z < b.z, a.w < b.w}; }
inline bvec4 greaterThanEqual(const vec4 &a, const vec4 &b) { return {a.x >= b.x, a.y >= b.y, a.z >= b.z, a.w >= b.w}; }
inline uint floatBitsToUint(float f)
{
	uint u;
	std::memcpy(&u, &f, 4);
	return u;
}
inline float uintBitsToFloat(uint u)
{
	float f;
	std::memcpy(&f, &u, 4);
	return f;
}
template <typename T> constexpr T pi() { return T(3.14159265358979323846264338327950288); }
template <typename T> constexpr T two_pi() { return T(6.28318530717958647692528676655900576); }
template <typename T> constexpr T one_over_pi() { return T(0.318309886183790671537767526745028724); }

struct mat3
{
	vec3 c[3];
	mat3() = default;
	mat3(float d) { c[0] = vec3(d, 0, 0), c[1] = vec3(0, d, 0), c[2] = vec3(0, 0, d); }
	vec3 &operator[](int i) { return c[i]; }
	const vec3 &operator[](int i) const { return c[i]; }
};
inline vec3 operator*(const mat3 &m, const vec3 &v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z; }
struct mat4
{
	vec4 c[4];
	mat4() = default;
	mat4(float d) { c[0] = vec4(d, 0, 0, 0), c[1] = vec4(0, d, 0, 0), c[2] = vec4(0, 0, d, 0), c[3] = vec4(0, 0, 0, d); }
	vec4 &operator[](int i) { return c[i]; }
	const vec4 &operator[](int i) const { return c[i]; }
};
inline vec4 operator*(const mat4 &m, const vec4 &v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z + m.c[3] * v.w; }
struct mat3x4
{
	vec4 c[3];
	operator mat3() const
	{
		mat3 m;
		for (int i = 0; i < 3; i++)
			m.c[i] = vec3(c[i]);
		return m;
	}
};
} // namespace glm
