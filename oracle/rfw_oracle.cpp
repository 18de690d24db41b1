// rfw_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A CPU restatement of the reference's (MeirBon/rendering-fw) hot path — generate -> extend ->
// shade(+NEE) -> connect -> compact -> finalize — used ONLY by tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference leg as the checker for the CUDA path in
// rendering-fw_b200/csrc.  Nothing under rendering-fw_b200/ includes, links or calls this file.
//
// PARITY PINNING: the reference holds no tests, golden vectors or fixtures for this path
// and neither of its renderers can be built in this image (Embree, TBB, glm, GL, cargo/rtbvh absent; CUDART
// uses CUDA-11-only surface references and CUDA launch syntax).  What CAN be built is the arithmetic the
// renderers are made of: oracle/ref_build compiles the reference's own headers — bsdf/{disney,tools,compat}.h,
// CUDART/src/{CUDAIntersect,getShadingData,lights}.h — from /root/reference against a small glm/CUDA stand-in
// into oracle/_ref/librfwref.so, and tests/test_ref_pin.py checks this oracle against it (live in the build
// container, and everywhere through the committed vectors tests/golden/ref_vectors.npz): hashes, RNG, normal
// packing, blue noise, Disney eval/sample/pdf, Moller-Trumbore, MBVH closest/any-hit traversal, getShadingData
// incl. trilinear texture fetch, light sampling and pick probabilities.  The glue is pinned as well: the same
// directory compiles the reference's whole CUDA wavefront file, CUDART/src/Kernels.cu, for the HOST (launch syntax
// rewritten at build time, CUDA threads run one after the other; ref_kernels_shim.cpp) into
// oracle/_ref/librfwref_kernels.so, and test_pt_pipeline_matches_reference_cudart_kernels compares this oracle's
// camera rays (bit-exact), primary hits, per-bounce queue sizes (identical) and accumulated image (1e-5) with it on a
// scene that keeps D1-D6 below out of play (vectors committed as tests/golden/ref_kernels_vectors.npz).
// The E-mode image is pinned the same way: oracle/_ref/librfwref_eframe.so is the per-pixel body of the reference's own
// EmbreeRT frame loop (Context::render_frame, EmbreeRT/src/Context.cpp:179-282, with retrieve_material :417-476) compiled
// from the reference tree; Embree's two calls on that path are answered by this oracle's traversal (ref_eframe_shim.cpp),
// everything around them runs as the reference wrote it, and this oracle's E-mode frame equals it bit for bit
// (test_emode_frame_matches_the_reference_frame_loop, vectors tests/golden/ref_eframe_vectors.npz).
// PINNED on reference code: the building blocks, the PT pipeline end to end (generatePrimaryRay, intersect_rays,
// shade_rays control flow, NEE, connect, the bounce loop) and the E-mode frame around Embree's intersection calls.
// The tree builder is pinned on the reference's in-tree node code: RFW/system/bvh {aabb,bvh_node,mbvh_node}.{h,cpp} compile from
// where they lie into oracle/_ref/librfwref_bvh.so (ref_bvh_shim.cpp), and bvh_partition / bvh_subdivide / mbvh_merge_node(s) below
// build the same trees byte for byte (tests/test_ref_pin_bvh.py, vectors tests/golden/ref_bvh_vectors.npz).  (The reference's tree
// CLASSES call the un-vendored Rust crate rtbvh instead of that code today; rtbvh itself is absent.)
// UNPINNED (restated from source, cited line by line, no executable reference): Embree's own traversal (absent; this
// oracle's two-level MBVH stands in its place).
//
// Two image models are restated:
//   PT-mode  = the wavefront estimator of backends/CUDART/src/Kernels.cu (+ getShadingData.h,
//              lights.h, system/context/rfw/bsdf/{disney,tools,compat}.h) driven by the host
//              loop of backends/CUDART/src/Context.cpp:65-159;
//   E-mode   = the image of backends/EmbreeRT/src/Context.cpp:104-300,417-476 (primary
//              visibility + centroid/point direct light + 0.1 ambient).
// Acceleration structure: binned-SAH BVH2 (system/bvh/include/bvh/bvh_node.h:56-233) collapsed
// to the 4-wide MBVH (system/bvh/src/mbvh_node.cpp:194-374), two-level with per-instance ray
// transform (backends/CUDART/src/Kernels.cu:226-303), traversed exactly as
// backends/CUDART/src/CUDAIntersect.h:155-197,270-322,391-439.
//
// Documented deviations from CUDART (all are reference bugs fixed the way its newer
// VulkanRTX/OptiX backends do; SURVEY.md §8a "quirks", appendix B):
//   D1 light-triangle index for MIS is read from u4.w (lightTriIdx), not v4.w
//      (device_structs.h:37 bug; VulkanRTX/shaders/rt_shade.comp:177);
//   D2 alpha-cutout continuation writes the throughput plane (Kernels.cu:643 bug;
//      rt_shade.comp:155);
//   D3 counters->samplesTaken equals sampleIndex while a sample renders
//      (OptiX6Context/src/OptiXContext.cpp:364; CUDART lags one frame after Reset);
//   D4 the two 16-bit barycentrics of a hit record are the weights of vertex 0 and vertex 1 as in CUDART, but computed
//      as (1 - u - v, u) from the Moller-Trumbore u, v instead of the area ratios of CUDAIntersect.h:82-87 (the same
//      weights up to rounding; "cudart_conventions"=on computes the area ratios);
//   D5 SampleBSDF's two RandomFloat(seed) arguments are drawn left to right (r3 then r4);
//      the C++ order is unspecified (bsdf/disney.h:278) — left to right is what nvcc compiles
//      (tests/test_ref_pin.py::test_nvcc_draws_samplebsdf_randoms_left_to_right), so this is the
//      reference's behaviour on a GPU; "bsdf_random_order"=rtl gives g++'s order;
//   D6 float->int conversions saturate (CUDA semantics) where the operand can be inf/NaN.
// D1 and D4 can be undone with the setting "cudart_conventions"=on (the pin test's "rich" case).
// E-mode determinism contract (the reference's shared xor128 is a data race):
//   r0..r3 of pixel p, sample s = first four outputs of xor128 with
//   x = 123456789 ^ WangHash(p*16789 + s*1791), y,z,w = reference defaults; shadow rays use
//   t in (1e-4, dist*(1-1e-4)) so a light's own triangle never shadows its centroid.
//
// Stage-level entries (rfworacle_generate_primary, _trace_closest, _trace_occluded, _shade_stage) expose the functions the
// frames above are made of, one stage at a time, over caller-supplied rays / paths: a frame composed from them equals
// rfworacle_render_frame's bit for bit (tests/test_shade_stage.py), so whatever is checked against a stage entry — the
// product's launches per ray and per path at the benchmarked scale, tests/test_parity_scale_gpu.py — is checked against
// the pinned pipeline.
//
// Build: g++ -O2 -ffp-contract=off -pthread -shared -fPIC (see oracle/Makefile).

#include "../include/rfwb200.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <functional>
#include <mutex>
#include <thread>

namespace
{

// ---------------------------------------------------------------------------------------------
// threading: plain std::thread work sharing (the image's default g++ has no libgomp spec)
// ---------------------------------------------------------------------------------------------
static int g_threads = 0;
static int num_threads()
{
	if (g_threads > 0)
		return g_threads;
	const unsigned hc = std::thread::hardware_concurrency();
	return hc ? int(hc) : 1;
}
template <typename F> static void parallel_for(int64_t n, int64_t chunk, const F &fn)
{
	const int nt = int(std::min<int64_t>(num_threads(), (n + chunk - 1) / std::max<int64_t>(chunk, 1)));
	if (nt <= 1)
	{
		for (int64_t i = 0; i < n; i++)
			fn(i);
		return;
	}
	std::atomic<int64_t> next{0};
	auto worker = [&]() {
		for (;;)
		{
			const int64_t b = next.fetch_add(chunk);
			if (b >= n)
				break;
			const int64_t e = std::min(n, b + chunk);
			for (int64_t i = b; i < e; i++)
				fn(i);
		}
	};
	std::vector<std::thread> th;
	for (int t = 1; t < nt; t++)
		th.emplace_back(worker);
	worker();
	for (auto &t : th)
		t.join();
}
static std::mutex g_probe_mutex;

// ---------------------------------------------------------------------------------------------
// minimal vector math (glm-equivalent operations only)
// ---------------------------------------------------------------------------------------------
struct vec2
{
	float x, y;
};
struct vec3
{
	float x, y, z;
	vec3() : x(0), y(0), z(0) {}
	vec3(float a) : x(a), y(a), z(a) {}
	vec3(float a, float b, float c) : x(a), y(b), z(c) {}
	explicit vec3(const float *p) : x(p[0]), y(p[1]), z(p[2]) {}
	float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
struct vec4
{
	float x, y, z, w;
	vec4() : x(0), y(0), z(0), w(0) {}
	vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
	vec4(const vec3 &v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
};
inline vec3 operator+(const vec3 &a, const vec3 &b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(const vec3 &a, const vec3 &b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(const vec3 &a, const vec3 &b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline vec3 operator*(const vec3 &a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, const vec3 &a) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator/(const vec3 &a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline vec3 operator-(const vec3 &a) { return {-a.x, -a.y, -a.z}; }
inline vec3 &operator+=(vec3 &a, const vec3 &b)
{
	a = a + b;
	return a;
}
inline vec3 &operator*=(vec3 &a, float s)
{
	a = a * s;
	return a;
}
inline vec4 operator+(const vec4 &a, const vec4 &b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline vec4 operator*(const vec4 &a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline float dot(const vec3 &a, const vec3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(const vec3 &a, const vec3 &b)
{
	return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
inline float length(const vec3 &a) { return std::sqrt(dot(a, a)); }
// glm::normalize = v * inversesqrt(dot(v, v))
inline vec3 normalize(const vec3 &a) { return a * (1.0f / std::sqrt(dot(a, a))); }
inline vec3 reflect(const vec3 &I, const vec3 &N) { return I - N * dot(N, I) * 2.0f; }
inline vec3 vmax(const vec3 &a, const vec3 &b) { return {std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)}; }
inline vec3 vmin(const vec3 &a, const vec3 &b) { return {std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)}; }
inline bool any_nan(const vec3 &a) { return std::isnan(a.x) || std::isnan(a.y) || std::isnan(a.z); }
inline float clampf(float v, float lo, float hi) { return std::min(std::max(v, lo), hi); }
inline float signf(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); } // glm::sign

inline uint32_t f2u(float f)
{
	uint32_t u;
	memcpy(&u, &f, 4);
	return u;
}
inline float u2f(uint32_t u)
{
	float f;
	memcpy(&f, &u, 4);
	return f;
}
// D6: CUDA float->uint / float->int conversion (round toward zero, saturating, NaN -> 0)
inline uint32_t cvt_u32(float f)
{
	if (!(f > 0.0f))
		return 0u;
	if (f >= 4294967296.0f)
		return 0xFFFFFFFFu;
	return static_cast<uint32_t>(f);
}
inline int32_t cvt_i32(float f)
{
	if (std::isnan(f))
		return 0;
	if (f >= 2147483648.0f)
		return INT32_MAX;
	if (f <= -2147483648.0f)
		return INT32_MIN;
	return static_cast<int32_t>(f);
}

// IEEE binary16 -> float (half_float::half conversion used by structs.h:88-128)
inline float half2float(uint16_t h)
{
	const uint32_t sign = (h >> 15) & 1u, exp = (h >> 10) & 31u, man = h & 1023u;
	uint32_t bits;
	if (exp == 0)
	{
		if (man == 0)
			bits = sign << 31;
		else
		{
			int e = -1;
			uint32_t m = man;
			do
			{
				e++;
				m <<= 1;
			} while ((m & 1024u) == 0);
			bits = (sign << 31) | (uint32_t(127 - 15 - e) << 23) | ((m & 1023u) << 13);
		}
	}
	else if (exp == 31)
		bits = (sign << 31) | 0x7F800000u | (man << 13);
	else
		bits = (sign << 31) | ((exp + 112u) << 23) | (man << 13);
	return u2f(bits);
}

struct mat4 // column-major, as glm
{
	float m[16];
	vec3 mul_point(const vec3 &p) const
	{
		return {m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
				m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]};
	}
	vec3 mul_dir(const vec3 &p) const
	{
		return {m[0] * p.x + m[4] * p.y + m[8] * p.z, m[1] * p.x + m[5] * p.y + m[9] * p.z,
				m[2] * p.x + m[6] * p.y + m[10] * p.z};
	}
};
struct mat3 // column-major
{
	float m[9];
	vec3 mul(const vec3 &p) const
	{
		return {m[0] * p.x + m[3] * p.y + m[6] * p.z, m[1] * p.x + m[4] * p.y + m[7] * p.z,
				m[2] * p.x + m[5] * p.y + m[8] * p.z};
	}
};

// general 4x4 inverse (glm::inverse, cofactor expansion) in double for a stable oracle
static mat4 inverse(const mat4 &a)
{
	double inv[16], m[16];
	for (int i = 0; i < 16; i++)
		m[i] = a.m[i];
	inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] +
			 m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
	inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] -
			 m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
	inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] +
			 m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
	inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] -
			  m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
	inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] -
			 m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
	inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] +
			 m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
	inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] -
			 m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
	inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] +
			  m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
	inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] +
			 m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
	inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] -
			 m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
	inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] +
			  m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
	inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] -
			  m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
	inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] -
			 m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
	inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] +
			 m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
	inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] -
			  m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
	inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] +
			  m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
	double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
	mat4 r;
	const double id = det != 0.0 ? 1.0 / det : 0.0;
	for (int i = 0; i < 16; i++)
		r.m[i] = float(inv[i] * id);
	return r;
}

// ---------------------------------------------------------------------------------------------
// constants — context/settings.h:3-12, CUDART/src/Kernels.cu:16-23, CUDART/src/Context.cpp:50-51
// ---------------------------------------------------------------------------------------------
constexpr int MIPLEVELCOUNT = 5;
constexpr float MIN_ROUGHNESS = 0.01f;
constexpr int MAX_IS_LIGHTS = 16;
constexpr float T_EPSILON = 1e-6f;
constexpr float GEOMETRY_EPSILON = 1e-5f;
constexpr uint32_t IS_SPECULAR = 1;
constexpr float INVPI = 0.318309886183790671537767526745028724f;
constexpr float PI = 3.14159265358979323846264338327950288f;
constexpr float INV2PI = 0.159154943091895335768883763372514362f;
constexpr float TWOPI = 6.28318530717958647692528676655900576f;

enum MatFlags
{
	HasDiffuseMap = 2,
	HasNormalMap = 3,
	HasSpecularityMap = 4,
	HasRoughnessMap = 5,
	Has2ndNormalMap = 7,
	Has3rdNormalMap = 8,
	Has2ndDiffuseMap = 9,
	Has3rdDiffuseMap = 10,
	HasSmoothNormals = 11,
	HasAlpha = 12,
	HasAlphaMap = 13
};
inline bool has_flag(uint32_t flags, int f) { return (flags & (1u << f)) != 0; }

// ---------------------------------------------------------------------------------------------
// hashing / RNG — bsdf/tools.h:218-235, utils/xor128.h:20-34, utils/rng.h:14
// ---------------------------------------------------------------------------------------------
inline uint32_t WangHash(uint32_t s)
{
	s = (s ^ 61u) ^ (s >> 16);
	s *= 9u;
	s = s ^ (s >> 4);
	s *= 0x27d4eb2du;
	s = s ^ (s >> 15);
	return s;
}
inline uint32_t RandomInt(uint32_t &s)
{
	s ^= s << 13;
	s ^= s >> 17;
	s ^= s << 5;
	return s;
}
inline float RandomFloat(uint32_t &s) { return float(RandomInt(s)) * 2.3283064365387e-10f; }

struct Xor128
{
	uint32_t x = 123456789u, y = 362436069u, z = 521288629u, w = 88675123u;
	uint32_t rand_uint()
	{
		uint32_t t = x ^ (x << 11);
		x = y;
		y = z;
		z = w;
		return w = w ^ (w >> 19) ^ (t ^ (t >> 8));
	}
	float rand() { return float(rand_uint()) * 2.3283064365387e-10f; }
};

// ---------------------------------------------------------------------------------------------
// AABB + BVH2 (binned SAH) + MBVH collapse
// ---------------------------------------------------------------------------------------------
struct AABB
{
	float bmin[3], bmax[3];
	static AABB invalid()
	{
		AABB a;
		for (int i = 0; i < 3; i++)
			a.bmin[i] = 1e34f, a.bmax[i] = -1e34f;
		return a;
	}
	void grow(const vec3 &p)
	{
		bmin[0] = std::min(bmin[0], p.x), bmin[1] = std::min(bmin[1], p.y), bmin[2] = std::min(bmin[2], p.z);
		bmax[0] = std::max(bmax[0], p.x), bmax[1] = std::max(bmax[1], p.y), bmax[2] = std::max(bmax[2], p.z);
	}
	void grow(const AABB &b)
	{
		for (int i = 0; i < 3; i++)
			bmin[i] = std::min(bmin[i], b.bmin[i]), bmax[i] = std::max(bmax[i], b.bmax[i]);
	}
	void offset_by(float o)
	{
		for (int i = 0; i < 3; i++)
			bmin[i] -= o, bmax[i] += o;
	}
	float centroid(int a) const { return (bmin[a] + bmax[a]) * 0.5f; }
	float area() const // half surface area, as system/bvh/src/aabb.cpp
	{
		const float ex = bmax[0] - bmin[0], ey = bmax[1] - bmin[1], ez = bmax[2] - bmin[2];
		const float v = ex * ey + ey * ez + ez * ex;
		return v > 0.0f ? v : 0.0f;
	}
};

struct BVHNode // 32 bytes, bvh_node.h:23-28
{
	AABB bounds;
	int left_first;
	int count; // >= 0: leaf
};

struct MBVHNode // 128 bytes, mbvh_node.h:60-106
{
	float bminx[4], bmaxx[4], bminy[4], bmaxy[4], bminz[4], bmaxz[4];
	int childs[4], counts[4];
	void set_bounds(int i, const AABB &b)
	{
		bminx[i] = b.bmin[0], bminy[i] = b.bmin[1], bminz[i] = b.bmin[2];
		bmaxx[i] = b.bmax[0], bmaxy[i] = b.bmax[1], bmaxz[i] = b.bmax[2];
	}
};

struct Bvh
{
	std::vector<BVHNode> nodes;
	std::vector<MBVHNode> mnodes;
	std::vector<uint32_t> prim_indices;
	AABB root_bounds;
};

// bvh_node.h:136-233 — evaluate 10 planes per axis (BINS=9 -> i in 1..10 of 11 slices), full
// scans, cost = area*count, keep when not worse than the parent; child boxes padded by 1e-5.
static bool bvh_partition(Bvh &bvh, int node_idx, const AABB *aabbs, int &pool_ptr)
{
	constexpr int BINS = 9;
	BVHNode &node = bvh.nodes[node_idx];
	uint32_t *prim = bvh.prim_indices.data();
	const int lFirst = node.left_first;
	const int count = node.count;
	float lowest = 1e34f, best_split = 0;
	int best_axis = 0;
	AABB best_l = AABB::invalid(), best_r = AABB::invalid();
	const float parent_cost = node.bounds.area() * float(count);
	const float lengths[3] = {node.bounds.bmax[0] - node.bounds.bmin[0], node.bounds.bmax[1] - node.bounds.bmin[1],
							  node.bounds.bmax[2] - node.bounds.bmin[2]};
	const float bin_size = 1.0f / float(BINS + 2);
	for (int axis = 0; axis < 3; axis++)
	{
		for (int i = 1; i < BINS + 2; i++)
		{
			const float split = node.bounds.bmin[axis] + lengths[axis] * (float(i) * bin_size);
			int lc = 0, rc = 0;
			AABB lb = AABB::invalid(), rb = AABB::invalid();
			for (int k = 0; k < count; k++)
			{
				const AABB &a = aabbs[prim[lFirst + k]];
				if (a.centroid(axis) <= split)
					lb.grow(a), lc++;
				else
					rb.grow(a), rc++;
			}
			const float cost = lb.area() * float(lc) + rb.area() * float(rc);
			if (lowest > cost)
				lowest = cost, best_split = split, best_axis = axis, best_l = lb, best_r = rb;
		}
	}
	if (parent_cost < lowest)
		return false;
	int lCount = 0;
	for (int k = 0; k < count; k++)
	{
		const AABB &a = aabbs[prim[lFirst + k]];
		if (a.centroid(best_axis) <= best_split)
		{
			std::swap(prim[lFirst + k], prim[lFirst + lCount]);
			lCount++;
		}
	}
	// a degenerate plane that moves nothing would recurse forever on the same set
	if (lCount == 0 || lCount == count)
		return false;
	const int left = pool_ptr;
	pool_ptr += 2;
	best_l.offset_by(1e-5f);
	best_r.offset_by(1e-5f);
	bvh.nodes[left] = {best_l, lFirst, lCount};
	bvh.nodes[left + 1] = {best_r, lFirst + lCount, count - lCount};
	bvh.nodes[node_idx].left_first = left;
	bvh.nodes[node_idx].count = -1;
	return true;
}

// bvh_node.h:56-81 (MAX_DEPTH 32, leaf when count < 3)
static void bvh_subdivide(Bvh &bvh, int node_idx, const AABB *aabbs, int depth, int &pool_ptr)
{
	depth++;
	if (bvh.nodes[node_idx].count < 3 || depth >= 32)
		return;
	if (!bvh_partition(bvh, node_idx, aabbs, pool_ptr))
		return;
	const int left = bvh.nodes[node_idx].left_first;
	if (bvh.nodes[left].count > 0)
		bvh_subdivide(bvh, left, aabbs, depth, pool_ptr);
	if (bvh.nodes[left + 1].count > 0)
		bvh_subdivide(bvh, left + 1, aabbs, depth, pool_ptr);
}

// mbvh_node.cpp:232-374 merge_node: pull the four grandchildren up; when only three children
// result and one is an inner node, split that one once more.
static void mbvh_merge_node(const Bvh &bvh, const BVHNode &node, MBVHNode &out, int &num)
{
	for (int i = 0; i < 4; i++)
		out.childs[i] = -1, out.counts[i] = -1;
	num = 0;
	auto put = [&](int slot, const BVHNode &n, int self_idx) {
		out.set_bounds(slot, n.bounds);
		if (n.count >= 0)
			out.childs[slot] = n.left_first, out.counts[slot] = n.count;
		else
			out.childs[slot] = self_idx, out.counts[slot] = -1;
	};
	for (int side = 0; side < 2; side++)
	{
		const int ci = node.left_first + side;
		const BVHNode &c = bvh.nodes[ci];
		if (c.count >= 0)
		{
			const int s = num++;
			out.childs[s] = c.left_first, out.counts[s] = c.count;
			out.set_bounds(s, c.bounds);
		}
		else
		{
			const int s1 = num++, s2 = num++;
			put(s1, bvh.nodes[c.left_first], c.left_first);
			put(s2, bvh.nodes[c.left_first + 1], c.left_first + 1);
		}
	}
	if (num == 3)
	{
		for (int i = 0; i < 3; i++)
		{
			if (out.counts[i] >= 0)
				continue;
			const int left = out.childs[i];
			put(i, bvh.nodes[left], left);
			put(num, bvh.nodes[left + 1], left + 1);
			num++;
			break;
		}
	}
}

// mbvh_node.cpp:194-230 merge_nodes
static void mbvh_merge_nodes(Bvh &bvh, int mnode_idx, const BVHNode &node, int &pool_ptr)
{
	int num = 0;
	{
		MBVHNode tmp;
		mbvh_merge_node(bvh, node, tmp, num);
		// unused slots keep the constructor's inverted bounds (mbvh_node.h:46-57)
		for (int i = num; i < 4; i++)
		{
			AABB inv = AABB::invalid();
			tmp.set_bounds(i, inv);
		}
		bvh.mnodes[mnode_idx] = tmp;
	}
	for (int idx = 0; idx < num; idx++)
	{
		MBVHNode &self = bvh.mnodes[mnode_idx];
		if (self.childs[idx] < 0)
		{
			AABB inv = AABB::invalid();
			self.set_bounds(idx, inv);
			self.childs[idx] = 0, self.counts[idx] = 0;
			continue;
		}
		if (self.counts[idx] < 0)
		{
			const BVHNode cur = bvh.nodes[self.childs[idx]];
			const int new_idx = pool_ptr++;
			bvh.mnodes[mnode_idx].childs[idx] = new_idx;
			mbvh_merge_nodes(bvh, new_idx, cur, pool_ptr);
		}
	}
}

static void build_bvh(Bvh &bvh, const std::vector<AABB> &aabbs)
{
	const int n = int(aabbs.size());
	bvh.prim_indices.resize(n);
	for (int i = 0; i < n; i++)
		bvh.prim_indices[i] = uint32_t(i);
	bvh.nodes.assign(std::max(2 * n, 2), BVHNode{AABB::invalid(), 0, 0});
	AABB root = AABB::invalid();
	for (int i = 0; i < n; i++)
		root.grow(aabbs[i]);
	bvh.nodes[0] = {root, 0, n};
	bvh.root_bounds = root;
	int pool = 2;
	if (n > 0)
		bvh_subdivide(bvh, 0, aabbs.data(), 0, pool);
	bvh.nodes.resize(pool);
	// MBVH (mbvh_tree.cpp constructs from the BVH root; a leaf root becomes one quad node with
	// a single leaf child)
	bvh.mnodes.assign(std::max(pool, 1), MBVHNode{});
	int mpool = 1;
	if (bvh.nodes[0].count >= 0)
	{
		MBVHNode m;
		for (int i = 0; i < 4; i++)
		{
			AABB inv = AABB::invalid();
			m.set_bounds(i, inv);
			m.childs[i] = 0, m.counts[i] = 0;
		}
		m.set_bounds(0, bvh.nodes[0].bounds);
		m.childs[0] = bvh.nodes[0].left_first, m.counts[0] = bvh.nodes[0].count;
		bvh.mnodes[0] = m;
	}
	else
		mbvh_merge_nodes(bvh, 0, bvh.nodes[0], mpool);
	bvh.mnodes.resize(mpool);
}

// CUDAIntersect.h:155-197 intersect_quad_node
struct MBVHHit
{
	float tmin[4];
	bool result[4];
};
inline MBVHHit intersect_quad_node(const MBVHNode &n, const vec3 &org, const vec3 &dirInv, float t)
{
	MBVHHit hit;

	for (int i = 0; i < 4; i++)
	{
		float t1 = (n.bminx[i] - org.x) * dirInv.x, t2 = (n.bmaxx[i] - org.x) * dirInv.x;
		float tmn = std::min(t1, t2), tmx = std::max(t1, t2);
		t1 = (n.bminy[i] - org.y) * dirInv.y, t2 = (n.bmaxy[i] - org.y) * dirInv.y;
		tmn = std::max(tmn, std::min(t1, t2)), tmx = std::min(tmx, std::max(t1, t2));
		t1 = (n.bminz[i] - org.z) * dirInv.z, t2 = (n.bmaxz[i] - org.z) * dirInv.z;
		tmn = std::max(tmn, std::min(t1, t2)), tmx = std::min(tmx, std::max(t1, t2));
		hit.tmin[i] = tmn;
		hit.result[i] = (tmx >= tmn) && (tmn < t);
	}
	for (int i = 0; i < 4; i++)
		hit.tmin[i] = u2f((f2u(hit.tmin[i]) & 0xFFFFFFFCu) | uint32_t(i));
	auto sw = [&](int a, int b) {
		if (hit.tmin[a] > hit.tmin[b])
			std::swap(hit.tmin[a], hit.tmin[b]);
	};
	sw(0, 1), sw(2, 3), sw(0, 2), sw(1, 3), sw(2, 3);
	return hit;
}

// D1 / D4 switch (setting "cudart_conventions" = off | on).  "on" follows CUDART to the letter: the barycentrics of a hit
// are computed as the area ratios of CUDAIntersect.h:82-87 (default: the same two weights from Moller-Trumbore's u, v)
// and the light index of the MIS pick probability is the material index (device_structs.h:37; default: the light
// index, as the reference's newer backends do).  Used by the pin test against the reference's host-compiled kernels.
static bool g_cudart_conventions = false;

// CUDAIntersect.h:48-94 intersect_triangle (Moller-Trumbore); D4: barycentrics returned are the
// Moller-Trumbore u,v (weights of vertex1, vertex2) instead of the area ratios of :82-87.
inline bool intersect_triangle(const vec3 &org, const vec3 &dir, float tmin, float *rayt, const vec3 &p0, const vec3 &p1,
							   const vec3 &p2, vec2 *bary, float epsilon)
{
	const vec3 e1 = p1 - p0, e2 = p2 - p0;
	const vec3 h = cross(dir, e2);
	const float a = dot(e1, h);
	if (a > -epsilon && a < epsilon)
		return false;
	const float f = 1.f / a;
	const vec3 s = org - p0;
	const float u = f * dot(s, h);
	if (u < 0.0f || u > 1.0f)
		return false;
	const vec3 q = cross(s, e1);
	const float v = f * dot(dir, q);
	if (v < 0.0f || u + v > 1.0f)
		return false;
	const float t = f * dot(e2, q);
	if (t > tmin && *rayt > t)
	{
		if (bary && g_cudart_conventions) // :82-87
		{
			const vec3 p = org + dir * t;
			const vec3 N = normalize(cross(e1, e2));
			const float areaABC = dot(N, cross(e1, e2));
			const float areaPBC = dot(N, cross(p1 - p, p2 - p));
			const float areaPCA = dot(N, cross(p2 - p, p0 - p));
			*bary = {areaPBC / areaABC, areaPCA / areaABC};
		}
		else if (bary)
			*bary = {u, v};
		*rayt = t;
		return true;
	}
	return false;
}

// ---------------------------------------------------------------------------------------------
// scene
// ---------------------------------------------------------------------------------------------
struct MeshData
{
	std::vector<vec4> vertices;
	std::vector<uint32_t> indices; // 3 per triangle, empty if unindexed
	std::vector<rfwb200_triangle> triangles;
	Bvh bvh;
	bool dirty = true;
	vec3 tri_vertex(size_t tri, int k) const
	{
		const size_t vi = indices.empty() ? tri * 3 + k : indices[tri * 3 + k];
		return {vertices[vi].x, vertices[vi].y, vertices[vi].z};
	}
};

struct Instance
{
	int mesh = -1;
	mat4 transform, inverse;
	mat3 normal;
	AABB world;
};

struct TextureDesc
{
	int type;
	uint32_t width, height, texel_count, addr;
};

struct ConnectRay // Shared.h:33-38 PotentialContribution
{
	vec4 O, D, E;
};

struct Counters
{
	uint64_t n_gen = 0, n_ext = 0, n_shade = 0, n_ext_out = 0, n_nee = 0, n_acc = 0;
};

} // namespace

struct rfworacle_context
{
	uint32_t width = 0, height = 0;
	std::vector<MeshData> meshes;
	std::vector<Instance> instances;
	Bvh tlas;
	std::vector<rfwb200_material> materials, materials_raw;
	std::vector<TextureDesc> textures;
	std::vector<uint32_t> uint_texels;
	std::vector<vec4> float_texels;
	std::vector<vec3> sky;
	uint32_t sky_w = 0, sky_h = 0;
	std::vector<rfwb200_area_light> area_lights;
	std::vector<rfwb200_point_light> point_lights;
	std::vector<rfwb200_spot_light> spot_lights;
	std::vector<rfwb200_directional_light> dir_lights;
	std::vector<uint32_t> blue_noise;

	// settings
	int spp = 1, max_path_length = 2, mode_pt = 1, survival_scale = 1;
	float clamp_value = 10.0f;

	// frame state
	std::vector<vec4> accumulator, framebuffer;
	uint32_t sample_index = 0;
	uint32_t probe_x = 0, probe_y = 0;
	uint32_t probed_inst = 0, probed_prim = 0;
	float probed_dist = 0;
	Counters counters;
	uint32_t last_spp = 0;
	float last_render_ms = 0;
	std::string error;
};

namespace
{
thread_local std::string g_error;
using Ctx = rfworacle_context;

// ---------------------------------------------------------------------------------------------
// extend / connect — Kernels.cu:226-303 intersect_scene, :305-381 is_occluded,
// CUDAIntersect.h:270-322 intersect_mbvh, :391-439 intersect_mbvh_shadow
// ---------------------------------------------------------------------------------------------
struct StackEntry
{
	int leftFirst, count;
};

static bool mesh_intersect(const MeshData &mesh, const vec3 &org, const vec3 &dir, float t_min, float *t, int *prim,
						   vec2 *bary)
{
	bool valid = false;
	StackEntry todo[64];
	int sp = 0;
	todo[0] = {0, -1};
	const vec3 dirInv = {1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z};
	const MBVHNode *nodes = mesh.bvh.mnodes.data();
	const uint32_t *prims = mesh.bvh.prim_indices.data();
	while (sp >= 0)
	{
		const int leftFirst = todo[sp].leftFirst, count = todo[sp].count;
		sp--;
		if (count >= 0)
		{
			for (int i = 0; i < count; i++)
			{
				const uint32_t id = prims[leftFirst + i];
				if (intersect_triangle(org, dir, t_min, t, mesh.tri_vertex(id, 0), mesh.tri_vertex(id, 1),
									   mesh.tri_vertex(id, 2), bary, T_EPSILON))
					valid = true, *prim = int(id);
			}
			continue;
		}
		const MBVHHit hit = intersect_quad_node(nodes[leftFirst], org, dirInv, *t);
		for (int i = 3; i >= 0; i--)
		{
			const int idx = int(f2u(hit.tmin[i]) & 3u);
			if (hit.result[idx] && nodes[leftFirst].childs[idx] >= 0)
			{
				sp++;
				todo[sp] = {nodes[leftFirst].childs[idx], nodes[leftFirst].counts[idx]};
			}
		}
	}
	return valid;
}

static bool mesh_occluded(const MeshData &mesh, const vec3 &org, const vec3 &dir, float t_min, float t_max)
{
	StackEntry todo[64];
	int sp = 0;
	todo[0] = {0, -1};
	const vec3 dirInv = {1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z};
	const MBVHNode *nodes = mesh.bvh.mnodes.data();
	const uint32_t *prims = mesh.bvh.prim_indices.data();
	while (sp >= 0)
	{
		const int leftFirst = todo[sp].leftFirst, count = todo[sp].count;
		sp--;
		if (count > -1)
		{
			for (int i = 0; i < count; i++)
			{
				const uint32_t id = prims[leftFirst + i];
				float tm = t_max;
				if (intersect_triangle(org, dir, t_min, &tm, mesh.tri_vertex(id, 0), mesh.tri_vertex(id, 1),
									   mesh.tri_vertex(id, 2), nullptr, T_EPSILON))
					return true;
			}
			continue;
		}
		const MBVHHit hit = intersect_quad_node(nodes[leftFirst], org, dirInv, t_max);
		for (int i = 3; i >= 0; i--)
		{
			const int idx = int(f2u(hit.tmin[i]) & 3u);
			if (hit.result[idx] && nodes[leftFirst].childs[idx] >= 0)
			{
				sp++;
				todo[sp] = {nodes[leftFirst].childs[idx], nodes[leftFirst].counts[idx]};
			}
		}
	}
	return false;
}

static bool intersect_scene(const Ctx &c, const vec3 &org, const vec3 &dir, int *inst, int *prim, float *t, vec2 *bary,
							float t_min)
{
	if (c.instances.empty())
		return false;
	bool valid = false;
	StackEntry todo[64];
	int sp = 0;
	todo[0] = {0, -1};
	const vec3 dirInv = {1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z};
	const MBVHNode *nodes = c.tlas.mnodes.data();
	const uint32_t *prims = c.tlas.prim_indices.data();
	while (sp >= 0)
	{
		const int leftFirst = todo[sp].leftFirst, count = todo[sp].count;
		sp--;
		if (count >= 0)
		{
			for (int i = 0; i < count; i++)
			{
				const uint32_t id = prims[leftFirst + i];
				const Instance &in = c.instances[id];
				if (in.mesh < 0)
					continue;
				// Kernels.cu:269-271: direction is not re-normalised => t stays in world units
				const vec3 o = in.inverse.mul_point(org), d = in.inverse.mul_dir(dir);
				if (mesh_intersect(c.meshes[in.mesh], o, d, t_min, t, prim, bary))
					valid = true, *inst = int(id);
			}
			continue;
		}
		const MBVHHit hit = intersect_quad_node(nodes[leftFirst], org, dirInv, *t);
		for (int i = 3; i >= 0; i--)
		{
			const int idx = int(f2u(hit.tmin[i]) & 3u);
			if (hit.result[idx] && nodes[leftFirst].childs[idx] >= 0)
			{
				sp++;
				todo[sp] = {nodes[leftFirst].childs[idx], nodes[leftFirst].counts[idx]};
			}
		}
	}
	return valid;
}

static bool is_occluded(const Ctx &c, const vec3 &org, const vec3 &dir, float t_min, float t_max)
{
	if (c.instances.empty())
		return false;
	StackEntry todo[64];
	int sp = 0;
	todo[0] = {0, -1};
	const vec3 dirInv = {1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z};
	const MBVHNode *nodes = c.tlas.mnodes.data();
	const uint32_t *prims = c.tlas.prim_indices.data();
	while (sp >= 0)
	{
		const int leftFirst = todo[sp].leftFirst, count = todo[sp].count;
		sp--;
		if (count > -1)
		{
			for (int i = 0; i < count; i++)
			{
				const uint32_t id = prims[leftFirst + i];
				const Instance &in = c.instances[id];
				if (in.mesh < 0)
					continue;
				const vec3 o = in.inverse.mul_point(org), d = in.inverse.mul_dir(dir);
				if (mesh_occluded(c.meshes[in.mesh], o, d, t_min, t_max))
					return true;
			}
			continue;
		}
		const MBVHHit hit = intersect_quad_node(nodes[leftFirst], org, dirInv, t_max);
		for (int i = 3; i >= 0; i--)
		{
			const int idx = int(f2u(hit.tmin[i]) & 3u);
			if (hit.result[idx] && nodes[leftFirst].childs[idx] >= 0)
			{
				sp++;
				todo[sp] = {nodes[leftFirst].childs[idx], nodes[leftFirst].counts[idx]};
			}
		}
	}
	return false;
}

// ---------------------------------------------------------------------------------------------
// bsdf/tools.h
// ---------------------------------------------------------------------------------------------
inline uint32_t PackNormal(const vec3 &N) // tools.h:10-21
{
	const float f = 65535.0f / std::max(std::sqrt(8.0f * N.z + 8.0f), 0.0001f);
	return cvt_u32(N.x * f + 32767.0f) + (cvt_u32(N.y * f + 32767.0f) << 16);
}
inline vec3 UnpackNormal(uint32_t p) // tools.h:22-29
{
	float nx = float(p & 65535u) * (2.0f / 65535.0f), ny = float(p >> 16) * (2.0f / 65535.0f);
	nx += -1, ny += -1;
	float nz = 1, nw = -1;
	float l = nx * -nx + ny * -ny + nz * -nw;
	nz = l;
	l = std::sqrt(l);
	nx *= l, ny *= l;
	return vec3(nx, ny, nz) * 2.0f + vec3(0, 0, -1);
}
inline float SurvivalProbability(const vec3 &d) { return std::min(1.0f, std::max(std::max(d.x, d.y), d.z)); } // :86
inline vec3 DiffuseReflectionUniform(float r0, float r1) // tools.h:102-108
{
	const float term1 = TWOPI * r0, term2 = std::sqrt(1 - r1 * r1);
	return vec3(std::cos(term1) * term2, std::sin(term1) * term2, r1);
}
inline vec3 DiffuseReflectionCosWeighted(float r0, float r1) // tools.h:110-117
{
	const float term1 = TWOPI * r0;
	const float term2 = float(std::sqrt(1.0 - double(r1)));
	return normalize(vec3(std::cos(term1) * term2, std::sin(term1) * term2, std::sqrt(r1)));
}
inline vec3 SafeOrigin(const vec3 &O, const vec3 &, const vec3 &N, float) { return O + N * 1e-5f; } // tools.h:119-123
inline void clampIntensity(vec3 &v, float clampValue) // tools.h:184-192
{
	const float m = std::max(v.x, std::max(v.y, v.z));
	if (m > clampValue)
		v = v * (clampValue / m);
}
inline void createTangentSpace(const vec3 &N, vec3 &T, vec3 &B) // tools.h:205-212
{
	const float s = signf(N.z);
	const float a = -1.0f / (s + N.z);
	const float b = N.x * N.y * a;
	T = vec3(1.0f + s * N.x * N.x * a, s * b, -s * N.x);
	B = vec3(b, s + N.y * N.y * a, -N.y);
}
inline vec3 tangentToWorld(const vec3 &s, const vec3 &N, const vec3 &T, const vec3 &B) { return T * s.x + B * s.y + N * s.z; }

// Kernels.cu:205-224 / tools.h:163-182 blueNoiseSampler
inline float blueNoiseSampler(const uint32_t *bn, int x, int y, int sampleIdx, int dim)
{
	x &= 127, y &= 127, sampleIdx &= 255, dim &= 255;
	const int ranked = sampleIdx ^ int(bn[dim + (x + y * 128) * 8 + 65536 * 3]);
	int value = int(bn[dim + ranked * 256]);
	value ^= int(bn[(dim & 7) + (x + y * 128) * 8 + 65536]);
	return (0.5f + float(value)) * (1.0f / 256.0f);
}

// ---------------------------------------------------------------------------------------------
// bsdf/compat.h:47-74 ShadingData, bsdf/disney.h
// ---------------------------------------------------------------------------------------------
struct ShadingData
{
	vec3 color;
	uint32_t flags = 0;
	vec3 absorption;
	uint32_t matID = 0;
	uint32_t parameters[4] = {0, 0, 0, 0};
	static float c2f(uint32_t v, int s) { return float((v >> s) & 255u) * (1.0f / 255.0f); }
	float metallic() const { return c2f(parameters[0], 0); }
	float subsurface() const { return c2f(parameters[0], 8); }
	float specular() const { return c2f(parameters[0], 16); }
	float roughness() const { return std::max(0.001f, c2f(parameters[0], 24)); }
	float spectint() const { return c2f(parameters[1], 0); }
	float clearcoat() const { return c2f(parameters[2], 0); }
	float clearcoatgloss() const { return c2f(parameters[2], 8); }
	float transmission() const { return c2f(parameters[2], 16); }
	float eta() const { return c2f(parameters[2], 24); }
	bool isEmissive() const { return color.x > 1.0f || color.y > 1.0f || color.z > 1.0f; }
};

inline float lerpf(float a, float b, float t) { return a + t * (b - a); }
inline vec3 lerp3(const vec3 &a, const vec3 &b, float t) { return a + (b - a) * t; }
inline float sqr(float x) { return x * x; }

inline bool Refract(const vec3 &wi, const vec3 &n, float eta, vec3 &wt) // disney.h:20-30
{
	const float cosThetaI = dot(n, wi);
	const float sin2ThetaI = std::max(0.0f, 1.0f - cosThetaI * cosThetaI);
	const float sin2ThetaT = eta * eta * sin2ThetaI;
	if (sin2ThetaT >= 1)
		return false;
	const float cosThetaT = std::sqrt(1.0f - sin2ThetaT);
	wt = (wi * -1.0f) * eta + n * (eta * cosThetaI - cosThetaT);
	return true;
}
inline float SchlickFresnel(float u) // disney.h:32-36
{
	const float m = clampf(1 - u, 0.0f, 1.0f);
	return float(m * m) * (m * m) * m;
}
inline float GTR1(float NDotH, float a) // disney.h:38-45
{
	if (a >= 1.0f)
		return INVPI;
	const float a2 = a * a;
	const float t = 1 + (a2 - 1) * NDotH * NDotH;
	return (a2 - 1) / (PI * std::log(a2) * t);
}
inline float GTR2(float NDotH, float a) // disney.h:47-52
{
	const float a2 = a * a;
	const float t = 1.0f + (a2 - 1.0f) * NDotH * NDotH;
	return a2 / (PI * t * t);
}
inline float SmithGGX(float NDotv, float alphaG) // disney.h:54-59
{
	const float a = alphaG * alphaG;
	const float b = NDotv * NDotv;
	return 1 / (NDotv + std::sqrt(a + b - a * b));
}
inline float Fr(float VDotN, float eio) // disney.h:61-72
{
	const float SinThetaT2 = sqr(eio) * (1.0f - VDotN * VDotN);
	if (SinThetaT2 > 1.0f)
		return 1.0f;
	const float LDotN = std::sqrt(1.0f - SinThetaT2);
	const float eta = 1.0f / eio;
	const float r1 = (VDotN - eta * LDotN) / (VDotN + eta * LDotN);
	const float r2 = (LDotN - eta * VDotN) / (LDotN + eta * VDotN);
	return 0.5f * (sqr(r1) + sqr(r2));
}
inline vec3 SafeNormalize(const vec3 &a) // disney.h:74-81
{
	const float ls = dot(a, a);
	if (ls > 0.0f)
		return a * (1.0f / std::sqrt(ls));
	return vec3(0.0f);
}

static float BSDFPdf(const ShadingData &sd, const vec3 &N, const vec3 &wo, const vec3 &wi) // disney.h:83-101
{
	float bsdfPdf = 0.0f, brdfPdf;
	if (dot(wi, N) <= 0.0f)
		brdfPdf = INV2PI * sd.subsurface() * 0.5f;
	else
	{
		const float F = Fr(dot(N, wo), sd.eta());
		const vec3 halfway = SafeNormalize(wi + wo);
		const float cosThetaHalf = std::fabs(dot(halfway, N));
		const float pdfHalf = GTR2(cosThetaHalf, sd.roughness()) * cosThetaHalf;
		const float pdfSpec = 0.25f * pdfHalf / std::max(1.e-6f, dot(wi, halfway));
		const float pdfDiff = std::fabs(dot(wi, N)) * INVPI * (1.0f - sd.subsurface());
		bsdfPdf = pdfSpec * F;
		brdfPdf = lerpf(pdfDiff, pdfSpec, 0.5f);
	}
	return lerpf(brdfPdf, bsdfPdf, sd.transmission());
}

static vec3 BSDFEval(const ShadingData &sd, const vec3 &N, const vec3 &wo, const vec3 &wi, float t,
					 bool backfacing) // disney.h:104-185
{
	const float NDotL = dot(N, wi);
	const float NDotV = dot(N, wo);
	const vec3 H = normalize(wi + wo);
	const float NDotH = dot(N, H);
	const float LDotH = dot(wi, H);
	const vec3 Cdlin = sd.color;
	const float Cdlum = .3f * Cdlin.x + .6f * Cdlin.y + .1f * Cdlin.z;
	const vec3 Ctint = Cdlum > 0.0f ? Cdlin / Cdlum : vec3(1.0f);
	const vec3 Cspec0 = lerp3(lerp3(vec3(1.0f), Ctint, sd.spectint()) * (sd.specular() * .08f), Cdlin, sd.metallic());
	vec3 bsdf = vec3(0.0f);
	vec3 brdf = vec3(0.0f);
	const float TRANSMISSION = sd.transmission(), METALLIC = sd.metallic(), SUBSURFACE = sd.subsurface();
	if (TRANSMISSION > 0.0f)
	{
		if (NDotL <= 0)
		{
			const float F = Fr(NDotV, sd.eta());
			bsdf = vec3((1.0f - F) / std::fabs(NDotL) * (1.0f - METALLIC) * TRANSMISSION);
		}
		else
		{
			const float a = sd.roughness();
			const float Ds = GTR2(NDotH, a);
			const float FH = Fr(LDotH, sd.eta());
			const vec3 Fs = lerp3(Cspec0, vec3(1.0f), FH);
			const float Gs = SmithGGX(NDotV, a) * SmithGGX(NDotL, a);
			bsdf = Fs * (Gs * Ds);
		}
	}
	if (TRANSMISSION < 1.0f)
	{
		if (NDotL <= 0)
		{
			if (SUBSURFACE > 0.0f)
			{
				const vec3 s = vec3(std::sqrt(sd.color.x), std::sqrt(sd.color.y), std::sqrt(sd.color.z));
				const float FL = SchlickFresnel(std::fabs(NDotL)), FV = SchlickFresnel(NDotV);
				const float Fd = (1.0f - 0.5f * FL) * (1.0f - 0.5f * FV);
				brdf = s * INVPI * SUBSURFACE * Fd * (1.0f - METALLIC);
			}
		}
		else
		{
			const float a = sd.roughness();
			const float Ds = GTR2(NDotH, a);
			const float FH = SchlickFresnel(LDotH);
			const vec3 Fs = lerp3(Cspec0, vec3(1.0f), FH);
			const float Gs = SmithGGX(NDotV, a) * SmithGGX(NDotL, a);
			const float FL = SchlickFresnel(NDotL), FV = SchlickFresnel(NDotV);
			const float Fd90 = float(0.5 + double(2.0f * LDotH * LDotH * a));
			const float Fd = lerpf(1.0f, Fd90, FL) * lerpf(1.0f, Fd90, FV);
			const float Dr = GTR1(NDotH, float(.1 + double(sd.clearcoatgloss()) * (.001 - .1)));
			const float Fc = lerpf(.04f, 1.0f, FH);
			const float Gr = SmithGGX(NDotL, .25f) * SmithGGX(NDotV, .25f);
			brdf = Cdlin * (INVPI * Fd) * (1.0f - METALLIC) * (1.0f - SUBSURFACE) + Fs * Gs * Ds +
				   vec3(sd.clearcoat() * Gr * Fc * Dr);
		}
	}
	const vec3 fin = lerp3(brdf, bsdf, TRANSMISSION);
	if (backfacing)
		return fin * vec3(std::exp(-sd.absorption.x * t), std::exp(-sd.absorption.y * t), std::exp(-sd.absorption.z * t));
	return fin;
}

static void BSDFSample(const ShadingData &sd, const vec3 &T, const vec3 &B, const vec3 &N, const vec3 &wo, vec3 &wi,
					   float &pdf, float r3, float r4) // disney.h:188-262
{
	const float transmission = sd.transmission();
	if (r3 < transmission)
	{
		const float F = Fr(dot(N, wo), sd.eta());
		if (r4 < F)
		{
			const float r1 = r3 / transmission;
			const float r2 = r4 / F;
			const float cosThetaHalf = std::sqrt((1.0f - r2) / (1.0f + (sqr(sd.roughness()) - 1.0f) * r2));
			const float sinThetaHalf = std::sqrt(std::max(0.0f, 1.0f - sqr(cosThetaHalf)));
			const float sinPhiHalf = std::sin(r1 * TWOPI);
			const float cosPhiHalf = std::cos(r1 * TWOPI);
			vec3 halfway = T * (sinThetaHalf * cosPhiHalf) + B * (sinThetaHalf * sinPhiHalf) + N * cosThetaHalf;
			if (dot(halfway, wo) <= 0.0f)
				halfway *= -1.0f;
			wi = reflect(wo * -1.0f, halfway);
		}
		else
		{
			pdf = 0;
			if (Refract(wo, N, sd.eta(), wi))
				pdf = (1.0f - F) * transmission;
		}
		return; // note: pdf is left untouched on the reflection branch (disney.h:196-220)
	}
	const float r1 = (r3 - transmission) / (1 - transmission);
	if (r4 < 0.5f)
	{
		const float r2 = r4 * 2;
		const float subsurface = sd.subsurface();
		vec3 d;
		if (r2 < subsurface)
		{
			const float r5 = r2 / subsurface;
			d = DiffuseReflectionUniform(r1, r5);
			d.z *= -1.0f;
		}
		else
		{
			const float r5 = (r2 - subsurface) / (1.0f - subsurface);
			d = DiffuseReflectionCosWeighted(r1, r5);
		}
		wi = T * d.x + B * d.y + N * d.z;
	}
	else
	{
		const float r2 = (r4 - 0.5f) * 2.0f;
		const float cosThetaHalf = std::sqrt((1.0f - r2) / (1.0f + (sqr(sd.roughness()) - 1.0f) * r2));
		const float sinThetaHalf = std::sqrt(std::max(0.0f, 1.0f - sqr(cosThetaHalf)));
		const float sinPhiHalf = std::sin(r1 * TWOPI);
		const float cosPhiHalf = std::cos(r1 * TWOPI);
		vec3 halfway = T * (sinThetaHalf * cosPhiHalf) + B * (sinThetaHalf * sinPhiHalf) + N * cosThetaHalf;
		if (dot(halfway, wo) <= 0.0f)
			halfway *= -1.0f;
		wi = reflect(wo * -1.0f, halfway);
	}
	pdf = BSDFPdf(sd, N, wo, wi);
}

// disney.h:266-280
static vec3 EvaluateBSDF(const ShadingData &sd, const vec3 &iN, const vec3 &wo, const vec3 &wi, float &pdf)
{
	const vec3 bsdf = BSDFEval(sd, iN, wo, wi, 0.0f, false);
	pdf = BSDFPdf(sd, iN, wo, wi);
	return bsdf;
}
// D5 switch (setting "bsdf_random_order" = ltr | rtl): C++ leaves the evaluation order of SampleBSDF's two RandomFloat(seed)
// arguments unspecified (bsdf/disney.h:278); g++ evaluates them right to left, which is what the host-compiled reference
// kernels of oracle/ref_build/ref_kernels_shim.cpp do, so the pin test selects rtl.
static bool g_bsdf_randoms_right_to_left = false;
static vec3 SampleBSDF(const ShadingData &sd, const vec3 &iN, const vec3 &T, const vec3 &B, const vec3 &wo, float t,
					   bool backfacing, vec3 &wi, float &pdf, uint32_t &seed)
{
	float r3 = RandomFloat(seed); // D5
	float r4 = RandomFloat(seed);
	if (g_bsdf_randoms_right_to_left) // the order a compiler that evaluates arguments right to left gives the reference
		std::swap(r3, r4);
	BSDFSample(sd, T, B, iN, wo, wi, pdf, r3, r4);
	return BSDFEval(sd, iN, wo, wi, t, backfacing);
}

// ---------------------------------------------------------------------------------------------
// CUDART/src/getShadingData.h
// ---------------------------------------------------------------------------------------------
inline vec4 uchar4_to_float4(uint32_t v) // :23-27
{
	const float r = 1.0f / 256.0f;
	return vec4(float(v & 255u) * r, float((v >> 8) & 255u) * r, float((v >> 16) & 255u) * r, float(v >> 24) * r);
}

static vec4 FetchTexel(const Ctx &c, float tcx, float tcy, int o, int w, int h) // :29-59, RGBA32 + BILINEAR
{
	if (w <= 0 || h <= 0)
		return vec4(0, 0, 0, 0);
	const float tx = (std::max(tcx + 1000, 0.0f) * float(w)) - 0.5f, ty = (std::max(tcy + 1000, 0.0f) * float(h)) - 0.5f;
	const int iu = cvt_i32(tx) % w;
	const int iv = cvt_i32(ty) % h;
	const float fu = tx - std::floor(tx);
	const float fv = ty - std::floor(ty);
	const float w0 = (1 - fu) * (1 - fv);
	const float w1 = fu * (1 - fv);
	const float w2 = (1 - fu) * fv;
	const float w3 = 1 - (w0 + w1 + w2);
	const uint32_t iu1 = uint32_t(iu + 1) % uint32_t(w), iv1 = uint32_t(iv + 1) % uint32_t(h);
	const uint32_t *t = c.uint_texels.data();
	const size_t n = c.uint_texels.size();
	auto fetch = [&](size_t i) { return i < n ? uchar4_to_float4(t[i]) : vec4(0, 0, 0, 0); };
	const vec4 p0 = fetch(size_t(o) + iu + size_t(iv) * w), p1 = fetch(size_t(o) + iu1 + size_t(iv) * w),
			   p2 = fetch(size_t(o) + iu + size_t(iv1) * w), p3 = fetch(size_t(o) + iu1 + size_t(iv1) * w);
	return p0 * w0 + p1 * w1 + p2 * w2 + p3 * w3;
}

static vec4 FetchTexelTrilinear(const Ctx &c, float lambda, float tcx, float tcy, int offset, int width,
								int height) // :61-98
{
	const int level0 = std::min(MIPLEVELCOUNT - 1, cvt_i32(lambda));
	const int level1 = std::min(MIPLEVELCOUNT - 1, level0 + 1);
	const float f = lambda - std::floor(lambda);
	uint32_t offset0 = offset, width0 = width, height0 = height;
	for (int i = 0; i < level0; i++)
		offset0 += width0 * height0, width0 >>= 1u, height0 >>= 1u;
	uint32_t offset1 = offset, width1 = width, height1 = height;
	for (int i = 0; i < level1; i++)
		offset1 += width1 * height1, width1 >>= 1u, height1 >>= 1u;
	const vec4 p0 = FetchTexel(c, tcx, tcy, int(offset0), int(width0), int(height0));
	const vec4 p1 = FetchTexel(c, tcx, tcy, int(offset1), int(width1), int(height1));
	return p0 * (1.0f - f) + p1 * f;
}

// :100-217 getShadingData (u, v: the weights of vertex 0 and vertex 1 as in the reference; w of vertex 2)
static ShadingData getShadingData(const Ctx &c, const vec3 &D, float u, float v, float coneWidth,
								  const rfwb200_triangle &tri, vec3 &N, vec3 &iN, vec3 &T, vec3 &B, const mat3 &invT)
{
	ShadingData r;
	r.matID = tri.material;
	const rfwb200_material &mat = c.materials[r.matID];
	const uint32_t flags = mat.flags;
	r.color = vec3(half2float(mat.diffuse[0]), half2float(mat.diffuse[1]), half2float(mat.diffuse[2]));
	r.absorption =
		vec3(half2float(mat.transmittance[0]), half2float(mat.transmittance[1]), half2float(mat.transmittance[2]));
	for (int i = 0; i < 4; i++)
		r.parameters[i] = mat.parameters[i];
	const float w = 1.0f - u - v;
	const vec3 n0(tri.vN0), n1(tri.vN1), n2(tri.vN2);
	N = vec3(tri.Nx, tri.Ny, tri.Nz);
	iN = N;
	if (has_flag(flags, HasSmoothNormals))
		iN = normalize(n0 * u + n1 * v + n2 * w); // :123
	N = normalize(invT.mul(N));
	iN = normalize(invT.mul(iN));
	createTangentSpace(iN, T, B);
	float tu = 0, tv = 0;
	if (has_flag(flags, HasDiffuseMap) || has_flag(flags, HasNormalMap) || has_flag(flags, HasSpecularityMap) ||
		has_flag(flags, HasRoughnessMap) || has_flag(flags, Has2ndDiffuseMap) || has_flag(flags, Has3rdDiffuseMap) ||
		has_flag(flags, Has2ndNormalMap) || has_flag(flags, Has3rdNormalMap))
	{
		tu = u * tri.u0 + v * tri.u1 + w * tri.u2; // :140-141
		tv = u * tri.v0 + v * tri.v1 + w * tri.v2;
	}
	if (has_flag(flags, HasDiffuseMap))
	{
		const float lambda = tri.LOD + std::log2(coneWidth * (1.0f / std::fabs(dot(-D, N))));
		auto layer = [&](const rfwb200_map_desc &m) {
			const float us = half2float(m.uscale), vs = half2float(m.vscale), uo = half2float(m.uoffs),
						vo = half2float(m.voffs);
			return FetchTexelTrilinear(c, lambda, us * (uo + tu), vs * (vo + tv), int(m.texaddr), int(m.width),
									   int(m.height));
		};
		auto nlayer = [&](const rfwb200_map_desc &m) {
			const float us = half2float(m.uscale), vs = half2float(m.vscale), uo = half2float(m.uoffs),
						vo = half2float(m.voffs);
			const vec4 t = FetchTexel(c, us * (uo + tu), vs * (vo + tv), int(m.texaddr), int(m.width), int(m.height));
			return (vec3(t.x, t.y, t.z) - vec3(0.5f)) * 2.0f;
		};
		const vec4 texel = layer(mat.tex0);
		if (has_flag(flags, HasAlpha) && texel.w < 0.5f)
		{
			r.flags |= 1;
			return r;
		}
		r.color = r.color * vec3(texel.x, texel.y, texel.z);
		if (has_flag(flags, Has2ndDiffuseMap))
		{
			const vec4 t = layer(mat.tex1);
			r.color += vec3(t.x, t.y, t.z);
		}
		if (has_flag(flags, Has3rdDiffuseMap))
		{
			const vec4 t = layer(mat.tex2);
			r.color += vec3(t.x, t.y, t.z);
		}
		if (has_flag(flags, HasNormalMap))
		{
			vec3 shadingNormal = nlayer(mat.nmap0);
			if (has_flag(flags, Has2ndNormalMap))
				shadingNormal += nlayer(mat.nmap1);
			if (has_flag(flags, Has3rdNormalMap))
				shadingNormal += nlayer(mat.nmap1); // reference re-reads layer 1 (:194-200) — kept
			shadingNormal = normalize(shadingNormal);
			iN = normalize(tangentToWorld(shadingNormal, iN, T, B));
		}
		r.color = r.color * vec3(texel.x, texel.y, texel.z); // second multiply (:213) — kept
	}
	return r;
}

// ---------------------------------------------------------------------------------------------
// CUDART/src/lights.h
// ---------------------------------------------------------------------------------------------
inline float area_light_energy(const rfwb200_area_light &l) { return length(vec3(l.radiance)); } // device_structs.h:116

static float PotentialAreaLightContribution(const Ctx &c, int idx, const vec3 &O, const vec3 &N, const vec3 &I,
											const vec3 &bary) // :17-36
{
	const rfwb200_area_light &light = c.area_lights[idx];
	const vec3 LN(light.normal);
	vec3 L = I;
	if (bary.x >= 0)
		L = vec3(light.vertex0) * bary.x + vec3(light.vertex1) * bary.y + vec3(light.vertex2) * bary.z;
	L = L - O;
	const float att = 1.0f / dot(L, L);
	L = normalize(L);
	const float LNdotL = std::max(0.0f, -dot(LN, L));
	const float NdotL = std::max(0.0f, dot(N, L));
	return light.energy * LNdotL * NdotL * att;
}
static float PotentialPointLightContribution(const Ctx &c, int idx, const vec3 &I, const vec3 &N) // :38-46
{
	const rfwb200_point_light &light = c.point_lights[idx];
	const vec3 L = vec3(light.position) - I;
	const float NdotL = std::max(0.0f, dot(N, L));
	const float att = 1.0f / dot(L, L);
	return light.energy * NdotL * att;
}
static float PotentialSpotLightContribution(const Ctx &c, int idx, const vec3 &I, const vec3 &N) // :48-68
{
	const rfwb200_spot_light &light = c.spot_lights[idx];
	vec3 L = vec3(light.position) - I;
	const float att = 1.0f / dot(L, L);
	L = normalize(L);
	const float d = (std::max(0.0f, -dot(L, vec3(light.direction))) - light.cos_outer) / (light.cos_inner - light.cos_outer);
	const float NdotL = std::max(0.0f, dot(N, L));
	const float LNdotL = std::max(0.0f, std::min(1.0f, d));
	return light.energy * LNdotL * NdotL * att;
}
static float PotentialDirectionalLightContribution(const Ctx &c, int idx, const vec3 &, const vec3 &N) // :70-76
{
	const rfwb200_directional_light &light = c.dir_lights[idx];
	const float LNdotL = std::max(0.0f, -dot(vec3(light.direction), N));
	return light.energy * LNdotL;
}
inline float CalculateLightPDF(const vec3 &D, float t, float lightArea, const vec3 &lightNormal) // :78-81
{
	return (t * t) / (-dot(D, lightNormal) * lightArea);
}
static float LightPickProb(const Ctx &c, int idx, const vec3 &O, const vec3 &N, const vec3 &I) // :83-116
{
	float potential[MAX_IS_LIGHTS];
	float sum = 0;
	const int na = int(c.area_lights.size());
	for (int i = 0; i < na; i++)
	{
		const float v = PotentialAreaLightContribution(c, i, O, N, I, vec3(-1.0f));
		if (i < MAX_IS_LIGHTS)
			potential[i] = v;
		sum += v;
	}
	for (size_t i = 0; i < c.point_lights.size(); i++)
		sum += PotentialPointLightContribution(c, int(i), O, N);
	for (size_t i = 0; i < c.spot_lights.size(); i++)
		sum += PotentialSpotLightContribution(c, int(i), O, N);
	for (size_t i = 0; i < c.dir_lights.size(); i++)
		sum += PotentialDirectionalLightContribution(c, int(i), O, N);
	if (sum <= 0)
		return 0;
	if (idx < 0 || idx >= na || idx >= MAX_IS_LIGHTS)
		return 0;
	return potential[idx] / sum;
}
static vec3 RandomBarycentrics(float r0) // :119-157
{
	const uint32_t uf = cvt_u32(r0 * float(4294967295u));
	vec2 A = {1.f, 0.f}, B = {0.f, 1.f}, C = {0.f, 0.f};
	auto mid = [](const vec2 &a, const vec2 &b) { return vec2{(a.x + b.x) * 0.5f, (a.y + b.y) * 0.5f}; };
	for (int i = 0; i < 16; ++i)
	{
		const int d = int((uf >> (2 * (15 - i))) & 0x3);
		vec2 An, Bn, Cn;
		switch (d)
		{
		case 0:
			An = mid(B, C), Bn = mid(A, C), Cn = mid(A, B);
			break;
		case 1:
			An = A, Bn = mid(A, B), Cn = mid(A, C);
			break;
		case 2:
			An = mid(B, A), Bn = B, Cn = mid(B, C);
			break;
		default:
			An = mid(C, A), Bn = mid(C, B), Cn = C;
			break;
		}
		A = An, B = Bn, C = Cn;
	}
	const float rx = (A.x + B.x + C.x) * 0.3333333f, ry = (A.y + B.y + C.y) * 0.3333333f;
	return vec3(rx, ry, 1.0f - rx - ry);
}
static vec3 RandomPointOnLight(const Ctx &c, float r0, float r1, const vec3 &I, const vec3 &N, float &pickProb,
							   float &lightPdf, vec3 &lightColor) // :159-265
{
	const int na = int(c.area_lights.size()), np = int(c.point_lights.size()), ns = int(c.spot_lights.size()),
			  nd = int(c.dir_lights.size());
	const float lightCount = float(na + np + ns + nd);
	const vec3 bary = RandomBarycentrics(r0);
	float potential[MAX_IS_LIGHTS];
	float sum = 0, total = 0;
	int lights = 0, lightIdx = 0;
	auto push = [&](float v) {
		if (lights < MAX_IS_LIGHTS)
			potential[lights] = v;
		lights++;
		sum += v;
	};
	for (int i = 0; i < na; i++)
		push(PotentialAreaLightContribution(c, i, I, N, vec3(0.0f), bary));
	for (int i = 0; i < np; i++)
		push(PotentialPointLightContribution(c, i, I, N));
	for (int i = 0; i < ns; i++)
		push(PotentialSpotLightContribution(c, i, I, N));
	for (int i = 0; i < nd; i++)
		push(PotentialDirectionalLightContribution(c, i, I, N));
	lights = std::min(lights, MAX_IS_LIGHTS);
	if (sum <= 0)
	{
		lightPdf = 0;
		return vec3(1.0f);
	}
	r1 *= sum;
	for (int i = 0; i < lights; i++)
	{
		total += potential[i];
		if (total >= r1)
		{
			lightIdx = i;
			break;
		}
	}
	pickProb = potential[lightIdx] / sum;
	lightIdx = std::min(std::max(lightIdx, 0), int(lightCount) - 1);
	if (lightIdx < na)
	{
		const rfwb200_area_light &light = c.area_lights[lightIdx];
		lightColor = vec3(light.radiance);
		const vec3 LN(light.normal);
		const vec3 P = vec3(light.vertex0) * bary.x + vec3(light.vertex1) * bary.y + vec3(light.vertex2) * bary.z;
		vec3 L = I - P;
		const float sqDist = dot(L, L);
		L = normalize(L);
		const float LNdotL = dot(L, LN);
		const float reciSolidAngle = sqDist / (light.area * LNdotL);
		lightPdf = (LNdotL > 0 && dot(L, N) < 0) ? (reciSolidAngle * (1.0f / area_light_energy(light))) : 0;
		return P;
	}
	if (lightIdx < na + np)
	{
		const rfwb200_point_light &light = c.point_lights[lightIdx - na];
		const vec3 pos(light.position);
		lightColor = vec3(light.radiance);
		const vec3 L = I - pos;
		const float sqDist = dot(L, L);
		lightPdf = dot(L, N) < 0 ? (sqDist / light.energy) : 0;
		return pos;
	}
	if (lightIdx < na + np + ns)
	{
		const rfwb200_spot_light &light = c.spot_lights[lightIdx - (na + np)];
		const vec3 P(light.position), D(light.direction);
		vec3 L = I - P;
		const float sqDist = dot(L, L);
		L = normalize(L);
		const float d = std::max(0.0f, dot(L, D) - light.cos_outer) / (light.cos_inner - light.cos_outer);
		const float LNdotL = std::min(1.0f, d);
		lightPdf = (LNdotL > 0 && dot(L, N) < 0) ? (sqDist / (LNdotL * light.energy)) : 0;
		lightColor = vec3(light.radiance);
		return P;
	}
	const rfwb200_directional_light &light = c.dir_lights[lightIdx - (na + np + ns)];
	const vec3 L(light.direction);
	lightColor = vec3(light.radiance);
	const float NdotL = dot(L, N);
	lightPdf = NdotL < 0 ? float(1 * (1.0 / double(light.energy))) : 0;
	return I - L * 1000.0f;
}

// ---------------------------------------------------------------------------------------------
// generate — Kernels.cu:383-426 generatePrimaryRay (PT) / EmbreeRT/src/Ray.cpp:16-47 (E-mode)
// ---------------------------------------------------------------------------------------------
static void generate_primary_pt(const Ctx &c, const rfwb200_camera_view &view, uint32_t pathID, uint32_t sampleIndex,
								vec3 &O, vec3 &D)
{
	const int sx = int(pathID % c.width), sy = int(pathID / c.width);
	const uint32_t *bn = c.blue_noise.data();
	const float r0 = blueNoiseSampler(bn, sx, sy, int(sampleIndex), 0);
	const float r1 = blueNoiseSampler(bn, sx, sy, int(sampleIndex), 1);
	float r2 = blueNoiseSampler(bn, sx, sy, int(sampleIndex), 2);
	float r3 = blueNoiseSampler(bn, sx, sy, int(sampleIndex), 3);
	const float blade = float(int(r0 * 9));
	r2 = (r2 - blade * (1.0f / 9.0f)) * 9.0f;
	constexpr float piOver4point5 = 3.14159265359f / 4.5f;
	// __sincosf(x, &s, &c): x1 receives the sine, y1 the cosine (:406-407)
	const float x1 = std::sin(blade * piOver4point5), y1 = std::cos(blade * piOver4point5);
	const float x2 = std::sin((blade + 1.0f) * piOver4point5), y2 = std::cos((blade + 1.0f) * piOver4point5);
	if ((r2 + r3) > 1.0f)
		r2 = 1.0f - r2, r3 = 1.0f - r3;
	const float xr = x1 * r2 + x2 * r3;
	const float yr = y1 * r2 + y2 * r3;
	const vec3 pos(view.pos), p1(view.p1);
	const vec3 right = vec3(view.p2) - p1, up = vec3(view.p3) - p1;
	O = pos + (right * xr + up * yr) * view.aperture;
	const float u = (float(sx) + r0) * (1.0f / float(c.width));
	const float v = (float(sy) + r1) * (1.0f / float(c.height));
	const vec3 pointOnPixel = p1 + right * u + up * v;
	D = normalize(pointOnPixel - O);
}

static void generate_primary_emode(const Ctx &c, const rfwb200_camera_view &view, uint32_t pixel, uint32_t sampleIndex,
								   vec3 &O, vec3 &D)
{
	Xor128 rng;
	rng.x = 123456789u ^ WangHash(pixel * 16789u + sampleIndex * 1791u);
	const float r0 = rng.rand(), r1 = rng.rand();
	float r2 = rng.rand(), r3 = rng.rand();
	const int x = int(pixel % c.width), y = int(pixel / c.width);
	const float blade = float(int(r0 * 9));
	r2 = (r2 - blade * (1.0f / 9.0f)) * 9.0f;
	constexpr float piOver4point5 = 3.14159265359f / 4.5f;
	const float x1 = std::cos(blade * piOver4point5), y1 = std::sin(blade * piOver4point5);
	const float x2 = std::cos((blade + 1.0f) * piOver4point5), y2 = std::sin((blade + 1.0f) * piOver4point5);
	if ((r2 + r3) > 1.0f)
		r2 = 1.0f - r2, r3 = 1.0f - r3;
	const float xr = x1 * r2 + x2 * r3;
	const float yr = y1 * r2 + y2 * r3;
	const vec3 pos(view.pos), p1(view.p1);
	const vec3 right = vec3(view.p2) - p1, up = vec3(view.p3) - p1;
	O = pos + (right * xr + up * yr) * view.aperture;
	const float u = (float(x) + r0) * (1.0f / float(c.width));
	const float v = (float(y) + r1) * (1.0f / float(c.height));
	D = normalize(p1 + right * u + up * v - O);
}

// ---------------------------------------------------------------------------------------------
// PT-mode frame — CUDART/src/Context.cpp:65-159 host loop over Kernels.cu:428-794
// ---------------------------------------------------------------------------------------------
struct PathState
{
	vec4 O, D, T, hit; // pathOrigins, pathDirections, pathThroughputs, pathStates
};

static vec4 trace_to_state(const Ctx &c, const vec3 &O, const vec3 &D)
{
	float t = 1e34f;
	int inst = 0, prim = -1;
	vec2 bary = {0, 0};
	if (intersect_scene(c, O, D, &inst, &prim, &t, &bary, 1e-5f)) // Kernels.cu:455-457
	{
		// D4: the hit record holds the weights of vertex 0 and vertex 1 like the reference's; they are taken from the
		// Moller-Trumbore u, v (1 - u - v, u) unless cudart_conventions asks for the area ratios of CUDAIntersect.h:82-87
		if (!g_cudart_conventions)
			bary = {1.0f - bary.x - bary.y, bary.x};
		return vec4(u2f(cvt_u32(65535.0f * bary.x) | (cvt_u32(65535.0f * bary.y) << 16)), u2f(uint32_t(inst)),
					u2f(uint32_t(prim)), t);
	}
	return vec4(0, 0, u2f(uint32_t(-1)), 0);
}

struct ShadeOut
{
	bool has_ext = false, has_shadow = false;
	PathState ext;
	ConnectRay shadow;
	bool acc = false;
	vec3 acc_value;
	uint32_t pixel = 0;
};

// Kernels.cu:571-794 shade_rays for one path
static void shade_path(Ctx &c, const rfwb200_camera_view &view, const PathState &p, uint32_t pathLength,
					   uint32_t samplesTaken, ShadeOut &out)
{
	const vec4 hitData = p.hit, O4 = p.O, D4 = p.D;
	const vec4 T4 = pathLength == 0 ? vec4(1, 1, 1, 1) : p.T;
	uint32_t flags = f2u(O4.w) & 0xFF;
	vec3 throughput(T4.x, T4.y, T4.z);
	const float bsdfPdf = T4.w;
	const vec3 D(D4.x, D4.y, D4.z);
	const uint32_t pathIndex = f2u(O4.w) >> 8;
	out.pixel = pathIndex;
	const int primIdx = int(f2u(hitData.z));
	if (primIdx < 0)
	{
		// :593-610 sky
		const uint32_t u = cvt_u32(float(c.sky_w) * 0.5f * (1.0f + std::atan2(D.x, -D.z) * INVPI));
		const uint32_t v = cvt_u32(float(c.sky_h) * std::acos(D.y) * INVPI);
		const uint32_t idx = u + v * c.sky_w;
		const vec3 sky = idx < c.sky_h * c.sky_w ? c.sky[idx] : vec3(0.0f);
		vec3 contribution = throughput * (1.0f / bsdfPdf) * sky;
		if (any_nan(contribution))
			return;
		clampIntensity(contribution, c.clamp_value);
		out.acc = true, out.acc_value = contribution;
		return;
	}
	const vec3 O(O4.x, O4.y, O4.z);
	const vec3 I = O + D * hitData.w;
	const uint32_t ub = f2u(hitData.x);
	const int instanceIdx = int(f2u(hitData.y));
	const Instance &instance = c.instances[instanceIdx];
	const rfwb200_triangle &triangle = c.meshes[instance.mesh].triangles[primIdx];
	const float bu = float(ub & 65535u) * (1.0f / 65535.0f), bv = float((ub >> 16) & 65535u) * (1.0f / 65535.0f);
	vec3 N, iN, T, B;
	const ShadingData sd =
		getShadingData(c, D, bu, bv, view.spread_angle * hitData.w, triangle, N, iN, T, B, instance.normal);
	if (pathLength == 0 && pathIndex == c.probe_x + c.probe_y * c.width) // :626-631
	{
		{ std::lock_guard<std::mutex> lk(g_probe_mutex);
			c.probed_inst = uint32_t(instanceIdx), c.probed_prim = uint32_t(primIdx), c.probed_dist = hitData.w;
		}
	}
	if (sd.flags & 1) // :634-647 alpha cut-out (D2)
	{
		if (int(pathLength) < c.max_path_length)
		{
			if (any_nan(throughput))
				return;
			out.has_ext = true;
			out.ext.O = vec4(I + D * GEOMETRY_EPSILON, O4.w);
			out.ext.D = D4;
			out.ext.T = T4;
		}
		return;
	}
	if (sd.isEmissive()) // :650-692
	{
		const float DdotNL = -dot(D, N);
		vec3 contribution(0.0f);
		if (DdotNL > 0)
		{
			if (pathLength == 0)
				contribution = sd.color;
			else if (flags & IS_SPECULAR)
				contribution = throughput * sd.color * (1.0f / bsdfPdf);
			else
			{
				const vec3 lastN = UnpackNormal(f2u(D4.w));
				const float lightPdf = CalculateLightPDF(D, hitData.w, triangle.area, N);
				const int triangleIdx = g_cudart_conventions ? int(triangle.material) : triangle.light_tri_idx; // D1
				const float pickProb = LightPickProb(c, triangleIdx, O, lastN, I);
				if ((bsdfPdf + lightPdf * pickProb) <= 0)
					return;
				contribution = throughput * sd.color * (1.0f / (bsdfPdf + lightPdf * pickProb));
			}
		}
		if (any_nan(contribution))
			contribution = vec3(0.0f);
		clampIntensity(contribution, c.clamp_value);
		out.acc = true, out.acc_value = contribution;
		return;
	}
	if (sd.roughness() < MIN_ROUGHNESS)
		flags |= IS_SPECULAR;
	else
		flags &= ~IS_SPECULAR;
	uint32_t seed = WangHash(pathIndex * 16789u + samplesTaken * 1791u + pathLength * 720898027u);
	const float flip = (dot(D, N) > 0) ? -1.0f : 1.0f;
	N *= flip;
	iN *= flip;
	throughput *= 1.0f / bsdfPdf;
	const size_t nlights = c.area_lights.size() + c.point_lights.size() + c.spot_lights.size() + c.dir_lights.size();
	if ((flags & IS_SPECULAR) == 0 && nlights > 0) // :706-755 NEE
	{
		vec3 lightColor;
		float r0, r1, pickProb = 0, lightPdf = 0;
		if (samplesTaken < 256)
		{
			const int x = int(pathIndex % c.width), y = int(pathIndex / c.width);
			r0 = blueNoiseSampler(c.blue_noise.data(), x, y, int(samplesTaken), 4);
			r1 = blueNoiseSampler(c.blue_noise.data(), x, y, int(samplesTaken), 5);
		}
		else
		{
			r0 = RandomFloat(seed);
			r1 = RandomFloat(seed);
		}
		vec3 L = RandomPointOnLight(c, r0, r1, I, iN, pickProb, lightPdf, lightColor) - I;
		const float dist = length(L);
		L *= 1.0f / dist;
		const float NdotL = dot(L, iN);
		if (NdotL > 0 && lightPdf > 0)
		{
			float shadowPdf;
			const vec3 sampledBSDF = EvaluateBSDF(sd, iN, D * -1.0f, L, shadowPdf);
			if (shadowPdf > 0)
			{
				vec3 contribution = throughput * sampledBSDF * lightColor * (NdotL / (shadowPdf + lightPdf * pickProb));
				clampIntensity(contribution, c.clamp_value);
				if (!any_nan(contribution))
				{
					out.has_shadow = true;
					out.shadow.O = vec4(SafeOrigin(I, L, N, GEOMETRY_EPSILON), 0);
					out.shadow.D = vec4(L, dist - 2.0f * GEOMETRY_EPSILON);
					out.shadow.E = vec4(contribution, u2f(pathIndex));
				}
			}
		}
	}
	if (int(pathLength) >= c.max_path_length) // :758-759
		return;
	vec3 R;
	float newBsdfPdf = 0.0f;
	const vec3 bsdf = SampleBSDF(sd, iN, T, B, D * -1.0f, hitData.w, flip < 0, R, newBsdfPdf, seed);
	if (c.survival_scale) // :783 (no kill test — only the 1/p scale)
		throughput = throughput * 1.0f / SurvivalProbability(throughput) * bsdf * std::fabs(dot(iN, R));
	else
		throughput = throughput * bsdf * std::fabs(dot(iN, R));
	if (newBsdfPdf < 1e-6f || std::isnan(newBsdfPdf) || throughput.x < 0.0f || throughput.y < 0.0f || throughput.z < 0.0f)
		return;
	out.has_ext = true;
	out.ext.O = vec4(SafeOrigin(I, R, N, GEOMETRY_EPSILON), u2f((pathIndex << 8) | flags));
	out.ext.D = vec4(R, u2f(PackNormal(iN)));
	out.ext.T = vec4(throughput, newBsdfPdf);
}

static void render_sample_pt(Ctx &c, const rfwb200_camera_view &view)
{
	const uint32_t P = c.width * c.height;
	const uint32_t sampleIndex = c.sample_index;
	std::vector<PathState> cur(P), next;
	// Primary: generate + extend (Kernels.cu:438-460)
	parallel_for(int64_t(int64_t(P)), 256, [&](int64_t i) {
	{
		vec3 O, D;
		generate_primary_pt(c, view, uint32_t(i), sampleIndex, O, D);
		cur[i].O = vec4(O, u2f((uint32_t(i) << 8) + 1));
		cur[i].D = vec4(D, 0);
		cur[i].T = vec4(1, 1, 1, 1);
		cur[i].hit = trace_to_state(c, O, D);
	}
	});
	c.counters.n_gen += P, c.counters.n_ext += P;
	uint32_t pathLength = 0;
	size_t active = P;
	std::vector<ConnectRay> shadows;
	for (;;)
	{
		// shade
		const size_t n = active;
		std::vector<ShadeOut> outs(n);
		parallel_for(int64_t(int64_t(n)), 256, [&](int64_t i) {
			shade_path(c, view, cur[i], pathLength, sampleIndex /* D3 */, outs[i]);
		});
		c.counters.n_shade += n;
		next.clear();
		shadows.clear();
		for (size_t i = 0; i < n; i++)
		{
			const ShadeOut &o = outs[i];
			if (o.acc)
			{
				vec4 &a = c.accumulator[o.pixel];
				a = a + vec4(o.acc_value, 0.0f);
				c.counters.n_acc++;
			}
			if (o.has_ext)
				next.push_back(o.ext);
			if (o.has_shadow)
				shadows.push_back(o.shadow);
		}
		c.counters.n_ext_out += next.size();
		// Context.cpp:109: while (activePaths > 0 && pathLength < MAX_PATH_LENGTH) — shadow rays of
		// the last shade, and of any shade that left no active path, are never traced.
		if (!(next.size() > 0 && int(pathLength) < c.max_path_length))
			break;
		pathLength++;
		// connect (Kernels.cu:484-498)
		const size_t ns = shadows.size();
		std::vector<uint8_t> visible(ns);
		parallel_for(int64_t(int64_t(ns)), 256, [&](int64_t i) {
		{
			const ConnectRay &s = shadows[i];
			visible[i] = !is_occluded(c, vec3(s.O.x, s.O.y, s.O.z), vec3(s.D.x, s.D.y, s.D.z), GEOMETRY_EPSILON, s.D.w);
		}
		});
		c.counters.n_nee += ns;
		for (size_t i = 0; i < ns; i++)
			if (visible[i])
			{
				vec4 &a = c.accumulator[f2u(shadows[i].E.w)];
				a = a + vec4(shadows[i].E.x, shadows[i].E.y, shadows[i].E.z, 1.0f);
				c.counters.n_acc++;
			}
		// extend (Kernels.cu:461-483)
		cur.swap(next);
		active = cur.size();
		parallel_for(int64_t(int64_t(active)), 256, [&](int64_t i) {
			cur[i].hit = trace_to_state(c, vec3(cur[i].O.x, cur[i].O.y, cur[i].O.z), vec3(cur[i].D.x, cur[i].D.y, cur[i].D.z));
		});
		c.counters.n_ext += active;
	}
	c.sample_index++;
}

// EmbreeRT/src/Context.cpp:417-476 retrieve_material: interpolated normal, material colour times the nearest texel of the
// first diffuse map (no mips, no filtering; the FLOAT4 case falls through into the UINT case)
static void emode_material(const Ctx &c, const Instance &inst, const rfwb200_triangle &tri, const vec3 &bary, vec3 &color, vec3 &iN)
{
	const rfwb200_material &material = c.materials_raw[tri.material];
	// retrieve_material :417-476
	const vec3 iNl = vec3(tri.vN0) * bary.x + vec3(tri.vN1) * bary.y + vec3(tri.vN2) * bary.z;
	iN = normalize(inst.normal.mul(iNl));
	color = vec3(half2float(material.diffuse[0]), half2float(material.diffuse[1]), half2float(material.diffuse[2]));
	float tu = 0.0f, tv = 0.0f;
	if (has_flag(material.flags, HasDiffuseMap) || has_flag(material.flags, HasNormalMap) ||
		has_flag(material.flags, HasRoughnessMap) || has_flag(material.flags, HasAlphaMap) ||
		has_flag(material.flags, HasSpecularityMap))
	{
		tu = bary.x * tri.u0 + bary.y * tri.u1 + bary.z * tri.u2;
		tv = bary.x * tri.v0 + bary.y * tri.v1 + bary.z * tri.v2;
	}
	if (has_flag(material.flags, HasDiffuseMap) && material.tex0.texaddr < c.textures.size())
	{
		const float u = (tu + half2float(material.tex0.uoffs)) * half2float(material.tex0.uscale);
		const float v = (tv + half2float(material.tex0.voffs)) * half2float(material.tex0.vscale);
		float txf = std::fmod(u, 1.0f), tyf = std::fmod(v, 1.0f);
		if (txf < 0.f)
			txf = 1.f + txf;
		if (tyf < 0.f)
			tyf = 1.f + tyf;
		const TextureDesc &tex = c.textures[material.tex0.texaddr]; // unpatched: texaddr0 = texture id
		const uint32_t ix = cvt_u32(txf * float(tex.width - 1)), iy = cvt_u32(tyf * float(tex.height - 1));
		const size_t id = size_t(iy) * tex.width + ix;
		if (tex.type == RFWB200_TEX_UINT)
		{
			const uint32_t tc = c.uint_texels[tex.addr + id];
			constexpr float sc = 1.0f / 256.0f;
			color = color * sc * vec3(float(tc & 0xFFu), float((tc >> 8) & 0xFFu), float((tc >> 16) & 0xFFu));
		}
		else
		{
			// FLOAT4 case falls through into the UINT case (:458-472): the texel is applied, then the word at index `id`
			// of the same buffer read as uints (float number id of the texture, not texel id) is applied as RGBA8.
			const vec4 tf = c.float_texels[tex.addr + id];
			color = color * vec3(tf.x, tf.y, tf.z);
			const vec4 &wq = c.float_texels[tex.addr + id / 4];
			const uint32_t tc = f2u(id % 4 == 0 ? wq.x : (id % 4 == 1 ? wq.y : (id % 4 == 2 ? wq.z : wq.w)));
			constexpr float sc = 1.0f / 256.0f;
			color = color * sc * vec3(float(tc & 0xFFu), float((tc >> 8) & 0xFFu), float((tc >> 16) & 0xFFu));
		}
	}
}

// ---------------------------------------------------------------------------------------------
// E-mode frame — EmbreeRT/src/Context.cpp:104-300, retrieve_material :417-476
// ---------------------------------------------------------------------------------------------
static void render_emode(Ctx &c, const rfwb200_camera_view &view)
{
	const uint32_t W = c.width, H = c.height;
	const uint32_t sampleIndex = c.sample_index;
	const int tilesY = int(H / 2), tilesX = int(W / 4); // :137-139 (remainder pixels are not rendered)
	parallel_for(int64_t(tilesY), 4, [&](int64_t ty) {
		for (int tx = 0; tx < tilesX; tx++)
			for (int j = 0; j < 8; j++)
			{
				const uint32_t x = uint32_t(tx * 4 + (j & 3)), y = uint32_t(ty * 2 + (j >> 2));
				const uint32_t pixel = y * W + x;
				vec3 origin, direction;
				generate_primary_emode(c, view, pixel, sampleIndex, origin, direction);
				float t = 1e34f; // Ray.cpp:182-183 tnear 1e-5, tfar 1e34
				int instID = 0, primID = -1;
				vec2 uv = {0, 0};
				if (!intersect_scene(c, origin, direction, &instID, &primID, &t, &uv, 1e-5f))
				{
					// :187-196
					const float su = 0.5f * (1.0f + std::atan2(direction.x, -direction.z) * INVPI);
					const float sv = std::acos(direction.y) * INVPI;
					const uint32_t px = cvt_u32(su * float(c.sky_w - 1)), py = cvt_u32(sv * float(c.sky_h - 1));
					const size_t si = size_t(py) * c.sky_w + px;
					const vec3 s = si < c.sky.size() ? c.sky[si] : vec3(0.0f);
					c.framebuffer[pixel] = vec4(s, 0.0f);
					continue;
				}
				if (pixel == c.probe_y * W + c.probe_x)
				{
#pragma omp critical
					{
						c.probed_dist = t, c.probed_inst = uint32_t(instID), c.probed_prim = uint32_t(primID);
					}
				}
				const Instance &inst = c.instances[instID];
				const rfwb200_triangle &tri = c.meshes[inst.mesh].triangles[primID];
				const vec3 bary(1.0f - uv.x - uv.y, uv.x, uv.y);
				const vec3 p = origin + direction * t;
				vec3 color, iN;
				emode_material(c, inst, tri, bary, color, iN); // retrieve_material :417-476
				if (color.x > 1 || color.y > 1 || color.z > 1) // :216-220
				{
					c.framebuffer[pixel] = vec4(color, 1.0f);
					continue;
				}
				vec3 contrib(0.1f);
				for (const auto &l : c.area_lights) // :229-250
				{
					vec3 L = vec3(l.position) - p;
					const float sq_dist = dot(L, L);
					const float dist = std::sqrt(sq_dist);
					L = L / dist;
					const float NdotL = dot(iN, L);
					const float LNdotL = -dot(vec3(l.normal), L);
					if (NdotL <= 0 || LNdotL <= 0)
						continue;
					if (!is_occluded(c, p, L, 1e-4f, dist * (1.0f - 1e-4f)))
						contrib += vec3(l.radiance) * l.area / sq_dist * NdotL * LNdotL;
				}
				for (const auto &l : c.point_lights) // :252-271
				{
					vec3 L = vec3(l.position) - p;
					const float sq_dist = dot(L, L);
					const float dist = std::sqrt(sq_dist);
					L = L / dist;
					const float NdotL = dot(iN, L);
					if (NdotL <= 0)
						continue;
					if (!is_occluded(c, p, L, 1e-4f, dist * (1.0f - 1e-4f)))
						contrib += vec3(l.radiance) / sq_dist * NdotL;
				}
				c.framebuffer[pixel] = vec4(color * contrib, 1.0f);
			}
	});
	c.counters.n_gen += uint64_t(tilesX) * tilesY * 8;
	c.counters.n_ext += uint64_t(tilesX) * tilesY * 8;
	c.sample_index++;
}

static int fail(const char *msg)
{
	g_error = msg;
	return RFWB200_ERR_INVALID;
}

} // namespace

// ---------------------------------------------------------------------------------------------
// C API — same shapes as include/rfwb200.h so tests feed both sides identically
// ---------------------------------------------------------------------------------------------
extern "C"
{
#define ORACLE_API __attribute__((visibility("default")))

	ORACLE_API const char *rfworacle_last_error(void) { return g_error.c_str(); }

	ORACLE_API int rfworacle_create(int, rfworacle_context **out)
	{
		if (!out)
			return fail("out is null");
		*out = new rfworacle_context();
		return RFWB200_OK;
	}
	ORACLE_API int rfworacle_destroy(rfworacle_context *c)
	{
		delete c;
		return RFWB200_OK;
	}
	// createBlueNoiseBuffer (context/blue_noise.h:8204-8219): 327,680 table bytes -> uint[5*65536]
	ORACLE_API int rfworacle_set_blue_noise(rfworacle_context *c, const uint8_t *table, size_t bytes)
	{
		if (!c || !table || bytes != 65536 + 2 * 131072)
			return fail("blue noise table must be 327680 bytes");
		c->blue_noise.assign(65536 * 5, 0);
		for (int i = 0; i < 65536; i++)
			c->blue_noise[i] = table[i];
		for (int i = 0; i < 131072; i++)
			c->blue_noise[i + 65536] = table[65536 + i];
		for (int i = 0; i < 131072; i++)
			c->blue_noise[i + 3 * 65536] = table[65536 + 131072 + i];
		return RFWB200_OK;
	}
	ORACLE_API int rfworacle_init(rfworacle_context *c, uint32_t w, uint32_t h)
	{
		if (!c || !w || !h)
			return fail("bad size");
		c->width = w, c->height = h;
		c->accumulator.assign(size_t(w) * h, vec4());
		c->framebuffer.assign(size_t(w) * h, vec4());
		c->sample_index = 0;
		return RFWB200_OK;
	}
	ORACLE_API int rfworacle_set_sky(rfworacle_context *c, const float *rgb, size_t w, size_t h)
	{
		c->sky.resize(w * h);
		for (size_t i = 0; i < w * h; i++)
			c->sky[i] = vec3(rgb + 3 * i);
		c->sky_w = uint32_t(w), c->sky_h = uint32_t(h);
		return RFWB200_OK;
	}
	// CUDART/src/Context.cpp:201-268
	ORACLE_API int rfworacle_set_textures(rfworacle_context *c, const rfwb200_texture_data *tex, size_t n)
	{
		c->textures.resize(n);
		c->uint_texels.clear();
		c->float_texels.clear();
		for (size_t i = 0; i < n; i++)
		{
			TextureDesc d{tex[i].type, tex[i].width, tex[i].height, tex[i].texel_count, 0};
			if (tex[i].type == RFWB200_TEX_UINT)
			{
				d.addr = uint32_t(c->uint_texels.size());
				const uint32_t *p = static_cast<const uint32_t *>(tex[i].data);
				c->uint_texels.insert(c->uint_texels.end(), p, p + tex[i].texel_count);
			}
			else
			{
				d.addr = uint32_t(c->float_texels.size());
				const vec4 *p = static_cast<const vec4 *>(tex[i].data);
				c->float_texels.insert(c->float_texels.end(), p, p + tex[i].texel_count);
			}
			c->textures[i] = d;
		}
		return RFWB200_OK;
	}
	// CUDART/src/Context.cpp:160-198
	ORACLE_API int rfworacle_set_materials(rfworacle_context *c, const rfwb200_material *mats,
										   const rfwb200_material_tex_ids *ids, size_t n)
	{
		c->materials.assign(mats, mats + n);
		c->materials_raw.assign(mats, mats + n);
		for (size_t i = 0; i < n; i++)
		{
			rfwb200_material &m = c->materials[i];
			rfwb200_map_desc *slots[11] = {&m.tex0, &m.tex1, &m.tex2, &m.nmap0, &m.nmap1, &m.nmap2,
										   &m.smap, &m.rmap, nullptr, &m.cmap,	&m.amap};
			for (int k = 0; k < 11; k++)
			{
				const int id = ids ? ids[i].texture[k] : -1;
				if (id != -1 && slots[k])
				{
					if (id < 0 || size_t(id) >= c->textures.size())
						return fail("material references unknown texture");
					slots[k]->texaddr = c->textures[id].addr;
				}
			}
		}
		return RFWB200_OK;
	}
	ORACLE_API int rfworacle_set_mesh(rfworacle_context *c, size_t index, const rfwb200_mesh *mesh)
	{
		if (!mesh || !mesh->vertices || !mesh->triangles)
			return fail("mesh needs vertices and triangles");
		if (index >= c->meshes.size())
			c->meshes.resize(index + 1);
		MeshData &m = c->meshes[index];
		m.vertices.resize(mesh->vertex_count);
		memcpy(static_cast<void *>(m.vertices.data()), mesh->vertices, mesh->vertex_count * sizeof(vec4));
		m.indices.clear();
		if (mesh->indices)
			m.indices.assign(mesh->indices, mesh->indices + 3 * mesh->triangle_count);
		m.triangles.assign(mesh->triangles, mesh->triangles + mesh->triangle_count);
		m.dirty = true;
		return RFWB200_OK;
	}
	ORACLE_API int rfworacle_set_instance(rfworacle_context *c, size_t i, size_t mesh_index, const float transform[16],
										  const float normal_matrix[9])
	{
		if (mesh_index >= c->meshes.size())
			return fail("instance references unknown mesh");
		if (i >= c->instances.size())
			c->instances.resize(i + 1);
		Instance &in = c->instances[i];
		in.mesh = int(mesh_index);
		memcpy(in.transform.m, transform, sizeof(float) * 16);
		in.inverse = inverse(in.transform); // top_level_bvh.cpp:309
		memcpy(in.normal.m, normal_matrix, sizeof(float) * 9);
		return RFWB200_OK;
	}
	ORACLE_API int rfworacle_set_lights(rfworacle_context *c, rfwb200_light_count n, const rfwb200_area_light *a,
										const rfwb200_point_light *p, const rfwb200_spot_light *s,
										const rfwb200_directional_light *d)
	{
		c->area_lights.assign(a, a + (a ? n.area : 0));
		c->point_lights.assign(p, p + (p ? n.point : 0));
		c->spot_lights.assign(s, s + (s ? n.spot : 0));
		c->dir_lights.assign(d, d + (d ? n.directional : 0));
		return RFWB200_OK;
	}
	ORACLE_API int rfworacle_update(rfworacle_context *c)
	{
		// per-mesh BVH (bvh_tree.cpp:419-452: per-triangle AABB grown by 1e-5)
		for (size_t mi = 0; mi < c->meshes.size(); mi++)
		{
			MeshData &m = c->meshes[mi];
			if (!m.dirty)
				continue;
			const size_t nt = m.triangles.size();
			std::vector<AABB> aabbs(nt);
			for (size_t t = 0; t < nt; t++)
			{
				AABB a = AABB::invalid();
				a.grow(m.tri_vertex(t, 0)), a.grow(m.tri_vertex(t, 1)), a.grow(m.tri_vertex(t, 2));
				a.offset_by(1e-5f);
				aabbs[t] = a;
			}
			build_bvh(m.bvh, aabbs);
			m.dirty = false;
		}
		// TLAS over world bounds of the 8 transformed corners + 1e-6 (top_level_bvh.cpp:314-342)
		std::vector<AABB> inst_aabbs(c->instances.size());
		for (size_t i = 0; i < c->instances.size(); i++)
		{
			Instance &in = c->instances[i];
			AABB w = AABB::invalid();
			if (in.mesh >= 0 && !c->meshes[in.mesh].triangles.empty())
			{
				const AABB &b = c->meshes[in.mesh].bvh.root_bounds;
				for (int k = 0; k < 8; k++)
					w.grow(in.transform.mul_point(
						vec3((k & 1) ? b.bmax[0] : b.bmin[0], (k & 2) ? b.bmax[1] : b.bmin[1], (k & 4) ? b.bmax[2] : b.bmin[2])));
				w.offset_by(1e-6f);
			}
			else
				for (int a = 0; a < 3; a++)
					w.bmin[a] = w.bmax[a] = 0;
			in.world = w;
			inst_aabbs[i] = w;
		}
		build_bvh(c->tlas, inst_aabbs);
		return RFWB200_OK;
	}
	ORACLE_API int rfworacle_set_setting(rfworacle_context *c, const char *key, const char *value)
	{
		const std::string k = key ? key : "", v = value ? value : "";
		if (k == "spp")
			c->spp = std::max(1, atoi(v.c_str()));
		else if (k == "max_path_length")
			c->max_path_length = std::max(0, atoi(v.c_str()));
		else if (k == "mode")
			c->mode_pt = (v == "pt");
		else if (k == "clamp")
			c->clamp_value = float(atof(v.c_str()));
		else if (k == "survival_scale")
			c->survival_scale = (v == "on" || v == "1");
		else if (k == "bsdf_random_order")
			g_bsdf_randoms_right_to_left = (v == "rtl");
		else if (k == "cudart_conventions")
			g_cudart_conventions = (v == "on" || v == "1");
		else if (k == "smem_nodes" || k == "threads")
		{
			if (k == "threads")
				g_threads = std::max(1, atoi(v.c_str()));
		}
		else
			return fail("unknown setting");
		return RFWB200_OK;
	}
	ORACLE_API int rfworacle_num_threads(void)
	{
		return num_threads();
	}
	ORACLE_API int rfworacle_render_frame(rfworacle_context *c, const rfwb200_camera_view *view, int status)
	{
		if (!c->width)
			return fail("init first");
		if (c->mode_pt && c->blue_noise.empty())
			return fail("blue noise table not set");
		const auto t0 = std::chrono::steady_clock::now();
		if (status == RFWB200_RESET)
		{
			std::fill(c->accumulator.begin(), c->accumulator.end(), vec4());
			c->sample_index = 0;
		}
		c->counters = Counters();
		c->last_spp = uint32_t(c->spp);
		for (int s = 0; s < c->spp; s++)
		{
			if (c->mode_pt)
				render_sample_pt(*c, *view);
			else
				render_emode(*c, *view); // EmbreeRT renders one sample per call, no accumulation
		}
		if (c->mode_pt)
		{
			// blit_buffer (Kernels.cu:181-203): accumulator * 1/samples
			const float scale = 1.0f / float(c->sample_index);
			for (size_t i = 0; i < c->accumulator.size(); i++)
				c->framebuffer[i] = c->accumulator[i] * scale;
			c->counters.n_acc += 0;
		}
		c->last_render_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
		return RFWB200_OK;
	}
	ORACLE_API int rfworacle_read_framebuffer(rfworacle_context *c, float *out, size_t capacity_pixels)
	{
		if (capacity_pixels < c->framebuffer.size())
			return fail("buffer too small");
		memcpy(out, static_cast<const void *>(c->framebuffer.data()), c->framebuffer.size() * sizeof(vec4));
		return RFWB200_OK;
	}
	ORACLE_API size_t rfworacle_local_pixel_count(const rfworacle_context *c) { return size_t(c->width) * c->height; }
	ORACLE_API int rfworacle_set_probe_index(rfworacle_context *c, uint32_t x, uint32_t y)
	{
		c->probe_x = x, c->probe_y = y;
		return RFWB200_OK;
	}
	ORACLE_API int rfworacle_get_probe_results(rfworacle_context *c, uint32_t *inst, uint32_t *prim, float *dist)
	{
		*inst = c->probed_inst, *prim = c->probed_prim, *dist = c->probed_dist;
		return RFWB200_OK;
	}
	ORACLE_API int rfworacle_get_frame_counters(rfworacle_context *c, rfwb200_frame_counters *o)
	{
		o->n_gen = c->counters.n_gen, o->n_ext = c->counters.n_ext, o->n_shade = c->counters.n_shade;
		o->n_ext_out = c->counters.n_ext_out, o->n_nee = c->counters.n_nee, o->n_acc = c->counters.n_acc;
		o->pixels = uint64_t(c->width) * c->height, o->samples = c->last_spp;
		return RFWB200_OK;
	}
	ORACLE_API float rfworacle_last_render_ms(const rfworacle_context *c) { return c->last_render_ms; }

	ORACLE_API int rfworacle_trace_closest(rfworacle_context *c, const float *origins, const float *directions, size_t n,
										   float t_min, rfwb200_hit *hits)
	{
		parallel_for(int64_t(int64_t(n)), 256, [&](int64_t i) {
		{
			float t = 1e34f;
			int inst = -1, prim = -1;
			vec2 b = {0, 0};
			const bool hit = intersect_scene(*c, vec3(origins + 4 * i), vec3(directions + 4 * i), &inst, &prim, &t, &b, t_min);
			hits[i].t = hit ? t : 1e34f;
			hits[i].u = hit ? b.x : 0, hits[i].v = hit ? b.y : 0;
			hits[i].inst_id = hit ? inst : -1, hits[i].prim_id = hit ? prim : -1;
		}
		});
		return RFWB200_OK;
	}
	ORACLE_API int rfworacle_trace_occluded(rfworacle_context *c, const float *origins, const float *directions,
											const float *t_max, size_t n, float t_min, uint8_t *occluded)
	{
		parallel_for(int64_t(int64_t(n)), 256, [&](int64_t i) {
			occluded[i] = is_occluded(*c, vec3(origins + 4 * i), vec3(directions + 4 * i), t_min, t_max[i]) ? 1 : 0;
		});
		return RFWB200_OK;
	}
	// single-triangle test, for tie analysis in the parity tests
	ORACLE_API int rfworacle_intersect_prim(rfworacle_context *c, const float *origin, const float *direction, int inst,
											int prim, float t_min, float *t_out)
	{
		if (inst < 0 || size_t(inst) >= c->instances.size())
			return fail("bad instance");
		const Instance &in = c->instances[inst];
		const MeshData &m = c->meshes[in.mesh];
		if (prim < 0 || size_t(prim) >= m.triangles.size())
			return fail("bad prim");
		const vec3 o = in.inverse.mul_point(vec3(origin)), d = in.inverse.mul_dir(vec3(direction));
		float t = 1e34f;
		vec2 b;
		const bool hit = intersect_triangle(o, d, t_min, &t, m.tri_vertex(prim, 0), m.tri_vertex(prim, 1), m.tri_vertex(prim, 2), &b, T_EPSILON);
		*t_out = hit ? t : 1e34f;
		return RFWB200_OK;
	}
	ORACLE_API int rfworacle_generate_primary(rfworacle_context *c, const rfwb200_camera_view *view, uint32_t sample_index,
											  float *origins, float *directions, size_t capacity)
	{
		const size_t P = size_t(c->width) * c->height;
		if (capacity < P)
			return fail("buffer too small");
		parallel_for(int64_t(int64_t(P)), 256, [&](int64_t i) {
		{
			vec3 O, D;
			if (c->mode_pt)
				generate_primary_pt(*c, *view, uint32_t(i), sample_index, O, D);
			else
				generate_primary_emode(*c, *view, uint32_t(i), sample_index, O, D);
			origins[4 * i] = O.x, origins[4 * i + 1] = O.y, origins[4 * i + 2] = O.z;
			origins[4 * i + 3] = u2f((uint32_t(i) << 8) + 1);
			directions[4 * i] = D.x, directions[4 * i + 1] = D.y, directions[4 * i + 2] = D.z, directions[4 * i + 3] = 0;
		}
		});
		return RFWB200_OK;
	}
	// Stage-level shade (Kernels.cu:571-794 over a caller-supplied wavefront): path i = (O, D, T, hit) in the layout of
	// the reference's path buffers — O.w = (pathIndex << 8) | flags with pathIndex = y * width + x, D.w = packed normal of
	// the previous vertex, T.w = pdf of the sampled direction, hit = (bits(w0_16 | w1_16 << 16), bits(instance),
	// bits(primitive) or -1 for a miss, t).  Runs shade_path — the very function the oracle's frames run — on every path
	// and returns its three outputs per path, unset ones zeroed: flags_out bit 0 = extension ray written (ext_O/D/T),
	// bit 1 = connect entry written (con_O/D/E), bit 2 = contribution accumulated (acc, xyz).  The context's own
	// max_path_length decides whether a path of this length may continue; the probe is left alone.
	ORACLE_API int rfworacle_shade_stage(rfworacle_context *c, const rfwb200_camera_view *view, const float *O, const float *D,
										 const float *T, const float *hit, size_t n, uint32_t path_length, uint32_t samples_taken,
										 uint32_t *flags_out, float *ext_O, float *ext_D, float *ext_T, float *con_O, float *con_D,
										 float *con_E, float *acc)
	{
		if (!c || !view || !O || !D || !T || !hit || !flags_out || !ext_O || !ext_D || !ext_T || !con_O || !con_D || !con_E || !acc)
			return fail("bad arguments");
		for (size_t i = 0; i < n; i++)
		{
			const int prim = int(f2u(hit[4 * i + 2]));
			if (prim < 0)
				continue;
			const uint32_t inst = f2u(hit[4 * i + 1]);
			if (inst >= c->instances.size() || size_t(prim) >= c->meshes[c->instances[inst].mesh].triangles.size())
				return fail("hit record names an unknown instance or primitive");
		}
		const uint32_t px = c->probe_x, py = c->probe_y;
		c->probe_x = c->probe_y = 0x7fffffffu; // no path index equals the probe pixel
		parallel_for(int64_t(n), 256, [&](int64_t i) {
			PathState p;
			p.O = vec4(O[4 * i], O[4 * i + 1], O[4 * i + 2], O[4 * i + 3]);
			p.D = vec4(D[4 * i], D[4 * i + 1], D[4 * i + 2], D[4 * i + 3]);
			p.T = vec4(T[4 * i], T[4 * i + 1], T[4 * i + 2], T[4 * i + 3]);
			p.hit = vec4(hit[4 * i], hit[4 * i + 1], hit[4 * i + 2], hit[4 * i + 3]);
			ShadeOut out;
			shade_path(*c, *view, p, path_length, samples_taken, out);
			const auto put = [i](float *dst, const vec4 &v) { dst[4 * i] = v.x, dst[4 * i + 1] = v.y, dst[4 * i + 2] = v.z, dst[4 * i + 3] = v.w; };
			const vec4 zero(0, 0, 0, 0);
			flags_out[i] = (out.has_ext ? 1u : 0u) | (out.has_shadow ? 2u : 0u) | (out.acc ? 4u : 0u);
			put(ext_O, out.has_ext ? out.ext.O : zero), put(ext_D, out.has_ext ? out.ext.D : zero), put(ext_T, out.has_ext ? out.ext.T : zero);
			put(con_O, out.has_shadow ? out.shadow.O : zero), put(con_D, out.has_shadow ? out.shadow.D : zero);
			put(con_E, out.has_shadow ? out.shadow.E : zero);
			put(acc, out.acc ? vec4(out.acc_value, 0.0f) : zero);
		});
		c->probe_x = px, c->probe_y = py;
		return RFWB200_OK;
	}


	// ---- hooks for pinning against the reference's own headers (oracle/ref_build, tests/test_ref_pin.py) ----
	ORACLE_API int rfworacle_export_mesh_mbvh(rfworacle_context *c, size_t mesh, void *nodes_out, size_t node_cap,
											  uint32_t *prims_out, size_t prim_cap, size_t *n_nodes, size_t *n_prims)
	{
		if (mesh >= c->meshes.size())
			return fail("bad mesh");
		const Bvh &b = c->meshes[mesh].bvh;
		*n_nodes = b.mnodes.size(), *n_prims = b.prim_indices.size();
		if (nodes_out && node_cap >= b.mnodes.size())
			memcpy(nodes_out, b.mnodes.data(), b.mnodes.size() * sizeof(MBVHNode));
		if (prims_out && prim_cap >= b.prim_indices.size())
			memcpy(prims_out, b.prim_indices.data(), b.prim_indices.size() * sizeof(uint32_t));
		return RFWB200_OK;
	}
	// test hook: this file's builder (bvh_partition / bvh_subdivide / mbvh_merge_nodes) over caller-supplied boxes (n x (min3, max3)),
	// for the pin against the reference's own in-tree node code (oracle/ref_build/ref_bvh_shim.cpp, tests/test_ref_pin_bvh.py).
	// nodes_out: 32-byte BVH2 nodes, mnodes_out: 128-byte 4-wide nodes, prims_out: the primitive order; capacities in elements.
	ORACLE_API int rfworacle_build_bvh_from_aabbs(const float *aabbs6, size_t n, void *nodes_out, size_t node_cap, uint32_t *prims_out,
												  void *mnodes_out, size_t mnode_cap, size_t *n_nodes, size_t *n_mnodes)
	{
		static_assert(sizeof(AABB) == 24 && sizeof(BVHNode) == 32 && sizeof(MBVHNode) == 128, "node layouts");
		if (!aabbs6 || !nodes_out || !prims_out || !mnodes_out || !n_nodes || !n_mnodes)
			return fail("bad arguments");
		std::vector<AABB> aabbs(n);
		memcpy(static_cast<void *>(aabbs.data()), aabbs6, n * sizeof(AABB));
		Bvh b;
		build_bvh(b, aabbs);
		*n_nodes = b.nodes.size(), *n_mnodes = b.mnodes.size();
		if (b.nodes.size() > node_cap || b.mnodes.size() > mnode_cap)
			return fail("buffer too small");
		memcpy(nodes_out, static_cast<const void *>(b.nodes.data()), b.nodes.size() * sizeof(BVHNode));
		memcpy(mnodes_out, static_cast<const void *>(b.mnodes.data()), b.mnodes.size() * sizeof(MBVHNode));
		memcpy(prims_out, b.prim_indices.data(), n * sizeof(uint32_t));
		return RFWB200_OK;
	}
	// the top-level MBVH over the instances, in the reference's node layout (for the host-compiled reference kernels)
	// test hook: the matrices this context traverses and shades instance i with (column-major mat4 / mat3), so that the
	// reference's host-compiled kernels (oracle/ref_build/ref_kernels_shim.cpp) can be given exactly the same ones
	ORACLE_API int rfworacle_export_instance(rfworacle_context *c, size_t i, float transform16[16], float inverse16[16],
											 float normal9[9])
	{
		if (i >= c->instances.size())
			return fail("unknown instance");
		const Instance &in = c->instances[i];
		memcpy(transform16, in.transform.m, sizeof(float) * 16);
		memcpy(inverse16, in.inverse.m, sizeof(float) * 16);
		memcpy(normal9, in.normal.m, sizeof(float) * 9);
		return RFWB200_OK;
	}
	// test hook: the E-mode material step for (instance, primitive) at Embree barycentrics (u, v) = weights of vertex 1, 2
	ORACLE_API int rfworacle_emode_material(rfworacle_context *c, int inst, int prim, float u, float v, float *color_out,
											float *iN_out)
	{
		if (inst < 0 || size_t(inst) >= c->instances.size())
			return fail("bad instance");
		const Instance &in = c->instances[inst];
		const MeshData &m = c->meshes[in.mesh];
		if (prim < 0 || size_t(prim) >= m.triangles.size())
			return fail("bad prim");
		vec3 color, iN;
		emode_material(*c, in, m.triangles[prim], vec3(1.0f - u - v, u, v), color, iN);
		color_out[0] = color.x, color_out[1] = color.y, color_out[2] = color.z;
		iN_out[0] = iN.x, iN_out[1] = iN.y, iN_out[2] = iN.z;
		return RFWB200_OK;
	}
	ORACLE_API int rfworacle_export_tlas_mbvh(rfworacle_context *c, void *nodes_out, size_t node_cap, uint32_t *prims_out,
											  size_t prim_cap, size_t *n_nodes, size_t *n_prims)
	{
		const Bvh &b = c->tlas;
		*n_nodes = b.mnodes.size(), *n_prims = b.prim_indices.size();
		if (nodes_out && node_cap >= b.mnodes.size())
			memcpy(nodes_out, b.mnodes.data(), b.mnodes.size() * sizeof(MBVHNode));
		if (prims_out && prim_cap >= b.prim_indices.size())
			memcpy(prims_out, b.prim_indices.data(), b.prim_indices.size() * sizeof(uint32_t));
		return RFWB200_OK;
	}
	ORACLE_API int rfworacle_shading_data(rfworacle_context *c, int inst, int prim, const float *D, float u, float v,
										  float cone_width, float *color_out, uint32_t *flags_out, float *N_out, float *iN_out)
	{
		if (inst < 0 || size_t(inst) >= c->instances.size())
			return fail("bad instance");
		const Instance &in = c->instances[inst];
		const MeshData &m = c->meshes[in.mesh];
		if (prim < 0 || size_t(prim) >= m.triangles.size())
			return fail("bad prim");
		vec3 N, iN, T, B;
		// the caller passes Moller-Trumbore u, v (weights of vertex 1, 2); getShadingData takes the weights of vertex 0, 1
		const ShadingData sd = getShadingData(*c, vec3(D), 1.0f - u - v, u, cone_width, m.triangles[prim], N, iN, T, B, in.normal);
		color_out[0] = sd.color.x, color_out[1] = sd.color.y, color_out[2] = sd.color.z;
		*flags_out = sd.flags;
		N_out[0] = N.x, N_out[1] = N.y, N_out[2] = N.z;
		iN_out[0] = iN.x, iN_out[1] = iN.y, iN_out[2] = iN.z;
		return RFWB200_OK;
	}
	ORACLE_API void rfworacle_random_point_on_light(rfworacle_context *c, float r0, float r1, const float *I, const float *N,
													float *P_out, float *pick_out, float *pdf_out, float *color_out)
	{
		float pick = 0, pdf = 0;
		vec3 color(0.0f);
		const vec3 P = RandomPointOnLight(*c, r0, r1, vec3(I), vec3(N), pick, pdf, color);
		P_out[0] = P.x, P_out[1] = P.y, P_out[2] = P.z;
		*pick_out = pick, *pdf_out = pdf;
		color_out[0] = color.x, color_out[1] = color.y, color_out[2] = color.z;
	}
	ORACLE_API float rfworacle_light_pick_prob(rfworacle_context *c, int idx, const float *O, const float *N, const float *I)
	{
		return LightPickProb(*c, idx, vec3(O), vec3(N), vec3(I));
	}
	ORACLE_API float rfworacle_light_pdf(const float *D, float t, float area, const float *LN)
	{
		return CalculateLightPDF(vec3(D), t, area, vec3(LN));
	}
	ORACLE_API void rfworacle_bsdf_sample_r(const float *color, const float *absorption, const uint32_t *params, const float *N,
											const float *wo, float t, int backfacing, float r3, float r4, float *wi_out,
											float *bsdf_out, float *pdf_out)
	{
		ShadingData sd;
		sd.color = vec3(color);
		sd.absorption = absorption ? vec3(absorption) : vec3(0.0f);
		for (int i = 0; i < 4; i++)
			sd.parameters[i] = params[i];
		vec3 T, B, wi(0.0f);
		createTangentSpace(vec3(N), T, B);
		float pdf = 0;
		BSDFSample(sd, T, B, vec3(N), vec3(wo), wi, pdf, r3, r4);
		const vec3 r = BSDFEval(sd, vec3(N), vec3(wo), wi, t, backfacing != 0);
		wi_out[0] = wi.x, wi_out[1] = wi.y, wi_out[2] = wi.z;
		bsdf_out[0] = r.x, bsdf_out[1] = r.y, bsdf_out[2] = r.z;
		*pdf_out = pdf;
	}
	ORACLE_API int rfworacle_intersect_triangle(const float *org, const float *dir, float tmin, float tmax, const float *p0,
												const float *p1, const float *p2, float eps, float *t_out, float *uv_out)
	{
		float t = tmax;
		vec2 b = {0, 0};
		const bool hit = intersect_triangle(vec3(org), vec3(dir), tmin, &t, vec3(p0), vec3(p1), vec3(p2), &b, eps);
		*t_out = t, uv_out[0] = b.x, uv_out[1] = b.y;
		return hit ? 1 : 0;
	}
	ORACLE_API void rfworacle_tangent_space(const float *n, float *T, float *B)
	{
		vec3 t, b;
		createTangentSpace(vec3(n), t, b);
		T[0] = t.x, T[1] = t.y, T[2] = t.z, B[0] = b.x, B[1] = b.y, B[2] = b.z;
	}

	// ---- scalar known-answer hooks (tests/test_oracle.py) ----
	ORACLE_API uint32_t rfworacle_wang_hash(uint32_t s) { return WangHash(s); }
	ORACLE_API uint32_t rfworacle_random_int(uint32_t *s) { return RandomInt(*s); }
	ORACLE_API uint32_t rfworacle_xor128(uint32_t x, uint32_t n)
	{
		Xor128 r;
		r.x = x;
		uint32_t v = 0;
		for (uint32_t i = 0; i < n; i++)
			v = r.rand_uint();
		return v;
	}
	ORACLE_API uint32_t rfworacle_pack_normal(const float *n) { return PackNormal(vec3(n)); }
	ORACLE_API void rfworacle_unpack_normal(uint32_t p, float *out)
	{
		const vec3 n = UnpackNormal(p);
		out[0] = n.x, out[1] = n.y, out[2] = n.z;
	}
	ORACLE_API float rfworacle_blue_noise(rfworacle_context *c, int x, int y, int s, int d)
	{
		return blueNoiseSampler(c->blue_noise.data(), x, y, s, d);
	}
	ORACLE_API float rfworacle_half_to_float(uint16_t h) { return half2float(h); }
	ORACLE_API void rfworacle_random_barycentrics(float r0, float *out)
	{
		const vec3 b = RandomBarycentrics(r0);
		out[0] = b.x, out[1] = b.y, out[2] = b.z;
	}
	// Disney BSDF evaluation with a 16-byte parameter block, for KATs against oracle/_ref
	ORACLE_API void rfworacle_bsdf_eval(const float *color, const uint32_t *params, const float *N, const float *wo,
										const float *wi, float *bsdf_out, float *pdf_out)
	{
		ShadingData sd;
		sd.color = vec3(color);
		for (int i = 0; i < 4; i++)
			sd.parameters[i] = params[i];
		float pdf;
		const vec3 r = EvaluateBSDF(sd, vec3(N), vec3(wo), vec3(wi), pdf);
		bsdf_out[0] = r.x, bsdf_out[1] = r.y, bsdf_out[2] = r.z;
		*pdf_out = pdf;
	}
	ORACLE_API void rfworacle_bsdf_sample(const float *color, const uint32_t *params, const float *N, const float *wo,
										  uint32_t seed, float *wi_out, float *bsdf_out, float *pdf_out)
	{
		ShadingData sd;
		sd.color = vec3(color);
		for (int i = 0; i < 4; i++)
			sd.parameters[i] = params[i];
		vec3 T, B, wi;
		createTangentSpace(vec3(N), T, B);
		float pdf = 0;
		const vec3 r = SampleBSDF(sd, vec3(N), T, B, vec3(wo), 0.0f, false, wi, pdf, seed);
		wi_out[0] = wi.x, wi_out[1] = wi.y, wi_out[2] = wi.z;
		bsdf_out[0] = r.x, bsdf_out[1] = r.y, bsdf_out[2] = r.z;
		*pdf_out = pdf;
	}
}
