// kernels.cu — sm_100a wavefront path-tracing kernels.
//
// Stage map (reference semantics in parentheses; paths relative to /root/reference/RFW):
//   k_wavefront_trace<true>   generate + extend for camera rays (backends/CUDART/src/Kernels.cu:383-460)
//   k_shade     material eval + NEE + BSDF sample + compaction + sort-bin ranks (Kernels.cu:571-794,
//               CUDART/src/getShadingData.h, CUDART/src/lights.h, system/context/rfw/bsdf/*.h)
//   k_sort_scan / k_sort_move  re-ordering of the bounce queue by (origin cell, direction bin) — no counterpart in the reference
//   k_wavefront_trace<false>  extend (closest hit) for extension rays and connect (any hit) for shadow
//               rays of one bounce in ONE persistent launch (Kernels.cu:461-498, CUDAIntersect.h);
//               <..., TL> the same over a two-level scene (Kernels.cu:226-303)
//   k_fold      per-sample radiance folded into the accumulator in sample order, accumulator / samples, and — for a
//               tile-sharded frame — the peer stores into the display device's image (Kernels.cu:181-203)
//   k_emode     the image model of backends/EmbreeRT/src/Context.cpp:104-300 as one fused kernel
//
// B200 design: persistent CTAs sized from the occupancy API (multiples of 148 SMs) pull work from
// device-resident cursors and queue sizes, so no launch depends on a host-read count and a frame never synchronises with
// the host; all samples of a frame travel in one wavefront; ray / hit / throughput state is float4 SoA moved with 128-bit
// coalesced loads and stores (marked evict-first: it streams past a BVH that should stay in L2); the traversal reads
// 80-byte packed nodes through L1 (a TMA-staged shared-memory prefix — cp.async.bulk + mbarrier, UBLKCP in the SASS —
// exists for both node formats, setting smem_nodes, and was measured slower than L1 on the default kernels: DESIGN.md
// 3a); queue compaction is warp-aggregated (ballot + popc, one 64-bit atomic per warp for both queues), the sort-bin ranks
// use __match_any_sync; tensor cores are not used (no dense contraction on this path).
#include "kernels.h"

#include "../../include/rfwb200.h"
#include "bvh_build.h"

#include <cuda_fp16.h>
#include <stdint.h>
#include <stdlib.h>

// This file is compiled twice (rendering-fw_b200/Makefile): RFW_PART=1 builds the trace side — traversal, generate,
// finalize, E-mode, stage-level kernels — with IEEE arithmetic, so hit points do not depend on approximate division;
// RFW_PART=2 builds k_shade with -use_fast_math, as the reference builds its whole CUDA backend
// (RFW/backends/CUDART/CMakeLists.txt:7-9).  RFW_PART=3 builds the same shade kernel once more WITHOUT fast math under
// the name k_shade_ieee (setting "shade_math" = "ieee"): the arithmetic of the CPU oracle, used by the parity tests that
// compare images at scene scales where one ulp decides a connect ray.  RFW_PART=0 (default) is everything in one unit.
#ifndef RFW_PART
#define RFW_PART 0
#endif
#if RFW_PART == 3
#define RFW_SHADE_ONLY 1
#define K_SHADE k_shade_ieee
#define LAUNCH_SHADE launch_shade_ieee
#define SHADE_OCCUPANCY shade_occupancy_ieee
#define SHADE_PRELOAD shade_preload_ieee
#undef RFW_PART
#define RFW_PART 2
#else
#define K_SHADE k_shade
#define LAUNCH_SHADE launch_shade
#define SHADE_OCCUPANCY shade_occupancy
#define SHADE_PRELOAD shade_preload_fast
#endif

namespace rfwb200
{
namespace // device code and kernels have internal linkage: the file is compiled into two objects
{

// ------------------------------------------------------------------------------------------------
// small math
// ------------------------------------------------------------------------------------------------
struct V3
{
	float x, y, z;
};
__device__ __forceinline__ V3 mk(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 mk(float a) { return V3{a, a, a}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b)
{
	return mk(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
__device__ __forceinline__ float length(V3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ V3 normalize(V3 a) { return a * (1.0f / sqrtf(dot(a, a))); }
__device__ __forceinline__ V3 reflect(V3 I, V3 N) { return I - N * (dot(N, I) * 2.0f); }
__device__ __forceinline__ bool any_nan(V3 a) { return isnan(a.x) || isnan(a.y) || isnan(a.z); }
__device__ __forceinline__ V3 ld3(const float *p) { return mk(p[0], p[1], p[2]); }
__device__ __forceinline__ float lerpf(float a, float b, float t) { return a + t * (b - a); }
__device__ __forceinline__ V3 lerp3(V3 a, V3 b, float t) { return a + (b - a) * t; }
__device__ __forceinline__ float sqr(float x) { return x * x; }
__device__ __forceinline__ float signf(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }

constexpr float INVPI = 0.318309886183790671537767526745028724f;
constexpr float PI = 3.14159265358979323846264338327950288f;
constexpr float INV2PI = 0.159154943091895335768883763372514362f;
constexpr float TWOPI = 6.28318530717958647692528676655900576f;
constexpr int MIPLEVELCOUNT = 5;		 // context/settings.h:3
constexpr float MIN_ROUGHNESS = 0.01f;	 // context/settings.h:4
constexpr int MAX_IS_LIGHTS = 16;		 // CUDART/src/Kernels.cu:21
constexpr uint32_t IS_SPECULAR = 1u;	 // CUDART/src/Kernels.cu:20
constexpr int32_t PRIM_MISS = -1;		 // hit.z on a miss (Kernels.cu:454)
constexpr int32_t PRIM_DEAD = -2;		 // padded pixel of an edge tile: never shaded
constexpr float T_EPSILON = 1e-6f;		 // CUDART/src/Kernels.cu:23

enum MatFlagBits
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
__device__ __forceinline__ bool has_flag(uint32_t f, int b) { return (f >> b) & 1u; }

// Wavefront state is written once and read once per bounce — gigabytes per frame streaming past a BVH of a few tens of
// megabytes that should stay in the 126 MB L2.  RFW_STREAM_HINTS marks those accesses evict-first (ld / st .cs):
// 1 = in the trace kernels, 2 = also in k_shade and the re-ordering pass.  Measured in profiles/r02 (sweep8).
#ifndef RFW_STREAM_HINTS
#define RFW_STREAM_HINTS 2
#endif
#if RFW_STREAM_HINTS >= 1
#define LD_TS(p) __ldcs(p)
#define ST_TS(p, v) __stcs((p), (v))
#else
#define LD_TS(p) (*(p))
#define ST_TS(p, v) (*(p) = (v))
#endif
#if RFW_STREAM_HINTS >= 2
#define LD_SS(p) __ldcs(p)
#define ST_SS(p, v) __stcs((p), (v))
#else
#define LD_SS(p) (*(p))
#define ST_SS(p, v) (*(p) = (v))
#endif

// ------------------------------------------------------------------------------------------------
// hashing / samplers — bsdf/tools.h:218-235, Kernels.cu:205-224, utils/xor128.h:20-27
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t WangHash(uint32_t s)
{
	s = (s ^ 61u) ^ (s >> 16);
	s *= 9u;
	s = s ^ (s >> 4);
	s *= 0x27d4eb2du;
	s = s ^ (s >> 15);
	return s;
}
__device__ __forceinline__ uint32_t RandomInt(uint32_t &s)
{
	s ^= s << 13;
	s ^= s >> 17;
	s ^= s << 5;
	return s;
}
__device__ __forceinline__ float RandomFloat(uint32_t &s) { return float(RandomInt(s)) * 2.3283064365387e-10f; }

__device__ __forceinline__ float blueNoiseSampler(const uint8_t *__restrict__ bn, int x, int y, int sampleIdx, int dim)
{
	x &= 127, y &= 127, sampleIdx &= 255, dim &= 255;
	const int ranked = sampleIdx ^ int(__ldg(bn + dim + (x + y * 128) * 8 + 65536 * 3));
	int value = int(__ldg(bn + dim + ranked * 256));
	value ^= int(__ldg(bn + (dim & 7) + (x + y * 128) * 8 + 65536));
	return (0.5f + float(value)) * (1.0f / 256.0f);
}

// ------------------------------------------------------------------------------------------------
// screen tiling: local work index -> pixel.  A tile is tile_w x tile_h pixels, tiles are dealt
// round-robin to ranks, and inside a tile consecutive 32 indices form an 8x4 pixel block so one
// warp traces a compact bundle of camera rays.
// ------------------------------------------------------------------------------------------------
// n / d for n, d < 2^24 with the reciprocal of d computed once on the host: the float estimate is within one of the
// quotient and is corrected by the remainder (a 32-bit division by a run-time divisor is ~20 instructions, and the index
// arithmetic below runs several of them per path)
__device__ __forceinline__ uint32_t fast_div(uint32_t n, uint32_t d, float inv_d)
{
	uint32_t q = __float2uint_rz(__uint2float_rn(n) * inv_d);
	const int32_t r = int32_t(n - q * d);
	q += r >= int32_t(d) ? 1u : 0u;
	q -= r < 0 ? 1u : 0u;
	return q;
}

__device__ __forceinline__ bool local_to_pixel(const ShardView &sh, uint32_t j, uint32_t &x, uint32_t &y)
{
	const uint32_t tp = sh.tile_w * sh.tile_h;
	const uint32_t lt = fast_div(j, tp, sh.inv_tile_pixels), w = j - lt * tp;
	const uint32_t gt = lt * sh.world + sh.rank;
	const uint32_t ty = fast_div(gt, sh.tiles_x, sh.inv_tiles_x), tx = gt - ty * sh.tiles_x;
	const uint32_t blk = w >> 5, lane = w & 31u;
	const uint32_t bpr = sh.tile_w >> 3;
	const uint32_t by = fast_div(blk, bpr, sh.inv_blocks_per_row), bx = blk - by * bpr;
	// inside the 8x4 block the 32 pixels follow a Z curve (bits x0 y0 x1 y1 x2): any aligned run of 4 / 8 / 16 consecutive pixels
	// is a compact 2x2 / 4x2 / 4x4 patch — the pixels a warp of camera rays holds when the samples of a pixel are neighbours
	x = tx * sh.tile_w + bx * 8u + ((lane & 1u) | ((lane >> 1) & 2u) | ((lane >> 2) & 4u));
	y = ty * sh.tile_h + by * 4u + (((lane >> 1) & 1u) | ((lane >> 2) & 2u));
	return (x < sh.width) & (y < sh.height) & (ty < sh.tiles_y);
}

// work item of a wavefront -> (local pixel, sample of the wavefront); see BatchView (device_types.h)
__device__ __forceinline__ void item_to_pixel_sample(const BatchView &bv, uint32_t item, uint32_t &j, uint32_t &s)
{
	const uint32_t grp = item >> 5, blk = fast_div(grp, bv.spp, bv.inv_spp);
	if (bv.sample_minor)
	{
		// the samples of a pixel are neighbours: a warp of camera rays is 32 / spp pixels x spp samples
		const uint32_t l = item - ((blk * bv.spp) << 5), pix = fast_div(l, bv.spp, bv.inv_spp);
		s = l - pix * bv.spp;
		j = (blk << 5) | pix;
		return;
	}
	s = grp - blk * bv.spp;
	j = (blk << 5) | (item & 31u);
}
// where sample s of local pixel j lives (the inverse of item_to_pixel_sample)
__device__ __forceinline__ uint32_t pixel_sample_to_item(const BatchView &bv, uint32_t j, uint32_t s)
{
	const uint32_t blk = j >> 5, lane = j & 31u;
	return bv.sample_minor ? (((blk * bv.spp) << 5) + lane * bv.spp + s) : ((((blk * bv.spp) + s) << 5) | lane);
}

// ------------------------------------------------------------------------------------------------
// bin of a bounce ray for the re-ordering pass: Morton code of the origin's grid cell + direction octant.
// Rays of one bin start in the same part of the scene and run towards the same octant, so the lanes of a warp walk
// the same nodes (Garanzha & Loop 2010, "Fast ray sorting and breadth-first packet traversal"; the reference traces
// its extension rays in the order shade happened to append them, Kernels.cu:788-793).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread3(uint32_t x)
{
	x = (x | (x << 16)) & 0x030000FFu;
	x = (x | (x << 8)) & 0x0300F00Fu;
	x = (x | (x << 4)) & 0x030C30C3u;
	x = (x | (x << 2)) & 0x09249249u;
	return x;
}
__device__ __forceinline__ uint32_t ray_bin(const SortGrid &g, int cell_bits, int dir_major, int dir_bits, float ox, float oy, float oz,
											float dx, float dy, float dz)
{
	const int hi = (1 << cell_bits) - 1;
	const int cx = min(max(__float2int_rd((ox - g.lo[0]) * g.scale[0]), 0), hi);
	const int cy = min(max(__float2int_rd((oy - g.lo[1]) * g.scale[1]), 0), hi);
	const int cz = min(max(__float2int_rd((oz - g.lo[2]) * g.scale[2]), 0), hi);
	const uint32_t cell = spread3(uint32_t(cx)) | (spread3(uint32_t(cy)) << 1) | (spread3(uint32_t(cz)) << 2);
	uint32_t oct = (dx < 0.0f ? 1u : 0u) | (dy < 0.0f ? 2u : 0u) | (dz < 0.0f ? 4u : 0u);
	if (dir_bits == 5) // octant x dominant axis: 24 direction bins
	{
		const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
		oct |= (ax >= ay ? (ax >= az ? 0u : 2u) : (ay >= az ? 1u : 2u)) << 3;
	}
	if (dir_bits == 6) // octahedral map of the direction, 8 x 8 bins
	{
		const float inv = 1.0f / (fabsf(dx) + fabsf(dy) + fabsf(dz) + 1e-30f);
		float u = dx * inv, v = dy * inv;
		if (dz < 0.0f)
		{
			const float tu = (1.0f - fabsf(v)) * (u < 0.0f ? -1.0f : 1.0f);
			v = (1.0f - fabsf(u)) * (v < 0.0f ? -1.0f : 1.0f);
			u = tu;
		}
		const int iu = min(max(int((u * 0.5f + 0.5f) * 8.0f), 0), 7), iv = min(max(int((v * 0.5f + 0.5f) * 8.0f), 0), 7);
		oct = uint32_t(iu) | (uint32_t(iv) << 3);
	}
	return dir_major ? ((oct << (3 * cell_bits)) | cell) : ((cell << dir_bits) | oct);
}

// ------------------------------------------------------------------------------------------------
// TMA bulk staging of the BVH prefix into shared memory
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void stage_bytes(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *mbar);
__device__ __forceinline__ void stage_nodes(float4 *smem_nodes, const BvhNode4 *gnodes, uint32_t n_stage, uint64_t *mbar)
{
	stage_bytes(smem_nodes, gnodes, n_stage * uint32_t(sizeof(BvhNode4)), mbar);
}
// one TMA bulk copy (cp.async.bulk, SASS: UBLKCP) of `bytes` (a multiple of 16) from global to shared memory, completion on an
// mbarrier every thread of the CTA then waits on
__device__ __forceinline__ void stage_bytes(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *mbar)
{
	float4 *smem_nodes = static_cast<float4 *>(smem_dst);
	const void *gnodes = gsrc;
	const uint32_t n_stage = bytes;
	if (n_stage == 0)
		return;
	const uint32_t bar = smem_u32(mbar);
	if (threadIdx.x == 0)
	{
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (threadIdx.x == 0)
	{
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
						 smem_u32(smem_nodes)),
					 "l"(gnodes), "r"(bytes), "r"(bar)
					 : "memory");
	}
	// every thread waits for the bytes to land (phase 0)
	uint32_t done = 0;
	while (!done)
	{
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
					 : "=r"(done)
					 : "r"(bar), "r"(0u)
					 : "memory");
	}
}

// ------------------------------------------------------------------------------------------------
// extend / connect: stack-based traversal of the flattened BVH4 + Moller-Trumbore
// (CUDAIntersect.h:48-94,155-197,270-322,391-439).  Children hit by the ray are ordered near to far
// with the reference's trick of stuffing the slot index into the two mantissa LSBs of tmin.
// ------------------------------------------------------------------------------------------------
struct NodeRegs
{
	float4 minx, maxx, miny, maxy, minz, maxz;
	int4 child;
};

__device__ __forceinline__ NodeRegs load_node(const SceneView &sc, const float4 *__restrict__ snodes, uint32_t n_smem,
											  uint32_t idx)
{
	// One generic 128-bit load path for both homes of a node (staged shared-memory prefix or global):
	// selecting the pointer instead of branching keeps the lanes of a warp together — a branch here split
	// every warp into a shared-memory half and a global half that executed one after the other
	// (profiles/r01: 11 + 11 of 32 threads on the two load sequences).
	const float4 *p = idx < n_smem ? snodes + size_t(idx) * 8 : reinterpret_cast<const float4 *>(sc.nodes) + size_t(idx) * 8;
	NodeRegs n;
	n.minx = p[0], n.maxx = p[1], n.miny = p[2], n.maxy = p[3], n.minz = p[4], n.maxz = p[5];
	n.child = *reinterpret_cast<const int4 *>(p + 6);
	return n;
}

__device__ __forceinline__ float slab(float lo, float hi, float idir, float ood, float &tfar_out)
{
	const float t1 = fmaf(lo, idir, -ood), t2 = fmaf(hi, idir, -ood);
	tfar_out = fmaxf(t1, t2);
	return fminf(t1, t2);
}

#define CHILD_T(K, C)                                                                                                   \
	{                                                                                                                   \
		float fx, fy, fz;                                                                                               \
		const float nx = slab(n.minx.C, n.maxx.C, idx, oodx, fx);                                                       \
		const float ny = slab(n.miny.C, n.maxy.C, idy, oody, fy);                                                       \
		const float nz = slab(n.minz.C, n.maxz.C, idz, oodz, fz);                                                       \
		const float tn = fmaxf(fmaxf(nx, ny), nz), tf = fminf(fminf(fx, fy), fz);                                       \
		const bool h = (tf >= tn) & (tn < tmax) & (tf >= tmin);                                                         \
		key##K = h ? ((__float_as_uint(fmaxf(tn, 0.0f)) & 0xFFFFFFFCu) | uint32_t(K)) : MISSKEY(K);      \
	}
// Sort keys are the BIT PATTERNS of the clamped (>= 0) entry distances compared as unsigned integers — monotonic for
// non-negative floats and, unlike float min/max, immune to flush-to-zero: a ray that starts inside a box has
// tn = 0 and its key is just the slot index, a denormal as a float (the reference compares them as floats,
// CUDAIntersect.h:185-194, which only works without -ftz).
#define MISSKEY(K) (0x7f000000u | uint32_t(K))
// same test with the near / far planes already selected by the sign of the ray direction (no per-axis min/max)
#define CHILD_N(K, C)                                                                                                   \
	{                                                                                                                   \
		const float tn = fmaxf(fmaxf(fmaf(nearx.C, idx, -oodx), fmaf(neary.C, idy, -oody)), fmaf(nearz.C, idz, -oodz));  \
		const float tf = fminf(fminf(fmaf(farx.C, idx, -oodx), fmaf(fary.C, idy, -oody)), fmaf(farz.C, idz, -oodz));     \
		const bool h = (tf >= tn) & (tn < tmax) & (tf >= tmin);                                                         \
		key##K = h ? ((__float_as_uint(fmaxf(tn, 0.0f)) & 0xFFFFFFFCu) | uint32_t(K)) : MISSKEY(K);      \
	}
#define NO_CHILD_HIT(K0, K1, K2, K3) (min(min(K0, K1), min(K2, K3)) >= 0x7f000000u) // MISSKEY is the largest key
#define CSWAP(A, B)                                                                                                     \
	{                                                                                                                   \
		const uint32_t lo_ = min(A, B), hi_ = max(A, B);                                                                \
		A = lo_, B = hi_;                                                                                               \
	}
#define PICK(KEY) ((((KEY)&3u) == 0u) ? n.child.x : ((((KEY)&3u) == 1u) ? n.child.y : ((((KEY)&3u) == 2u) ? n.child.z : n.child.w)))

// Hit-record barycentrics as CUDART stores them (CUDAIntersect.h:82-87, Kernels.cu:457): 16-bit weights of vertex 0 and
// vertex 1 (the reference computes them as area ratios; 1 - u - v and u of Moller-Trumbore are the same weights), so the
// 16-bit quantisation falls on the same two weights as in the reference and shading interpolates as getShadingData.h:123,140.
__device__ __forceinline__ uint32_t pack_barycentrics(float mt_u, float mt_v)
{
	return uint32_t(65535.0f * (1.0f - mt_u - mt_v)) | (uint32_t(65535.0f * mt_u) << 16);
}

// one ray through the compressed 8-wide BVH (cwbvh.h): the per-thread form of k_wavefront_trace_cw's loop, used by the
// stage-level kernels and E-mode
template <bool ANY_HIT>
__device__ bool traverse_cw(const SceneView &sc, V3 o, V3 d, float tmin, float &tmax, uint32_t &hit_tri, float &hit_u, float &hit_v)
{
	const float tiny = 1e-30f;
	CwRay ray;
	ray.ox = o.x, ray.oy = o.y, ray.oz = o.z;
	ray.idx = 1.0f / (fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x));
	ray.idy = 1.0f / (fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y));
	ray.idz = 1.0f / (fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z));
	ray.octinv = cw_octinv(ray.idx, ray.idy, ray.idz);
	const float4 *__restrict__ tris = reinterpret_cast<const float4 *>(sc.tris);
	uint2 stack[CW_STACK];
	int sp = 0;
	uint2 ng = make_uint2(0u, 0x80000000u), tg = make_uint2(0u, 0u);
	bool found = false;
	for (;;)
	{
		if (ng.y > 0x00ffffffu)
		{
			const uint32_t bit = 31u - uint32_t(__clz(ng.y));
			const uint32_t imask = ng.y;
			ng.y &= ~(1u << bit);
			if (ng.y > 0x00ffffffu)
				stack[sp++] = ng;
			const uint32_t slot = (bit - 24u) ^ ray.octinv;
			const uint32_t node = ng.x + uint32_t(__popc(imask & ~(0xffffffffu << slot)));
			const uint4 *np_ = sc.cw_nodes + size_t(node) * 5;
			const uint4 q0 = __ldg(np_), q1 = __ldg(np_ + 1), q2 = __ldg(np_ + 2), q3 = __ldg(np_ + 3), q4 = __ldg(np_ + 4);
			const uint32_t w[20] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y,
									q2.z, q2.w, q3.x, q3.y, q3.z, q3.w, q4.x, q4.y, q4.z, q4.w};
			const uint32_t hm = cw_intersect_children(w, ray, tmin, tmax);
			ng = make_uint2(q1.x, (hm & 0xff000000u) | (q0.w >> 24));
			tg = make_uint2(q1.y, hm & 0x00ffffffu);
		}
		while (tg.y != 0u)
		{
			const uint32_t bit = 31u - uint32_t(__clz(tg.y));
			tg.y &= ~(1u << bit);
			const size_t ti = size_t(tg.x + bit) * 3;
			const float4 a = __ldg(tris + ti + 0), b = __ldg(tris + ti + 1), c = __ldg(tris + ti + 2);
			const V3 p0 = mk(a.x, a.y, a.z), e1 = mk(a.w, b.x, b.y), e2 = mk(b.z, b.w, c.x);
			const V3 h = cross(d, e2);
			const float det = dot(e1, h);
			const float eps = c.z;
			if (det > -eps && det < eps)
				continue;
			const float f = 1.0f / det;
			const V3 s = o - p0;
			const float u = f * dot(s, h);
			if (u < 0.0f || u > 1.0f)
				continue;
			const V3 q = cross(s, e1);
			const float vv = f * dot(d, q);
			if (vv < 0.0f || u + vv > 1.0f)
				continue;
			const float t = f * dot(e2, q);
			if (t > tmin && tmax > t)
			{
				if (ANY_HIT)
					return true;
				tmax = t, hit_u = u, hit_v = vv, hit_tri = __float_as_uint(c.y);
				found = true;
			}
		}
		if (ng.y <= 0x00ffffffu)
		{
			if (sp == 0)
				break;
			ng = stack[--sp];
		}
	}
	return found;
}

// One ray through a two-level scene (TlInstance, device_types.h): the top-level tree in world space; a top-level leaf names one
// instance, whose inverse transform moves the ray into object space (direction NOT re-normalised: t stays the world distance,
// Kernels.cu:229-232,270-272) for the walk of that mesh's tree; a marker on the stack brings the world ray back when the
// instance's subtree is exhausted.  One stack serves both levels.
template <bool ANY_HIT>
__device__ __noinline__ bool traverse_tl(const SceneView &sc, const V3 wo, const V3 wd, const float tmin, float &tmax, uint32_t &hit_tri,
										 float &hit_u, float &hit_v, uint32_t &hit_inst)
{
	constexpr int RETURN_TO_WORLD = 0x7ffffffe;
	constexpr uint32_t NO_INSTANCE = 0xffffffffu;
	const float tiny = 1e-30f;
	V3 o = wo, d = wd;
	float idx, idy, idz, oodx, oody, oodz;
#define RFW_RAY_SETUP()                                                                                                 \
	idx = 1.0f / (fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x));                                                      \
	idy = 1.0f / (fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y));                                                      \
	idz = 1.0f / (fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z));                                                      \
	oodx = o.x * idx, oody = o.y * idy, oodz = o.z * idz;
	RFW_RAY_SETUP()
	int stack[TRAVERSAL_STACK];
	int sp = 0;
	int cur = 0;
	uint32_t inst = NO_INSTANCE; // the instance whose tree is being walked
	bool found = false;
	const float4 *__restrict__ tris = reinterpret_cast<const float4 *>(sc.tris);
	if (sc.node_count == 0)
		return false;
	for (;;)
	{
		if (cur == RETURN_TO_WORLD)
		{
			o = wo, d = wd, inst = NO_INSTANCE;
			RFW_RAY_SETUP()
			if (sp == 0)
				break;
			cur = stack[--sp];
		}
		else if (cur >= 0)
		{
			const NodeRegs n = load_node(sc, nullptr, 0u, uint32_t(cur));
			uint32_t key0, key1, key2, key3;
			CHILD_T(0, x)
			CHILD_T(1, y)
			CHILD_T(2, z)
			CHILD_T(3, w)
			if (NO_CHILD_HIT(key0, key1, key2, key3))
			{
				if (sp == 0)
					break;
				cur = stack[--sp];
				continue;
			}
			CSWAP(key0, key1)
			CSWAP(key2, key3)
			CSWAP(key0, key2)
			CSWAP(key1, key3)
			CSWAP(key1, key2)
			cur = PICK(key0);
			if (key1 < 0x7f000000u) // sorted: misses are last
			{
				if (key3 < 0x7f000000u)
					stack[sp++] = PICK(key3);
				if (key2 < 0x7f000000u)
					stack[sp++] = PICK(key2);
				stack[sp++] = PICK(key1);
			}
		}
		else if (inst == NO_INSTANCE)
		{
			// top-level leaf: enter the instance
			inst = uint32_t(~cur) >> 2;
			const float4 *ip = reinterpret_cast<const float4 *>(sc.tl_instances + inst);
			const float4 r0 = __ldg(ip), r1 = __ldg(ip + 1), r2 = __ldg(ip + 2);
			o = mk(r0.x * wo.x + r0.y * wo.y + r0.z * wo.z + r0.w, r1.x * wo.x + r1.y * wo.y + r1.z * wo.z + r1.w,
				   r2.x * wo.x + r2.y * wo.y + r2.z * wo.z + r2.w);
			d = mk(r0.x * wd.x + r0.y * wd.y + r0.z * wd.z, r1.x * wd.x + r1.y * wd.y + r1.z * wd.z, r2.x * wd.x + r2.y * wd.y + r2.z * wd.z);
			RFW_RAY_SETUP()
			stack[sp++] = RETURN_TO_WORLD;
			cur = int(__ldg(&sc.tl_instances[inst].blas_root));
		}
		else
		{
			const uint32_t v = uint32_t(~cur), first = v >> 2, cnt = (v & 3u) + 1u;
			for (uint32_t i = 0; i < cnt; i++)
			{
				const float4 a = __ldg(tris + size_t(first + i) * 3 + 0);
				const float4 b = __ldg(tris + size_t(first + i) * 3 + 1);
				const float4 c = __ldg(tris + size_t(first + i) * 3 + 2);
				const V3 p0 = mk(a.x, a.y, a.z), e1 = mk(a.w, b.x, b.y), e2 = mk(b.z, b.w, c.x);
				const V3 h = cross(d, e2);
				const float det = dot(e1, h);
				const float eps = c.z;
				if (det > -eps && det < eps)
					continue;
				const float f = 1.0f / det;
				const V3 s = o - p0;
				const float u = f * dot(s, h);
				if (u < 0.0f || u > 1.0f)
					continue;
				const V3 q = cross(s, e1);
				const float vv = f * dot(d, q);
				if (vv < 0.0f || u + vv > 1.0f)
					continue;
				const float t = f * dot(e2, q);
				if (t > tmin && tmax > t)
				{
					if (ANY_HIT)
						return true;
					tmax = t, hit_u = u, hit_v = vv, hit_tri = __float_as_uint(c.y), hit_inst = inst;
					found = true;
				}
			}
			cur = stack[--sp]; // at least the marker is below
		}
	}
#undef RFW_RAY_SETUP
	return found;
}

// the normals of a mesh's shading record taken to world space with the instance's normal matrix, in the arithmetic of the
// flattening pass (context.cpp flatten_scene / geometry.cu k_flatten_shade), so both scene forms shade alike
__device__ __forceinline__ void apply_instance_normals(const SceneView &sc, uint32_t inst, ShadeTri &tri)
{
	const float *m = sc.tl_instances[inst].normal;
	const float m0 = __ldg(m), m1 = __ldg(m + 1), m2 = __ldg(m + 2), m3 = __ldg(m + 3), m4 = __ldg(m + 4), m5 = __ldg(m + 5), m6 = __ldg(m + 6),
				m7 = __ldg(m + 7), m8 = __ldg(m + 8);
#define RFW_NMUL(X, Y, Z)                                                                                               \
	{                                                                                                                   \
		const float x_ = m0 * X + m3 * Y + m6 * Z, y_ = m1 * X + m4 * Y + m7 * Z, z_ = m2 * X + m5 * Y + m8 * Z;           \
		X = x_, Y = y_, Z = z_;                                                                                         \
	}
	RFW_NMUL(tri.n0x, tri.n0y, tri.n0z)
	RFW_NMUL(tri.n1x, tri.n1y, tri.n1z)
	RFW_NMUL(tri.n2x, tri.n2y, tri.n2z)
	RFW_NMUL(tri.Nx, tri.Ny, tri.Nz)
#undef RFW_NMUL
	const float il = 1.0f / sqrtf(tri.Nx * tri.Nx + tri.Ny * tri.Ny + tri.Nz * tri.Nz);
	tri.Nx *= il, tri.Ny *= il, tri.Nz *= il;
	// the caller's instance index: row of the top-level instance, column of the triangle's mesh (context.cpp update_two_level)
	tri.inst_id = __ldg(sc.tl_inst_map + __ldg(&sc.tl_instances[inst].pad[0]) + tri.inst_id);
}

template <bool ANY_HIT>
__device__ __forceinline__ bool traverse(const SceneView &sc, const float4 *__restrict__ snodes, uint32_t n_smem, V3 o,
										 V3 d, float tmin, float &tmax, uint32_t &hit_tri, float &hit_u, float &hit_v, uint32_t &hit_inst)
{
	if (sc.tl_instances != nullptr)
		return traverse_tl<ANY_HIT>(sc, o, d, tmin, tmax, hit_tri, hit_u, hit_v, hit_inst);
	if (sc.cw_nodes != nullptr)
		return traverse_cw<ANY_HIT>(sc, o, d, tmin, tmax, hit_tri, hit_u, hit_v);
	const float tiny = 1e-30f;
	const float idx = 1.0f / (fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x));
	const float idy = 1.0f / (fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y));
	const float idz = 1.0f / (fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z));
	const float oodx = o.x * idx, oody = o.y * idy, oodz = o.z * idz;
	int stack[TRAVERSAL_STACK];
	int sp = 0;
	int cur = 0;
	bool found = false;
	const float4 *__restrict__ tris = reinterpret_cast<const float4 *>(sc.tris);
	for (;;)
	{
		if (cur >= 0)
		{
			const NodeRegs n = load_node(sc, snodes, n_smem, uint32_t(cur));
			uint32_t key0, key1, key2, key3;
			CHILD_T(0, x)
			CHILD_T(1, y)
			CHILD_T(2, z)
			CHILD_T(3, w)
			if (NO_CHILD_HIT(key0, key1, key2, key3))
			{
				if (sp == 0)
					break;
				cur = stack[--sp];
				continue;
			}
			CSWAP(key0, key1)
			CSWAP(key2, key3)
			CSWAP(key0, key2)
			CSWAP(key1, key3)
			CSWAP(key1, key2)
			cur = PICK(key0);
			if (key1 < 0x7f000000u) // sorted: misses are last
			{
				if (key3 < 0x7f000000u)
					stack[sp++] = PICK(key3);
				if (key2 < 0x7f000000u)
					stack[sp++] = PICK(key2);
				stack[sp++] = PICK(key1);
			}
		}
		else
		{
			const uint32_t v = uint32_t(~cur), first = v >> 2, cnt = (v & 3u) + 1u;
			for (uint32_t i = 0; i < cnt; i++)
			{
				const float4 a = __ldg(tris + size_t(first + i) * 3 + 0);
				const float4 b = __ldg(tris + size_t(first + i) * 3 + 1);
				const float4 c = __ldg(tris + size_t(first + i) * 3 + 2);
				const V3 p0 = mk(a.x, a.y, a.z), e1 = mk(a.w, b.x, b.y), e2 = mk(b.z, b.w, c.x);
				const V3 h = cross(d, e2);
				const float det = dot(e1, h);
				const float eps = c.z;
				if (det > -eps && det < eps)
					continue;
				const float f = 1.0f / det;
				const V3 s = o - p0;
				const float u = f * dot(s, h);
				if (u < 0.0f || u > 1.0f)
					continue;
				const V3 q = cross(s, e1);
				const float vv = f * dot(d, q);
				if (vv < 0.0f || u + vv > 1.0f)
					continue;
				const float t = f * dot(e2, q);
				if (t > tmin && tmax > t)
				{
					if (ANY_HIT)
						return true;
					tmax = t, hit_u = u, hit_v = vv, hit_tri = __float_as_uint(c.y);
					found = true;
				}
			}
			if (sp == 0)
				break;
			cur = stack[--sp];
		}
	}
	return found;
}
template <bool ANY_HIT>
__device__ __forceinline__ bool traverse(const SceneView &sc, const float4 *__restrict__ snodes, uint32_t n_smem, V3 o,
										 V3 d, float tmin, float &tmax, uint32_t &hit_tri, float &hit_u, float &hit_v)
{
	uint32_t inst = 0;
	return traverse<ANY_HIT>(sc, snodes, n_smem, o, d, tmin, tmax, hit_tri, hit_u, hit_v, inst);
}

// ------------------------------------------------------------------------------------------------
// generate — Kernels.cu:383-426 (PT: blue noise dims 0-3), EmbreeRT/src/Ray.cpp:16-47 (E-mode: xor128)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void camera_ray(const FrameParams &fp, const ShardView &sh, uint32_t x, uint32_t y, float r0,
										   float r1, float r2, float r3, bool cuda_sincos_order, V3 &O, V3 &D)
{
	const float blade = float(int(r0 * 9));
	r2 = (r2 - blade * (1.0f / 9.0f)) * 9.0f;
	constexpr float piOver4point5 = 3.14159265359f / 4.5f;
	float s1, c1, s2, c2;
	sincosf(blade * piOver4point5, &s1, &c1);
	sincosf((blade + 1.0f) * piOver4point5, &s2, &c2);
	// CUDART: __sincosf(a,&x1,&y1) => x = sin, y = cos (Kernels.cu:406-407); EmbreeRT: x = cos, y = sin
	const float x1 = cuda_sincos_order ? s1 : c1, y1 = cuda_sincos_order ? c1 : s1;
	const float x2 = cuda_sincos_order ? s2 : c2, y2 = cuda_sincos_order ? c2 : s2;
	if ((r2 + r3) > 1.0f)
		r2 = 1.0f - r2, r3 = 1.0f - r3;
	const float xr = x1 * r2 + x2 * r3;
	const float yr = y1 * r2 + y2 * r3;
	const V3 pos = ld3(fp.pos), p1 = ld3(fp.p1), right = ld3(fp.right), up = ld3(fp.up);
	O = pos + (right * xr + up * yr) * fp.aperture;
	const float u = (float(x) + r0) * (1.0f / float(sh.width));
	const float v = (float(y) + r1) * (1.0f / float(sh.height));
	D = normalize(p1 + right * u + up * v - O);
}

__device__ __forceinline__ void generate_pt(const SceneView &sc, const FrameParams &fp, const ShardView &sh, uint32_t x,
											uint32_t y, uint32_t sampleIndex, V3 &O, V3 &D)
{
	const float r0 = blueNoiseSampler(sc.blue_noise, int(x), int(y), int(sampleIndex), 0);
	const float r1 = blueNoiseSampler(sc.blue_noise, int(x), int(y), int(sampleIndex), 1);
	const float r2 = blueNoiseSampler(sc.blue_noise, int(x), int(y), int(sampleIndex), 2);
	const float r3 = blueNoiseSampler(sc.blue_noise, int(x), int(y), int(sampleIndex), 3);
	camera_ray(fp, sh, x, y, r0, r1, r2, r3, true, O, D);
}

__device__ __forceinline__ void generate_emode(const FrameParams &fp, const ShardView &sh, uint32_t x, uint32_t y,
											   uint32_t sampleIndex, V3 &O, V3 &D)
{
	// determinism contract: xor128 with x seeded per (pixel, sample)
	const uint32_t pixel = y * sh.width + x;
	uint32_t sx = 123456789u ^ WangHash(pixel * 16789u + sampleIndex * 1791u), sy = 362436069u, sz = 521288629u,
			 sw = 88675123u;
	float r[4];
#pragma unroll
	for (int i = 0; i < 4; i++)
	{
		const uint32_t t = sx ^ (sx << 11);
		sx = sy, sy = sz, sz = sw;
		sw = sw ^ (sw >> 19) ^ (t ^ (t >> 8));
		r[i] = float(sw) * 2.3283064365387e-10f;
	}
	camera_ray(fp, sh, x, y, r[0], r[1], r[2], r[3], false, O, D);
}

#if RFW_PART != 2 // ---- trace part: compiled with IEEE arithmetic ----
// ------------------------------------------------------------------------------------------------
// k_wavefront_trace — the persistent-threads traversal kernel of both trace stages.
//
//   PRIMARY = true : work item = local pixel; generate the camera ray (Kernels.cu:383-426), closest hit
//   PRIMARY = false: work items [0, n_ext) = extension rays (closest hit, Kernels.cu:461-483),
//                    [n_ext, n_ext + n_shadow) = connect rays (any hit + accumulate, Kernels.cu:484-498)
//
// Every lane owns one ray at a time.  "While-while" traversal: all lanes descend inner nodes until
// each holds a leaf, then the leaves are intersected together; a lane whose ray is finished stays idle
// only until at least rs.fetch_threshold lanes of its warp are idle, then the idle lanes pull new work
// items from the device cursor with ONE atomic for the warp.  This keeps SIMD lanes busy when rays of
// very different cost (incoherent bounces, early-out shadow rays) share a warp — the first version
// with one ray per lane per 32-ray chunk ran at 9-12 of 32 active threads per instruction
// (profiles/r01).
// ------------------------------------------------------------------------------------------------

extern __shared__ __align__(128) unsigned char g_dyn_smem[];

#ifndef TRACE_MINB
#define TRACE_MINB 4 // resident CTAs per SM the register allocation is sized for (measured)
#endif
// Template parameters: LQ = leaves a lane may hold back before it has to wait for the warp's leaf phase (1 = the
// classic single postponed leaf); LEAN = node step without the four-key sorting network: the nearest hit child is
// entered, the others are pushed in slot order (about 45 fewer instructions per node visit for a slightly less
// ordered traversal; LEAN = 2 applies it to connect rays only, which stop at any hit and gain nothing from the order);
// PACKED = read the 80-byte BvhNode4Packed form of a node (five 128-bit loads per visit instead of seven, bfloat16
// planes decoded with one shift / mask each); the staged shared-memory prefix (setting smem_nodes) exists only in
// the <LQ 1, LEAN 0, !PACKED> variant.
// What one trace launch reads: queue sizes (device-resident, written by the shade launch in front of it) and planes.
struct TraceQueue
{
	uint32_t n_ext, n_shadow, total;
	const float4 *Oin, *Din;
	const uint32_t *seen; // ext_seen row of the depth that emitted these rays
};
template <bool PRIMARY>
__device__ __forceinline__ TraceQueue open_queue(const ShardView &sh, const WavefrontView &wf, const BatchView &bv, uint32_t depth,
												 uint32_t in_buf)
{
	TraceQueue q;
	q.Oin = wf.O[in_buf], q.Din = wf.D[in_buf];
	q.seen = nullptr;
	if (PRIMARY)
		q.n_ext = bv.items, q.n_shadow = 0;
	else
	{
		const DepthCounters *prev = &wf.counters[bv.index * MAX_DEPTH_SLOTS + depth - 1];
		q.n_ext = prev->ext;
		q.n_shadow = prev->shadow;
		q.seen = wf.ext_seen + size_t(bv.index * MAX_DEPTH_SLOTS + depth - 1) * MAX_BATCH_SPP;
	}
	q.total = q.n_ext + q.n_shadow;
	return q;
}

// Moller-Trumbore against one triangle record, all conditions folded into one predicate (used for the cached candidates)
__device__ __forceinline__ bool hits_record(const float4 *__restrict__ tris, uint32_t rec, V3 o, V3 d, float tmin, float tmax, float &t)
{
	const float4 a = __ldg(tris + size_t(rec) * 3 + 0);
	const float4 b = __ldg(tris + size_t(rec) * 3 + 1);
	const float4 c = __ldg(tris + size_t(rec) * 3 + 2);
	const V3 p0 = mk(a.x, a.y, a.z), e1 = mk(a.w, b.x, b.y), e2 = mk(b.z, b.w, c.x);
	const V3 h = cross(d, e2);
	const float det = dot(e1, h);
	const float f = 1.0f / det;
	const V3 s = o - p0;
	const float u = f * dot(s, h);
	const V3 q = cross(s, e1);
	const float vv = f * dot(d, q);
	t = f * dot(e2, q);
	return !(det > -c.z && det < c.z) && u >= 0.0f && u <= 1.0f && vv >= 0.0f && u + vv <= 1.0f && t > tmin && tmax > t;
}

// One work item of a trace launch becomes a ray.  Returns false when the item needs no traversal (padded pixel, connect
// ray the reference never traces, connect ray stopped by a remembered occluder).
//   camera ray  : item = ((block * spp + s) << 5) | lane; generates the ray (Kernels.cu:383-426), writes O/D
//   extension   : item < n_ext, reads O/D
//   connect     : item - n_ext indexes the connect queue; `pidx` is the path index its contribution goes to
template <bool PRIMARY>
__device__ __forceinline__ bool fetch_ray(const SceneView &sc, const ShardView &sh, const WavefrontView &wf, const RenderSettings &rs,
										  const BatchView &bv, const FrameParams &fp, const TraceQueue &q, uint32_t depth, uint32_t in_buf,
										  uint32_t item, V3 &o, V3 &d, float &tmin, float &tmax, bool &shadow, uint32_t &pidx,
										  uint32_t &occluder, uint32_t &occ_slot, uint32_t &n_traced)
{
	const float4 *__restrict__ tris = reinterpret_cast<const float4 *>(sc.tris);
	if (PRIMARY)
	{
		uint32_t j, s, x, y;
		item_to_pixel_sample(bv, item, j, s);
		pidx = j;
		if (!local_to_pixel(sh, j, x, y))
		{
			ST_TS(&wf.hit[item], make_float4(0.f, 0.f, __int_as_float(PRIM_DEAD), 0.f));
			return false;
		}
		generate_pt(sc, fp, sh, x, y, fp.sample_base + bv.first_sample + s, o, d);
		ST_TS(&wf.O[in_buf][item], make_float4(o.x, o.y, o.z, __uint_as_float((item << 8) + 1u)));
		ST_TS(&wf.D[in_buf][item], make_float4(d.x, d.y, d.z, 0.0f));
		tmin = 1e-5f, tmax = 1e34f, shadow = false;
		// Bound from an earlier sample of the pixel: the samples of a pixel differ by a sub-pixel jitter, so the triangle
		// one of them hit is almost always hit by the next.  Its distance (plus a few ulps) only LIMITS the search — the
		// triangle is found again by the traversal itself, in the usual order — so the result is bit-identical with and
		// without the cache, whatever stale value a concurrent lane left here.
		const uint32_t cand = rs.primary_cache ? wf.prim_cache[j] : 0xffffffffu;
		float t;
		if (cand < sc.tri_count && hits_record(tris, cand, o, d, tmin, 1e33f, t))
			tmax = t * 1.000004f + 1e-30f;
		return true;
	}
	if (item < q.n_ext)
	{
		const float4 O4 = LD_TS(&q.Oin[item]), D4 = LD_TS(&q.Din[item]);
		o = mk(O4.x, O4.y, O4.z), d = mk(D4.x, D4.y, D4.z);
		tmin = 1e-5f, tmax = 1e34f, shadow = false;
		return true;
	}
	const uint32_t k = item - q.n_ext;
	pidx = __float_as_uint(LD_TS(&wf.sE[k].w));
	// The reference only traces the connect queue of a sample when its bounce loop continues, i.e. when that sample
	// emitted at least one extension ray at this depth anywhere in the frame (CUDART/src/Context.cpp:109-120: the host
	// loop leaves before the pending shadow rays are traced).  The ranks of a sharded frame merge their flags between the
	// shade launch and this one (k_shard_sync), so the assembled frame does not depend on how it was sharded.
	{
		uint32_t pj, ps;
		item_to_pixel_sample(bv, pidx, pj, ps);
		if (q.seen[ps] == 0u)
			return false;
	}
	n_traced++;
	const float4 O4 = LD_TS(&wf.sO[k]), D4 = LD_TS(&wf.sD[k]);
	o = mk(O4.x, O4.y, O4.z), d = mk(D4.x, D4.y, D4.z);
	tmin = rs.geometry_epsilon, tmax = D4.w, shadow = true;
	if (rs.shadow_cache == 2)
	{
		uint32_t j, s;
		item_to_pixel_sample(bv, pidx, j, s);
		occ_slot = ((depth - 1u) & 1u) * sh.local_pixels + j;
		occluder = wf.occ_cache[occ_slot];
	}
	float t;
	if (rs.shadow_cache && occluder < sc.tri_count && hits_record(tris, occluder, o, d, tmin, tmax, t))
		return false; // occluded: nothing to accumulate, the lane stays idle and takes the next item
	return true;
}

// end of a ray: hit record of a closest-hit ray, or accumulator[path] += (contribution, 1) of an unoccluded connect ray
// (Kernels.cu:457-459, 495-497).  Every path of a wavefront owns its accumulator slot (one per pixel and sample), so the
// read-modify-write needs no atomic, exactly like the reference's one-sample-per-launch `+=`.
// (hit.z of a hit: 0 in a flattened scene, the instance index in a two-level scene — any value >= 0 means "hit")
__device__ __forceinline__ void retire_ray(const WavefrontView &wf, const TraceQueue &q, uint32_t item, bool shadow, uint32_t hit_tri,
										   float hit_u, float hit_v, float tmax, uint32_t pidx, uint32_t &acc_count, uint32_t hit_inst = 0u)
{
	if (!shadow)
	{
		float4 hit = make_float4(0.f, 0.f, __int_as_float(PRIM_MISS), 0.f);
		if (hit_tri != 0xffffffffu)
			hit = make_float4(__uint_as_float(pack_barycentrics(hit_u, hit_v)), __uint_as_float(hit_tri), __int_as_float(int(hit_inst)), tmax);
		ST_TS(&wf.hit[item], hit);
	}
	else if (hit_tri == 0xffffffffu)
	{
		const float4 E = LD_TS(&wf.sE[item - q.n_ext]);
		float4 a = LD_TS(&wf.sample_acc[pidx]);
		a.x += E.x, a.y += E.y, a.z += E.z, a.w += 1.0f;
		ST_TS(&wf.sample_acc[pidx], a);
		acc_count++;
	}
}

// per-warp {start, end, rays, smid} record of the debug timeline + the launch's bookkeeping counters
__device__ __forceinline__ void close_launch(const RenderSettings &rs, DepthCounters *curc, bool primary, bool dbg, unsigned long long dbg_t0,
											 unsigned long long dbg_rays, uint32_t acc_count, uint32_t n_traced)
{
	const uint32_t lane = threadIdx.x & 31u;
	if (dbg)
	{
		unsigned long long t1;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
		const unsigned long long n = __reduce_add_sync(0xffffffffu, uint32_t(dbg_rays));
		if (lane == 0)
		{
			unsigned long long *rec = rs.debug + size_t(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 4;
			uint32_t smid;
			asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
			rec[0] = dbg_t0, rec[1] = t1, rec[2] = n, rec[3] = smid;
		}
	}
	if (!primary)
	{
		// bookkeeping for the algorithmic-bytes formula: one atomic per warp
		acc_count = __reduce_add_sync(0xffffffffu, acc_count);
		n_traced = __reduce_add_sync(0xffffffffu, n_traced);
		if (lane == 0 && acc_count)
			atomicAdd(&curc->acc, acc_count);
		if (lane == 0 && n_traced)
			atomicAdd(&curc->shadow_traced, n_traced);
	}
}

// STAGE_P: the first rs.smem_nodes PACKED nodes (80 B each: 256 nodes = 20 KB) are staged into shared memory with one TMA bulk copy
// per CTA, and node fetches below that index read shared memory (setting smem_nodes with the packed variants; measured in
// profiles/r02 against the L1-only default).
// TL: the scene is two-level (TlInstance, device_types.h).  A top-level leaf names one instance; reaching it means moving the
// ray into the instance's object space, pushing a marker that brings the world ray back, and descending the tree below —
// all in the node phase of the same while-while loop, on the same stack.  A lane that holds a postponed leaf of an instance
// stops at that instance's marker (the leaf's triangles must be tested with the object-space ray), exactly as it stops at
// a second leaf.
template <bool PRIMARY, int LQ, int LEAN, bool PACKED, bool STAGE_P = false, bool TL = false>
__global__ void __launch_bounds__(256, TRACE_MINB) k_wavefront_trace(const SceneView sc, const ShardView sh, const WavefrontView wf,
															const RenderSettings rs, const BatchView bv, const uint32_t depth,
															const uint32_t in_buf)
{
	constexpr bool STAGED = (LQ == 1 && LEAN == 0 && !PACKED);
	__shared__ uint64_t mbar;
	float4 *snodes = reinterpret_cast<float4 *>(g_dyn_smem);
	const uint4 *snodes16 = reinterpret_cast<const uint4 *>(g_dyn_smem);
	const uint32_t n_smem = (STAGED || STAGE_P) ? min(uint32_t(rs.smem_nodes), sc.node_count) : 0u;
	if (STAGED)
		stage_nodes(snodes, sc.nodes, n_smem, &mbar);
	if (STAGE_P)
		stage_bytes(g_dyn_smem, sc.nodes16, n_smem * 80u, &mbar);

	DepthCounters *curc = &wf.counters[bv.index * MAX_DEPTH_SLOTS + depth];
	const TraceQueue q = open_queue<PRIMARY>(sh, wf, bv, depth, in_buf);
	const uint32_t total = q.total;
	const float4 *__restrict__ tris = reinterpret_cast<const float4 *>(sc.tris);
	uint32_t *cursor = &curc->trace_cursor;
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t lt_mask = (1u << lane) - 1u;
	FrameParams fp;
	if (PRIMARY)
		fp = *wf.frame;

	unsigned long long dbg_t0 = 0, dbg_rays = 0;
	const bool dbg = !PRIMARY && rs.debug != nullptr && int(depth) == rs.debug_depth;
	if (dbg)
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));

	// per-lane ray state
	bool alive = false;
	bool shadow = false;
	uint32_t item = 0;
	V3 o = mk(0.f), d = mk(0.f);
	float idx = 0.f, idy = 0.f, idz = 0.f, oodx = 0.f, oody = 0.f, oodz = 0.f;
	uint32_t sgx = 0, sgy = 0, sgz = 0; // 1 when the ray runs towards -axis: the max plane is the near plane
	float tmin = 0.f, tmax = 0.f, hit_u = 0.f, hit_v = 0.f;
	uint32_t hit_tri = 0xffffffffu;
	int stack[TRAVERSAL_STACK];
	constexpr int SENTINEL = 0x7fffffff; // bottom-of-stack marker: popping it ends the ray
	constexpr int NO_LEAF = 0;			 // leaf references are negative, so 0 can mean "none postponed"
	int sp = 0, cur = SENTINEL, leaf = NO_LEAF;
	int leaf1 = NO_LEAF, leaf2 = NO_LEAF; // further held-back leaves (LQ > 1)
	// Occluder cache (setting shadow_cache): the triangle record that stopped this lane's previous connect ray is tested
	// first; lanes take neighbouring queue entries (neighbouring pixels aiming at the same light), so a wall or a roof that
	// occluded one ray usually occludes the next, which then never enters the tree.  Occlusion is a yes/no answer, so the
	// result does not depend on which occluder is found.
	uint32_t occluder = 0xffffffffu;
	uint32_t hit_pos = 0xffffffffu; // record position of the closest hit (camera rays: remembered per pixel)
	uint32_t occ_slot = 0;			// where this connect ray's occluder is remembered (shadow_cache = 2)
	uint32_t pidx = 0;				// camera ray: local pixel; connect ray: path index its contribution belongs to
	bool exhausted = false; // warp-uniform: the queue has no more items
	uint32_t acc_count = 0, n_traced = 0;
	constexpr int TL_MARKER = 0x7ffffffe;		// stack entry: back to the world ray (below SENTINEL, so the node loop sees it)
	constexpr uint32_t NO_INSTANCE = 0xffffffffu;
	uint32_t inst = NO_INSTANCE, hit_inst = 0; // TL: instance whose tree the lane is in; instance of the closest hit
	float world_ray[TL ? 6 : 1];			   // TL: the world ray while the lane is inside an instance (indexed dynamically nowhere: registers
											   // would do, but the kernel has none to spare — the compiler keeps it in local memory, L1)
	uint32_t chunk_pos = 0, chunk_end = 0; // warp-uniform: the run of queue entries this warp is working through
	const uint32_t chunk_len = uint32_t(rs.fetch_chunk);

	for (;;)
	{
		// ---- refill idle lanes ---------------------------------------------------------------------------
		const uint32_t idle_mask = __ballot_sync(0xffffffffu, !alive);
		if (idle_mask == 0xffffffffu && exhausted)
			break;
		const int n_idle = __popc(idle_mask);
		if (!exhausted && (n_idle >= rs.fetch_threshold))
		{
			// Default (fetch_chunk = 0): the idle lanes take the next n_idle queue entries with one atomic, so all warps of the
			// GPU work on one narrow front of the queue and share its nodes in L1 / L2.  fetch_chunk > 0 gives each warp a
			// private run of that many consecutive entries instead (its lanes then hold neighbours of the queue, but the
			// warps of an SM are spread over 4,736 runs): measured slower at every run length, profiles/r02.
			if (chunk_pos >= chunk_end)
			{
				const uint32_t want = chunk_len ? chunk_len : uint32_t(n_idle);
				uint32_t base = 0;
				if (lane == 0)
					base = atomicAdd(cursor, want);
				base = __shfl_sync(0xffffffffu, base, 0);
				chunk_pos = min(base, total), chunk_end = min(base + want, total);
				if (base + want >= total)
					exhausted = chunk_len == 0u || base >= total;
			}
			const uint32_t take = min(uint32_t(n_idle), chunk_end - chunk_pos);
			const uint32_t mine = __popc(idle_mask & lt_mask);
			const uint32_t first = chunk_pos;
			chunk_pos += take;
			if (!alive)
			{
				item = first + mine;
				if (mine < take &&
					fetch_ray<PRIMARY>(sc, sh, wf, rs, bv, fp, q, depth, in_buf, item, o, d, tmin, tmax, shadow, pidx, occluder, occ_slot, n_traced))
				{
					dbg_rays++;
#define RFW_LANE_RAY_SETUP()                                                                                            \
	{                                                                                                                   \
		const float tiny = 1e-30f;                                                                                      \
		idx = 1.0f / (fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x));                                                  \
		idy = 1.0f / (fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y));                                                  \
		idz = 1.0f / (fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z));                                                  \
		oodx = o.x * idx, oody = o.y * idy, oodz = o.z * idz;                                                           \
		sgx = idx < 0.0f ? 1u : 0u, sgy = idy < 0.0f ? 1u : 0u, sgz = idz < 0.0f ? 1u : 0u;                             \
	}
					RFW_LANE_RAY_SETUP()
					stack[0] = SENTINEL;
					inst = NO_INSTANCE, hit_inst = 0;
					sp = 1, cur = 0, leaf = leaf1 = leaf2 = NO_LEAF, hit_tri = 0xffffffffu, hit_u = 0.f, hit_v = 0.f;
					alive = true;
				}
			}
		}

		// ---- traverse (Aila/Laine "while-while" with a postponed leaf, adapted to 4-wide nodes) ------------------------
		bool finished = false;
		if (alive)
		{
			while (cur != SENTINEL || leaf != NO_LEAF)
			{
				// descend inner nodes; the first leaf a lane meets is postponed and the lane keeps descending
				// (speculatively) so it stays useful until every lane of the warp holds a leaf
				while (uint32_t(cur) < uint32_t(SENTINEL) && !(TL && cur == TL_MARKER && leaf != NO_LEAF))
				{
					if (TL && cur == TL_MARKER)
					{
						// the instance's subtree is exhausted: back to the world ray — if the top level has anything left for it
						// (the ray was parked in thread-local memory when it entered the instance: an L1 hit, not a trip to the queue)
						inst = NO_INSTANCE;
						cur = stack[--sp];
						if (cur != SENTINEL)
						{
							o = mk(world_ray[0], world_ray[1], world_ray[2]), d = mk(world_ray[3], world_ray[4], world_ray[5]);
							RFW_LANE_RAY_SETUP()
						}
					}
					else
					{
					struct
					{
						int4 child;
					} n;
					uint32_t key0, key1, key2, key3;
					if (PACKED)
					{
						// five 16-byte pieces: min corner, three axes of bfloat16 planes (lo0 lo1 | lo2 lo3 | hi0 hi1 | hi2 hi3), children
						const uint4 *np_ = (STAGE_P && uint32_t(cur) < n_smem) ? snodes16 + size_t(cur) * 5 : sc.nodes16 + size_t(cur) * 5;
						const uint4 hp = np_[0], px = np_[1], py = np_[2], pz = np_[3];
						n.child = *reinterpret_cast<const int4 *>(np_ + 4);
						// t = plane * idir - o * idir with plane = p + v:  v * idir + (p * idir - o * idir)
						const float cx = fmaf(__uint_as_float(hp.x), idx, -oodx), cy = fmaf(__uint_as_float(hp.y), idy, -oody),
									cz = fmaf(__uint_as_float(hp.z), idz, -oodz);
						const uint32_t nx01 = sgx ? px.z : px.x, nx23 = sgx ? px.w : px.y, fx01 = sgx ? px.x : px.z, fx23 = sgx ? px.y : px.w;
						const uint32_t ny01 = sgy ? py.z : py.x, ny23 = sgy ? py.w : py.y, fy01 = sgy ? py.x : py.z, fy23 = sgy ? py.y : py.w;
						const uint32_t nz01 = sgz ? pz.z : pz.x, nz23 = sgz ? pz.w : pz.y, fz01 = sgz ? pz.x : pz.z, fz23 = sgz ? pz.y : pz.w;
						// The plane in the upper half of a word is used as it stands: the lower half (its neighbour's bits) only
						// extends the mantissa, i.e. makes the offset larger by less than one bfloat16 step, and k_pack_nodes
						// (geometry.cu) rounds the planes stored there with that step to spare.  The interval test is folded into
						// the min / max chain: [max(tn, tmin), min(tf, tmax)] non-empty — a superset of the fp32 kernel's test
						// (it admits tn == tmax), so no node is skipped that the exact test enters; unused slots are inverted
						// boxes of +-3e38 here (not NaN boxes: fmaxf(NaN, tmin) would admit them).
#define BF_LO(W) __uint_as_float((W) << 16)
#define BF_HI(W) __uint_as_float(W)
#define CHILD_P(K, NX, FX, NY, FY, NZ, FZ)                                                                              \
	{                                                                                                                   \
		const float tn = fmaxf(fmaxf(fmaxf(fmaf(NX, idx, cx), fmaf(NY, idy, cy)), fmaf(NZ, idz, cz)), tmin);            \
		const float tf = fminf(fminf(fminf(fmaf(FX, idx, cx), fmaf(FY, idy, cy)), fmaf(FZ, idz, cz)), tmax);            \
		key##K = tn <= tf ? ((__float_as_uint(tn) & 0xFFFFFFFCu) | uint32_t(K)) : MISSKEY(K);                           \
	}
						CHILD_P(0, BF_LO(nx01), BF_LO(fx01), BF_LO(ny01), BF_LO(fy01), BF_LO(nz01), BF_LO(fz01))
						CHILD_P(1, BF_HI(nx01), BF_HI(fx01), BF_HI(ny01), BF_HI(fy01), BF_HI(nz01), BF_HI(fz01))
						CHILD_P(2, BF_LO(nx23), BF_LO(fx23), BF_LO(ny23), BF_LO(fy23), BF_LO(nz23), BF_LO(fz23))
						CHILD_P(3, BF_HI(nx23), BF_HI(fx23), BF_HI(ny23), BF_HI(fy23), BF_HI(nz23), BF_HI(fz23))
#undef CHILD_P
#undef BF_LO
#undef BF_HI
					}
					else
					{
						// one node = one 128-byte line: six plane quads picked by the ray's octant + the child words
						const float4 *np_ = (STAGED && uint32_t(cur) < n_smem) ? snodes + size_t(cur) * 8
																			   : reinterpret_cast<const float4 *>(sc.nodes) + size_t(cur) * 8;
						const float4 nearx = np_[sgx], farx = np_[sgx ^ 1u];
						const float4 neary = np_[2u + sgy], fary = np_[3u - sgy];
						const float4 nearz = np_[4u + sgz], farz = np_[5u - sgz];
						n.child = *reinterpret_cast<const int4 *>(np_ + 6);
						CHILD_N(0, x)
						CHILD_N(1, y)
						CHILD_N(2, z)
						CHILD_N(3, w)
					}
					if (LEAN == 1 || ((LEAN == 2 || LEAN == 3) && shadow))
					{
						const uint32_t km = min(min(key0, key1), min(key2, key3));
						if (km >= 0x7f000000u)
							cur = stack[--sp];
						else
						{
							cur = PICK(km);
							// the other hit children in slot order (predicated stores, no sorting network)
							if (key0 < 0x7f000000u && key0 != km)
								stack[sp++] = n.child.x;
							if (key1 < 0x7f000000u && key1 != km)
								stack[sp++] = n.child.y;
							if (key2 < 0x7f000000u && key2 != km)
								stack[sp++] = n.child.z;
							if (key3 < 0x7f000000u && key3 != km)
								stack[sp++] = n.child.w;
						}
					}
					else if (LEAN >= 3)
					{
						// the same order as the key-only network below, but the child words travel with their keys (no
						// slot -> child selects afterwards) and the pushes are predicated stores instead of nested branches
						int c0 = n.child.x, c1 = n.child.y, c2 = n.child.z, c3 = n.child.w;
#define CSWAP2(KA, KB, CA, CB)                                                                                          \
	{                                                                                                                   \
		const bool sw_ = KA > KB;                                                                                       \
		const uint32_t ka_ = sw_ ? KB : KA, kb_ = sw_ ? KA : KB;                                                        \
		const int ca_ = sw_ ? CB : CA, cb_ = sw_ ? CA : CB;                                                             \
		KA = ka_, KB = kb_, CA = ca_, CB = cb_;                                                                         \
	}
						CSWAP2(key0, key1, c0, c1)
						CSWAP2(key2, key3, c2, c3)
						CSWAP2(key0, key2, c0, c2)
						CSWAP2(key1, key3, c1, c3)
						CSWAP2(key1, key2, c1, c2)
#undef CSWAP2
						const int nh = (key0 < 0x7f000000u ? 1 : 0) + (key1 < 0x7f000000u ? 1 : 0) + (key2 < 0x7f000000u ? 1 : 0) +
									   (key3 < 0x7f000000u ? 1 : 0);
						if (nh > 3)
							stack[sp++] = c3;
						if (nh > 2)
							stack[sp++] = c2;
						if (nh > 1)
							stack[sp++] = c1;
						cur = nh > 0 ? c0 : stack[sp - 1];
						sp -= nh > 0 ? 0 : 1;
					}
					else if (NO_CHILD_HIT(key0, key1, key2, key3))
						cur = stack[--sp];
					else
					{
						CSWAP(key0, key1)
						CSWAP(key2, key3)
						CSWAP(key0, key2)
						CSWAP(key1, key3)
						CSWAP(key1, key2)
						cur = PICK(key0);
						if (key1 < 0x7f000000u) // sorted: misses are last
						{
							if (key3 < 0x7f000000u)
								stack[sp++] = PICK(key3);
							if (key2 < 0x7f000000u)
								stack[sp++] = PICK(key2);
							stack[sp++] = PICK(key1);
						}
					}
					} // node step
					if (TL && cur < 0 && inst == NO_INSTANCE)
					{
						// top-level leaf: enter the instance right here, in the node phase (Kernels.cu:229-232,270-272: the direction
						// is not re-normalised, so t stays the world distance).  A lane in world space never holds a postponed leaf
						// (it cannot pop an instance's marker while it holds one), so nothing has to wait for the warp's leaf phase —
						// a freshly fetched ray would otherwise idle from its first node visit to the next leaf phase of its warp.
						inst = uint32_t(~cur) >> 2;
						const float4 *ip = reinterpret_cast<const float4 *>(sc.tl_instances + inst);
						const float4 r0 = __ldg(ip), r1 = __ldg(ip + 1), r2 = __ldg(ip + 2);
						const V3 wo = o, wd = d;
						world_ray[0] = wo.x, world_ray[1] = wo.y, world_ray[2] = wo.z, world_ray[3] = wd.x, world_ray[4] = wd.y, world_ray[5] = wd.z;
						o = mk(r0.x * wo.x + r0.y * wo.y + r0.z * wo.z + r0.w, r1.x * wo.x + r1.y * wo.y + r1.z * wo.z + r1.w,
							   r2.x * wo.x + r2.y * wo.y + r2.z * wo.z + r2.w);
						d = mk(r0.x * wd.x + r0.y * wd.y + r0.z * wd.z, r1.x * wd.x + r1.y * wd.y + r1.z * wd.z,
							   r2.x * wd.x + r2.y * wd.y + r2.z * wd.z);
						RFW_LANE_RAY_SETUP()
						stack[sp++] = TL_MARKER;
						cur = int(__ldg(&sc.tl_instances[inst].blas_root));
					}
					// hold back up to LQ leaves and keep descending
					if (cur < 0 && leaf == NO_LEAF)
					{
						leaf = cur;
						cur = stack[--sp];
					}
					if (LQ > 1 && cur < 0 && leaf1 == NO_LEAF)
					{
						leaf1 = cur;
						cur = stack[--sp];
					}
					if (LQ > 2 && cur < 0 && leaf2 == NO_LEAF)
					{
						leaf2 = cur;
						cur = stack[--sp];
					}
					if (!__any_sync(__activemask(), leaf == NO_LEAF))
						break;
				}
				// intersect the postponed leaf (and any leaf that directly follows it on the stack)
				while (leaf != NO_LEAF)
				{
					const uint32_t v = uint32_t(~leaf), first = v >> 2, cnt = (v & 3u) + 1u;
					for (uint32_t i = 0; i < cnt; i++)
					{
						const float4 a = __ldg(tris + size_t(first + i) * 3 + 0);
						const float4 b = __ldg(tris + size_t(first + i) * 3 + 1);
						const float4 c = __ldg(tris + size_t(first + i) * 3 + 2);
						const V3 p0 = mk(a.x, a.y, a.z), e1 = mk(a.w, b.x, b.y), e2 = mk(b.z, b.w, c.x);
						const V3 h = cross(d, e2);
						const float det = dot(e1, h);
						const float eps = c.z;
						if (det > -eps && det < eps)
							continue;
						const float f = 1.0f / det;
						const V3 s = o - p0;
						const float u = f * dot(s, h);
						if (u < 0.0f || u > 1.0f)
							continue;
						const V3 q = cross(s, e1);
						const float vv = f * dot(d, q);
						if (vv < 0.0f || u + vv > 1.0f)
							continue;
						const float t = f * dot(e2, q);
						if (t > tmin && tmax > t)
						{
							tmax = t, hit_u = u, hit_v = vv, hit_tri = __float_as_uint(c.y);
							if (TL)
								hit_inst = inst;
							if (PRIMARY)
								hit_pos = first + i;
							if (shadow)
							{
								cur = SENTINEL; // any hit ends a connect ray
								leaf1 = leaf2 = NO_LEAF;
								occluder = first + i;
								if (!PRIMARY && rs.shadow_cache == 2)
									wf.occ_cache[occ_slot] = occluder;
								break;
							}
						}
					}
					leaf = NO_LEAF;
					if (LQ > 1)
						leaf = leaf1, leaf1 = leaf2, leaf2 = NO_LEAF;
					if (leaf == NO_LEAF && cur < 0)
					{
						leaf = cur;
						cur = stack[--sp];
					}
				}
				// too few lanes left in this loop: leave it so the idle lanes can fetch new rays
				if (__popc(__activemask()) <= 32 - rs.fetch_threshold)
					break;
			}
			finished = (cur == SENTINEL && leaf == NO_LEAF);
		}
		// ---- retire --------------------------------------------------------------------------------------------------
		if (alive && finished)
		{
			alive = false;
			retire_ray(wf, q, item, shadow, hit_tri, hit_u, hit_v, tmax, pidx, acc_count, hit_inst);
			if (PRIMARY && rs.primary_cache)
				wf.prim_cache[pidx] = hit_tri != 0xffffffffu ? hit_pos : 0xffffffffu;
		}
	}
	close_launch(rs, curc, PRIMARY, dbg, dbg_t0, dbg_rays, acc_count, n_traced);
}

// ------------------------------------------------------------------------------------------------
// k_wavefront_trace_tl — the trace stages over a two-level scene (TlInstance): the fallback for scenes that are not
// flattened.  Persistent CTAs, 32 queue entries per warp and fetch, one ray per lane walked by traverse_tl; the same
// queue semantics, hit records and accumulator updates as k_wavefront_trace.
// ------------------------------------------------------------------------------------------------
template <bool PRIMARY>
__global__ void __launch_bounds__(256, TRACE_MINB) k_wavefront_trace_tl(const SceneView sc, const ShardView sh, const WavefrontView wf,
															   const RenderSettings rs, const BatchView bv, const uint32_t depth,
															   const uint32_t in_buf)
{
	DepthCounters *curc = &wf.counters[bv.index * MAX_DEPTH_SLOTS + depth];
	const TraceQueue q = open_queue<PRIMARY>(sh, wf, bv, depth, in_buf);
	const uint32_t total = q.total;
	uint32_t *cursor = &curc->trace_cursor;
	const uint32_t lane = threadIdx.x & 31u;
	FrameParams fp;
	if (PRIMARY)
		fp = *wf.frame;
	uint32_t acc_count = 0, n_traced = 0;
	for (;;)
	{
		uint32_t base = 0;
		if (lane == 0)
			base = atomicAdd(cursor, 32u);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (base >= total)
			break;
		const uint32_t item = base + lane;
		if (item >= total)
			continue;
		V3 o = mk(0.f), d = mk(0.f);
		float tmin = 0.f, tmax = 0.f, hit_u = 0.f, hit_v = 0.f;
		bool shadow = false;
		uint32_t pidx = 0, occluder = 0xffffffffu, occ_slot = 0, hit_tri = 0xffffffffu, hit_inst = 0;
		if (!fetch_ray<PRIMARY>(sc, sh, wf, rs, bv, fp, q, depth, in_buf, item, o, d, tmin, tmax, shadow, pidx, occluder, occ_slot, n_traced))
			continue;
		if (shadow)
		{
			if (traverse_tl<true>(sc, o, d, tmin, tmax, hit_tri, hit_u, hit_v, hit_inst))
				hit_tri = 0u; // occluded: anything but "no hit"
		}
		else
			traverse_tl<false>(sc, o, d, tmin, tmax, hit_tri, hit_u, hit_v, hit_inst);
		retire_ray(wf, q, item, shadow, hit_tri, hit_u, hit_v, tmax, pidx, acc_count, hit_inst);
	}
	close_launch(rs, curc, PRIMARY, false, 0ull, 0ull, acc_count, n_traced);
}

// ------------------------------------------------------------------------------------------------
// k_wavefront_trace_cw — the same persistent kernel over the compressed 8-wide BVH (cwbvh.h, setting bvh=8).
// Per node visit a lane loads 80 B (five 128-bit loads) instead of 112 B and tests eight children; children come out
// in octant order, so there is no sorting network, and the stack holds one 8-byte node group per level.
// ------------------------------------------------------------------------------------------------
template <bool PRIMARY>
__global__ void __launch_bounds__(256, TRACE_MINB) k_wavefront_trace_cw(const SceneView sc, const ShardView sh, const WavefrontView wf,
															const RenderSettings rs, const BatchView bv, const uint32_t depth,
															const uint32_t in_buf)
{
	const uint4 *__restrict__ cwn = sc.cw_nodes;

	DepthCounters *curc = &wf.counters[bv.index * MAX_DEPTH_SLOTS + depth];
	const TraceQueue q = open_queue<PRIMARY>(sh, wf, bv, depth, in_buf);
	const uint32_t total = q.total;
	const float4 *__restrict__ tris = reinterpret_cast<const float4 *>(sc.tris);
	uint32_t *cursor = &curc->trace_cursor;
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t lt_mask = (1u << lane) - 1u;
	FrameParams fp;
	if (PRIMARY)
		fp = *wf.frame;

	unsigned long long dbg_t0 = 0, dbg_rays = 0;
	const bool dbg = !PRIMARY && rs.debug != nullptr && int(depth) == rs.debug_depth;
	if (dbg)
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));

	// per-lane ray state
	bool alive = false;
	bool shadow = false;
	uint32_t item = 0;
	V3 o = mk(0.f), d = mk(0.f);
	CwRay ray;
	ray.ox = ray.oy = ray.oz = 0.f, ray.idx = ray.idy = ray.idz = 1.f, ray.octinv = 0u;
	float tmin = 0.f, tmax = 0.f, hit_u = 0.f, hit_v = 0.f;
	uint32_t hit_tri = 0xffffffffu;
	// node group (x = index of the first inner child, y = hits in bits 24..31 | imask), the triangle group held back
	// (x = first triangle record, y = one bit per triangle) and the one found after it; stack of node groups
	uint2 stack[CW_STACK];
	uint2 ng = make_uint2(0u, 0u), held = make_uint2(0u, 0u), tg = make_uint2(0u, 0u);
	int sp = 0;
	bool trav_done = true;
	bool exhausted = false; // warp-uniform: the queue has no more items
	uint32_t acc_count = 0, n_traced = 0, pidx = 0, occluder = 0xffffffffu, occ_slot = 0;
	uint32_t chunk_pos = 0, chunk_end = 0; // warp-uniform: the run of queue entries this warp is working through
	const uint32_t chunk_len = uint32_t(rs.fetch_chunk);

	for (;;)
	{
		// ---- refill idle lanes ---------------------------------------------------------------------------
		const uint32_t idle_mask = __ballot_sync(0xffffffffu, !alive);
		if (idle_mask == 0xffffffffu && exhausted)
			break;
		const int n_idle = __popc(idle_mask);
		if (!exhausted && (n_idle >= rs.fetch_threshold))
		{
			// Default (fetch_chunk = 0): the idle lanes take the next n_idle queue entries with one atomic, so all warps of the
			// GPU work on one narrow front of the queue and share its nodes in L1 / L2.  fetch_chunk > 0 gives each warp a
			// private run of that many consecutive entries instead (its lanes then hold neighbours of the queue, but the
			// warps of an SM are spread over 4,736 runs): measured slower at every run length, profiles/r02.
			if (chunk_pos >= chunk_end)
			{
				const uint32_t want = chunk_len ? chunk_len : uint32_t(n_idle);
				uint32_t base = 0;
				if (lane == 0)
					base = atomicAdd(cursor, want);
				base = __shfl_sync(0xffffffffu, base, 0);
				chunk_pos = min(base, total), chunk_end = min(base + want, total);
				if (base + want >= total)
					exhausted = chunk_len == 0u || base >= total;
			}
			const uint32_t take = min(uint32_t(n_idle), chunk_end - chunk_pos);
			const uint32_t mine = __popc(idle_mask & lt_mask);
			const uint32_t first = chunk_pos;
			chunk_pos += take;
			if (!alive)
			{
				item = first + mine;
				if (mine < take &&
					fetch_ray<PRIMARY>(sc, sh, wf, rs, bv, fp, q, depth, in_buf, item, o, d, tmin, tmax, shadow, pidx, occluder, occ_slot, n_traced))
				{
					dbg_rays++;
					const float tiny = 1e-30f;
					ray.ox = o.x, ray.oy = o.y, ray.oz = o.z;
					ray.idx = 1.0f / (fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x));
					ray.idy = 1.0f / (fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y));
					ray.idz = 1.0f / (fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z));
					ray.octinv = cw_octinv(ray.idx, ray.idy, ray.idz);
					ng = make_uint2(0u, 0x80000000u), held = tg = make_uint2(0u, 0u); // the root is "child 7 ^ octinv of nothing"
					sp = 0, trav_done = false, hit_tri = 0xffffffffu, hit_u = 0.f, hit_v = 0.f;
					alive = true;
				}
			}
		}

		// ---- traverse: while-while over node groups with one triangle group held back ------------------------------------
		bool finished = false;
		if (alive)
		{
			while (!trav_done || held.y != 0u)
			{
				// node phase: every lane enters the nearest hit child of its node group; the triangles found there are held
				// back (the lane keeps descending) until every lane of the warp holds some
				while (!trav_done)
				{
					if (ng.y <= 0x00ffffffu)
					{
						if (sp == 0)
						{
							trav_done = true;
							break;
						}
						ng = stack[--sp];
					}
					const uint32_t bit = 31u - uint32_t(__clz(ng.y));
					const uint32_t imask = ng.y;
					ng.y &= ~(1u << bit);
					if (ng.y > 0x00ffffffu)
						stack[sp++] = ng;
					const uint32_t slot = (bit - 24u) ^ ray.octinv;
					const uint32_t node = ng.x + uint32_t(__popc(imask & ~(0xffffffffu << slot)));
					const uint4 *np_ = cwn + size_t(node) * 5;
					const uint4 q0 = __ldg(np_), q1 = __ldg(np_ + 1), q2 = __ldg(np_ + 2), q3 = __ldg(np_ + 3), q4 = __ldg(np_ + 4);
					const uint32_t w[20] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y,
											q2.z, q2.w, q3.x, q3.y, q3.z, q3.w, q4.x, q4.y, q4.z, q4.w};
					const uint32_t hm = cw_intersect_children(w, ray, tmin, tmax);
					ng = make_uint2(q1.x, (hm & 0xff000000u) | (q0.w >> 24));
					const uint32_t tbits = hm & 0x00ffffffu;
					if (tbits)
					{
						if (held.y == 0u)
							held = make_uint2(q1.y, tbits);
						else
						{
							tg = make_uint2(q1.y, tbits); // a second group: this lane waits for the triangle phase
							break;
						}
					}
					if (!__any_sync(__activemask(), held.y == 0u))
						break;
				}
				// triangle phase: the held group, then the one that stopped the lane
				while (held.y != 0u)
				{
					const uint32_t bit = 31u - uint32_t(__clz(held.y));
					held.y &= ~(1u << bit);
					const size_t ti = size_t(held.x + bit) * 3;
					const float4 a = __ldg(tris + ti + 0);
					const float4 b = __ldg(tris + ti + 1);
					const float4 c = __ldg(tris + ti + 2);
					const V3 p0 = mk(a.x, a.y, a.z), e1 = mk(a.w, b.x, b.y), e2 = mk(b.z, b.w, c.x);
					const V3 h = cross(d, e2);
					const float det = dot(e1, h);
					const float eps = c.z;
					const float f = 1.0f / det;
					const V3 s = o - p0;
					const float u = f * dot(s, h);
					const V3 q = cross(s, e1);
					const float vv = f * dot(d, q);
					const float t = f * dot(e2, q);
					const bool ok = !(det > -eps && det < eps) && !(u < 0.0f || u > 1.0f) && !(vv < 0.0f || u + vv > 1.0f) && t > tmin && tmax > t;
					if (ok)
					{
						tmax = t, hit_u = u, hit_v = vv, hit_tri = __float_as_uint(c.y);
						if (shadow)
							trav_done = true, sp = 0, held.y = 0u, tg.y = 0u, ng.y = 0u; // any hit ends a connect ray
					}
					if (held.y == 0u)
						held = tg, tg.y = 0u;
				}
				// too few lanes left in this loop: leave it so the idle lanes can fetch new rays
				if (__popc(__activemask()) <= 32 - rs.fetch_threshold)
					break;
			}
			finished = trav_done && held.y == 0u;
		}
		// ---- retire --------------------------------------------------------------------------------------------------
		if (alive && finished)
		{
			alive = false;
			retire_ray(wf, q, item, shadow, hit_tri, hit_u, hit_v, tmax, pidx, acc_count);
		}
	}
	close_launch(rs, curc, PRIMARY, dbg, dbg_t0, dbg_rays, acc_count, n_traced);
}


#endif // RFW_PART != 2 (trace part)

// ------------------------------------------------------------------------------------------------
// shading helpers — bsdf/tools.h, bsdf/compat.h, bsdf/disney.h, CUDART/src/getShadingData.h,
// CUDART/src/lights.h
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t PackNormal(V3 N) // tools.h:10-21
{
	const float f = 65535.0f / fmaxf(sqrtf(8.0f * N.z + 8.0f), 0.0001f);
	return uint32_t(N.x * f + 32767.0f) + (uint32_t(N.y * f + 32767.0f) << 16);
}
__device__ __forceinline__ V3 UnpackNormal(uint32_t p) // tools.h:22-29
{
	float nx = float(p & 65535u) * (2.0f / 65535.0f), ny = float(p >> 16) * (2.0f / 65535.0f);
	nx += -1.f, ny += -1.f;
	float l = nx * -nx + ny * -ny + 1.0f;
	const float nz = l;
	l = sqrtf(l);
	nx *= l, ny *= l;
	return mk(nx, ny, nz) * 2.0f + mk(0.f, 0.f, -1.f);
}
__device__ __forceinline__ float SurvivalProbability(V3 d) { return fminf(1.0f, fmaxf(fmaxf(d.x, d.y), d.z)); }
__device__ __forceinline__ void clampIntensity(V3 &v, float clampValue) // tools.h:184-192
{
	const float m = fmaxf(v.x, fmaxf(v.y, v.z));
	if (m > clampValue)
		v = v * (clampValue / m);
}
__device__ __forceinline__ void createTangentSpace(V3 N, V3 &T, V3 &B) // tools.h:205-212
{
	const float s = signf(N.z);
	const float a = -1.0f / (s + N.z);
	const float b = N.x * N.y * a;
	T = mk(1.0f + s * N.x * N.x * a, s * b, -s * N.x);
	B = mk(b, s + N.y * N.y * a, -N.y);
}
__device__ __forceinline__ V3 DiffuseReflectionUniform(float r0, float r1) // tools.h:102-108
{
	const float term1 = TWOPI * r0, term2 = sqrtf(1 - r1 * r1);
	float s, c;
	sincosf(term1, &s, &c);
	return mk(c * term2, s * term2, r1);
}
__device__ __forceinline__ V3 DiffuseReflectionCosWeighted(float r0, float r1) // tools.h:110-117
{
	const float term1 = TWOPI * r0;
	const float term2 = sqrtf(1.0f - r1);
	float s, c;
	sincosf(term1, &s, &c);
	return normalize(mk(c * term2, s * term2, sqrtf(r1)));
}

struct ShadingData // bsdf/compat.h:47-74
{
	V3 color;
	uint32_t flags;
	V3 absorption;
	uint32_t p0, p1, p2;
	__device__ __forceinline__ static float c2f(uint32_t v, int s) { return float((v >> s) & 255u) * (1.0f / 255.0f); }
	__device__ __forceinline__ float metallic() const { return c2f(p0, 0); }
	__device__ __forceinline__ float subsurface() const { return c2f(p0, 8); }
	__device__ __forceinline__ float specular() const { return c2f(p0, 16); }
	__device__ __forceinline__ float roughness() const { return fmaxf(0.001f, c2f(p0, 24)); }
	__device__ __forceinline__ float spectint() const { return c2f(p1, 0); }
	__device__ __forceinline__ float clearcoat() const { return c2f(p2, 0); }
	__device__ __forceinline__ float clearcoatgloss() const { return c2f(p2, 8); }
	__device__ __forceinline__ float transmission() const { return c2f(p2, 16); }
	__device__ __forceinline__ float eta() const { return c2f(p2, 24); }
	__device__ __forceinline__ bool isEmissive() const { return color.x > 1.0f || color.y > 1.0f || color.z > 1.0f; }
};

__device__ __forceinline__ bool Refract(V3 wi, V3 n, float eta, V3 &wt) // disney.h:20-30
{
	const float cosThetaI = dot(n, wi);
	const float sin2ThetaI = fmaxf(0.0f, 1.0f - cosThetaI * cosThetaI);
	const float sin2ThetaT = eta * eta * sin2ThetaI;
	if (sin2ThetaT >= 1)
		return false;
	const float cosThetaT = sqrtf(1.0f - sin2ThetaT);
	wt = (wi * -1.0f) * eta + n * (eta * cosThetaI - cosThetaT);
	return true;
}
__device__ __forceinline__ float SchlickFresnel(float u) // disney.h:32-36
{
	const float m = fminf(fmaxf(1 - u, 0.0f), 1.0f);
	return (m * m) * (m * m) * m;
}
__device__ __forceinline__ float GTR1(float NDotH, float a) // disney.h:38-45
{
	if (a >= 1.0f)
		return INVPI;
	const float a2 = a * a;
	const float t = 1 + (a2 - 1) * NDotH * NDotH;
	return (a2 - 1) / (PI * logf(a2) * t);
}
__device__ __forceinline__ float GTR2(float NDotH, float a) // disney.h:47-52
{
	const float a2 = a * a;
	const float t = 1.0f + (a2 - 1.0f) * NDotH * NDotH;
	return a2 / (PI * t * t);
}
__device__ __forceinline__ float SmithGGX(float NDotv, float alphaG) // disney.h:54-59
{
	const float a = alphaG * alphaG;
	const float b = NDotv * NDotv;
	return 1 / (NDotv + sqrtf(a + b - a * b));
}
__device__ __forceinline__ float Fr(float VDotN, float eio) // disney.h:61-72
{
	const float SinThetaT2 = sqr(eio) * (1.0f - VDotN * VDotN);
	if (SinThetaT2 > 1.0f)
		return 1.0f;
	const float LDotN = sqrtf(1.0f - SinThetaT2);
	const float eta = 1.0f / eio;
	const float r1 = (VDotN - eta * LDotN) / (VDotN + eta * LDotN);
	const float r2 = (LDotN - eta * VDotN) / (LDotN + eta * VDotN);
	return 0.5f * (sqr(r1) + sqr(r2));
}
__device__ __forceinline__ V3 SafeNormalize(V3 a) // disney.h:74-81
{
	const float ls = dot(a, a);
	if (ls > 0.0f)
		return a * (1.0f / sqrtf(ls));
	return mk(0.f);
}

__device__ float BSDFPdf(const ShadingData &sd, V3 N, V3 wo, V3 wi) // disney.h:83-101
{
	float bsdfPdf = 0.0f, brdfPdf;
	if (dot(wi, N) <= 0.0f)
		brdfPdf = INV2PI * sd.subsurface() * 0.5f;
	else
	{
		const float F = Fr(dot(N, wo), sd.eta());
		const V3 halfway = SafeNormalize(wi + wo);
		const float cosThetaHalf = fabsf(dot(halfway, N));
		const float pdfHalf = GTR2(cosThetaHalf, sd.roughness()) * cosThetaHalf;
		const float pdfSpec = 0.25f * pdfHalf / fmaxf(1.e-6f, dot(wi, halfway));
		const float pdfDiff = fabsf(dot(wi, N)) * INVPI * (1.0f - sd.subsurface());
		bsdfPdf = pdfSpec * F;
		brdfPdf = lerpf(pdfDiff, pdfSpec, 0.5f);
	}
	return lerpf(brdfPdf, bsdfPdf, sd.transmission());
}

#ifndef RFW_BSDF_NOINLINE
#define RFW_BSDF_NOINLINE
#endif
RFW_BSDF_NOINLINE __device__ V3 BSDFEval(const ShadingData &sd, V3 N, V3 wo, V3 wi, float t, bool backfacing) // disney.h:104-185
{
	const float NDotL = dot(N, wi);
	const float NDotV = dot(N, wo);
	const V3 H = normalize(wi + wo);
	const float NDotH = dot(N, H);
	const float LDotH = dot(wi, H);
	const V3 Cdlin = sd.color;
	const float Cdlum = .3f * Cdlin.x + .6f * Cdlin.y + .1f * Cdlin.z;
	const V3 Ctint = Cdlum > 0.0f ? Cdlin / Cdlum : mk(1.0f);
	const float TRANSMISSION = sd.transmission(), METALLIC = sd.metallic(), SUBSURFACE = sd.subsurface();
	const V3 Cspec0 = lerp3(lerp3(mk(1.0f), Ctint, sd.spectint()) * (sd.specular() * .08f), Cdlin, METALLIC);
	V3 bsdf = mk(0.f), brdf = mk(0.f);
	if (TRANSMISSION > 0.0f)
	{
		if (NDotL <= 0)
		{
			const float F = Fr(NDotV, sd.eta());
			bsdf = mk((1.0f - F) / fabsf(NDotL) * (1.0f - METALLIC) * TRANSMISSION);
		}
		else
		{
			const float a = sd.roughness();
			const float Ds = GTR2(NDotH, a);
			const float FH = Fr(LDotH, sd.eta());
			const V3 Fs = lerp3(Cspec0, mk(1.0f), FH);
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
				const V3 s = mk(sqrtf(sd.color.x), sqrtf(sd.color.y), sqrtf(sd.color.z));
				const float FL = SchlickFresnel(fabsf(NDotL)), FV = SchlickFresnel(NDotV);
				const float Fd = (1.0f - 0.5f * FL) * (1.0f - 0.5f * FV);
				brdf = s * INVPI * SUBSURFACE * Fd * (1.0f - METALLIC);
			}
		}
		else
		{
			const float a = sd.roughness();
			const float Ds = GTR2(NDotH, a);
			const float FH = SchlickFresnel(LDotH);
			const V3 Fs = lerp3(Cspec0, mk(1.0f), FH);
			const float Gs = SmithGGX(NDotV, a) * SmithGGX(NDotL, a);
			const float FL = SchlickFresnel(NDotL), FV = SchlickFresnel(NDotV);
			const float Fd90 = 0.5f + 2.0f * LDotH * LDotH * a;
			const float Fd = lerpf(1.0f, Fd90, FL) * lerpf(1.0f, Fd90, FV);
			const float Dr = GTR1(NDotH, lerpf(.1f, .001f, sd.clearcoatgloss()));
			const float Fc = lerpf(.04f, 1.0f, FH);
			const float Gr = SmithGGX(NDotL, .25f) * SmithGGX(NDotV, .25f);
			brdf = Cdlin * (INVPI * Fd) * (1.0f - METALLIC) * (1.0f - SUBSURFACE) + Fs * Gs * Ds +
				   mk(sd.clearcoat() * Gr * Fc * Dr);
		}
	}
	const V3 fin = lerp3(brdf, bsdf, TRANSMISSION);
	if (backfacing)
		return fin * mk(expf(-sd.absorption.x * t), expf(-sd.absorption.y * t), expf(-sd.absorption.z * t));
	return fin;
}

__device__ void BSDFSample(const ShadingData &sd, V3 T, V3 B, V3 N, V3 wo, V3 &wi, float &pdf, float r3,
						   float r4) // disney.h:188-262
{
	const float transmission = sd.transmission();
	if (r3 < transmission)
	{
		const float F = Fr(dot(N, wo), sd.eta());
		if (r4 < F)
		{
			const float r1 = r3 / transmission;
			const float r2 = r4 / F;
			const float cosThetaHalf = sqrtf((1.0f - r2) / (1.0f + (sqr(sd.roughness()) - 1.0f) * r2));
			const float sinThetaHalf = sqrtf(fmaxf(0.0f, 1.0f - sqr(cosThetaHalf)));
			float sinPhiHalf, cosPhiHalf;
			sincosf(r1 * TWOPI, &sinPhiHalf, &cosPhiHalf);
			V3 halfway = T * (sinThetaHalf * cosPhiHalf) + B * (sinThetaHalf * sinPhiHalf) + N * cosThetaHalf;
			if (dot(halfway, wo) <= 0.0f)
				halfway = halfway * -1.0f;
			wi = reflect(wo * -1.0f, halfway);
		}
		else
		{
			pdf = 0;
			if (Refract(wo, N, sd.eta(), wi))
				pdf = (1.0f - F) * transmission;
		}
		return;
	}
	const float r1 = (r3 - transmission) / (1 - transmission);
	if (r4 < 0.5f)
	{
		const float r2 = r4 * 2;
		const float subsurface = sd.subsurface();
		V3 d;
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
		const float cosThetaHalf = sqrtf((1.0f - r2) / (1.0f + (sqr(sd.roughness()) - 1.0f) * r2));
		const float sinThetaHalf = sqrtf(fmaxf(0.0f, 1.0f - sqr(cosThetaHalf)));
		float sinPhiHalf, cosPhiHalf;
		sincosf(r1 * TWOPI, &sinPhiHalf, &cosPhiHalf);
		V3 halfway = T * (sinThetaHalf * cosPhiHalf) + B * (sinThetaHalf * sinPhiHalf) + N * cosThetaHalf;
		if (dot(halfway, wo) <= 0.0f)
			halfway = halfway * -1.0f;
		wi = reflect(wo * -1.0f, halfway);
	}
	pdf = BSDFPdf(sd, N, wo, wi);
}

// ---- textures: CUDART/src/getShadingData.h:23-98 ------------------------------------------------
__device__ __forceinline__ float4 texel_rgba8(const SceneView &sc, uint32_t i)
{
	if (i >= sc.uint_texel_count)
		return make_float4(0.f, 0.f, 0.f, 0.f);
	const uint32_t v = __ldg(sc.uint_texels + i);
	const float r = 1.0f / 256.0f;
	return make_float4(float(v & 255u) * r, float((v >> 8) & 255u) * r, float((v >> 16) & 255u) * r, float(v >> 24) * r);
}
__device__ float4 FetchTexel(const SceneView &sc, float tcx, float tcy, int o, int w, int h)
{
	if (w <= 0 || h <= 0)
		return make_float4(0.f, 0.f, 0.f, 0.f);
	const float tx = (fmaxf(tcx + 1000, 0.0f) * float(w)) - 0.5f, ty = (fmaxf(tcy + 1000, 0.0f) * float(h)) - 0.5f;
	// wrap: tx, ty >= -0.5 so the truncated coordinates are >= 0 and `%` is a mask for power-of-two sizes (every mip of
	// a power-of-two texture); other sizes take the division
	const bool pow2 = ((w & (w - 1)) | (h & (h - 1))) == 0;
	const int iu = pow2 ? (__float2int_rz(tx) & (w - 1)) : (__float2int_rz(tx) % w);
	const int iv = pow2 ? (__float2int_rz(ty) & (h - 1)) : (__float2int_rz(ty) % h);
	const float fu = tx - floorf(tx);
	const float fv = ty - floorf(ty);
	const float w0 = (1 - fu) * (1 - fv);
	const float w1 = fu * (1 - fv);
	const float w2 = (1 - fu) * fv;
	const float w3 = 1 - (w0 + w1 + w2);
	const uint32_t iu1 = pow2 ? (uint32_t(iu + 1) & uint32_t(w - 1)) : (uint32_t(iu + 1) % uint32_t(w));
	const uint32_t iv1 = pow2 ? (uint32_t(iv + 1) & uint32_t(h - 1)) : (uint32_t(iv + 1) % uint32_t(h));
	const float4 p0 = texel_rgba8(sc, uint32_t(o) + iu + uint32_t(iv) * w), p1 = texel_rgba8(sc, uint32_t(o) + iu1 + uint32_t(iv) * w);
	const float4 p2 = texel_rgba8(sc, uint32_t(o) + iu + iv1 * w), p3 = texel_rgba8(sc, uint32_t(o) + iu1 + iv1 * w);
	return make_float4(p0.x * w0 + p1.x * w1 + p2.x * w2 + p3.x * w3, p0.y * w0 + p1.y * w1 + p2.y * w2 + p3.y * w3,
					   p0.z * w0 + p1.z * w1 + p2.z * w2 + p3.z * w3, p0.w * w0 + p1.w * w1 + p2.w * w2 + p3.w * w3);
}
#ifndef RFW_TEX_NOINLINE
#define RFW_TEX_NOINLINE // experiment hook: -DRFW_TEX_NOINLINE=__noinline__ shrinks k_shade
#endif
RFW_TEX_NOINLINE __device__ float4 FetchTexelTrilinear(const SceneView &sc, float lambda, float tcx, float tcy, int offset, int width,
									  int height)
{
	const int level0 = min(MIPLEVELCOUNT - 1, __float2int_rz(lambda));
	const int level1 = min(MIPLEVELCOUNT - 1, level0 + 1);
	const float f = lambda - floorf(lambda);
	uint32_t offset0 = offset, width0 = width, height0 = height;
	for (int i = 0; i < level0; i++)
		offset0 += width0 * height0, width0 >>= 1u, height0 >>= 1u;
	// level1 = min(4, level0 + 1) differs from level0 only for 0 <= level0 < 4 (negative levels select
	// the base level for both fetches, exactly like the reference's zero-trip loops)
	const bool second = level0 >= 0 && level0 < MIPLEVELCOUNT - 1;
	const float4 p0 = FetchTexel(sc, tcx, tcy, int(offset0), int(width0), int(height0));
	float4 p1 = p0;
	if (second)
		p1 = FetchTexel(sc, tcx, tcy, int(offset0 + width0 * height0), int(width0 >> 1u), int(height0 >> 1u));
	return make_float4((1.0f - f) * p0.x + f * p1.x, (1.0f - f) * p0.y + f * p1.y, (1.0f - f) * p0.z + f * p1.z,
					   (1.0f - f) * p0.w + f * p1.w);
}

struct MapDesc
{
	int w, h;
	float us, vs, uo, vo;
	uint32_t addr;
};
__device__ __forceinline__ MapDesc load_map(const rfwb200_map_desc *m)
{
	const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(m));
	MapDesc d;
	d.w = int(short(raw.x & 0xffffu)), d.h = int(short(raw.x >> 16));
	d.us = __half2float(__ushort_as_half((unsigned short)(raw.y & 0xffffu)));
	d.vs = __half2float(__ushort_as_half((unsigned short)(raw.y >> 16)));
	d.uo = __half2float(__ushort_as_half((unsigned short)(raw.z & 0xffffu)));
	d.vo = __half2float(__ushort_as_half((unsigned short)(raw.z >> 16)));
	d.addr = raw.w;
	return d;
}

// getShadingData.h:100-217 on the repacked ShadeTri (normals are already in world space)
__device__ ShadingData getShadingData(const SceneView &sc, V3 D, float u, float v, float coneWidth, const ShadeTri &tri, V3 &N,
									  V3 &iN, V3 &T, V3 &B)
{
	ShadingData r;
	const rfwb200_material *mat = reinterpret_cast<const rfwb200_material *>(sc.materials) + tri.material;
	const uint4 base = __ldg(reinterpret_cast<const uint4 *>(mat));
	const uint4 par = __ldg(reinterpret_cast<const uint4 *>(mat) + 1);
	const uint32_t flags = base.w;
	r.flags = 0;
	r.color = mk(__half2float(__ushort_as_half((unsigned short)(base.x & 0xffffu))),
				 __half2float(__ushort_as_half((unsigned short)(base.x >> 16))),
				 __half2float(__ushort_as_half((unsigned short)(base.y & 0xffffu))));
	r.absorption = mk(__half2float(__ushort_as_half((unsigned short)(base.y >> 16))),
					  __half2float(__ushort_as_half((unsigned short)(base.z & 0xffffu))),
					  __half2float(__ushort_as_half((unsigned short)(base.z >> 16))));
	r.p0 = par.x, r.p1 = par.y, r.p2 = par.z;
	const float w = 1.0f - u - v; // u, v: weights of vertex 0 and vertex 1 (the reference's convention), w: of vertex 2
	N = mk(tri.Nx, tri.Ny, tri.Nz);
	iN = N;
	if (has_flag(flags, HasSmoothNormals))
		iN = normalize(mk(tri.n0x, tri.n0y, tri.n0z) * u + mk(tri.n1x, tri.n1y, tri.n1z) * v + mk(tri.n2x, tri.n2y, tri.n2z) * w);
	createTangentSpace(iN, T, B);
	if (has_flag(flags, HasDiffuseMap))
	{
		const float tu = u * tri.u0 + v * tri.u1 + w * tri.u2;
		const float tv = u * tri.v0 + v * tri.v1 + w * tri.v2;
		const float lambda = tri.lod + log2f(coneWidth * (1.0f / fabsf(dot(-D, N))));
		const MapDesc m0 = load_map(&mat->tex0);
		const float4 texel = FetchTexelTrilinear(sc, lambda, m0.us * (m0.uo + tu), m0.vs * (m0.vo + tv), int(m0.addr), m0.w, m0.h);
		if (has_flag(flags, HasAlpha) && texel.w < 0.5f)
		{
			r.flags |= 1;
			return r;
		}
		r.color = r.color * mk(texel.x, texel.y, texel.z);
		if (has_flag(flags, Has2ndDiffuseMap))
		{
			const MapDesc m = load_map(&mat->tex1);
			const float4 t = FetchTexelTrilinear(sc, lambda, m.us * (m.uo + tu), m.vs * (m.vo + tv), int(m.addr), m.w, m.h);
			r.color = r.color + mk(t.x, t.y, t.z);
		}
		if (has_flag(flags, Has3rdDiffuseMap))
		{
			const MapDesc m = load_map(&mat->tex2);
			const float4 t = FetchTexelTrilinear(sc, lambda, m.us * (m.uo + tu), m.vs * (m.vo + tv), int(m.addr), m.w, m.h);
			r.color = r.color + mk(t.x, t.y, t.z);
		}
		if (has_flag(flags, HasNormalMap))
		{
			const MapDesc m = load_map(&mat->nmap0);
			const float4 t = FetchTexel(sc, m.us * (m.uo + tu), m.vs * (m.vo + tv), int(m.addr), m.w, m.h);
			V3 sn = (mk(t.x, t.y, t.z) - mk(0.5f)) * 2.0f;
			if (has_flag(flags, Has2ndNormalMap) || has_flag(flags, Has3rdNormalMap))
			{
				// the reference reads layer 1 for both the 2nd and the 3rd layer (getShadingData.h:194-200)
				const MapDesc m1 = load_map(&mat->nmap1);
				const float4 t1 = FetchTexel(sc, m1.us * (m1.uo + tu), m1.vs * (m1.vo + tv), int(m1.addr), m1.w, m1.h);
				const V3 l1 = (mk(t1.x, t1.y, t1.z) - mk(0.5f)) * 2.0f;
				if (has_flag(flags, Has2ndNormalMap))
					sn = sn + l1;
				if (has_flag(flags, Has3rdNormalMap))
					sn = sn + l1;
			}
			sn = normalize(sn);
			iN = normalize(T * sn.x + B * sn.y + iN * sn.z);
		}
		r.color = r.color * mk(texel.x, texel.y, texel.z); // second multiply, getShadingData.h:213
	}
	return r;
}

// ---- lights: CUDART/src/lights.h ---------------------------------------------------------------
struct AreaLightRegs
{
	V3 normal, radiance, v0, v1, v2;
	float energy, area;
};
__device__ __forceinline__ AreaLightRegs load_area_light(const SceneView &sc, int idx)
{
	const float4 *p = reinterpret_cast<const float4 *>(sc.area_lights) + size_t(idx) * 6;
	const float4 a = __ldg(p + 0), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3), e = __ldg(p + 4), f = __ldg(p + 5);
	AreaLightRegs l;
	l.energy = a.w;
	l.normal = mk(b.x, b.y, b.z), l.area = b.w;
	l.radiance = mk(c.x, c.y, c.z);
	l.v0 = mk(d.x, d.y, d.z), l.v1 = mk(e.x, e.y, e.z), l.v2 = mk(f.x, f.y, f.z);
	return l;
}
__device__ float PotentialAreaLightContribution(const SceneView &sc, int idx, V3 O, V3 N, V3 I, V3 bary) // :17-36
{
	const AreaLightRegs light = load_area_light(sc, idx);
	V3 L = I;
	if (bary.x >= 0)
		L = light.v0 * bary.x + light.v1 * bary.y + light.v2 * bary.z;
	L = L - O;
	const float att = 1.0f / dot(L, L);
	L = normalize(L);
	const float LNdotL = fmaxf(0.0f, -dot(light.normal, L));
	const float NdotL = fmaxf(0.0f, dot(N, L));
	return light.energy * LNdotL * NdotL * att;
}
__device__ float PotentialPointLightContribution(const SceneView &sc, int idx, V3 I, V3 N) // :38-46
{
	const float4 *p = reinterpret_cast<const float4 *>(sc.point_lights) + size_t(idx) * 2;
	const float4 a = __ldg(p);
	const V3 L = mk(a.x, a.y, a.z) - I;
	const float NdotL = fmaxf(0.0f, dot(N, L));
	const float att = 1.0f / dot(L, L);
	return a.w * NdotL * att;
}
__device__ float PotentialSpotLightContribution(const SceneView &sc, int idx, V3 I, V3 N) // :48-68
{
	const float4 *p = reinterpret_cast<const float4 *>(sc.spot_lights) + size_t(idx) * 3;
	const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
	V3 L = mk(a.x, a.y, a.z) - I;
	const float att = 1.0f / dot(L, L);
	L = normalize(L);
	const float d = (fmaxf(0.0f, -dot(L, mk(c.x, c.y, c.z))) - b.w) / (a.w - b.w);
	const float NdotL = fmaxf(0.0f, dot(N, L));
	const float LNdotL = fmaxf(0.0f, fminf(1.0f, d));
	return c.w * LNdotL * NdotL * att;
}
__device__ float PotentialDirectionalLightContribution(const SceneView &sc, int idx, V3 N) // :70-76
{
	const float4 a = __ldg(reinterpret_cast<const float4 *>(sc.dir_lights) + size_t(idx) * 2);
	const float LNdotL = fmaxf(0.0f, -dot(mk(a.x, a.y, a.z), N));
	return a.w * LNdotL;
}
__device__ float LightPickProb(const SceneView &sc, int idx, V3 O, V3 N, V3 I) // :83-116
{
	float sum = 0, mine = 0;
	const int na = int(sc.lights.area);
	for (int i = 0; i < na; i++)
	{
		const float c = PotentialAreaLightContribution(sc, i, O, N, I, mk(-1.0f));
		if (i == idx)
			mine = c;
		sum += c;
	}
	for (uint32_t i = 0; i < sc.lights.point; i++)
		sum += PotentialPointLightContribution(sc, int(i), O, N);
	for (uint32_t i = 0; i < sc.lights.spot; i++)
		sum += PotentialSpotLightContribution(sc, int(i), O, N);
	for (uint32_t i = 0; i < sc.lights.directional; i++)
		sum += PotentialDirectionalLightContribution(sc, int(i), N);
	if (sum <= 0)
		return 0;
	if (idx < 0 || idx >= na || idx >= MAX_IS_LIGHTS)
		return 0;
	return mine / sum;
}
// lights.h:119-157 subdivides the unit triangle A=(1,0) B=(0,1) C=(0,0) sixteen times — digit d of r0 (two bits, most
// significant first) keeps corner d's sub-triangle (d = 0: the middle one) — and returns the centroid of what is left.
// Every step is linear in the corners, so the sum of the final corners is w . (A,B,C) with the row vector
// w = [1 1 1] T(d15) ... T(d0); with h = w / 2 and H = hA + hB + hC (constant: the components always sum to 3):
//   d = 0: w' = H - h      d = 1: w' = h + (H,0,0)      d = 2: w' = h + (0,H,0)      d = 3: w' = h + (0,0,H)
// All corner coordinates of the reference's loop are dyadic rationals with at most 16 fractional bits, i.e. exact in
// float, so the same sums in 16.16 fixed point give the identical centroid — without a 16-trip loop around a four-way
// branch (17 % of k_shade's instructions at depth 0, profiles/r02).
__device__ __forceinline__ V3 RandomBarycentrics(float r0)
{
	uint32_t uf = __float2uint_rz(r0 * 4294967295.0f);
	constexpr uint32_t H = 3u << 15;
	uint32_t wA = 1u << 16, wB = 1u << 16;
#pragma unroll
	for (int i = 0; i < 16; ++i, uf >>= 2) // least significant digit first: it is the last subdivision
	{
		const uint32_t d = uf & 3u, hA = wA >> 1, hB = wB >> 1;
		wA = d == 0u ? H - hA : (d == 1u ? hA + H : hA);
		wB = d == 0u ? H - hB : (d == 2u ? hB + H : hB);
	}
	const float rx = (float(wA) * (1.0f / 65536.0f)) * 0.3333333f, ry = (float(wB) * (1.0f / 65536.0f)) * 0.3333333f;
	return mk(rx, ry, 1.0f - rx - ry);
}
__device__ V3 RandomPointOnLight(const SceneView &sc, float r0, float r1, V3 I, V3 N, float &pickProb, float &lightPdf,
								 V3 &lightColor) // :159-265
{
	const int na = int(sc.lights.area), np = int(sc.lights.point), ns = int(sc.lights.spot), nd = int(sc.lights.directional);
	const int lightCount = na + np + ns + nd;
	const V3 bary = RandomBarycentrics(r0);
	float potential[MAX_IS_LIGHTS];
	float sum = 0, total = 0;
	int lights = 0, lightIdx = 0;
	for (int i = 0; i < na; i++)
	{
		const float c = PotentialAreaLightContribution(sc, i, I, N, mk(0.f), bary);
		if (lights < MAX_IS_LIGHTS)
			potential[lights] = c;
		lights++, sum += c;
	}
	for (int i = 0; i < np; i++)
	{
		const float c = PotentialPointLightContribution(sc, i, I, N);
		if (lights < MAX_IS_LIGHTS)
			potential[lights] = c;
		lights++, sum += c;
	}
	for (int i = 0; i < ns; i++)
	{
		const float c = PotentialSpotLightContribution(sc, i, I, N);
		if (lights < MAX_IS_LIGHTS)
			potential[lights] = c;
		lights++, sum += c;
	}
	for (int i = 0; i < nd; i++)
	{
		const float c = PotentialDirectionalLightContribution(sc, i, N);
		if (lights < MAX_IS_LIGHTS)
			potential[lights] = c;
		lights++, sum += c;
	}
	lights = min(lights, MAX_IS_LIGHTS);
	if (sum <= 0)
	{
		lightPdf = 0;
		return mk(1.0f);
	}
	r1 *= sum;
	float picked = potential[0];
	for (int i = 0; i < lights; i++)
	{
		total += potential[i];
		if (total >= r1)
		{
			lightIdx = i, picked = potential[i];
			break;
		}
	}
	pickProb = picked / sum;
	lightIdx = min(max(lightIdx, 0), lightCount - 1);
	if (lightIdx < na)
	{
		const AreaLightRegs light = load_area_light(sc, lightIdx);
		lightColor = light.radiance;
		const V3 P = light.v0 * bary.x + light.v1 * bary.y + light.v2 * bary.z;
		V3 L = I - P;
		const float sqDist = dot(L, L);
		L = normalize(L);
		const float LNdotL = dot(L, light.normal);
		const float reciSolidAngle = sqDist / (light.area * LNdotL);
		// DeviceAreaLight::getEnergy() is length(radiance), not the stored field (device_structs.h:116)
		lightPdf = (LNdotL > 0 && dot(L, N) < 0) ? (reciSolidAngle * (1.0f / length(light.radiance))) : 0;
		return P;
	}
	if (lightIdx < na + np)
	{
		const float4 *p = reinterpret_cast<const float4 *>(sc.point_lights) + size_t(lightIdx - na) * 2;
		const float4 a = __ldg(p), b = __ldg(p + 1);
		const V3 pos = mk(a.x, a.y, a.z);
		lightColor = mk(b.x, b.y, b.z);
		const V3 L = I - pos;
		const float sqDist = dot(L, L);
		lightPdf = dot(L, N) < 0 ? (sqDist / a.w) : 0;
		return pos;
	}
	if (lightIdx < na + np + ns)
	{
		const float4 *p = reinterpret_cast<const float4 *>(sc.spot_lights) + size_t(lightIdx - (na + np)) * 3;
		const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
		const V3 P = mk(a.x, a.y, a.z), Dl = mk(c.x, c.y, c.z);
		V3 L = I - P;
		const float sqDist = dot(L, L);
		L = normalize(L);
		const float d = fmaxf(0.0f, dot(L, Dl) - b.w) / (a.w - b.w);
		const float LNdotL = fminf(1.0f, d);
		lightPdf = (LNdotL > 0 && dot(L, N) < 0) ? (sqDist / (LNdotL * c.w)) : 0;
		lightColor = mk(b.x, b.y, b.z);
		return P;
	}
	const float4 *p = reinterpret_cast<const float4 *>(sc.dir_lights) + size_t(lightIdx - (na + np + ns)) * 2;
	const float4 a = __ldg(p), b = __ldg(p + 1);
	const V3 L = mk(a.x, a.y, a.z);
	lightColor = mk(b.x, b.y, b.z);
	const float NdotL = dot(L, N);
	lightPdf = NdotL < 0 ? (1.0f / a.w) : 0;
	return I - L * 1000.0f;
}

#if RFW_PART != 1 // ---- shade part: compiled with -use_fast_math (Makefile) ----
// ------------------------------------------------------------------------------------------------
// k_shade — Kernels.cu:571-794.  Each thread shades one path into registers; the three outputs
// (accumulate, connect-queue entry, extension-queue entry) are committed at one converged point
// with warp-aggregated slot allocation.
// ------------------------------------------------------------------------------------------------
#ifndef SHADE_MINB
#define SHADE_MINB 7 // 72 registers: shade 3.37 ms/frame vs 3.45 at 6 (80 registers) and 3.44 at 5 (96); IEEE build of an
					 // earlier kernel: 4.94 at 6, 5.67 at 4, 4.92 at 8.  Final kernel of round 2 (profiles/r02/sweep19): 2.93 / 2.89 /
					 // 2.88 / 2.92 ms at 6 / 7 / 8 / 9 CTAs per SM (80 / 72 / 64 / 56 registers) — flat
#endif
__global__ void __launch_bounds__(128, SHADE_MINB) K_SHADE(const SceneView sc, const ShardView sh, const WavefrontView wf,
											  const RenderSettings rs, const BatchView bv, const uint32_t pathLength,
											  const uint32_t buf, const uint32_t nbuf, const ShardSync sync)
{
	DepthCounters *curc = &wf.counters[bv.index * MAX_DEPTH_SLOTS + pathLength];
	const uint32_t n_paths = pathLength == 0 ? bv.items : wf.counters[bv.index * MAX_DEPTH_SLOTS + pathLength - 1].ext;
	const FrameParams fp = *wf.frame;
	const uint32_t lane = threadIdx.x & 31u;
	uint32_t acc_count = 0;
	// extension rays leave with the bin they are re-ordered by before the next trace launch (ray_bin)
	const bool binning = rs.sort_mode != 0 && int(pathLength) < rs.max_path_length;
	SortGrid grid{};
	if (binning)
		grid = *wf.grid;
	uint32_t *seen = wf.ext_seen + size_t(bv.index * MAX_DEPTH_SLOTS + pathLength) * MAX_BATCH_SPP;
	// rs.shade_static (setting shade_loop=static): no work cursor — warp w of the grid shades jobs [32 w, 32 w + 32), then the same
	// 32 jobs one grid further on, and so on.  The queue size is known when the launch starts and paths cost about the same, so
	// nothing needs balancing; what goes away is one L2 round trip (the cursor's atomic) in front of every 32 paths of a kernel
	// that waits on memory most of the time, and knowing the next job early lets the warp pull that job's state lines towards L2
	// while it shades this one (prefetch: no register held across the body).
	const uint32_t stride = gridDim.x * blockDim.x;
	uint32_t job = blockIdx.x * blockDim.x + threadIdx.x;
	for (;; job += stride)
	{
		// (issuing the atomic for the NEXT chunk before shading the current one was measured: one more live register, 430
		// instead of 274 bytes of spills, shade 3.13 instead of 3.00 ms per frame — profiles/r02/sweep6)
		if (rs.shade_static)
		{
			if (job - lane >= n_paths)
				break;
			const uint32_t nj = job + stride;
			if (nj < n_paths)
			{
				asm volatile("prefetch.global.L2 [%0];" ::"l"(&wf.hit[nj]));
				asm volatile("prefetch.global.L2 [%0];" ::"l"(&wf.O[buf][nj]));
				asm volatile("prefetch.global.L2 [%0];" ::"l"(&wf.D[buf][nj]));
				if (pathLength != 0)
					asm volatile("prefetch.global.L2 [%0];" ::"l"(&wf.T[buf][nj]));
			}
		}
		else
		{
			uint32_t base = 0;
			if (lane == 0)
				base = atomicAdd(&curc->shade_cursor, 32u);
			base = __shfl_sync(0xffffffffu, base, 0);
			if (base >= n_paths)
				break;
			job = base + lane;
		}

		bool do_acc = false, do_shadow = false, do_ext = false, dead = false;
		V3 accv = mk(0.f);
		V3 aov_albedo = mk(0.f), aov_normal = mk(0.f); // depth 0 only (setting "aov")
		uint32_t pathIndex = 0, pixelIndex = 0, sampleInBatch = 0;
		float4 eO, eD, eT, cO, cD, cE;
		eO = eD = eT = cO = cD = cE = make_float4(0.f, 0.f, 0.f, 0.f);

		if (job < n_paths)
		{
			const float4 hitData = LD_SS(&wf.hit[job]);
			// shade_static: the path's state is requested together with its hit record (one trip to memory instead of two in a
			// row; a padded work item's state is allocated like any other, so the loads need not wait for the test below)
			float4 O4, D4, T4 = make_float4(1.f, 1.f, 1.f, 1.f);
			if (rs.shade_static)
			{
				O4 = LD_SS(&wf.O[buf][job]), D4 = LD_SS(&wf.D[buf][job]);
				if (pathLength != 0)
					T4 = LD_SS(&wf.T[buf][job]);
			}
			const int primIdx = __float_as_int(hitData.z);
			dead = primIdx == PRIM_DEAD;
			if (!dead)
			{
				if (!rs.shade_static)
				{
					O4 = LD_SS(&wf.O[buf][job]), D4 = LD_SS(&wf.D[buf][job]);
					if (pathLength != 0)
						T4 = LD_SS(&wf.T[buf][job]);
				}
				uint32_t flags = __float_as_uint(O4.w) & 0xFFu;
				V3 throughput = mk(T4.x, T4.y, T4.z);
				const float bsdfPdf = T4.w;
				const V3 D = mk(D4.x, D4.y, D4.z);
				pathIndex = __float_as_uint(O4.w) >> 8; // work item the path started as: (local pixel, sample) of the wavefront
				item_to_pixel_sample(bv, pathIndex, pixelIndex, sampleInBatch);
				const uint32_t samplesTaken = fp.sample_base + bv.first_sample + sampleInBatch;
				if (primIdx == PRIM_MISS)
				{
					// Kernels.cu:593-610
					const uint32_t u = __float2uint_rz(float(sc.sky_w) * 0.5f * (1.0f + atan2f(D.x, -D.z) * INVPI));
					const uint32_t v = __float2uint_rz(float(sc.sky_h) * acosf(D.y) * INVPI);
					const uint32_t idx = u + v * sc.sky_w;
					V3 sky = mk(0.f);
					if (idx < sc.sky_h * sc.sky_w)
					{
						const float4 s4 = __ldg(reinterpret_cast<const float4 *>(sc.sky) + idx);
						sky = mk(s4.x, s4.y, s4.z);
					}
					V3 contribution = throughput * (1.0f / bsdfPdf) * sky;
					if (!any_nan(contribution))
					{
						aov_albedo = contribution; // OptiX6 kernels.cu:122-133: the sky as seen, no normal
						clampIntensity(contribution, rs.clamp_value);
						do_acc = true, accv = contribution;
					}
				}
				else
				{
					const V3 O = mk(O4.x, O4.y, O4.z);
					const V3 I = O + D * hitData.w;
					const uint32_t ub = __float_as_uint(hitData.x);
					const uint32_t shadeIdx = __float_as_uint(hitData.y);
					ShadeTri tri;
					{
						const float4 *tp = reinterpret_cast<const float4 *>(sc.shade_tris + shadeIdx);
						float4 *dst = reinterpret_cast<float4 *>(&tri);
#pragma unroll
						for (int k = 0; k < 6; k++)
							dst[k] = __ldg(tp + k);
					}
					if (sc.tl_instances != nullptr) // two-level scene: the record is the mesh's, hit.z names the instance
						apply_instance_normals(sc, uint32_t(primIdx), tri);
					const float bu = float(ub & 65535u) * (1.0f / 65535.0f), bv = float((ub >> 16) & 65535u) * (1.0f / 65535.0f);
					V3 N, iN, T, B;
					const ShadingData sd = getShadingData(sc, D, bu, bv, fp.spread_angle * hitData.w, tri, N, iN, T, B);

					if (pathLength == 0)
					{
						uint32_t px, py;
						local_to_pixel(sh, pixelIndex, px, py);
						if (py * sh.width + px == fp.probe_pixel && sampleInBatch == 0u) // Kernels.cu:626-631
						{
							wf.probe->inst = int(tri.inst_id), wf.probe->prim = int(tri.prim_id), wf.probe->dist = hitData.w;
						}
					}

					if (sd.flags & 1u)
					{
						// alpha cut-out: continue through the surface (Kernels.cu:634-647; throughput plane per
						// VulkanRTX/shaders/rt_shade.comp:155)
						if (int(pathLength) < rs.max_path_length && !any_nan(throughput))
						{
							do_ext = true;
							const V3 no = I + D * rs.geometry_epsilon;
							eO = make_float4(no.x, no.y, no.z, O4.w);
							eD = D4;
							eT = T4;
						}
					}
					else if (sd.isEmissive())
					{
						// Kernels.cu:650-692
						const float DdotNL = -dot(D, N);
						V3 contribution = mk(0.f);
						bool skip = false;
						if (DdotNL > 0)
						{
							if (pathLength == 0)
								contribution = sd.color;
							else if (flags & IS_SPECULAR)
								contribution = throughput * sd.color * (1.0f / bsdfPdf);
							else
							{
								const V3 lastN = UnpackNormal(__float_as_uint(D4.w));
								const float lightPdf = (hitData.w * hitData.w) / (-dot(D, N) * tri.area); // lights.h:78-81
								const float pickProb = LightPickProb(sc, tri.light_tri_idx, O, lastN, I);
								if ((bsdfPdf + lightPdf * pickProb) <= 0)
									skip = true;
								else
									contribution = throughput * sd.color * (1.0f / (bsdfPdf + lightPdf * pickProb));
							}
						}
						if (!skip)
						{
							if (any_nan(contribution))
								contribution = mk(0.f);
							aov_albedo = mk(fminf(contribution.x, 1.0f), fminf(contribution.y, 1.0f), fminf(contribution.z, 1.0f)); // :206-221
							aov_normal = iN;
							clampIntensity(contribution, rs.clamp_value);
							do_acc = true, accv = contribution;
						}
					}
					else
					{
						if (sd.roughness() < MIN_ROUGHNESS)
							flags |= IS_SPECULAR;
						else
							flags &= ~IS_SPECULAR;
						uint32_t gx, gy;
						local_to_pixel(sh, pixelIndex, gx, gy);
						const uint32_t globalPixel = gy * sh.width + gx;
						uint32_t seed = WangHash(globalPixel * 16789u + samplesTaken * 1791u + pathLength * 720898027u);
						const float flip = (dot(D, N) > 0) ? -1.0f : 1.0f;
						N = N * flip;
						iN = iN * flip;
						throughput = throughput * (1.0f / bsdfPdf);
						const uint32_t nlights = sc.lights.area + sc.lights.point + sc.lights.spot + sc.lights.directional;
						// shadow rays emitted at the last depth are never traced by the reference's host loop
						// (CUDART/src/Context.cpp:109-120); skipping them is exact, except that the seed must
						// advance identically when samplesTaken >= 256
						const bool nee = (flags & IS_SPECULAR) == 0 && nlights > 0;
						if (nee)
						{
							float r0, r1;
							if (samplesTaken < 256)
							{
								r0 = blueNoiseSampler(sc.blue_noise, int(gx), int(gy), int(samplesTaken), 4);
								r1 = blueNoiseSampler(sc.blue_noise, int(gx), int(gy), int(samplesTaken), 5);
							}
							else
							{
								r0 = RandomFloat(seed);
								r1 = RandomFloat(seed);
							}
							if (int(pathLength) < rs.max_path_length)
							{
								V3 lightColor = mk(0.f);
								float pickProb = 0, lightPdf = 0;
								V3 L = RandomPointOnLight(sc, r0, r1, I, iN, pickProb, lightPdf, lightColor) - I;
								// The connect ray's direction and length decide whether it stops 2e-5 short of the light it aims at or
								// inside it (one float ulp at Sponza scale is larger than that, DESIGN.md "Epsilons"): they are computed
								// with correctly rounded sqrt / division in the fast-math build too, so both builds and the CPU oracle
								// aim the same ray; everything else of the shade kernel keeps -use_fast_math.
								const float dist = __fsqrt_rn(dot(L, L));
								L = L * __fdiv_rn(1.0f, dist);
								const float NdotL = dot(L, iN);
								if (NdotL > 0 && lightPdf > 0)
								{
									const V3 wo = D * -1.0f;
									const V3 sampledBSDF = BSDFEval(sd, iN, wo, L, 0.0f, false);
									const float shadowPdf = BSDFPdf(sd, iN, wo, L);
									if (shadowPdf > 0)
									{
										V3 contribution = throughput * sampledBSDF * lightColor * (NdotL / (shadowPdf + lightPdf * pickProb));
										clampIntensity(contribution, rs.clamp_value);
										if (!any_nan(contribution))
										{
											do_shadow = true;
											const V3 so = I + N * 1e-5f; // SafeOrigin, tools.h:119-123
											cO = make_float4(so.x, so.y, so.z, 0.f);
											cD = make_float4(L.x, L.y, L.z, dist - 2.0f * rs.geometry_epsilon);
											cE = make_float4(contribution.x, contribution.y, contribution.z, __uint_as_float(pathIndex));
										}
									}
								}
							}
						}
						if (int(pathLength) < rs.max_path_length)
						{
							V3 R = mk(0.f);
							float newBsdfPdf = 0.0f;
							const V3 wo = D * -1.0f;
							const float r3 = RandomFloat(seed);
							const float r4 = RandomFloat(seed);
							BSDFSample(sd, T, B, iN, wo, R, newBsdfPdf, r3, r4);
							const V3 bsdf = BSDFEval(sd, iN, wo, R, hitData.w, flip < 0);
							aov_albedo = sd.color * fabsf(dot(iN, R)), aov_normal = iN; // :316-330
							if (rs.survival_scale) // Kernels.cu:783
								throughput = throughput * 1.0f / SurvivalProbability(throughput) * bsdf * fabsf(dot(iN, R));
							else
								throughput = throughput * bsdf * fabsf(dot(iN, R));
							if (!(newBsdfPdf < 1e-6f || isnan(newBsdfPdf) || throughput.x < 0.0f || throughput.y < 0.0f ||
								  throughput.z < 0.0f))
							{
								do_ext = true;
								const V3 so = I + N * 1e-5f;
								eO = make_float4(so.x, so.y, so.z, __uint_as_float((pathIndex << 8) | flags));
								eD = make_float4(R.x, R.y, R.z, __uint_as_float(PackNormal(iN)));
								eT = make_float4(throughput.x, throughput.y, throughput.z, newBsdfPdf);
							}
						}
					}
				}
			}
		}

		// ---- converged commit: warp-aggregated compaction --------------------------------------
		// accumulator[path] += contribution (Kernels.cu:607,690).  A path owns its slot (one per pixel and sample of the
		// wavefront); the first shade launch of a wavefront visits every live path exactly once and initialises the slot,
		// so no clear is needed between frames.
		if (pathLength == 0)
		{
			if (job < n_paths && !dead)
			{
				ST_SS(&wf.sample_acc[job], make_float4(accv.x, accv.y, accv.z, 0.0f));
				if (wf.sample_albedo != nullptr)
				{
					// a BSDF sample can be NaN (the path is then dropped, Kernels.cu:785); a feature plane must stay finite
					if (any_nan(aov_albedo))
						aov_albedo = mk(0.f);
					if (any_nan(aov_normal))
						aov_normal = mk(0.f);
					const float *m = fp.to_eye;
					wf.sample_albedo[job] = make_float4(aov_albedo.x, aov_albedo.y, aov_albedo.z, 0.0f);
					wf.sample_normal[job] = make_float4(m[0] * aov_normal.x + m[3] * aov_normal.y + m[6] * aov_normal.z,
														m[1] * aov_normal.x + m[4] * aov_normal.y + m[7] * aov_normal.z,
														m[2] * aov_normal.x + m[5] * aov_normal.y + m[8] * aov_normal.z, 0.0f);
				}
			}
		}
		else if (do_acc)
		{
			float4 a = LD_SS(&wf.sample_acc[pathIndex]);
			a.x += accv.x, a.y += accv.y, a.z += accv.z;
			ST_SS(&wf.sample_acc[pathIndex], a);
		}
		if (do_acc)
			acc_count++;
		const uint32_t lt_mask = (1u << lane) - 1u;
		const uint32_t m_sh = __ballot_sync(0xffffffffu, do_shadow);
		const uint32_t m_ex = __ballot_sync(0xffffffffu, do_ext);
		// slots in both queues from ONE atomic: DepthCounters::ext and ::shadow are the two halves of an aligned 64-bit word
		// (neither half can carry into the other: a queue holds fewer than 2^32 entries), so the warp waits for one round
		// trip to L2 instead of two in a row
		uint32_t ebase = 0, sbase = 0;
		if (m_sh | m_ex)
		{
			unsigned long long both = 0ull;
			if (lane == 0)
				both = atomicAdd(reinterpret_cast<unsigned long long *>(&curc->ext),
								 (unsigned long long)(__popc(m_ex)) | ((unsigned long long)(__popc(m_sh)) << 32));
			both = __shfl_sync(0xffffffffu, both, 0);
			ebase = uint32_t(both), sbase = uint32_t(both >> 32);
		}
		if (m_sh)
		{
			if (do_shadow)
			{
				const uint32_t slot = sbase + __popc(m_sh & lt_mask);
				ST_SS(&wf.sO[slot], cO), ST_SS(&wf.sD[slot], cD), ST_SS(&wf.sE[slot], cE);
			}
		}
		if (m_ex)
		{
			if (do_ext)
			{
				const uint32_t slot = ebase + __popc(m_ex & lt_mask);
				ST_SS(&wf.O[nbuf][slot], eO), ST_SS(&wf.D[nbuf][slot], eD), ST_SS(&wf.T[nbuf][slot], eT);
				if (seen[sampleInBatch] == 0u) // this sample's bounce loop continues (CUDART/src/Context.cpp:109-120)
				{
					seen[sampleInBatch] = 1u;
					if (sync.seen != nullptr) // tell the other ranks of a sharded frame (a handful of peer stores per sample and depth)
						*reinterpret_cast<volatile uint32_t *>(sync.seen + pathLength * MAX_BATCH_SPP + sampleInBatch) = sync.stamp;
				}
				if (binning)
				{
					// counting sort, first half: the lanes of the warp that emit into the same bin are found with one
					// match instruction and take consecutive ranks from ONE atomic on the bin's counter
					const uint32_t bin = ray_bin(grid, rs.sort_cell_bits, rs.sort_dir_major, rs.sort_dir_bits, eO.x, eO.y, eO.z, eD.x, eD.y, eD.z);
					const uint32_t peers = __match_any_sync(m_ex, bin);
					const int leader = __ffs(peers) - 1;
					uint32_t rank = 0;
					if (int(lane) == leader)
						rank = atomicAdd(&wf.sort_hist[bin], uint32_t(__popc(peers)));
					rank = __shfl_sync(m_ex, rank, leader) + __popc(peers & lt_mask);
					ST_SS(&wf.sort_key[slot], make_uint2(bin, rank));
				}
			}
		}
	}
	acc_count = __reduce_add_sync(0xffffffffu, acc_count);
	if (lane == 0 && acc_count)
		atomicAdd(&curc->acc, acc_count);
}

#endif // RFW_PART != 1 (shade part)

#if RFW_PART != 2 // ---- trace part ----
// ------------------------------------------------------------------------------------------------
// k_fold — Kernels.cu:181-203: accumulator += the wavefront's samples (in sample order, so a frame does not depend on how
// its samples were split into wavefronts or render_frame calls), framebuffer = accumulator * 1/samples.  world == 1
// writes the row-major image; world > 1 keeps the tile-major shard for the gather.
// ------------------------------------------------------------------------------------------------
__global__ void k_fold(const ShardView sh, const WavefrontView wf, const BatchView bv, const float scale, const int write_fb,
					   const DisplayTarget dt)
{
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j < sh.local_pixels)
	{
		uint32_t x, y;
		const bool live = local_to_pixel(sh, j, x, y);
		float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
		if (live)
		{
			a = wf.accumulator[j];
			for (uint32_t s = 0; s < bv.spp; s++)
			{
				const float4 b = wf.sample_acc[pixel_sample_to_item(bv, j, s)];
				a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
			}
			wf.accumulator[j] = a;
			if (wf.sample_albedo != nullptr)
			{
				float4 al = wf.albedo_acc[j], no = wf.normal_acc[j];
				for (uint32_t s = 0; s < bv.spp; s++)
				{
					const uint32_t it = pixel_sample_to_item(bv, j, s);
					const float4 b = wf.sample_albedo[it], c = wf.sample_normal[it];
					al.x += b.x, al.y += b.y, al.z += b.z, no.x += c.x, no.y += c.y, no.z += c.z;
				}
				wf.albedo_acc[j] = al, wf.normal_acc[j] = no;
			}
		}
		if (write_fb)
		{
			a.x *= scale, a.y *= scale, a.z *= scale, a.w *= scale;
			if (sh.world == 1)
			{
				if (live)
					wf.framebuffer[size_t(y) * sh.width + x] = a;
			}
			else
				wf.framebuffer[j] = a;
			if (dt.image != nullptr && live)
				dt.image[size_t(y) * sh.width + x] = a; // the display rank's image: a peer store over NVLink on the other ranks
		}
	}
	if (write_fb && dt.image != nullptr)
	{
		// the last CTA of the launch tells the display rank that this rank's tiles of the frame are in place
		__threadfence_system();
		__syncthreads();
		if (threadIdx.x == 0)
		{
			const uint32_t done = atomicAdd(dt.local_done, 1u);
			if (done == gridDim.x - 1u)
			{
				*dt.local_done = 0u;
				__threadfence_system();
				atomicAdd_system(dt.arrivals, 1u);
			}
		}
	}
}

// Flow control of the display image, one thread each.  Counters only grow: frame f is complete when arrivals reaches
// f * world; a rank may overwrite the image with frame f + 1 once the display rank has released frame f (consumed >= f).
// Every wait gives up after four seconds and raises *err instead of hanging the device.
__global__ void k_display_spin(const volatile uint32_t *counter, const uint32_t need, uint32_t *err)
{
	unsigned long long t0, t1;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
	while (*counter < need)
	{
		__nanosleep(200);
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
		if (t1 - t0 > 4000000000ull)
		{
			*err = 1u;
			break;
		}
	}
	__threadfence_system();
}
// Between shade(depth) and trace(depth + 1) of a sharded frame: this rank arrives (its own flags were published by the
// shade launch) and takes over the merged "sample s emitted an extension ray" flags (ShardSync).  It only has to WAIT for
// the other ranks' shade launches when one of its own flags is still clear — a set flag cannot be unset by anybody — so
// in an ordinary frame the ranks do not run in lockstep.  `need_lagged` (arrivals of an earlier sync) is waited for in
// any case: it bounds how far a rank may run ahead, so that the flag rows of the wavefront after next (same rows, see
// DISPLAY_TAIL_BYTES) are not written while a slower rank still reads this wavefront's.
__global__ void k_shard_sync(const ShardSync sync, const uint32_t depth, const uint32_t need_full, const uint32_t need_lagged,
							 const uint32_t spp, uint32_t *local_seen_row, uint32_t *err)
{
	__shared__ uint32_t any_clear;
	if (threadIdx.x == 0)
		any_clear = 0u;
	__syncthreads();
	if (threadIdx.x < spp && local_seen_row[threadIdx.x] == 0u)
		any_clear = 1u;
	__syncthreads();
	if (threadIdx.x == 0)
	{
		__threadfence_system();
		atomicAdd_system(sync.arrivals, 1u);
		const uint32_t need = any_clear ? need_full : need_lagged;
		unsigned long long t0, t1;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
		while (*reinterpret_cast<const volatile uint32_t *>(sync.arrivals) < need)
		{
			__nanosleep(200);
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
			if (t1 - t0 > 4000000000ull)
			{
				*err = 1u; // the frame goes on with this rank's own flags
				break;
			}
		}
		__threadfence_system();
	}
	__syncthreads();
	if (any_clear && threadIdx.x < spp)
	{
		const uint32_t g = *reinterpret_cast<const volatile uint32_t *>(sync.seen + depth * MAX_BATCH_SPP + threadIdx.x);
		if (g == sync.stamp)
			local_seen_row[threadIdx.x] = 1u;
	}
}

__global__ void k_display_release(uint32_t *consumed, const uint32_t value)
{
	__threadfence_system();
	*reinterpret_cast<volatile uint32_t *>(consumed) = value;
	__threadfence_system();
}

// ------------------------------------------------------------------------------------------------
// Re-ordering of the bounce queue: a counting sort whose first half (bin + rank inside the bin, one warp-aggregated
// atomic per bin and warp) is done by k_shade while it appends the rays.
//   k_sort_setup  the grid the bins are cells of, from the root of the current tree (runs once per frame: one warp)
//   k_sort_scan   exclusive prefix of the bin counts inside 4096-bin chunks + the chunk totals; clears the counts
//   k_sort_move   ray i of the staging queue goes to position chunk_prefix + base[bin] + rank of the trace queue
// ------------------------------------------------------------------------------------------------
__global__ void k_sort_setup(const SceneView sc, SortGrid *grid, const int cell_bits)
{
	if (threadIdx.x != 0 || blockIdx.x != 0)
		return;
	float lo[3] = {0.f, 0.f, 0.f}, hi[3] = {1.f, 1.f, 1.f};
	if (sc.cw_nodes != nullptr && sc.cw_node_count > 0)
	{
		// compressed 8-wide root: grid origin p and step 2^(e - 127) per axis, 256 steps span the node
		const uint4 q0 = sc.cw_nodes[0];
		const float p[3] = {__uint_as_float(q0.x), __uint_as_float(q0.y), __uint_as_float(q0.z)};
		for (int a = 0; a < 3; a++)
		{
			const uint32_t e = (q0.w >> (8 * a)) & 255u;
			lo[a] = p[a], hi[a] = p[a] + 256.0f * __uint_as_float(e << 23);
		}
	}
	else if (sc.node_count > 0)
	{
		const BvhNode4 &n = sc.nodes[0];
		for (int a = 0; a < 3; a++)
			lo[a] = 3e38f, hi[a] = -3e38f;
		for (int k = 0; k < 4; k++) // fminf / fmaxf drop the NaN boxes of unused slots
		{
			lo[0] = fminf(lo[0], n.minx[k]), lo[1] = fminf(lo[1], n.miny[k]), lo[2] = fminf(lo[2], n.minz[k]);
			hi[0] = fmaxf(hi[0], n.maxx[k]), hi[1] = fmaxf(hi[1], n.maxy[k]), hi[2] = fmaxf(hi[2], n.maxz[k]);
		}
	}
	for (int a = 0; a < 3; a++)
	{
		const float ext = hi[a] - lo[a];
		grid->lo[a] = lo[a];
		grid->scale[a] = (ext > 0.0f && ext < 3e38f) ? float(1 << cell_bits) / ext : 0.0f;
	}
}

__global__ void __launch_bounds__(1024) k_sort_scan(const WavefrontView wf)
{
	__shared__ uint32_t warp_sum[32];
	const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
	uint4 *hist = reinterpret_cast<uint4 *>(wf.sort_hist + size_t(blockIdx.x) * SORT_CHUNK) + t;
	const uint4 c = *hist;
	*hist = make_uint4(0u, 0u, 0u, 0u);
	const uint32_t mine = c.x + c.y + c.z + c.w;
	uint32_t incl = mine;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
		if (int(lane) >= o)
			incl += v;
	}
	if (lane == 31u)
		warp_sum[warp] = incl;
	__syncthreads();
	if (warp == 0)
	{
		const uint32_t w = warp_sum[lane];
		uint32_t wi = w;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
			if (int(lane) >= o)
				wi += v;
		}
		warp_sum[lane] = wi - w; // exclusive
		if (lane == 31u)
			wf.sort_chunk[blockIdx.x] = wi;
	}
	__syncthreads();
	const uint32_t excl = warp_sum[warp] + incl - mine;
	reinterpret_cast<uint4 *>(wf.sort_base + size_t(blockIdx.x) * SORT_CHUNK)[t] = make_uint4(excl, excl + c.x, excl + c.x + c.y, excl + c.x + c.y + c.z);
}

__global__ void __launch_bounds__(256) k_sort_move(const WavefrontView wf, const BatchView bv, const uint32_t depth, const uint32_t n_chunks)
{
	// prefix of the chunk totals (at most 2^21 bins / 4096 = 512 chunks), recomputed by every CTA: cheaper than a third launch
	__shared__ uint32_t chunk_prefix[512];
	if (threadIdx.x < 32)
	{
		uint32_t carry = 0;
		for (uint32_t b0 = 0; b0 < n_chunks; b0 += 32)
		{
			const uint32_t i = b0 + threadIdx.x;
			const uint32_t v = i < n_chunks ? wf.sort_chunk[i] : 0u;
			uint32_t incl = v;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
				if (int(threadIdx.x) >= o)
					incl += u;
			}
			if (i < n_chunks)
				chunk_prefix[i] = carry + incl - v;
			carry += __shfl_sync(0xffffffffu, incl, 31);
		}
	}
	__syncthreads();
	DepthCounters *row = &wf.counters[bv.index * MAX_DEPTH_SLOTS];
	const uint32_t n = row[depth - 1].ext;
	const float4 *__restrict__ sO = wf.O[1], *__restrict__ sD = wf.D[1], *__restrict__ sT = wf.T[1];
	// two rays per thread and iteration: all eight loads of a pair are in flight before the first dependent store (the pass is
	// a latency-bound scatter: 3.7 TB/s with one ray per iteration)
	const uint32_t stride = gridDim.x * blockDim.x;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 2u * stride)
	{
		const uint32_t j = i + stride;
		const bool two = j < n;
		const uint2 kr0 = LD_SS(&wf.sort_key[i]);
		const uint2 kr1 = two ? LD_SS(&wf.sort_key[j]) : make_uint2(0u, 0u);
		const float4 o0 = LD_SS(&sO[i]), d0 = LD_SS(&sD[i]), t0 = LD_SS(&sT[i]);
		float4 o1 = o0, d1 = d0, t1 = t0;
		if (two)
			o1 = LD_SS(&sO[j]), d1 = LD_SS(&sD[j]), t1 = LD_SS(&sT[j]);
		const uint32_t dst0 = chunk_prefix[kr0.x / SORT_CHUNK] + wf.sort_base[kr0.x] + kr0.y;
		ST_SS(&wf.O[0][dst0], o0), ST_SS(&wf.D[0][dst0], d0), ST_SS(&wf.T[0][dst0], t0);
		if (two)
		{
			const uint32_t dst1 = chunk_prefix[kr1.x / SORT_CHUNK] + wf.sort_base[kr1.x] + kr1.y;
			ST_SS(&wf.O[0][dst1], o1), ST_SS(&wf.D[0][dst1], d1), ST_SS(&wf.T[0][dst1], t1);
		}
	}
}

__global__ void k_assemble(const ShardView sh, const float4 *__restrict__ gathered, const size_t stride,
						   float4 *__restrict__ image)
{
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t r = blockIdx.y;
	ShardView s = sh;
	s.rank = r;
	const uint32_t total_tiles = s.tiles_x * s.tiles_y;
	const uint32_t ltiles = (total_tiles > r) ? (total_tiles - r + s.world - 1) / s.world : 0;
	if (j >= ltiles * s.tile_w * s.tile_h)
		return;
	uint32_t x, y;
	if (local_to_pixel(s, j, x, y))
		image[size_t(y) * s.width + x] = gathered[size_t(r) * stride + j];
}

// feature plane / samples, row-major for a single shard (tile-major otherwise, like the framebuffer)
__global__ void k_aov_finalize(const ShardView sh, const float4 *__restrict__ acc, const float scale, float4 *__restrict__ out)
{
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= sh.local_pixels)
		return;
	uint32_t x, y;
	const bool live = local_to_pixel(sh, j, x, y);
	float4 a = live ? acc[j] : make_float4(0.f, 0.f, 0.f, 0.f);
	a.x *= scale, a.y *= scale, a.z *= scale, a.w = 0.f;
	if (sh.world == 1)
	{
		if (live)
			out[size_t(y) * sh.width + x] = a;
	}
	else
		out[j] = a;
}

// ------------------------------------------------------------------------------------------------
// k_tone_map — the display pass of rfw::system::render_frame(camera, status, toneMap = true) (system.cpp:694-713), i.e.
// assets/shaders/tone-map.frag: rgb' = ACESFitted(max(0, rgb - 0.5 * contrast + 0.5 + brightness)), alpha passed through,
// written as RGBA8 (round to nearest, the UNORM conversion of an 8-bit GL target) so a display consumer reads back a
// quarter of the float framebuffer.  Every operation is written out (no contraction: __fmul_rn / __fadd_rn) in the
// order of the shader, so a float32 restatement of the shader reproduces the bytes exactly (tests).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float rrt_odt_fit(float v)
{
	const float a = __fadd_rn(__fmul_rn(v, __fadd_rn(v, 0.0245786f)), -0.000090537f);
	const float b = __fadd_rn(__fmul_rn(v, __fadd_rn(__fmul_rn(0.983729f, v), 0.4329510f)), 0.238081f);
	return __fdiv_rn(a, b);
}
__device__ __forceinline__ float mat_row(float m0, float m1, float m2, float x, float y, float z)
{
	return __fadd_rn(__fadd_rn(__fmul_rn(m0, x), __fmul_rn(m1, y)), __fmul_rn(m2, z)); // column sum c0*x + c1*y + c2*z
}
__device__ __forceinline__ uint32_t unorm8(float v)
{
	v = fminf(fmaxf(v, 0.0f), 1.0f);
	return uint32_t(__fadd_rn(__fmul_rn(v, 255.0f), 0.5f));
}
__global__ void k_tone_map(const float4 *__restrict__ fb, uint32_t *__restrict__ out, const uint32_t n, const float contrast,
						   const float brightness)
{
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n)
		return;
	const float4 c = fb[j];
	const float half_contrast = __fmul_rn(0.5f, contrast);
	const float r = fmaxf(0.0f, __fadd_rn(__fadd_rn(__fadd_rn(c.x, -half_contrast), 0.5f), brightness));
	const float g = fmaxf(0.0f, __fadd_rn(__fadd_rn(__fadd_rn(c.y, -half_contrast), 0.5f), brightness));
	const float b = fmaxf(0.0f, __fadd_rn(__fadd_rn(__fadd_rn(c.z, -half_contrast), 0.5f), brightness));
	// ACESInputMat (columns as written in the shader), RRT+ODT fit, ACESOutputMat
	const float ir = rrt_odt_fit(mat_row(0.59719f, 0.35458f, 0.04823f, r, g, b));
	const float ig = rrt_odt_fit(mat_row(0.07600f, 0.90834f, 0.01566f, r, g, b));
	const float ib = rrt_odt_fit(mat_row(0.02840f, 0.13383f, 0.83777f, r, g, b));
	const float orr = mat_row(1.60475f, -0.53108f, -0.07367f, ir, ig, ib);
	const float og = mat_row(-0.10208f, 1.10813f, -0.00605f, ir, ig, ib);
	const float ob = mat_row(-0.00327f, -0.07276f, 1.07602f, ir, ig, ib);
	out[j] = unorm8(orr) | (unorm8(og) << 8) | (unorm8(ob) << 16) | (unorm8(c.w) << 24);
}

// ------------------------------------------------------------------------------------------------
// k_emode — EmbreeRT/src/Context.cpp:104-300 + retrieve_material :417-476, one thread per pixel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 3) k_emode(const SceneView sc, const ShardView sh, const WavefrontView wf,
											  const RenderSettings rs, const rfwb200_material *__restrict__ raw_materials,
											  const uint32_t *__restrict__ tex_desc, const uint32_t tex_count)
{
	__shared__ uint64_t mbar;
	float4 *snodes = reinterpret_cast<float4 *>(g_dyn_smem);
	const uint32_t n_smem = min(uint32_t(rs.smem_nodes), sc.node_count);
	stage_nodes(snodes, sc.nodes, n_smem, &mbar);
	const FrameParams fp = *wf.frame;
	uint32_t *cursor = &wf.counters[0].trace_cursor;
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t total = sh.local_pixels;
	// only whole 4x2 tiles are rendered by the reference (Context.cpp:137-139)
	const uint32_t wlim = (sh.width / 4u) * 4u, hlim = (sh.height / 2u) * 2u;
	for (;;)
	{
		uint32_t base = 0;
		if (lane == 0)
			base = atomicAdd(cursor, 32u);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (base >= total)
			break;
		const uint32_t j = base + lane;
		uint32_t x, y;
		if (j >= total || !local_to_pixel(sh, j, x, y))
			continue;
		float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
		if (x < wlim && y < hlim)
		{
			V3 O, D;
			generate_emode(fp, sh, x, y, fp.sample_base, O, D);
			float t = 1e34f, u = 0.f, v = 0.f;
			uint32_t triIdx = 0, hitInst = 0;
			if (!traverse<false>(sc, snodes, n_smem, O, D, 1e-5f, t, triIdx, u, v, hitInst))
			{
				const float su = 0.5f * (1.0f + atan2f(D.x, -D.z) * INVPI);
				const float sv = acosf(D.y) * INVPI;
				const uint32_t px = __float2uint_rz(su * float(sc.sky_w - 1)), py = __float2uint_rz(sv * float(sc.sky_h - 1));
				const size_t si = size_t(py) * sc.sky_w + px;
				if (si < size_t(sc.sky_w) * sc.sky_h)
				{
					const float4 s4 = __ldg(reinterpret_cast<const float4 *>(sc.sky) + si);
					out = make_float4(s4.x, s4.y, s4.z, 0.f);
				}
			}
			else
			{
				ShadeTri tri;
				{
					const float4 *tp = reinterpret_cast<const float4 *>(sc.shade_tris + triIdx);
					float4 *dst = reinterpret_cast<float4 *>(&tri);
#pragma unroll
					for (int k = 0; k < 6; k++)
						dst[k] = __ldg(tp + k);
				}
				if (sc.tl_instances != nullptr)
					apply_instance_normals(sc, hitInst, tri);
				if (y * sh.width + x == fp.probe_pixel)
					wf.probe->inst = int(tri.inst_id), wf.probe->prim = int(tri.prim_id), wf.probe->dist = t;
				const float b0 = 1.0f - u - v;
				const V3 p = O + D * t;
				const V3 iN = normalize(mk(tri.n0x, tri.n0y, tri.n0z) * b0 + mk(tri.n1x, tri.n1y, tri.n1z) * u + mk(tri.n2x, tri.n2y, tri.n2z) * v);
				const rfwb200_material *mat = raw_materials + tri.material;
				const uint4 basew = __ldg(reinterpret_cast<const uint4 *>(mat));
				const uint32_t flags = basew.w;
				V3 color = mk(__half2float(__ushort_as_half((unsigned short)(basew.x & 0xffffu))),
							  __half2float(__ushort_as_half((unsigned short)(basew.x >> 16))),
							  __half2float(__ushort_as_half((unsigned short)(basew.y & 0xffffu))));
				if (has_flag(flags, HasDiffuseMap))
				{
					const MapDesc m = load_map(&mat->tex0);
					if (m.addr < tex_count)
					{
						const float tu = b0 * tri.u0 + u * tri.u1 + v * tri.u2;
						const float tv = b0 * tri.v0 + u * tri.v1 + v * tri.v2;
						const float uu = (tu + m.uo) * m.us, vv2 = (tv + m.vo) * m.vs;
						float tx = fmodf(uu, 1.0f), ty = fmodf(vv2, 1.0f);
						if (tx < 0.f)
							tx = 1.f + tx;
						if (ty < 0.f)
							ty = 1.f + ty;
						const uint32_t ttype = tex_desc[m.addr * 4 + 0], tw = tex_desc[m.addr * 4 + 1], th = tex_desc[m.addr * 4 + 2],
									   taddr = tex_desc[m.addr * 4 + 3];
						const uint32_t ix = __float2uint_rz(tx * float(tw - 1)), iy = __float2uint_rz(ty * float(th - 1));
						const uint32_t id = iy * tw + ix;
						constexpr float tsc = 1.0f / 256.0f;
						if (ttype == RFWB200_TEX_UINT)
						{
							const uint32_t tc = __ldg(sc.uint_texels + taddr + id);
							color = color * tsc * mk(float(tc & 0xFFu), float((tc >> 8) & 0xFFu), float((tc >> 16) & 0xFFu));
						}
						else
						{
							// FLOAT4 falls through into the UINT case in the reference (Context.cpp:458-472): the texel is
							// applied, then the 32-bit word at index `id` of the SAME buffer read as uints — float number id
							// of the texture, not texel id — is applied as RGBA8
							const float4 tf = __ldg(reinterpret_cast<const float4 *>(sc.float_texels) + taddr + id);
							color = color * mk(tf.x, tf.y, tf.z);
							const uint32_t tc = __ldg(reinterpret_cast<const uint32_t *>(reinterpret_cast<const float4 *>(sc.float_texels) + taddr) + id);
							color = color * tsc * mk(float(tc & 0xFFu), float((tc >> 8) & 0xFFu), float((tc >> 16) & 0xFFu));
						}
					}
				}
				if (color.x > 1 || color.y > 1 || color.z > 1)
					out = make_float4(color.x, color.y, color.z, 1.0f);
				else
				{
					V3 contrib = mk(0.1f);
					for (uint32_t li = 0; li < sc.lights.area; li++)
					{
						const float4 *lp = reinterpret_cast<const float4 *>(sc.area_lights) + size_t(li) * 6;
						const float4 la = __ldg(lp), lb = __ldg(lp + 1), lc = __ldg(lp + 2);
						V3 L = mk(la.x, la.y, la.z) - p;
						const float sq_dist = dot(L, L);
						const float dist = sqrtf(sq_dist);
						L = L / dist;
						const float NdotL = dot(iN, L);
						const float LNdotL = -dot(mk(lb.x, lb.y, lb.z), L);
						if (NdotL <= 0 || LNdotL <= 0)
							continue;
						float tmax = dist * (1.0f - 1e-4f), uu, vv2;
						uint32_t tt;
						if (!traverse<true>(sc, snodes, n_smem, p, L, 1e-4f, tmax, tt, uu, vv2))
							contrib = contrib + mk(lc.x, lc.y, lc.z) * lb.w / sq_dist * NdotL * LNdotL;
					}
					for (uint32_t li = 0; li < sc.lights.point; li++)
					{
						const float4 *lp = reinterpret_cast<const float4 *>(sc.point_lights) + size_t(li) * 2;
						const float4 la = __ldg(lp), lb = __ldg(lp + 1);
						V3 L = mk(la.x, la.y, la.z) - p;
						const float sq_dist = dot(L, L);
						const float dist = sqrtf(sq_dist);
						L = L / dist;
						const float NdotL = dot(iN, L);
						if (NdotL <= 0)
							continue;
						float tmax = dist * (1.0f - 1e-4f), uu, vv2;
						uint32_t tt;
						if (!traverse<true>(sc, snodes, n_smem, p, L, 1e-4f, tmax, tt, uu, vv2))
							contrib = contrib + mk(lb.x, lb.y, lb.z) / sq_dist * NdotL;
					}
					const V3 c = color * contrib;
					out = make_float4(c.x, c.y, c.z, 1.0f);
				}
			}
		}
		if (sh.world == 1)
			wf.framebuffer[size_t(y) * sh.width + x] = out;
		else
			wf.framebuffer[j] = out;
	}
}

// ------------------------------------------------------------------------------------------------
// stage-level kernels on caller rays
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_trace_closest(const SceneView sc, const RenderSettings rs,
													  const float4 *__restrict__ origins, const float4 *__restrict__ dirs,
													  const uint32_t n, const float t_min, float4 *__restrict__ hits,
													  uint32_t *cursor, uint32_t *__restrict__ inst_out)
{
	__shared__ uint64_t mbar;
	float4 *snodes = reinterpret_cast<float4 *>(g_dyn_smem);
	const uint32_t n_smem = min(uint32_t(rs.smem_nodes), sc.node_count);
	stage_nodes(snodes, sc.nodes, n_smem, &mbar);
	const uint32_t lane = threadIdx.x & 31u;
	for (;;)
	{
		uint32_t base = 0;
		if (lane == 0)
			base = atomicAdd(cursor, 32u);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (base >= n)
			break;
		const uint32_t i = base + lane;
		if (i >= n)
			continue;
		const float4 O4 = origins[i], D4 = dirs[i];
		float t = 1e34f, u = 0.f, v = 0.f;
		uint32_t tri = 0xffffffffu, inst = 0;
		const bool h = traverse<false>(sc, snodes, n_smem, mk(O4.x, O4.y, O4.z), mk(D4.x, D4.y, D4.z), t_min, t, tri, u, v, inst);
		hits[i] = make_float4(h ? t : 1e34f, u, v, __uint_as_float(h ? tri : 0xffffffffu));
		if (inst_out != nullptr) // two-level scene: the instance the hit was found in
			inst_out[i] = inst;
	}
}

__global__ void __launch_bounds__(256) k_trace_occluded(const SceneView sc, const RenderSettings rs,
													   const float4 *__restrict__ origins,
													   const float4 *__restrict__ dirs_tmax, const uint32_t n,
													   const float t_min, uint8_t *__restrict__ occluded, uint32_t *cursor)
{
	__shared__ uint64_t mbar;
	float4 *snodes = reinterpret_cast<float4 *>(g_dyn_smem);
	const uint32_t n_smem = min(uint32_t(rs.smem_nodes), sc.node_count);
	stage_nodes(snodes, sc.nodes, n_smem, &mbar);
	const uint32_t lane = threadIdx.x & 31u;
	for (;;)
	{
		uint32_t base = 0;
		if (lane == 0)
			base = atomicAdd(cursor, 32u);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (base >= n)
			break;
		const uint32_t i = base + lane;
		if (i >= n)
			continue;
		const float4 O4 = origins[i], D4 = dirs_tmax[i];
		float tmax = D4.w, u, v;
		uint32_t tri;
		occluded[i] = traverse<true>(sc, snodes, n_smem, mk(O4.x, O4.y, O4.z), mk(D4.x, D4.y, D4.z), t_min, tmax, tri, u, v) ? 1 : 0;
	}
}

__global__ void k_generate_only(const SceneView sc, const ShardView sh, const WavefrontView wf, const uint32_t sample_index,
								const int emode, float4 *__restrict__ origins, float4 *__restrict__ dirs)
{
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= sh.local_pixels)
		return;
	uint32_t x, y;
	if (!local_to_pixel(sh, j, x, y))
		return;
	const FrameParams fp = *wf.frame;
	V3 O, D;
	if (emode)
		generate_emode(fp, sh, x, y, sample_index, O, D);
	else
		generate_pt(sc, fp, sh, x, y, sample_index, O, D);
	const uint32_t pixel = y * sh.width + x;
	origins[pixel] = make_float4(O.x, O.y, O.z, __uint_as_float((pixel << 8) + 1u));
	dirs[pixel] = make_float4(D.x, D.y, D.z, 0.f);
}

#endif // RFW_PART != 2
} // namespace

#if RFW_PART != 2 // ---- trace part: launchers ----
// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
cudaError_t shade_preload_fast();
cudaError_t shade_preload_ieee();
static cudaError_t shade_preload()
{
	const cudaError_t e = shade_preload_fast();
	return e != cudaSuccess ? e : shade_preload_ieee();
}

cudaError_t configure_launches(const RenderSettings &rs, uint32_t node_count, LaunchDims &dims)
{
	int dev = 0, sms = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess)
		return e;
	e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	if (e != cudaSuccess)
		return e;
	const uint32_t staged = rs.smem_nodes > 0 ? (uint32_t(rs.smem_nodes) < node_count ? uint32_t(rs.smem_nodes) : node_count) : 0u;
	dims.trace_smem = size_t(staged) * sizeof(BvhNode4);
	const void *trace_kernels[] = {(const void *)k_wavefront_trace<true, 1, 0, false>, (const void *)k_wavefront_trace<false, 1, 0, false>,
								   (const void *)k_emode, (const void *)k_trace_closest, (const void *)k_trace_occluded};
	for (const void *k : trace_kernels)
	{
		e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(dims.trace_smem));
		if (e != cudaSuccess)
			return e;
	}
	// kernels that use no shared memory ask for the whole unified array as L1 (traversal is bound by L1 hits on nodes)
	if (getenv("RFWB200_NO_CARVEOUT") == nullptr)
	{
		const void *l1_kernels[] = {(const void *)k_wavefront_trace<true, 2, 0, false>,  (const void *)k_wavefront_trace<false, 2, 0, false>,
									(const void *)k_wavefront_trace<true, 1, 1, false>,  (const void *)k_wavefront_trace<false, 1, 1, false>,
									(const void *)k_wavefront_trace<true, 1, 2, false>,  (const void *)k_wavefront_trace<false, 1, 2, false>,
									(const void *)k_wavefront_trace<true, 1, 0, true>,	 (const void *)k_wavefront_trace<false, 1, 0, true>,
									(const void *)k_wavefront_trace<true, 1, 2, true>,	 (const void *)k_wavefront_trace<false, 1, 2, true>,
									(const void *)k_wavefront_trace<true, 1, 1, true>,	 (const void *)k_wavefront_trace<false, 1, 1, true>,
									(const void *)k_wavefront_trace<true, 1, 3, true>,	 (const void *)k_wavefront_trace<false, 1, 3, true>,
									(const void *)k_wavefront_trace<true, 1, 4, true>,	 (const void *)k_wavefront_trace<false, 1, 4, true>,
									(const void *)k_wavefront_trace_cw<true>,			 (const void *)k_wavefront_trace_cw<false>,
									(const void *)k_wavefront_trace_tl<true>,			 (const void *)k_wavefront_trace_tl<false>,
									(const void *)k_wavefront_trace<true, 1, 2, true, false, true>, (const void *)k_wavefront_trace<false, 1, 2, true, false, true>};
		for (const void *k : l1_kernels)
		{
			e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
			if (e != cudaSuccess)
				return e;
		}
	}
	// Load every kernel of the frame now.  With CUDA's lazy module loading the first launch of a kernel may have to wait
	// for the device to drain, and the ranks of a sharded frame wait for each other INSIDE kernels (k_display_spin,
	// k_shard_sync): a first frame would sit in that wait until its time-out.
	{
		const void *all[] = {(const void *)k_fold,		   (const void *)k_sort_setup,		(const void *)k_sort_scan,	  (const void *)k_sort_move,
							 (const void *)k_display_spin, (const void *)k_display_release, (const void *)k_shard_sync, (const void *)k_assemble,
							 (const void *)k_tone_map,	   (const void *)k_generate_only,	(const void *)k_emode,		  (const void *)k_trace_closest,
							 (const void *)k_trace_occluded};
		cudaFuncAttributes fa;
		for (const void *k : all)
			if ((e = cudaFuncGetAttributes(&fa, k)) != cudaSuccess)
				return e;
		if ((e = shade_preload()) != cudaSuccess)
			return e;
	}
	{
		dims.trace_smem16 = size_t(staged) * 80;
		const void *ks[] = {(const void *)k_wavefront_trace<true, 1, 2, true, true>, (const void *)k_wavefront_trace<false, 1, 2, true, true>};
		for (const void *k : ks)
			if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(dims.trace_smem16))) != cudaSuccess)
				return e;
		int p16 = 0;
		e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p16, k_wavefront_trace<false, 1, 2, true, true>, dims.trace_block, dims.trace_smem16);
		if (e != cudaSuccess)
			return e;
		dims.trace_grid_staged16 = sms * (p16 < 1 ? 1 : p16);
	}
	int per_sm = 0;
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wavefront_trace<false, 1, 0, false>, dims.trace_block, dims.trace_smem);
	if (e != cudaSuccess)
		return e;
	if (per_sm < 1)
		per_sm = 1;
	dims.trace_grid = sms * per_sm; // a whole number of CTAs per SM: no partial wave
	int shade_per_sm = 0;
	e = shade_occupancy(dims.shade_block, &shade_per_sm);
	if (e != cudaSuccess)
		return e;
	if (shade_per_sm < 1)
		shade_per_sm = 1;
	dims.shade_grid = sms * shade_per_sm;
	dims.move_grid = sms * 8; // k_sort_move: 8 CTAs of 256 threads per SM, grid-stride over the device-resident queue size
	return cudaSuccess;
}

cudaError_t launch_primary(const SceneView &sc, const ShardView &sh, const WavefrontView &wf, const RenderSettings &rs,
						   const BatchView &bv, const LaunchDims &dims, cudaStream_t stream)
{
	if (sc.tl_instances != nullptr)
	{
		// two-level scene: the while-while kernel with the instance step (packed nodes), or — primary_variant 0 — the plain
		// one-ray-per-lane kernel around traverse_tl
		if (rs.primary_variant == 0 || sc.nodes16 == nullptr)
			k_wavefront_trace_tl<true><<<dims.trace_grid, dims.trace_block, 0, stream>>>(sc, sh, wf, rs, bv, 0u, 0u);
		else
			k_wavefront_trace<true, 1, 2, true, false, true><<<dims.trace_grid, dims.trace_block, 0, stream>>>(sc, sh, wf, rs, bv, 0u, 0u);
		return cudaGetLastError();
	}
	if (sc.cw_nodes != nullptr)
	{
		k_wavefront_trace_cw<true><<<dims.trace_grid, dims.trace_block, 0, stream>>>(sc, sh, wf, rs, bv, 0u, 0u);
		return cudaGetLastError();
	}
	switch ((rs.smem_nodes > 0 && rs.primary_variant != 13) ? 0 : rs.primary_variant) // the staged fp32 prefix only exists in variant 0, the staged packed one in 13
	{
#define RFW_TRACE_CASE(V, LQ, LEAN, PACKED)                                                                                     \
	case V:                                                                                                             \
		k_wavefront_trace<true, LQ, LEAN, PACKED><<<dims.trace_grid, dims.trace_block, (LQ == 1 && LEAN == 0 && !PACKED) ? dims.trace_smem : 0, stream>>>( \
			sc, sh, wf, rs, bv, 0u, 0u);                                                                                \
		break;
		RFW_TRACE_CASE(1, 2, 0, false)
		RFW_TRACE_CASE(3, 1, 1, false)
		RFW_TRACE_CASE(5, 1, 2, false)
		RFW_TRACE_CASE(8, 1, 0, true)
		RFW_TRACE_CASE(9, 1, 2, true)
		RFW_TRACE_CASE(10, 1, 1, true)
		RFW_TRACE_CASE(11, 1, 3, true)
		RFW_TRACE_CASE(12, 1, 4, true)
	case 13: // packed nodes with the TMA-staged prefix
		k_wavefront_trace<true, 1, 2, true, true><<<dims.trace_grid_staged16, dims.trace_block, dims.trace_smem16, stream>>>(sc, sh, wf, rs, bv, 0u, 0u);
		break;
	default:
		RFW_TRACE_CASE(0, 1, 0, false)
#undef RFW_TRACE_CASE
	}
	return cudaGetLastError();
}
cudaError_t launch_trace(const SceneView &sc, const ShardView &sh, const WavefrontView &wf, const RenderSettings &rs,
						 const BatchView &bv, uint32_t depth, uint32_t in_buf, const LaunchDims &dims, cudaStream_t stream)
{
	if (sc.tl_instances != nullptr)
	{
		if (rs.trace_variant == 0 || sc.nodes16 == nullptr)
			k_wavefront_trace_tl<false><<<dims.trace_grid, dims.trace_block, 0, stream>>>(sc, sh, wf, rs, bv, depth, in_buf);
		else
			k_wavefront_trace<false, 1, 2, true, false, true><<<dims.trace_grid, dims.trace_block, 0, stream>>>(sc, sh, wf, rs, bv, depth, in_buf);
		return cudaGetLastError();
	}
	if (sc.cw_nodes != nullptr)
	{
		k_wavefront_trace_cw<false><<<dims.trace_grid, dims.trace_block, 0, stream>>>(sc, sh, wf, rs, bv, depth, in_buf);
		return cudaGetLastError();
	}
	switch ((rs.smem_nodes > 0 && rs.trace_variant != 13) ? 0 : rs.trace_variant)
	{
#define RFW_TRACE_CASE(V, LQ, LEAN, PACKED)                                                                                     \
	case V:                                                                                                             \
		k_wavefront_trace<false, LQ, LEAN, PACKED><<<dims.trace_grid, dims.trace_block, (LQ == 1 && LEAN == 0 && !PACKED) ? dims.trace_smem : 0, stream>>>( \
			sc, sh, wf, rs, bv, depth, in_buf);                                                                         \
		break;
		RFW_TRACE_CASE(1, 2, 0, false)
		RFW_TRACE_CASE(3, 1, 1, false)
		RFW_TRACE_CASE(5, 1, 2, false)
		RFW_TRACE_CASE(8, 1, 0, true)
		RFW_TRACE_CASE(9, 1, 2, true)
		RFW_TRACE_CASE(10, 1, 1, true)
		RFW_TRACE_CASE(11, 1, 3, true)
		RFW_TRACE_CASE(12, 1, 4, true)
	case 13: // packed nodes with the TMA-staged prefix
		k_wavefront_trace<false, 1, 2, true, true><<<dims.trace_grid_staged16, dims.trace_block, dims.trace_smem16, stream>>>(sc, sh, wf, rs, bv, depth, in_buf);
		break;
	default:
		RFW_TRACE_CASE(0, 1, 0, false)
#undef RFW_TRACE_CASE
	}
	return cudaGetLastError();
}
cudaError_t launch_fold(const ShardView &sh, const WavefrontView &wf, const BatchView &bv, float scale, int write_fb, const DisplayTarget &dt,
						cudaStream_t stream)
{
	const uint32_t n = sh.local_pixels > 0 ? sh.local_pixels : 1u; // a rank without tiles still reports its arrival
	k_fold<<<(n + 255) / 256, 256, 0, stream>>>(sh, wf, bv, scale, write_fb, dt);
	return cudaGetLastError();
}
cudaError_t launch_display_spin(const uint32_t *counter, uint32_t need, uint32_t *err, cudaStream_t stream)
{
	k_display_spin<<<1, 1, 0, stream>>>(counter, need, err);
	return cudaGetLastError();
}
cudaError_t launch_shard_sync(const ShardSync &sync, uint32_t depth, uint32_t need_full, uint32_t need_lagged, uint32_t spp,
							  uint32_t *local_seen_row, uint32_t *err, cudaStream_t stream)
{
	k_shard_sync<<<1, MAX_BATCH_SPP, 0, stream>>>(sync, depth, need_full, need_lagged, spp, local_seen_row, err);
	return cudaGetLastError();
}
cudaError_t launch_display_release(uint32_t *consumed, uint32_t value, cudaStream_t stream)
{
	k_display_release<<<1, 1, 0, stream>>>(consumed, value);
	return cudaGetLastError();
}
cudaError_t launch_sort_setup(const SceneView &sc, const WavefrontView &wf, const RenderSettings &rs, cudaStream_t stream)
{
	k_sort_setup<<<1, 32, 0, stream>>>(sc, const_cast<SortGrid *>(wf.grid), rs.sort_cell_bits);
	return cudaGetLastError();
}
cudaError_t launch_sort(const WavefrontView &wf, const RenderSettings &rs, const BatchView &bv, uint32_t depth, const LaunchDims &dims,
						cudaStream_t stream)
{
	const uint32_t bins = 1u << (3 * rs.sort_cell_bits + rs.sort_dir_bits), n_chunks = bins / SORT_CHUNK;
	k_sort_scan<<<n_chunks, 1024, 0, stream>>>(wf);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess)
		return e;
	k_sort_move<<<dims.move_grid, 256, 0, stream>>>(wf, bv, depth, n_chunks);
	return cudaGetLastError();
}
cudaError_t launch_aov_finalize(const ShardView &sh, const float4 *acc, float scale, float4 *out, cudaStream_t stream)
{
	if (sh.local_pixels)
		k_aov_finalize<<<(sh.local_pixels + 255) / 256, 256, 0, stream>>>(sh, acc, scale, out);
	return cudaGetLastError();
}
cudaError_t launch_tone_map(const float4 *framebuffer, uint32_t *rgba8_out, uint32_t n, float contrast, float brightness,
							cudaStream_t stream)
{
	if (n)
		k_tone_map<<<(n + 255) / 256, 256, 0, stream>>>(framebuffer, rgba8_out, n, contrast, brightness);
	return cudaGetLastError();
}
cudaError_t launch_emode(const SceneView &sc, const ShardView &sh, const WavefrontView &wf, const RenderSettings &rs,
						 const void *raw_materials, const uint32_t *tex_desc, uint32_t tex_count, const LaunchDims &dims,
						 cudaStream_t stream)
{
	k_emode<<<dims.trace_grid, dims.trace_block, dims.trace_smem, stream>>>(
		sc, sh, wf, rs, reinterpret_cast<const rfwb200_material *>(raw_materials), tex_desc, tex_count);
	return cudaGetLastError();
}
cudaError_t launch_trace_closest(const SceneView &sc, const RenderSettings &rs, const float4 *origins, const float4 *directions,
								 uint32_t n, float t_min, float4 *hits_out, uint32_t *cursor, const LaunchDims &dims,
								 cudaStream_t stream, uint32_t *inst_out)
{
	k_trace_closest<<<dims.trace_grid, dims.trace_block, dims.trace_smem, stream>>>(sc, rs, origins, directions, n, t_min,
																				   hits_out, cursor, inst_out);
	return cudaGetLastError();
}
cudaError_t launch_trace_occluded(const SceneView &sc, const RenderSettings &rs, const float4 *origins,
								  const float4 *directions_tmax, uint32_t n, float t_min, uint8_t *occluded_out,
								  uint32_t *cursor, const LaunchDims &dims, cudaStream_t stream)
{
	k_trace_occluded<<<dims.trace_grid, dims.trace_block, dims.trace_smem, stream>>>(sc, rs, origins, directions_tmax, n,
																					t_min, occluded_out, cursor);
	return cudaGetLastError();
}
cudaError_t launch_generate_only(const SceneView &sc, const ShardView &sh, const WavefrontView &wf, uint32_t sample_index,
								 int emode, float4 *origins_out, float4 *directions_out, cudaStream_t stream)
{
	const uint32_t n = sh.local_pixels;
	k_generate_only<<<(n + 255) / 256, 256, 0, stream>>>(sc, sh, wf, sample_index, emode, origins_out, directions_out);
	return cudaGetLastError();
}
cudaError_t launch_assemble(const ShardView &sh, const float4 *gathered, size_t stride, float4 *image, cudaStream_t stream)
{
	const uint32_t total_tiles = sh.tiles_x * sh.tiles_y;
	const uint32_t max_local = ((total_tiles + sh.world - 1) / sh.world) * sh.tile_w * sh.tile_h;
	dim3 grid((max_local + 255) / 256, sh.world);
	k_assemble<<<grid, 256, 0, stream>>>(sh, gathered, stride, image);
	return cudaGetLastError();
}

#endif // RFW_PART != 2 (trace part)

#if RFW_PART != 1 // ---- shade part ----
cudaError_t LAUNCH_SHADE(const SceneView &sc, const ShardView &sh, const WavefrontView &wf, const RenderSettings &rs,
						 const BatchView &bv, uint32_t depth, uint32_t in_buf, uint32_t out_buf, const ShardSync &sync, const LaunchDims &dims,
						 cudaStream_t stream)
{
	K_SHADE<<<dims.shade_grid, dims.shade_block, 0, stream>>>(sc, sh, wf, rs, bv, depth, in_buf, out_buf, sync);
	return cudaGetLastError();
}
cudaError_t SHADE_OCCUPANCY(int block, int *per_sm)
{
	return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, K_SHADE, block, 0);
}
cudaError_t SHADE_PRELOAD()
{
	cudaFuncAttributes fa;
	return cudaFuncGetAttributes(&fa, K_SHADE);
}
#endif // RFW_PART != 1 (shade part)

} // namespace rfwb200
