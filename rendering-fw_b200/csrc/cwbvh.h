// cwbvh.h — the compressed 8-wide BVH the trace kernels walk by default (setting "bvh" = 8), shared by the host builder
// (bvh_build.cpp), the device refit (geometry.cu) and the traversal kernels (kernels.cu).
//
// Why: ncu shows the 4-wide kernel bound by L1TEX throughput (76-82 % of peak, profiles/r01b) — every node visit of every
// lane is 7 x 16 B from a different line.  An 8-wide node with child boxes quantised to 8 bits per plane relative to the
// node's own box (Ylitie, Karras, Laine: "Efficient Incoherent Ray Traversal on GPUs Through Compressed Wide BVHs",
// HPG 2017) is 80 B for 8 children instead of 128 B for 4, and a ray visits about half as many nodes.  The reference's
// own traversal is the uncompressed 4-wide MBVH of RFW/system/bvh/include/bvh/mbvh_node.h:60-106 walked by
// CUDART/src/CUDAIntersect.h:155-197; quantised boxes are conservative (they only grow), so the set of triangles a ray
// tests is a superset and the closest hit is the same triangle.
//
// Node, 80 B = 5 x 16 B:
//   w0..w2   p            origin of the quantisation grid = min corner of the node
//   w3       ex ey ez imask   biased exponents: grid step on axis a is 2^(e_a - 127); imask bit s = slot s is an inner node
//   w4       child_base   index of the first inner child; inner children are stored consecutively in slot order
//   w5       tri_base     first triangle record of this node's leaf children
//   w6,w7    meta[8]      per slot: 0 = empty; inner: 0x20 | (24 + slot); leaf: (unary triangle count << 5) | offset from tri_base
//   w8,w9    qlo_x[8]     w10,w11 qlo_y[8]   w12,w13 qlo_z[8]
//   w14,w15  qhi_x[8]     w16,w17 qhi_y[8]   w18,w19 qhi_z[8]
// A leaf slot holds 1-3 triangles, a node at most 24; slots are assigned so that slot s lies towards the octant
// (s&4 ? +x : -x, s&2 ? +y : -y, s&1 ? +z : -z) of the node: XOR-ing the slot number with the ray's octant then yields a
// front-to-back order without sorting.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CW_HD __host__ __device__ __forceinline__
#else
#define CW_HD inline
#endif

namespace rfwb200
{

struct alignas(16) CwNode
{
	float p[3];
	uint8_t e[3], imask;
	uint32_t child_base, tri_base;
	uint8_t meta[8];
	uint8_t qlo[3][8];
	uint8_t qhi[3][8];
};
static_assert(sizeof(CwNode) == 80, "compressed wide node is 5 x 16 bytes");

// full-precision child boxes of a node, kept beside the compressed nodes for refits (192 B)
struct alignas(16) CwAux
{
	float lo[3][8];
	float hi[3][8];
};
static_assert(sizeof(CwAux) == 192, "");

constexpr int CW_STACK = 32;	// uint2 entries per ray: one per level at most (builder bounds the depth)
constexpr int CW_MAX_DEPTH = 30;

CW_HD float cw_bits_to_float(uint32_t b)
{
#if defined(__CUDA_ARCH__)
	return __uint_as_float(b);
#else
	union
	{
		uint32_t u;
		float f;
	} c;
	c.u = b;
	return c.f;
#endif
}
CW_HD uint32_t cw_float_to_bits(float f)
{
#if defined(__CUDA_ARCH__)
	return __float_as_uint(f);
#else
	union
	{
		uint32_t u;
		float f;
	} c;
	c.f = f;
	return c.u;
#endif
}

// Quantise the (already padded) child boxes of one node.  Every operation is a single correctly rounded IEEE operation
// or exact (power-of-two scaling, floor, ceil), evaluated in this order on the host and in geometry.cu (-fmad=false), so a
// device refit reproduces the host's nodes bit for bit.
CW_HD void cw_quantize(CwNode &n, const CwAux &a)
{
	float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
	for (int s = 0; s < 8; s++)
		if (n.meta[s])
			for (int k = 0; k < 3; k++)
			{
				lo[k] = a.lo[k][s] < lo[k] ? a.lo[k][s] : lo[k];
				hi[k] = a.hi[k][s] > hi[k] ? a.hi[k][s] : hi[k];
			}
	for (int k = 0; k < 3; k++)
	{
		if (!(lo[k] <= hi[k]))
			lo[k] = hi[k] = 0.0f; // node without children
		n.p[k] = lo[k];
		const float ext = hi[k] - lo[k];
		const uint32_t sb = cw_float_to_bits(ext / 255.0f);
		uint32_t be = (sb >> 23) & 0xffu;
		if (sb & 0x7fffffu)
			be += 1; // ceil(log2(ext / 255))
		be = be < 1u ? 1u : (be > 253u ? 253u : be);
		while (be < 253u && lo[k] + 255.0f * cw_bits_to_float(be << 23) < hi[k])
			be += 1; // rounding of ext / 255 or of lo + 255 * step
		n.e[k] = uint8_t(be);
		const float step = cw_bits_to_float(be << 23), inv_step = cw_bits_to_float((254u - be) << 23);
		for (int s = 0; s < 8; s++)
		{
			if (!n.meta[s])
			{
				n.qlo[k][s] = 255, n.qhi[k][s] = 0; // never hit: child bits of an empty slot are 0 anyway
				continue;
			}
			float ql = floorf((a.lo[k][s] - lo[k]) * inv_step);
			ql = ql < 0.0f ? 0.0f : (ql > 255.0f ? 255.0f : ql);
			while (ql > 0.0f && lo[k] + ql * step > a.lo[k][s])
				ql -= 1.0f;
			float qh = ceilf((a.hi[k][s] - lo[k]) * inv_step);
			qh = qh < 0.0f ? 0.0f : (qh > 255.0f ? 255.0f : qh);
			while (qh < 255.0f && lo[k] + qh * step < a.hi[k][s])
				qh += 1.0f;
			n.qlo[k][s] = uint8_t(ql), n.qhi[k][s] = uint8_t(qh);
		}
	}
}

// per-ray constants of the compressed traversal
struct CwRay
{
	float ox, oy, oz;		// origin
	float idx, idy, idz;	// 1 / direction (never 0 or inf: callers clamp tiny components)
	uint32_t octinv;		// 7 - octant, octant = (dx<0)<<2 | (dy<0)<<1 | (dz<0)
};

CW_HD uint32_t cw_octinv(float dx, float dy, float dz) { return 7u - ((dx < 0.0f ? 4u : 0u) | (dy < 0.0f ? 2u : 0u) | (dz < 0.0f ? 1u : 0u)); }

CW_HD float cw_byte_to_float(uint32_t word, int byte)
{
	// exact: 0x4B000000 | b is the float 2^23 + b
	return cw_bits_to_float(0x4B000000u | ((word >> (8 * byte)) & 0xffu)) - 8388608.0f;
}

// Intersect the 8 children of a node (given as its 20 words) with the ray segment [tmin, tmax].
// Returns the hit mask: bits 24..31 = inner children, already permuted into traversal order (highest bit = nearest
// octant), bits 0..23 = triangles of leaf children (bit = offset from tri_base).
CW_HD uint32_t cw_intersect_children(const uint32_t *w, const CwRay &r, float tmin, float tmax)
{
	const uint32_t e = w[3];
	const float sx = cw_bits_to_float((e & 0xffu) << 23) * r.idx;
	const float sy = cw_bits_to_float(((e >> 8) & 0xffu) << 23) * r.idy;
	const float sz = cw_bits_to_float(((e >> 16) & 0xffu) << 23) * r.idz;
	const float bx = (cw_bits_to_float(w[0]) - r.ox) * r.idx;
	const float by = (cw_bits_to_float(w[1]) - r.oy) * r.idy;
	const float bz = (cw_bits_to_float(w[2]) - r.oz) * r.idz;
	// near / far plane words per axis, chosen once per node by the sign of the direction
	const bool nx = r.idx < 0.0f, ny = r.idy < 0.0f, nz = r.idz < 0.0f;
	uint32_t hitmask = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int half = 0; half < 2; half++)
	{
		const uint32_t meta4 = w[6 + half];
		const uint32_t nearx = nx ? w[14 + half] : w[8 + half], farx = nx ? w[8 + half] : w[14 + half];
		const uint32_t neary = ny ? w[16 + half] : w[10 + half], fary = ny ? w[10 + half] : w[16 + half];
		const uint32_t nearz = nz ? w[18 + half] : w[12 + half], farz = nz ? w[12 + half] : w[18 + half];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (int i = 0; i < 4; i++)
		{
			const uint32_t meta = (meta4 >> (8 * i)) & 0xffu;
			const float t0x = cw_byte_to_float(nearx, i) * sx + bx, t1x = cw_byte_to_float(farx, i) * sx + bx;
			const float t0y = cw_byte_to_float(neary, i) * sy + by, t1y = cw_byte_to_float(fary, i) * sy + by;
			const float t0z = cw_byte_to_float(nearz, i) * sz + bz, t1z = cw_byte_to_float(farz, i) * sz + bz;
			const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));
			const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tmax));
			if (tn <= tf)
			{
				const bool inner = (meta & 0x18u) == 0x18u; // low five bits 24..31
				const uint32_t bit = inner ? (24u + ((meta & 7u) ^ r.octinv)) : (meta & 31u);
				hitmask |= (meta >> 5) << bit;
			}
		}
	}
	return hitmask;
}

} // namespace rfwb200
