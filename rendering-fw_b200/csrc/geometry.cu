// geometry.cu — device-side geometry pipeline in front of the trace kernels:
//
//   k_skin_vertices    linear-blend skinning of a mesh's vertices and normals from joint matrices
//                      (reference: CPU + TBB, RFW/system/src/rfw/geometry/gltf/mesh.cpp:18-48)
//   k_update_triangles rebuild the 160-B triangle records of a skinned mesh (mesh.cpp:428-449)
//   k_flatten_shade    instance flattening of the shading records (ShadeTri, device_types.h); the reference
//                      applies the instance normal matrix per path (CUDART/src/getShadingData.h:129-130)
//   k_refit            ONE launch: every leaf slot transforms its triangles to world space, writes their
//                      intersection records (TriRec) and its own box, then walks towards the root — the last
//                      thread to finish a node's slots (atomic arrival counter) carries the node's union into its
//                      parent's slot.  Replaces the host refit (RFW/system/bvh/src/bvh_tree.cpp:104-114,
//                      top_level_bvh.cpp:46-52) + full re-upload (CUDART/src/Context.cpp:270-311).
//
// This file is compiled with -fmad=false and IEEE division/sqrt: every expression below is evaluated in the same
// order as the host builder's (bvh_build.cpp tri_box/padded, context.cpp mul_point/mul_mat3), so a device refit is
// BIT-IDENTICAL to the host refit of the same vertices (tests/test_parity_gpu.py::test_device_refit_bit_exact).
#include "geometry.h"
#include "lbvh.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

namespace rfwb200
{

namespace
{

struct F3
{
	float x, y, z;
};

__device__ __forceinline__ F3 xf_point(const float *__restrict__ m, const float4 p)
{
	F3 o; // context.cpp mul_point: m[r]*p0 + m[4+r]*p1 + m[8+r]*p2 + m[12+r], left to right
	o.x = m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12];
	o.y = m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13];
	o.z = m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14];
	return o;
}

__device__ __forceinline__ F3 xf_normal(const float *__restrict__ m, const float x, const float y, const float z)
{
	F3 o; // context.cpp mul_mat3
	o.x = m[0] * x + m[3] * y + m[6] * z;
	o.y = m[1] * x + m[4] * y + m[7] * z;
	o.z = m[2] * x + m[5] * y + m[8] * z;
	return o;
}

constexpr float BOX_PAD = 1e-5f; // bvh_build.cpp

__device__ __forceinline__ void pad_axis(float &lo, float &hi)
{
	const float m = fmaxf(fabsf(lo), fabsf(hi)); // bvh_build.cpp padded()
	const float pad = fmaxf(BOX_PAD, m * 2.4e-7f);
	lo -= pad, hi += pad;
}

struct Box3
{
	float lo[3], hi[3];
};

__device__ __forceinline__ void box_reset(Box3 &b)
{
	b.lo[0] = b.lo[1] = b.lo[2] = 1e30f; // bvh_build.cpp Box::reset
	b.hi[0] = b.hi[1] = b.hi[2] = -1e30f;
}

__device__ __forceinline__ void store_slot(BvhNode4 *n, int s, const Box3 &b)
{
	n->minx[s] = b.lo[0], n->miny[s] = b.lo[1], n->minz[s] = b.lo[2];
	n->maxx[s] = b.hi[0], n->maxy[s] = b.hi[1], n->maxz[s] = b.hi[2];
}

} // namespace

// ------------------------------------------------------------------------------------------------
// k_refit — one thread per (node, slot)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_refit(const GeometryView g)
{
	const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t node = gid >> 2;
	int slot = int(gid & 3u);
	if (node >= g.node_count)
		return;
	BvhNode4 *nodes = g.nodes;
	if (slot >= nodes[node].pad[0])
		return;
	const int32_t c = nodes[node].child[slot];
	if (c >= 0)
		return; // inner slots are filled by whichever thread completes the child node
	Box3 b;
	box_reset(b);
	{
		const uint32_t v = uint32_t(~c), first = v >> 2, cnt = (v & 3u) + 1u;
		for (uint32_t i = first; i < first + cnt; i++)
		{
			const uint32_t src = g.tri_order[i];
			const uint32_t ii = g.flat_inst[src];
			const DeviceInstance &in = g.instances[ii];
			const uint32_t prim = src - in.flat_off;
			const uint32_t *ix = g.indices + size_t(in.tri_off + prim) * 3;
			const float4 *vb = g.verts + in.vert_off;
			const F3 v0 = xf_point(in.transform, vb[ix[0]]);
			const F3 v1 = xf_point(in.transform, vb[ix[1]]);
			const F3 v2 = xf_point(in.transform, vb[ix[2]]);
			float4 *rec = reinterpret_cast<float4 *>(g.out_tris + i);
			rec[0] = make_float4(v0.x, v0.y, v0.z, v1.x - v0.x);
			rec[1] = make_float4(v1.y - v0.y, v1.z - v0.z, v2.x - v0.x, v2.y - v0.y);
			rec[2] = make_float4(v2.z - v0.z, __uint_as_float(src), in.det_eps, 0.0f);
			float lo[3] = {fminf(fminf(v0.x, v1.x), v2.x), fminf(fminf(v0.y, v1.y), v2.y), fminf(fminf(v0.z, v1.z), v2.z)};
			float hi[3] = {fmaxf(fmaxf(v0.x, v1.x), v2.x), fmaxf(fmaxf(v0.y, v1.y), v2.y), fmaxf(fmaxf(v0.z, v1.z), v2.z)};
			if (g.ref_boxes != nullptr && in.moved == 0u)
			{
				// a triangle that has not moved since the build keeps the builder's box of this reference (clipped by
				// spatial splits); refitting it from the whole triangle would undo the splits (bvh_build.cpp refit_ref_box)
				const float *rb = g.ref_boxes + size_t(i) * 6;
				lo[0] = rb[0], lo[1] = rb[1], lo[2] = rb[2], hi[0] = rb[3], hi[1] = rb[4], hi[2] = rb[5];
			}
#pragma unroll
			for (int a = 0; a < 3; a++)
			{
				pad_axis(lo[a], hi[a]);
				b.lo[a] = fminf(b.lo[a], lo[a]), b.hi[a] = fmaxf(b.hi[a], hi[a]);
			}
		}
	}
	if (!g.write_boxes)
		return;
	// walk up: write the slot, announce it, and continue only as the last arrival of the node
	for (;;)
	{
		store_slot(&nodes[node], slot, b);
		__threadfence();
		const uint32_t used = uint32_t(nodes[node].pad[0]);
		const uint32_t arrived = atomicAdd(&g.arrivals[node], 1u) + 1u;
		if (arrived < used)
			return;
		const uint32_t up = g.parent_slot[node];
		if (up == 0xffffffffu)
			return; // root complete
		__threadfence();
		// union of the finished node (its other slots were written by other threads: read past L1)
		box_reset(b);
		const volatile BvhNode4 *vn = nodes + node;
		for (uint32_t k = 0; k < used; k++)
		{
			b.lo[0] = fminf(b.lo[0], vn->minx[k]), b.lo[1] = fminf(b.lo[1], vn->miny[k]), b.lo[2] = fminf(b.lo[2], vn->minz[k]);
			b.hi[0] = fmaxf(b.hi[0], vn->maxx[k]), b.hi[1] = fmaxf(b.hi[1], vn->maxy[k]), b.hi[2] = fmaxf(b.hi[2], vn->maxz[k]);
		}
		node = up >> 2, slot = int(up & 3u);
	}
}

// ------------------------------------------------------------------------------------------------
// k_pack_nodes — BvhNode4 (128 B, fp32 planes) -> BvhNode4Packed (80 B, bfloat16 planes relative to the node's min
// corner, rounded outwards).  One thread per node.
// ------------------------------------------------------------------------------------------------
// The trace kernel decodes the plane in the LOWER half of a word exactly (bits << 16) and uses the one in the UPPER half
// as the word stands, i.e. with up to 0xffff of foreign mantissa bits below it (kernels.cu CHILD_P): `slack` is that
// worst case, and the rounding loops make sure the box still contains the exact one with it.
__device__ __forceinline__ uint32_t bf16_down(float v, float p, float exact_lo, uint32_t slack)
{
	// largest bfloat16 <= v (v >= 0: truncation), then make sure p + decoded <= the exact plane
	uint32_t b = __float_as_uint(v) >> 16;
	while (b > 0u && p + __uint_as_float((b << 16) | slack) > exact_lo)
		b -= 1u;
	return b;
}
__device__ __forceinline__ uint32_t bf16_up(float v, float p, float exact_hi)
{
	const uint32_t bits = __float_as_uint(v);
	uint32_t b = (bits >> 16) + ((bits & 0xffffu) ? 1u : 0u);
	while (b < 0x7f7fu && p + __uint_as_float(b << 16) < exact_hi)
		b += 1u;
	return b;
}

__global__ void __launch_bounds__(256) k_pack_nodes(const BvhNode4 *__restrict__ nodes, BvhNode4Packed *__restrict__ out, uint32_t n)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const BvhNode4 nd = nodes[i];
	const int used = nd.pad[0];
	const float *lo[3] = {nd.minx, nd.miny, nd.minz}, *hi[3] = {nd.maxx, nd.maxy, nd.maxz};
	BvhNode4Packed o;
	o.pad = uint32_t(used);
	for (int a = 0; a < 3; a++)
	{
		float p = 3.0e38f;
		for (int k = 0; k < used; k++)
			p = fminf(p, lo[a][k]);
		// the corner sits a hair below the smallest plane, so that an offset of zero with foreign low bits (a denormal)
		// still decodes to a plane at or below the exact one
		p = used == 0 ? 0.0f : p - (fabsf(p) * 1.2e-7f + 1e-30f);
		o.p[a] = p;
		uint32_t l[4], h[4];
		for (int k = 0; k < 4; k++)
		{
			if (k < used)
				l[k] = bf16_down(lo[a][k] - p, p, lo[a][k], (k & 1) ? 0xffffu : 0u), h[k] = bf16_up(hi[a][k] - p, p, hi[a][k]);
			else
				l[k] = 0x7f7fu, h[k] = 0xff7fu; // unused slot: inverted box (+3.4e38, -3.4e38), empty for every ray direction
		}
		o.plane[a][0] = l[0] | (l[1] << 16), o.plane[a][1] = l[2] | (l[3] << 16);
		o.plane[a][2] = h[0] | (h[1] << 16), o.plane[a][3] = h[2] | (h[3] << 16);
	}
	for (int k = 0; k < 4; k++)
		o.child[k] = nd.child[k];
	out[i] = o;
}

// ------------------------------------------------------------------------------------------------
// k_records — one thread per leaf reference: the intersection record only (compressed 8-wide layout, whose boxes are
// refitted and re-quantised by the host builder)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_records(const GeometryView g)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= g.ref_count)
		return;
	const uint32_t src = g.tri_order[i];
	const uint32_t ii = g.flat_inst[src];
	const DeviceInstance &in = g.instances[ii];
	const uint32_t prim = src - in.flat_off;
	const uint32_t *ix = g.indices + size_t(in.tri_off + prim) * 3;
	const float4 *vb = g.verts + in.vert_off;
	const F3 v0 = xf_point(in.transform, vb[ix[0]]);
	const F3 v1 = xf_point(in.transform, vb[ix[1]]);
	const F3 v2 = xf_point(in.transform, vb[ix[2]]);
	float4 *rec = reinterpret_cast<float4 *>(g.out_tris + i);
	rec[0] = make_float4(v0.x, v0.y, v0.z, v1.x - v0.x);
	rec[1] = make_float4(v1.y - v0.y, v1.z - v0.z, v2.x - v0.x, v2.y - v0.y);
	rec[2] = make_float4(v2.z - v0.z, __uint_as_float(src), in.det_eps, 0.0f);
}

// ------------------------------------------------------------------------------------------------
// k_flatten_shade — one thread per flattened triangle (context.cpp flatten_scene, shading half)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_flatten_shade(const GeometryView g)
{
	const uint32_t src = blockIdx.x * blockDim.x + threadIdx.x;
	if (src >= g.flat_count)
		return;
	const uint32_t ii = g.flat_inst[src];
	const DeviceInstance &in = g.instances[ii];
	const uint32_t prim = src - in.flat_off;
	const float4 *t = reinterpret_cast<const float4 *>(g.mesh_tris) + size_t(in.tri_off + prim) * 10; // 160-B rfwb200_triangle
	const float4 q0 = t[0], q1 = t[1], q2 = t[2], q3 = t[3], q4 = t[4], q5 = t[5], q6 = t[6];
	const F3 n0 = xf_normal(in.normal, q2.x, q2.y, q2.z);
	const F3 n1 = xf_normal(in.normal, q3.x, q3.y, q3.z);
	const F3 n2 = xf_normal(in.normal, q4.x, q4.y, q4.z);
	const F3 gn = xf_normal(in.normal, q2.w, q3.w, q4.w);
	const float il = 1.0f / sqrtf(gn.x * gn.x + gn.y * gn.y + gn.z * gn.z);
	float4 *o = reinterpret_cast<float4 *>(g.out_shade + src);
	o[0] = q0; // u0 u1 u2 light_tri_idx
	o[1] = q1; // v0 v1 v2 material
	o[2] = make_float4(n0.x, n0.y, n0.z, gn.x * il);
	o[3] = make_float4(n1.x, n1.y, n1.z, gn.y * il);
	o[4] = make_float4(n2.x, n2.y, n2.z, gn.z * il);
	o[5] = make_float4(q5.w, q6.w, __uint_as_float(ii), __uint_as_float(prim)); // area, LOD, instance, primitive
}

// ------------------------------------------------------------------------------------------------
// skinning — gltf/mesh.cpp:18-48: S = sum_k w_k J[j_k];  v' = S v;  n' = (n^T S^-1).xyz / |n^T S^-1|_4
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_skin_vertices(const SkinView s)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= s.vertex_count)
		return;
	const uint4 j4 = s.joints[i];
	const float4 w4 = s.weights[i];
	float S[16]; // column-major like glm
	{
		const float *a = s.joint_matrices + size_t(j4.x) * 16, *b = s.joint_matrices + size_t(j4.y) * 16;
		const float *c = s.joint_matrices + size_t(j4.z) * 16, *d = s.joint_matrices + size_t(j4.w) * 16;
#pragma unroll
		for (int k = 0; k < 16; k++)
			S[k] = ((a[k] * w4.x + b[k] * w4.y) + c[k] * w4.z) + d[k] * w4.w;
	}
	const float4 p = s.base_vertices[i];
	float4 r;
	r.x = S[0] * p.x + S[4] * p.y + S[8] * p.z + S[12] * p.w;
	r.y = S[1] * p.x + S[5] * p.y + S[9] * p.z + S[13] * p.w;
	r.z = S[2] * p.x + S[6] * p.y + S[10] * p.z + S[14] * p.w;
	r.w = S[3] * p.x + S[7] * p.y + S[11] * p.z + S[15] * p.w;
	s.out_vertices[i] = r;

	// n^T S^-1 with the w of the base normal = 0: only the upper-left 3x3 rows of the inverse matter for xyz; use
	// the adjugate / determinant of the full 4x4 the way a general inverse does (joint matrices are affine, last
	// row (0,0,0,1) up to rounding, but keep the general formula)
	const float a00 = S[0], a01 = S[4], a02 = S[8], a03 = S[12];
	const float a10 = S[1], a11 = S[5], a12 = S[9], a13 = S[13];
	const float a20 = S[2], a21 = S[6], a22 = S[10], a23 = S[14];
	const float a30 = S[3], a31 = S[7], a32 = S[11], a33 = S[15];
	const float b00 = a00 * a11 - a01 * a10, b01 = a00 * a12 - a02 * a10, b02 = a00 * a13 - a03 * a10;
	const float b03 = a01 * a12 - a02 * a11, b04 = a01 * a13 - a03 * a11, b05 = a02 * a13 - a03 * a12;
	const float b06 = a20 * a31 - a21 * a30, b07 = a20 * a32 - a22 * a30, b08 = a20 * a33 - a23 * a30;
	const float b09 = a21 * a32 - a22 * a31, b10 = a21 * a33 - a23 * a31, b11 = a22 * a33 - a23 * a32;
	const float det = b00 * b11 - b01 * b10 + b02 * b09 + b03 * b08 - b04 * b07 + b05 * b06;
	const float id = 1.0f / det;
	// inverse entries inv[r][c] (row r, column c), rows 0..2 only
	const float i00 = (a11 * b11 - a12 * b10 + a13 * b09) * id, i01 = (-a01 * b11 + a02 * b10 - a03 * b09) * id;
	const float i02 = (a31 * b05 - a32 * b04 + a33 * b03) * id;
	const float i10 = (-a10 * b11 + a12 * b08 - a13 * b07) * id, i11 = (a00 * b11 - a02 * b08 + a03 * b07) * id;
	const float i12 = (-a30 * b05 + a32 * b02 - a33 * b01) * id;
	const float i20 = (a10 * b10 - a11 * b08 + a13 * b06) * id, i21 = (-a00 * b10 + a01 * b08 - a03 * b06) * id;
	const float i22 = (a30 * b04 - a31 * b02 + a33 * b00) * id;
	const float i03 = (a22 * b04 - a21 * b05 - a23 * b03) * id, i13 = (a20 * b05 - a22 * b02 + a23 * b01) * id;
	const float i23 = (a21 * b02 - a20 * b04 - a23 * b00) * id;
	const float4 n = s.base_normals[i];
	// row vector times matrix: out[c] = sum_r n[r] * inv[r][c]
	float nx = n.x * i00 + n.y * i10 + n.z * i20;
	float ny = n.x * i01 + n.y * i11 + n.z * i21;
	float nz = n.x * i02 + n.y * i12 + n.z * i22;
	// Reference behaviour kept: the product is a vec4 whose w = n . (translation column of the inverse) is NOT zero,
	// and `result / result.length()` (mesh.cpp:43) divides by the length of all four components (math.h:797-806), so
	// skinned normals come out slightly shorter than 1 wherever the skin matrix translates.
	const float nw = n.x * i03 + n.y * i13 + n.z * i23;
	const float len = sqrtf(nx * nx + ny * ny + nz * nz + nw * nw);
	nx /= len, ny /= len, nz /= len;
	s.out_normals[i] = make_float4(nx, ny, nz, 0.0f);
}

// morph targets — gltf/mesh.cpp:126-148: vertex = (pose0, 1) + sum_j w_j * (pose_j, 0), normal = pose0.n + sum_j w_j * pose_j.n
__global__ void __launch_bounds__(128) k_morph_vertices(const MorphView m)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m.vertex_count)
		return;
	float4 p = m.pose_positions[i], n = m.pose_normals[i];
	p.w = 1.0f, n.w = 0.0f;
	for (uint32_t j = 0; j < m.n_weights; j++)
	{
		const float w = m.weights[j];
		const float4 dp = m.pose_positions[size_t(j + 1) * m.vertex_count + i], dn = m.pose_normals[size_t(j + 1) * m.vertex_count + i];
		p.x += w * dp.x, p.y += w * dp.y, p.z += w * dp.z;
		n.x += w * dn.x, n.y += w * dn.y, n.z += w * dn.z;
	}
	m.out_vertices[i] = p;
	m.out_normals[i] = n;
}

// mesh.cpp:428-449 (indexed) — vertex0..2, vN0..2 and the geometric normal of every triangle record
__global__ void __launch_bounds__(128) k_update_triangles(const SkinView s)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= s.triangle_count)
		return;
	const uint32_t *ix = s.indices + size_t(t) * 3;
	const float4 v0 = s.out_vertices[ix[0]], v1 = s.out_vertices[ix[1]], v2 = s.out_vertices[ix[2]];
	const float4 n0 = s.out_normals[ix[0]], n1 = s.out_normals[ix[1]], n2 = s.out_normals[ix[2]];
	const float e1x = v1.x - v0.x, e1y = v1.y - v0.y, e1z = v1.z - v0.z;
	const float e2x = v2.x - v0.x, e2y = v2.y - v0.y, e2z = v2.z - v0.z;
	float Nx = e1y * e2z - e1z * e2y, Ny = e1z * e2x - e1x * e2z, Nz = e1x * e2y - e1y * e2x;
	const float il = 1.0f / sqrtf(Nx * Nx + Ny * Ny + Nz * Nz);
	Nx *= il, Ny *= il, Nz *= il;
	float4 *q = reinterpret_cast<float4 *>(s.mesh_tris) + size_t(t) * 10;
	q[2] = make_float4(n0.x, n0.y, n0.z, Nx);
	q[3] = make_float4(n1.x, n1.y, n1.z, Ny);
	q[4] = make_float4(n2.x, n2.y, n2.z, Nz);
	q[7] = make_float4(v0.x, v0.y, v0.z, q[7].w);
	q[8] = make_float4(v1.x, v1.y, v1.z, q[8].w);
	q[9] = make_float4(v2.x, v2.y, v2.z, q[9].w);
}

// ------------------------------------------------------------------------------------------------
// LBVH build on the device (lbvh.h): bounds + Morton keys -> cub radix sort -> radix tree -> boxes -> 4-wide collapse
// ------------------------------------------------------------------------------------------------
namespace
{
__device__ __forceinline__ int32_t float_to_ordered(float f)
{
	const int32_t i = __float_as_int(f);
	return i >= 0 ? i : i ^ 0x7fffffff;
}
__host__ __device__ __forceinline__ float ordered_to_float(int32_t i)
{
	const int32_t b = i >= 0 ? i : i ^ 0x7fffffff;
#if defined(__CUDA_ARCH__)
	return __int_as_float(b);
#else
	float f;
	memcpy(&f, &b, 4);
	return f;
#endif
}
} // namespace

// per flattened triangle: exact world-space box (the same vertex arithmetic as k_refit) + scene bounds of the centroids
__device__ __forceinline__ void lbvh_world_triangle(const GeometryView &g, uint32_t src, F3 &v0, F3 &v1, F3 &v2)
{
	const uint32_t ii = g.flat_inst[src];
	const DeviceInstance &in = g.instances[ii];
	const uint32_t prim = src - in.flat_off;
	const uint32_t *ix = g.indices + size_t(in.tri_off + prim) * 3;
	const float4 *vb = g.verts + in.vert_off;
	v0 = xf_point(in.transform, vb[ix[0]]);
	v1 = xf_point(in.transform, vb[ix[1]]);
	v2 = xf_point(in.transform, vb[ix[2]]);
}

__global__ void __launch_bounds__(256) k_lbvh_bounds(const GeometryView g, LbvhBox *__restrict__ boxes, int32_t *__restrict__ scene_bounds)
{
	const uint32_t src = blockIdx.x * blockDim.x + threadIdx.x;
	float c[3] = {0.f, 0.f, 0.f};
	const bool live = src < g.flat_count;
	if (live)
	{
		F3 v0, v1, v2;
		lbvh_world_triangle(g, src, v0, v1, v2);
		LbvhBox b;
		b.lo[0] = fminf(fminf(v0.x, v1.x), v2.x), b.lo[1] = fminf(fminf(v0.y, v1.y), v2.y), b.lo[2] = fminf(fminf(v0.z, v1.z), v2.z);
		b.hi[0] = fmaxf(fmaxf(v0.x, v1.x), v2.x), b.hi[1] = fmaxf(fmaxf(v0.y, v1.y), v2.y), b.hi[2] = fmaxf(fmaxf(v0.z, v1.z), v2.z);
		for (int a = 0; a < 3; a++)
			c[a] = 0.5f * (b.lo[a] + b.hi[a]);
		b.pad0 = b.pad1 = 0.0f;
		boxes[src] = b;
	}
	// warp-reduced atomics on the centroid bounds
	for (int a = 0; a < 3; a++)
	{
		const int32_t lo = __reduce_min_sync(0xffffffffu, live ? float_to_ordered(c[a]) : 0x7fffffff);
		const int32_t hi = __reduce_max_sync(0xffffffffu, live ? float_to_ordered(c[a]) : int32_t(0x80000000));
		if ((threadIdx.x & 31u) == 0u)
		{
			atomicMin(&scene_bounds[a], lo);
			atomicMax(&scene_bounds[3 + a], hi);
		}
	}
}

// early split clipping (lbvh.h step 0): pieces per triangle, then one reference per piece
__global__ void __launch_bounds__(256) k_lbvh_counts(const LbvhBox *__restrict__ boxes, uint32_t n, float cell, uint32_t *__restrict__ counts)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
		counts[i] = uint32_t(lb_piece_count(boxes[i], cell));
}

__global__ void __launch_bounds__(256) k_lbvh_emit(const GeometryView g, const LbvhBox *__restrict__ boxes, const uint32_t *__restrict__ counts,
												   const uint32_t *__restrict__ offsets, LbvhBox *__restrict__ ref_box,
												   uint32_t *__restrict__ ref_tri)
{
	const uint32_t src = blockIdx.x * blockDim.x + threadIdx.x;
	if (src >= g.flat_count)
		return;
	const LbvhBox full = boxes[src];
	const uint32_t cnt = counts ? counts[src] : 1u, off = offsets ? offsets[src] : src;
	if (cnt <= 1u)
	{
		LbvhBox b = full;
		for (int a = 0; a < 3; a++)
			pad_axis(b.lo[a], b.hi[a]);
		ref_box[off] = b, ref_tri[off] = src;
		return;
	}
	F3 v0, v1, v2;
	lbvh_world_triangle(g, src, v0, v1, v2);
	const float p0[3] = {v0.x, v0.y, v0.z}, p1[3] = {v1.x, v1.y, v1.z}, p2[3] = {v2.x, v2.y, v2.z};
	for (uint32_t j = 0; j < cnt; j++)
	{
		int axis;
		float lo, hi;
		lb_piece_slab(full, int(cnt), int(j), axis, lo, hi);
		LbvhBox b = lb_clip_to_slab(p0, p1, p2, full, axis, lo, hi);
		for (int a = 0; a < 3; a++)
			pad_axis(b.lo[a], b.hi[a]);
		ref_box[off + j] = b, ref_tri[off + j] = src;
	}
}

__global__ void __launch_bounds__(256) k_lbvh_keys(const LbvhBox *__restrict__ ref_box, uint32_t n, const int32_t *__restrict__ scene_bounds,
												   uint64_t *__restrict__ keys)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	float lo[3], inv[3];
	for (int a = 0; a < 3; a++)
	{
		lo[a] = ordered_to_float(scene_bounds[a]);
		const float ext = ordered_to_float(scene_bounds[3 + a]) - lo[a];
		inv[a] = ext > 0.0f ? 1.0f / ext : 0.0f;
	}
	const LbvhBox b = ref_box[i];
	// centroid of the padded box = centroid of the exact box up to rounding: only the ordering matters
	const uint32_t m = lb_morton30(0.5f * (b.lo[0] + b.hi[0]), 0.5f * (b.lo[1] + b.hi[1]), 0.5f * (b.lo[2] + b.hi[2]), lo, inv);
	keys[i] = (uint64_t(m) << 32) | uint64_t(i);
}

__global__ void __launch_bounds__(256) k_lbvh_gather(const uint64_t *__restrict__ keys, const LbvhBox *__restrict__ ref_box,
													 const uint32_t *__restrict__ ref_tri, uint32_t n, LbvhBox *__restrict__ leaf_box,
													 uint32_t *__restrict__ tri_order, float *__restrict__ ref_boxes_out)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const uint32_t r = uint32_t(keys[i] & 0xffffffffull);
	const LbvhBox b = ref_box[r];
	leaf_box[i] = b;
	tri_order[i] = ref_tri[r];
	if (ref_boxes_out) // kept for refits: triangles that do not move keep their clipped boxes (k_refit)
	{
		float *o = ref_boxes_out + size_t(i) * 6;
		o[0] = b.lo[0], o[1] = b.lo[1], o[2] = b.lo[2], o[3] = b.hi[0], o[4] = b.hi[1], o[5] = b.hi[2];
	}
}

__global__ void __launch_bounds__(256) k_lbvh_inner(const Lbvh2View t)
{
	const int i = int(blockIdx.x * blockDim.x + threadIdx.x);
	if (i < t.n - 1)
		lb_build_inner(t, i);
}

__global__ void __launch_bounds__(256) k_lbvh_boxes(const Lbvh2View t)
{
	const int i = int(blockIdx.x * blockDim.x + threadIdx.x);
	if (i >= t.n)
		return;
	int cur = t.parent_leaf[i];
	while (cur >= 0)
	{
		__threadfence();
		if (atomicAdd(&t.arrivals[cur], 1u) == 0u)
			return; // the sibling subtree is not done yet: its last thread continues from here
		__threadfence();
		const int32_t l = t.left[cur], r = t.right[cur];
		const volatile LbvhBox *bl = l < 0 ? &t.leaf_box[~l] : &t.inner_box[l];
		const volatile LbvhBox *br = r < 0 ? &t.leaf_box[~r] : &t.inner_box[r];
		LbvhBox u;
		for (int k = 0; k < 3; k++)
			u.lo[k] = fminf(bl->lo[k], br->lo[k]), u.hi[k] = fmaxf(bl->hi[k], br->hi[k]);
		u.pad0 = u.pad1 = 0.0f;
		t.inner_box[cur] = u;
		cur = t.parent_inner[cur];
	}
}

// ---- builder=ploc: parallel locally-ordered clustering (lbvh.h ploc_*), one launch per step and round -----------------------
__global__ void __launch_bounds__(256) k_ploc_init(int32_t *__restrict__ id, LbvhBox *__restrict__ box, const LbvhBox *__restrict__ leaf_box, int n)
{
	const int i = int(blockIdx.x * blockDim.x + threadIdx.x);
	if (i < n)
		id[i] = ~i, box[i] = leaf_box[i];
}
__global__ void __launch_bounds__(256) k_ploc_nearest(const PlocRound p)
{
	const int i = int(blockIdx.x * blockDim.x + threadIdx.x);
	if (i < p.c)
		ploc_nearest(p, i, PLOC_RADIUS);
}
__global__ void __launch_bounds__(256) k_ploc_merge(const PlocRound p, const Lbvh2View t, int32_t *counts, uint32_t *made)
{
	const int i = int(blockIdx.x * blockDim.x + threadIdx.x);
	if (i < p.c)
		ploc_merge(p, t, counts, i, [made] __device__() { return atomicAdd(made, 1u); });
}
__global__ void __launch_bounds__(256) k_ploc_compact(const PlocRound p, const uint32_t *__restrict__ offsets, int32_t *__restrict__ id_next,
													  LbvhBox *__restrict__ box_next)
{
	const int i = int(blockIdx.x * blockDim.x + threadIdx.x);
	if (i < p.c && p.keep[i])
		id_next[offsets[i]] = p.out_id[i], box_next[offsets[i]] = p.out_box[i];
}
__global__ void __launch_bounds__(256) k_ploc_positions(const Lbvh2View t, const int32_t *__restrict__ counts, int32_t *__restrict__ position)
{
	const int i = int(blockIdx.x * blockDim.x + threadIdx.x);
	if (i < t.n)
		position[i] = ploc_leaf_position(t, counts, i);
}
__global__ void __launch_bounds__(256) k_ploc_ranges(const Lbvh2View t, const int32_t *__restrict__ counts, const int32_t *__restrict__ position)
{
	const int k = int(blockIdx.x * blockDim.x + threadIdx.x);
	if (k < t.n - 1)
		ploc_node_range(t, counts, position, k);
}
// references to depth-first order (the leaf encoding needs a subtree's references contiguous), leaf children relabelled
__global__ void __launch_bounds__(256) k_ploc_permute(const Lbvh2View t, const int32_t *__restrict__ position, const uint32_t *__restrict__ tri_order_in,
													  LbvhBox *__restrict__ leaf_box_out, uint32_t *__restrict__ tri_order_out,
													  int32_t *__restrict__ parent_leaf_out, float *__restrict__ ref_boxes_out)
{
	const int i = int(blockIdx.x * blockDim.x + threadIdx.x);
	if (i >= t.n)
		return;
	const int32_t q = position[i];
	const LbvhBox b = t.leaf_box[i];
	leaf_box_out[q] = b, tri_order_out[q] = tri_order_in[i], parent_leaf_out[q] = t.parent_leaf[i];
	if (ref_boxes_out)
	{
		float *o = ref_boxes_out + size_t(q) * 6;
		o[0] = b.lo[0], o[1] = b.lo[1], o[2] = b.lo[2], o[3] = b.hi[0], o[4] = b.hi[1], o[5] = b.hi[2];
	}
	if (i < t.n - 1)
	{
		const int32_t l = t.left[i], r = t.right[i];
		if (l < 0)
			t.left[i] = ~position[~l];
		if (r < 0)
			t.right[i] = ~position[~r];
	}
}

__global__ void __launch_bounds__(128) k_lbvh_collapse(const Lbvh2View t, const LbvhPending *__restrict__ queue, uint32_t count, uint32_t level_base,
													   uint32_t next_level_base, BvhNode4 *nodes, uint32_t *parent_slot, LbvhPending *next_queue,
													   uint32_t *next_count)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	lb_collapse_node(t, queue[i], level_base + i, next_level_base, nodes, parent_slot, next_queue,
					 [next_count] __device__(uint32_t k) { return atomicAdd(next_count, k); });
}

// scratch for `refs` references over `tris` triangles (refs >= tris)
size_t lbvh_scratch_bytes(size_t tris, size_t refs)
{
	size_t sort_tmp = 0, scan_tmp = 0;
	cub::DeviceRadixSort::SortKeys(nullptr, sort_tmp, (const uint64_t *)nullptr, (uint64_t *)nullptr, int(refs), 0, 62);
	cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (const uint32_t *)nullptr, (uint32_t *)nullptr, int(refs));
	auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
	return al(sort_tmp) + al(scan_tmp) + 2 * al(refs * 8) + al(tris * sizeof(LbvhBox)) + 3 * al(refs * sizeof(LbvhBox)) + 8 * al(refs * 4) +
		   2 * al(tris * 4) + 2 * al(refs * sizeof(LbvhPending)) + al(64) + 8192 +
		   2 * al(refs * sizeof(LbvhBox)) + 8 * al(refs * 4); // builder=ploc: cluster ids / boxes (double-buffered), partner, keep, offsets, counts, positions, parent links
}

// Builds nodes[0..*node_count), tri_order[0..*ref_count), parent_slot[0..*node_count) (and ref_boxes_out, 6 floats per
// reference, when given) on `stream`; synchronises once per tree level to read the size of the next level, and once or
// twice more when `presplit` is on (scene extent, reference count).
cudaError_t lbvh_build(const GeometryView &g, int presplit_and_flags, void *scratch, size_t scratch_bytes, size_t ref_capacity, BvhNode4 *nodes,
					   uint32_t *tri_order, uint32_t *parent_slot, float *ref_boxes_out, uint32_t *node_count, uint32_t *ref_count, int *depth,
					   int *launches, cudaStream_t stream)
{
	const int presplit = presplit_and_flags & 1;
	const bool ploc = (presplit_and_flags & 2) != 0;
	const uint32_t nt = g.flat_count;
	*node_count = 0, *ref_count = 0, *depth = 0;
	if (nt == 0 || ref_capacity < nt || scratch_bytes < lbvh_scratch_bytes(nt, ref_capacity))
		return cudaErrorInvalidValue;
	auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
	char *p = static_cast<char *>(scratch);
	auto take = [&](size_t b) {
		char *r = p;
		p += al(b);
		return r;
	};
	const size_t cap = ref_capacity;
	size_t sort_tmp = 0, scan_tmp = 0;
	cub::DeviceRadixSort::SortKeys(nullptr, sort_tmp, (const uint64_t *)nullptr, (uint64_t *)nullptr, int(cap), 0, 62);
	cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (const uint32_t *)nullptr, (uint32_t *)nullptr, int(cap));
	void *d_sort = take(sort_tmp), *d_scan = take(scan_tmp);
	uint64_t *keys_a = reinterpret_cast<uint64_t *>(take(cap * 8)), *keys_b = reinterpret_cast<uint64_t *>(take(cap * 8));
	LbvhBox *boxes = reinterpret_cast<LbvhBox *>(take(size_t(nt) * sizeof(LbvhBox)));
	LbvhBox *ref_box = reinterpret_cast<LbvhBox *>(take(cap * sizeof(LbvhBox)));
	Lbvh2View t{};
	t.leaf_box = reinterpret_cast<LbvhBox *>(take(cap * sizeof(LbvhBox)));
	t.inner_box = reinterpret_cast<LbvhBox *>(take(cap * sizeof(LbvhBox)));
	t.left = reinterpret_cast<int32_t *>(take(cap * 4)), t.right = reinterpret_cast<int32_t *>(take(cap * 4));
	t.first = reinterpret_cast<int32_t *>(take(cap * 4)), t.last = reinterpret_cast<int32_t *>(take(cap * 4));
	t.parent_inner = reinterpret_cast<int32_t *>(take(cap * 4)), t.parent_leaf = reinterpret_cast<int32_t *>(take(cap * 4));
	t.arrivals = reinterpret_cast<uint32_t *>(take(cap * 4));
	uint32_t *ref_tri = reinterpret_cast<uint32_t *>(take(cap * 4));
	uint32_t *counts = reinterpret_cast<uint32_t *>(take(size_t(nt) * 4)), *offsets = reinterpret_cast<uint32_t *>(take(size_t(nt) * 4));
	LbvhPending *queue_a = reinterpret_cast<LbvhPending *>(take(cap * sizeof(LbvhPending)));
	LbvhPending *queue_b = reinterpret_cast<LbvhPending *>(take(cap * sizeof(LbvhPending)));
	int32_t *bounds = reinterpret_cast<int32_t *>(take(64)); // 6 ordered ints + the level counter + the clustering's node counter
	uint32_t *next_count = reinterpret_cast<uint32_t *>(bounds + 8);
	uint32_t *made = reinterpret_cast<uint32_t *>(bounds + 9);
	const uint32_t tblocks = (nt + 255u) / 256u;
	cudaError_t e;
	const int32_t init[12] = {0x7fffffff, 0x7fffffff, 0x7fffffff, int32_t(0x80000000), int32_t(0x80000000), int32_t(0x80000000), 0, 0, 0, 0, 0, 0};
	if ((e = cudaMemcpyAsync(bounds, init, sizeof(init), cudaMemcpyHostToDevice, stream)) != cudaSuccess)
		return e;
	k_lbvh_bounds<<<tblocks, 256, 0, stream>>>(g, boxes, bounds);
	*launches += 1;
	uint32_t n = nt;
	bool split = false;
	if (presplit && cap > nt)
	{
		int32_t hb[6];
		if ((e = cudaMemcpyAsync(hb, bounds, sizeof(hb), cudaMemcpyDeviceToHost, stream)) != cudaSuccess)
			return e;
		if ((e = cudaStreamSynchronize(stream)) != cudaSuccess)
			return e;
		float ext = 0.0f;
		for (int a = 0; a < 3; a++)
			ext = fmaxf(ext, ordered_to_float(hb[3 + a]) - ordered_to_float(hb[a]));
		float cell = ext * (1.0f / 64.0f);
		for (int attempt = 0; attempt < 8 && cell > 0.0f; attempt++, cell *= 2.0f)
		{
			k_lbvh_counts<<<tblocks, 256, 0, stream>>>(boxes, nt, cell, counts);
			if ((e = cub::DeviceScan::ExclusiveSum(d_scan, scan_tmp, counts, offsets, int(nt), stream)) != cudaSuccess)
				return e;
			*launches += 2;
			uint32_t last_off = 0, last_cnt = 0;
			if ((e = cudaMemcpyAsync(&last_off, offsets + (nt - 1), 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess)
				return e;
			if ((e = cudaMemcpyAsync(&last_cnt, counts + (nt - 1), 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess)
				return e;
			if ((e = cudaStreamSynchronize(stream)) != cudaSuccess)
				return e;
			if (size_t(last_off) + last_cnt <= cap)
			{
				n = last_off + last_cnt, split = true;
				break;
			}
		}
	}
	k_lbvh_emit<<<tblocks, 256, 0, stream>>>(g, boxes, split ? counts : nullptr, split ? offsets : nullptr, ref_box, ref_tri);
	const uint32_t blocks = (n + 255u) / 256u;
	k_lbvh_keys<<<blocks, 256, 0, stream>>>(ref_box, n, bounds, keys_a);
	if ((e = cub::DeviceRadixSort::SortKeys(d_sort, sort_tmp, keys_a, keys_b, int(n), 0, 62, stream)) != cudaSuccess)
		return e;
	t.keys = keys_b, t.n = int32_t(n);
	k_lbvh_gather<<<blocks, 256, 0, stream>>>(keys_b, ref_box, ref_tri, n, t.leaf_box, tri_order, ref_boxes_out);
	*launches += 3 + 3; // + the sort's passes, roughly
	LbvhPending root{n == 1 ? ~0 : 0, 0xffffffffu};
	if (n > 1 && ploc)
	{
		int32_t *id_a = reinterpret_cast<int32_t *>(take(cap * 4)), *id_b = reinterpret_cast<int32_t *>(take(cap * 4));
		LbvhBox *box_a = reinterpret_cast<LbvhBox *>(take(cap * sizeof(LbvhBox))), *box_b = reinterpret_cast<LbvhBox *>(take(cap * sizeof(LbvhBox)));
		int32_t *nn = reinterpret_cast<int32_t *>(take(cap * 4));
		uint32_t *keep = reinterpret_cast<uint32_t *>(take(cap * 4)), *offs = reinterpret_cast<uint32_t *>(take(cap * 4));
		int32_t *cnts = reinterpret_cast<int32_t *>(take(cap * 4)), *position = reinterpret_cast<int32_t *>(take(cap * 4));
		int32_t *parent_leaf2 = reinterpret_cast<int32_t *>(take(cap * 4));
		k_ploc_init<<<blocks, 256, 0, stream>>>(id_a, box_a, t.leaf_box, int(n));
		*launches += 1;
		// out_id / out_box of a round: the round's own scratch (id_b / box_b); the compaction writes back into id_a / box_a only
		// after every thread of the merge kernel has read them (separate launches)
		int c = int(n), rounds = 0;
		while (c > 1)
		{
			const PlocRound pr{c, id_a, box_a, nn, id_b, box_b, keep};
			const uint32_t cb = (uint32_t(c) + 255u) / 256u;
			k_ploc_nearest<<<cb, 256, 0, stream>>>(pr);
			k_ploc_merge<<<cb, 256, 0, stream>>>(pr, t, cnts, made);
			if ((e = cub::DeviceScan::ExclusiveSum(d_scan, scan_tmp, keep, offs, c, stream)) != cudaSuccess)
				return e;
			k_ploc_compact<<<cb, 256, 0, stream>>>(pr, offs, id_a, box_a);
			*launches += 5;
			uint32_t last_off = 0, last_keep = 0;
			if ((e = cudaMemcpyAsync(&last_off, offs + (c - 1), 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess)
				return e;
			if ((e = cudaMemcpyAsync(&last_keep, keep + (c - 1), 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess)
				return e;
			if ((e = cudaStreamSynchronize(stream)) != cudaSuccess)
				return e;
			const int c2 = int(last_off + last_keep);
			if (c2 >= c || ++rounds > 4096)
				return cudaErrorUnknown; // a round always merges the globally closest pair
			c = c2;
		}
		k_ploc_positions<<<blocks, 256, 0, stream>>>(t, cnts, position);
		k_ploc_ranges<<<blocks, 256, 0, stream>>>(t, cnts, position);
		if ((e = cudaMemcpyAsync(ref_tri, tri_order, size_t(n) * 4, cudaMemcpyDeviceToDevice, stream)) != cudaSuccess)
			return e;
		k_ploc_permute<<<blocks, 256, 0, stream>>>(t, position, ref_tri, ref_box, tri_order, parent_leaf2, ref_boxes_out);
		t.leaf_box = ref_box, t.parent_leaf = parent_leaf2;
		*launches += 3;
	}
	else if (n > 1)
	{
		if ((e = cudaMemsetAsync(t.arrivals, 0, size_t(n) * 4, stream)) != cudaSuccess)
			return e;
		k_lbvh_inner<<<blocks, 256, 0, stream>>>(t);
		k_lbvh_boxes<<<blocks, 256, 0, stream>>>(t);
		*launches += 2;
	}
	if ((e = cudaMemcpyAsync(queue_a, &root, sizeof(root), cudaMemcpyHostToDevice, stream)) != cudaSuccess)
		return e;
	uint32_t level_base = 0, count = 1;
	LbvhPending *qin = queue_a, *qout = queue_b;
	int levels = 0;
	while (count > 0)
	{
		if (size_t(level_base) + count > cap)
			return cudaErrorMemoryAllocation;
		if ((e = cudaMemsetAsync(next_count, 0, 4, stream)) != cudaSuccess)
			return e;
		k_lbvh_collapse<<<(count + 127u) / 128u, 128, 0, stream>>>(t, qin, count, level_base, level_base + count, nodes, parent_slot, qout, next_count);
		*launches += 1;
		uint32_t next = 0;
		if ((e = cudaMemcpyAsync(&next, next_count, 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess)
			return e;
		if ((e = cudaStreamSynchronize(stream)) != cudaSuccess)
			return e;
		level_base += count, count = next, levels++;
		LbvhPending *tmp = qin;
		qin = qout, qout = tmp;
		if (levels > 64)
			return cudaErrorUnknown;
	}
	*node_count = level_base, *ref_count = n, *depth = levels;
	return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
cudaError_t launch_refit(const GeometryView &g, cudaStream_t stream)
{
	if (g.node_count == 0)
		return cudaSuccess;
	cudaError_t e = cudaMemsetAsync(g.arrivals, 0, size_t(g.node_count) * sizeof(uint32_t), stream);
	if (e != cudaSuccess)
		return e;
	const uint32_t threads = g.node_count * 4u;
	k_refit<<<(threads + 255u) / 256u, 256, 0, stream>>>(g);
	return cudaGetLastError();
}

cudaError_t launch_pack_nodes(const BvhNode4 *nodes, BvhNode4Packed *out, uint32_t n, cudaStream_t stream)
{
	if (n == 0)
		return cudaSuccess;
	k_pack_nodes<<<(n + 255u) / 256u, 256, 0, stream>>>(nodes, out, n);
	return cudaGetLastError();
}

cudaError_t launch_records(const GeometryView &g, cudaStream_t stream)
{
	if (g.ref_count == 0)
		return cudaSuccess;
	k_records<<<(g.ref_count + 255u) / 256u, 256, 0, stream>>>(g);
	return cudaGetLastError();
}

cudaError_t launch_flatten_shade(const GeometryView &g, cudaStream_t stream)
{
	if (g.flat_count == 0)
		return cudaSuccess;
	k_flatten_shade<<<(g.flat_count + 255u) / 256u, 256, 0, stream>>>(g);
	return cudaGetLastError();
}

cudaError_t launch_morph(const MorphView &m, cudaStream_t stream)
{
	if (m.vertex_count)
		k_morph_vertices<<<(m.vertex_count + 127u) / 128u, 128, 0, stream>>>(m);
	if (m.triangle_count)
	{
		SkinView s{};
		s.indices = m.indices, s.out_vertices = m.out_vertices, s.out_normals = m.out_normals, s.mesh_tris = m.mesh_tris;
		s.vertex_count = m.vertex_count, s.triangle_count = m.triangle_count;
		k_update_triangles<<<(m.triangle_count + 127u) / 128u, 128, 0, stream>>>(s);
	}
	return cudaGetLastError();
}

cudaError_t launch_skin(const SkinView &s, cudaStream_t stream)
{
	if (s.vertex_count)
		k_skin_vertices<<<(s.vertex_count + 127u) / 128u, 128, 0, stream>>>(s);
	if (s.triangle_count)
		k_update_triangles<<<(s.triangle_count + 127u) / 128u, 128, 0, stream>>>(s);
	return cudaGetLastError();
}

} // namespace rfwb200
