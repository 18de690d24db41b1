// geometry.h — device-side geometry pipeline (geometry.cu): instance flattening, BVH refit and skinning on the GPU
#pragma once
#include "device_types.h"

#include <cuda_runtime.h>

namespace rfwb200
{

// one instance of a mesh inside the device geometry arena, 128 B
struct alignas(16) DeviceInstance
{
	float transform[16]; // column-major mat4 (rfw::RenderContext::set_instance, context.h:91)
	float normal[9];	 // column-major normal matrix
	float det_eps;		 // 1e-6 * |det(linear part)| (TriRec::det_eps)
	uint32_t vert_off;	 // first vertex of the mesh in GeometryView::verts
	uint32_t tri_off;	 // first triangle of the mesh in GeometryView::mesh_tris / indices
	uint32_t flat_off;	 // first flattened triangle of this instance
	uint32_t tri_count;
	uint32_t moved; // 1: transform or mesh changed since the last build (refit bounds its triangles whole)
	uint32_t pad;
};
static_assert(sizeof(DeviceInstance) == 128, "");

struct GeometryView
{
	// arena (all meshes back to back)
	const float4 *verts;	 // mesh-local vertices
	const uint32_t *indices; // 3 per mesh triangle, mesh-local
	const void *mesh_tris;	 // rfwb200_triangle[ ] (160 B)
	const DeviceInstance *instances;
	const uint32_t *flat_inst; // flattened triangle -> instance
	uint32_t flat_count;
	// BVH topology (static between builds)
	BvhNode4 *nodes;
	const uint32_t *tri_order;	 // leaf-ordered reference -> flattened triangle
	const uint32_t *parent_slot; // (parent << 2) | slot, root 0xffffffff
	const float *ref_boxes;		 // 6 floats per leaf reference: the builder's box (null: bound every triangle whole)
	uint32_t *arrivals;			 // one counter per node (zeroed by launch_refit)
	uint32_t node_count, ref_count;
	int write_boxes; // 1 = refit (recompute every box), 0 = only the records (fresh build: keep the builder's boxes)
	// outputs
	TriRec *out_tris;	 // [ref_count]
	ShadeTri *out_shade; // [flat_count]
};

struct SkinView
{
	const float4 *base_vertices, *base_normals; // bind pose
	const uint4 *joints;
	const float4 *weights;
	const float *joint_matrices; // mat4 column-major per joint
	const uint32_t *indices;	 // of this mesh
	float4 *out_vertices;		 // the mesh's slice of the arena
	float4 *out_normals;		 // scratch, vertex_count
	void *mesh_tris;			 // the mesh's slice of the arena's triangle records
	uint32_t vertex_count, triangle_count;
};

cudaError_t launch_refit(const GeometryView &g, cudaStream_t stream);
// GPU BVH build (lbvh.h): nodes / tri_order / parent_slot (and the per-reference boxes for later refits) from the
// geometry arena; outputs are sized for ref_capacity references (>= triangles; > triangles enables early split clipping);
// *launches is incremented
size_t lbvh_scratch_bytes(size_t triangles, size_t ref_capacity);
cudaError_t lbvh_build(const GeometryView &g, int presplit_and_flags /* bit 0: early split clipping, bit 1: clustering (builder=ploc) */, void *scratch, size_t scratch_bytes, size_t ref_capacity, BvhNode4 *nodes,
					   uint32_t *tri_order, uint32_t *parent_slot, float *ref_boxes_out, uint32_t *node_count, uint32_t *ref_count, int *depth,
					   int *launches, cudaStream_t stream);
cudaError_t launch_pack_nodes(const BvhNode4 *nodes, BvhNode4Packed *out, uint32_t n, cudaStream_t stream);
cudaError_t launch_records(const GeometryView &g, cudaStream_t stream); // TriRec only (bvh=8)
cudaError_t launch_flatten_shade(const GeometryView &g, cudaStream_t stream);
cudaError_t launch_skin(const SkinView &s, cudaStream_t stream);

// morph targets (gltf/mesh.cpp:126-148): v = pose0 + sum_j w_j * pose_j, n likewise (not renormalised), then triangle update
struct MorphView
{
	const float4 *pose_positions; // [(n_weights + 1)][vertex_count], xyz
	const float4 *pose_normals;	  // [(n_weights + 1)][vertex_count], xyz
	const float *weights;		  // [n_weights]
	uint32_t n_weights;
	const uint32_t *indices;
	float4 *out_vertices, *out_normals;
	void *mesh_tris;
	uint32_t vertex_count, triangle_count;
};
cudaError_t launch_morph(const MorphView &m, cudaStream_t stream);

} // namespace rfwb200
