// device_types.h — HBM-resident data layouts shared by the host side (context.cpp, bvh_build.cpp)
// and the kernels (kernels.cu).  See DESIGN.md "Data layout in HBM".
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <vector_types.h>

namespace rfwb200
{

// 4-wide BVH node, 128 B = one L2 line.  SoA child boxes like the reference's MBVHNode
// (RFW/system/bvh/include/bvh/mbvh_node.h:60-106) so one slab test is four independent FMAs per
// plane, but with a single signed child word: >= 0 inner-node index, < 0 leaf with
// ~child = (first_triangle << 2) | (count - 1).  Unused slots carry a NaN box (every comparison of the
// hit test is false); pad[0] is the number of used slots.
// Nodes are stored breadth-first, so the top of the tree is one contiguous prefix that a CTA can
// stage into shared memory with a single bulk (TMA) copy.
struct alignas(16) BvhNode4
{
	float minx[4], maxx[4], miny[4], maxy[4], minz[4], maxz[4];
	int32_t child[4];
	int32_t pad[4];
};
static_assert(sizeof(BvhNode4) == 128, "node must be one 128-byte line");

// The same node in 80 B = five 16-byte pieces (instead of eight), the form the wavefront trace kernel reads by default:
// ncu shows the kernel bound by L1 wavefronts — a divergent 128-bit load costs about the same whatever it returns
// (profiles/r01b: 6.4 wavefronts per LDG.128 at 16 active lanes) — so what counts is the NUMBER of loads per node visit,
// 7 for BvhNode4.  Child planes are stored relative to the node's min corner `p` as bfloat16 (the upper half of a
// float: one shift or mask rebuilds the float), rounded outwards, so the packed box always contains the exact one and
// the closest hit is unchanged.  Derived from BvhNode4 by k_pack_nodes (geometry.cu) after every build / refit.
struct alignas(16) BvhNode4Packed
{
	float p[3];
	uint32_t pad;
	uint32_t plane[3][4]; // per axis: lo0|lo1<<16, lo2|lo3<<16, hi0|hi1<<16, hi2|hi3<<16
	int32_t child[4];
};
static_assert(sizeof(BvhNode4Packed) == 80, "packed node is 5 x 16 bytes");

// Intersection record of one (instance-flattened, world-space) triangle, 48 B, stored in BVH leaf
// order so a leaf is one contiguous run: p0 and the two edges Moller-Trumbore needs
// (the reference precomputes the same triple on the CPU, RFW/system/bvh/src/bvh_tree.cpp:413-415).
struct alignas(16) TriRec
{
	float p0x, p0y, p0z, e1x;
	float e1y, e1z, e2x, e2y;
	float e2z;
	uint32_t shade_idx; // index into ShadeTri[]
	float det_eps;		// T_EPSILON * |det(instance linear part)|: the reference tests the determinant in
						// object space (CUDART/src/Kernels.cu:269-271 + CUDAIntersect.h:61-63)
	uint32_t pad0;
};
static_assert(sizeof(TriRec) == 48, "triangle record is 3 x 16 bytes");

// What shade needs from rfw::DeviceTriangle (160 B, device_structs.h:22-34), repacked to 96 B with
// the instance's normal matrix already applied to the four normals (the reference applies it per
// path, CUDART/src/getShadingData.h:129-130).
struct alignas(16) ShadeTri
{
	float u0, u1, u2;
	int32_t light_tri_idx;
	float v0, v1, v2;
	uint32_t material;
	float n0x, n0y, n0z, Nx; // normal-matrix * vN0 (not normalised), world geometric normal x
	float n1x, n1y, n1z, Ny;
	float n2x, n2y, n2z, Nz;
	float area, lod;
	uint32_t inst_id, prim_id;
};
static_assert(sizeof(ShadeTri) == 96, "shading triangle is 6 x 16 bytes");

// per (wavefront batch, depth) queue sizes and per-launch work cursors; zeroed once per frame.
struct DepthCounters
{
	uint32_t ext;		  // extension rays emitted by shade(depth)          (counters->extensionRays)
	uint32_t shadow;	  // shadow rays emitted by shade(depth)             (counters->shadowRays)
	uint32_t trace_cursor; // persistent-thread work cursor of the trace launch consuming them
	uint32_t shade_cursor; // work cursor of the shade launch at this depth
	uint32_t acc;		  // accumulator read-modify-writes at this depth
	uint32_t shadow_traced; // shadow rays actually traced by connect (reference drops some)
	uint32_t move_cursor;  // work cursor of the re-ordering pass in front of the trace launch at this depth
	uint32_t pad1;
};
static_assert(sizeof(DepthCounters) == 32, "");
static_assert(offsetof(DepthCounters, ext) % 8 == 0 && offsetof(DepthCounters, shadow) == offsetof(DepthCounters, ext) + 4,
			  "k_shade advances ext and shadow with one 64-bit atomic");

constexpr int MAX_DEPTH_SLOTS = 8; // max_path_length + 1 <= 8
constexpr int MAX_BATCH_SPP = 64;  // samples of a frame that travel in one wavefront (BatchView)

// One wavefront carries `spp` samples of every local pixel.  Work item (default, sample_minor = 1) = block * 32 * spp +
// pixel_in_block * spp + s, where block is an 8x4-pixel block of the shard (32 consecutive local pixels): the samples of a
// pixel are neighbours, so a warp of camera rays is 32 / spp pixels x spp samples — rays that differ by a sub-pixel jitter
// and walk the same nodes (measured against one sample of 32 pixels per warp, sample_minor = 0: work item =
// ((block * spp + s) << 5) | lane: camera-ray launch -7 %, frame -2 %, bit-identical frames).  A frame is 3 + 4 * bounces
// launches however many samples it has.  The reference renders one sample per render_frame call
// (CUDART/src/Context.cpp:75-80,149).
struct BatchView
{
	uint32_t spp;		   // samples in this wavefront (<= MAX_BATCH_SPP)
	uint32_t first_sample; // index within the frame of its first sample
	uint32_t index;		   // wavefront number within the frame: selects the DepthCounters row
	uint32_t items;		   // local_pixels * spp
	float inv_spp;		   // 1 / spp (fast_div)
	uint32_t sample_minor; // 1 (setting sample_layout=pixel): item = block * 32 * spp + pixel_in_block * spp + s instead — the
						   // samples of a pixel sit next to each other, a warp of camera rays is 32 / spp pixels x spp samples
};

// Display target of a tile-sharded frame: ONE row-major image in the display rank's memory that the fold
// kernel of every rank writes its pixels into — peer stores over NVLink for the other ranks — followed by one arrival per
// rank and frame on a counter beside it.  No collective and no de-tiling pass.
struct DisplayTarget
{
	float4 *image;		  // float4[width * height], row-major (local or peer-mapped); null = no display target
	uint32_t *arrivals;	  // in the display rank's memory: += 1 by every rank when its tiles of a frame are written
	uint32_t *local_done; // this rank's CTA counter for the last-CTA-signals pattern
};

// What the ranks of a sharded frame agree on inside a frame (all of it in the display rank's memory, beside the image): whether
// sample s of the wavefront emitted an extension ray at depth d ANYWHERE in the frame — the reference's host loop leaves a
// sample's bounce loop, without tracing its pending shadow rays, as soon as a bounce emits none (CUDART/src/Context.cpp:109-120),
// and a rank only sees its own tiles.  A flag is set by writing the wavefront's stamp (no clearing between frames).
struct ShardSync
{
	uint32_t *seen;		// [MAX_DEPTH_SLOTS][MAX_BATCH_SPP]; null = this context renders alone (or has no display target)
	uint32_t *arrivals; // += 1 by every rank after each shade launch below the last depth
	uint32_t stamp;		// value that means "set" for the wavefront in flight
};

// uniform grid over the scene box the bounce rays are binned in before they are traced (k_shade emits the key)
struct SortGrid
{
	float lo[3];
	float scale[3]; // cells per unit length
};

struct ProbeResult
{
	int32_t inst, prim;
	float dist;
	uint32_t pad;
};

// constant per frame; lives in device memory so a captured CUDA graph stays valid across frames
struct FrameParams
{
	float pos[3], p1[3], right[3], up[3]; // CameraView with right = p2-p1, up = p3-p1
	float aperture, spread_angle;
	uint32_t sample_base; // sample index of the first sample of this frame
	uint32_t probe_pixel; // global pixel id, 0xffffffff = none
	float to_eye[9];	  // column-major mat3 applied to the normal AOV (OptiX6Context: transpose(inverse(camera matrix))); identity by default
};

// Two-level scene (setting "levels" = 2, or "auto" when the flattened triangle count exceeds "flatten_budget"): instances
// stay instances — a top-level tree over their world boxes whose leaves name ONE instance each, and one object-space tree
// per mesh, all in SceneView::nodes (top-level nodes first).  The reference traverses exactly this shape
// (CUDART/src/Kernels.cu:226-303: ray moved into object space with the instance's inverse transform, un-normalised, so
// distances stay world distances) and applies the normal matrix at shading time (getShadingData.h:129-130).
struct alignas(16) TlInstance
{
	float inv[12];		// object = inv * world: rows (m00 m01 m02 m03), (m10 ...), (m20 ...)
	float normal[9];	// column-major normal matrix
	uint32_t blas_root; // node index of the tree below this instance
	uint32_t pad[2];	// pad[0]: first entry of this instance's row in SceneView::tl_inst_map
};
static_assert(sizeof(TlInstance) == 96, "");

struct LightCounts
{
	uint32_t area, point, spot, directional;
};

// everything a kernel needs; passed by value (fits the 4 KB parameter space easily)
struct SceneView
{
	const BvhNode4 *nodes;
	const TriRec *tris;
	const ShadeTri *shade_tris;
	uint32_t node_count, tri_count;
	const uint4 *nodes16;  // BvhNode4Packed[node_count] (5 x 16 B per node), derived from `nodes`
	const uint4 *cw_nodes; // compressed 8-wide BVH (cwbvh.h, 5 x 16 B per node) when setting bvh=8, else null
	uint32_t cw_node_count;
	const void *materials; // rfwb200_material[ ] (192 B, texaddr patched)
	uint32_t material_count;
	const uint32_t *uint_texels;
	uint32_t uint_texel_count;
	const float *float_texels; // float4
	uint32_t float_texel_count;
	const float *sky; // float4 per texel (rgb, 0)
	uint32_t sky_w, sky_h;
	const void *area_lights, *point_lights, *spot_lights, *dir_lights;
	LightCounts lights;
	const uint8_t *blue_noise; // 327,680 table bytes at the offsets of createBlueNoiseBuffer (blue_noise.h:8204-8219)
	// two-level scene (null = flattened, the default): `tris` / `shade_tris` are then per MESH triangle in object space and a
	// hit carries the instance it was found in
	const TlInstance *tl_instances;
	uint32_t tl_instance_count;
	const uint32_t *tl_inst_map; // [tl_instances[i].pad[0] + ShadeTri::inst_id (rank of the triangle's mesh in its group)] -> the caller's instance index
};

struct ShardView
{
	uint32_t width, height;		// full image
	uint32_t rank, world;		// this context's shard
	uint32_t tile_w, tile_h;	// multiples of 8 / 4
	uint32_t tiles_x, tiles_y;	// tile grid of the full image
	uint32_t local_tiles;		// tiles owned by this rank
	uint32_t local_pixels;		// local_tiles * tile_w * tile_h (padded: edge tiles carry dead pixels)
	float inv_tile_pixels, inv_tiles_x, inv_blocks_per_row; // reciprocals for the index divisions of local_to_pixel (fast_div)
};

struct WavefrontView
{
	float4 *O[2], *D[2], *T[2]; // ray planes: origin|pathIdx<<8|flags, direction|packed normal, throughput|pdf.  With re-ordering
								// on, [0] is the queue the trace / shade launches read and [1] the staging queue shade appends
								// to; with it off the two alternate per depth like Kernels.cu:578-584
	float4 *hit;				// bits(u16|v16<<16), bits(shade_idx), bits(prim or -1), t
	float4 *sO, *sD, *sE;		// connect queue: origin, direction|tmax, contribution|bits(path index)
	float4 *sample_acc;			// per work item (= per pixel and sample of the wavefront): radiance of that one sample
	float4 *accumulator;		// per local pixel: sum over all samples since the last Reset (k_fold adds the wavefront's samples in order)
	float4 *framebuffer;		// finalised
	DepthCounters *counters;	// [wavefront][MAX_DEPTH_SLOTS]
	uint32_t *ext_seen;			// [wavefront][MAX_DEPTH_SLOTS][MAX_BATCH_SPP]: sample s emitted an extension ray at that depth
	ProbeResult *probe;
	uint32_t *occ_cache;		// [2][local pixels]: occluder of the pixel's previous connect ray at this depth parity (shadow_cache = 2)
	uint32_t *prim_cache;		// per local pixel: triangle record hit by the previous camera ray of that pixel (a bound only)
	const FrameParams *frame;
	// depth-0 feature planes for a denoiser (OptiX6Context/assets/kernels/kernels.cu:122-133,206-221,316-330), setting "aov":
	// per work item like sample_acc, folded into per-pixel sums by k_fold; null when the setting is off
	float4 *sample_albedo, *sample_normal, *albedo_acc, *normal_acc;
	// re-ordering of the bounce queue (counting sort by origin cell + direction octant)
	uint2 *sort_key;	 // per staging slot: (bin, rank inside the bin)
	uint32_t *sort_hist; // per bin: rays emitted into it (zeroed again by k_sort_scan)
	uint32_t *sort_base; // per bin: exclusive prefix inside its 4096-bin chunk
	uint32_t *sort_chunk; // per chunk: rays in the chunk
	const SortGrid *grid;
};

struct RenderSettings
{
	int max_path_length;
	float clamp_value;
	float geometry_epsilon;
	int survival_scale;
	int smem_nodes;
	int fetch_threshold; // idle lanes per warp that trigger a refill from the work cursor
	unsigned long long *debug; // optional per-warp {start ns, end ns, rays} records of one trace launch (tools/diag)
	int debug_depth;
	int trace_variant; // which instantiation of k_wavefront_trace runs (kernels.cu launch_trace; tuning)
	int primary_cache; // 1: a camera ray first tests the triangle its pixel hit last time and starts with that distance as bound
	int shadow_cache; // connect rays first test a remembered occluder: 1 = of the lane's previous connect ray, 2 = of the same pixel's
					  // previous connect ray at this depth (another sample); occlusion is yes/no, so frames do not change
	int primary_variant; // the same for camera rays (coherent: fp32 nodes measured faster there than packed ones)
	int sort_mode;		 // 1: bounce rays are re-ordered by (origin cell, direction octant) before they are traced
	int sort_cell_bits;	 // grid resolution per axis: 2^bits cells (3..6)
	int sort_dir_major;	 // 1: the octant is the most significant part of the key, 0: the cell is
	int fetch_chunk;	 // 0: idle lanes take the next entries of one shared front; > 0: a warp claims a private run of that many entries
	int sort_dir_bits;	 // direction part of the re-ordering key: 3 = octant, 5 = octant x dominant axis (24 bins)
	int shade_static;	 // 1: k_shade walks the queue in grid strides and prefetches the next job's state (no work cursor)
};
constexpr int SORT_CHUNK = 4096; // bins scanned by one CTA of k_sort_scan

} // namespace rfwb200
