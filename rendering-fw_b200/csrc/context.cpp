// context.cpp — implementation of the C ABI declared in include/rfwb200.h.
//
// Host side of the hot path: keeps a host copy of what the reference's RenderSystem uploads
// (meshes, instances, materials, textures, lights, sky), flattens instances into one world-space
// triangle soup on update() (memory is laid out for 180 GB of HBM, so instancing is traded for a
// single-level BVH — DESIGN.md), builds / refits the 4-wide BVH, owns the SoA wavefront state and
// enqueues the per-sample kernel sequence without any host synchronisation.
//
// Reference counterparts: backends/CUDART/src/Context.cpp (render_frame :65-159, set_materials
// :160-198, set_textures :201-268, set_mesh :270-311, set_instance :313-321, set_lights :331-383,
// update :394-456, buffers :462-483).
#include "../../include/rfwb200.h"

#include "bvh_build.h"
#include "device_types.h"
#include "geometry.h"
#include "lbvh.h"
#include "kernels.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

extern "C" const unsigned char rfwb200_bluenoise_table[];
extern "C" const unsigned int rfwb200_bluenoise_table_size;

using namespace rfwb200;

namespace
{
thread_local std::string g_last_error;

int set_error(int code, const std::string &msg)
{
	g_last_error = msg;
	return code;
}

#define CK(expr)                                                                                                        \
	do                                                                                                                  \
	{                                                                                                                   \
		const cudaError_t e__ = (expr);                                                                                 \
		if (e__ != cudaSuccess)                                                                                         \
			return set_error(e__ == cudaErrorMemoryAllocation ? RFWB200_ERR_OOM : RFWB200_ERR_CUDA,                     \
							 std::string(#expr) + ": " + cudaGetErrorString(e__));                                      \
	} while (0)

#define REQUIRE(cond, msg)                                                                                              \
	do                                                                                                                  \
	{                                                                                                                   \
		if (!(cond))                                                                                                    \
			return set_error(RFWB200_ERR_INVALID, msg);                                                                 \
	} while (0)

// growable device buffer
struct DevBuf
{
	void *ptr = nullptr;
	size_t bytes = 0;
	~DevBuf() { release(); }
	void release()
	{
		if (ptr)
			cudaFree(ptr);
		ptr = nullptr, bytes = 0;
	}
	cudaError_t reserve(size_t n)
	{
		if (n <= bytes)
			return cudaSuccess;
		release();
		const cudaError_t e = cudaMalloc(&ptr, std::max<size_t>(n, 16));
		if (e == cudaSuccess)
			bytes = std::max<size_t>(n, 16);
		return e;
	}
	template <typename T> T *as() const { return static_cast<T *>(ptr); }
};

struct HostMesh
{
	std::vector<float> vertices;   // vec4
	std::vector<uint32_t> indices; // 3 per tri or empty
	std::vector<rfwb200_triangle> triangles;
	uint64_t version = 0;
};

struct HostInstance
{
	int mesh = -1;
	float transform[16];
	float normal[9];
};

struct HostTexture
{
	int type;
	uint32_t width, height, texel_count, addr;
};

inline void mul_point(const float *m, const float *p, float *o)
{
	for (int r = 0; r < 3; r++)
		o[r] = m[r] * p[0] + m[4 + r] * p[1] + m[8 + r] * p[2] + m[12 + r];
}
inline void mul_mat3(const float *m, const float *p, float *o)
{
	for (int r = 0; r < 3; r++)
		o[r] = m[r] * p[0] + m[3 + r] * p[1] + m[6 + r] * p[2];
}
} // namespace

namespace
{
// Two-level scenes: the unit below the top level is a GROUP of meshes that are placed by exactly the same list of (transform,
// normal matrix) pairs (update_two_level).  placed[mesh] = that mesh's placements sorted bytewise (ties by instance index);
// groups = member meshes (ascending) of every group, groups ordered by their first member.  Entry j of every member's list is
// the same matrix: it becomes top-level instance j of the group.
struct Placement
{
	float m[25]; // transform[16] + normal[9]
	uint32_t instance;
};
void group_by_placement(const std::vector<HostInstance> &instances, const std::vector<uint8_t> &mesh_has_triangles,
						std::vector<std::vector<Placement>> &placed, std::vector<std::vector<uint32_t>> &groups)
{
	const size_t nm = mesh_has_triangles.size();
	placed.assign(nm, {});
	groups.clear();
	for (size_t ii = 0; ii < instances.size(); ii++)
	{
		const HostInstance &in = instances[ii];
		if (in.mesh < 0 || size_t(in.mesh) >= nm || !mesh_has_triangles[in.mesh])
			continue;
		Placement pl;
		memcpy(pl.m, in.transform, sizeof(float) * 16);
		memcpy(pl.m + 16, in.normal, sizeof(float) * 9);
		pl.instance = uint32_t(ii);
		placed[in.mesh].push_back(pl);
	}
	for (auto &v : placed)
		std::sort(v.begin(), v.end(), [](const Placement &x, const Placement &y) {
			const int r = memcmp(x.m, y.m, sizeof(x.m));
			return r != 0 ? r < 0 : x.instance < y.instance;
		});
	std::vector<size_t> order; // meshes that are placed at all
	for (size_t mi = 0; mi < nm; mi++)
		if (!placed[mi].empty())
			order.push_back(mi);
	auto same_list = [&](size_t x, size_t y) {
		if (placed[x].size() != placed[y].size())
			return false;
		for (size_t k = 0; k < placed[x].size(); k++)
			if (memcmp(placed[x][k].m, placed[y][k].m, sizeof(placed[x][k].m)) != 0)
				return false;
		return true;
	};
	std::sort(order.begin(), order.end(), [&](size_t x, size_t y) {
		if (placed[x].size() != placed[y].size())
			return placed[x].size() < placed[y].size();
		for (size_t k = 0; k < placed[x].size(); k++)
		{
			const int r = memcmp(placed[x][k].m, placed[y][k].m, sizeof(placed[x][k].m));
			if (r != 0)
				return r < 0;
		}
		return x < y;
	});
	for (size_t k = 0; k < order.size(); k++)
	{
		if (k == 0 || !same_list(order[k - 1], order[k]))
			groups.emplace_back();
		groups.back().push_back(uint32_t(order[k]));
	}
	std::sort(groups.begin(), groups.end(), [](const std::vector<uint32_t> &x, const std::vector<uint32_t> &y) { return x[0] < y[0]; });
}

} // namespace

// One host thread per further device of an in-process group: it enqueues that device's frame while the calling thread
// enqueues its own (a frame is ~20 asynchronous API calls per device; issued one device after the other they would
// stagger the start of eight GPUs by half a millisecond).
class DeviceWorker
{
  public:
	DeviceWorker() : th([this] { loop(); }) {}
	~DeviceWorker()
	{
		{
			std::lock_guard<std::mutex> l(m);
			quit = true;
		}
		cv.notify_all();
		th.join();
	}
	void submit(std::function<int(std::string &)> f)
	{
		{
			std::lock_guard<std::mutex> l(m);
			job = std::move(f), busy = true;
		}
		cv.notify_all();
	}
	int wait(std::string &msg)
	{
		std::unique_lock<std::mutex> l(m);
		cv.wait(l, [this] { return !busy; });
		msg = message;
		return code;
	}

  private:
	void loop()
	{
		for (;;)
		{
			std::function<int(std::string &)> f;
			{
				std::unique_lock<std::mutex> l(m);
				cv.wait(l, [this] { return quit || (busy && job); });
				if (quit)
					return;
				f = std::move(job), job = nullptr;
			}
			std::string msg;
			const int r = f(msg);
			{
				std::lock_guard<std::mutex> l(m);
				code = r, message = msg, busy = false;
			}
			cv.notify_all();
		}
	}
	std::mutex m;
	std::condition_variable cv;
	std::function<int(std::string &)> job;
	bool busy = false, quit = false;
	int code = 0;
	std::string message;
	std::thread th;
};

struct rfwb200_context
{
	int device = 0;
	cudaStream_t stream = nullptr;
	bool owns_stream = false;
	bool initialised = false;
	uint32_t width = 0, height = 0;
	ShardView shard{};

	// ---- host copies of the scene -------------------------------------------------------------
	std::vector<HostMesh> meshes;
	std::vector<HostInstance> instances;
	std::vector<rfwb200_material> materials, materials_raw;
	std::vector<rfwb200_material_tex_ids> material_tex_ids;
	std::vector<HostTexture> textures;
	bool geometry_dirty = true, topology_dirty = true;
	std::vector<std::pair<int, size_t>> built_layout; // (mesh, tri count) per instance at last build
	size_t built_tri_count = 0;
	bool spatial_splits = true;

	// ---- device scene ------------------------------------------------------------------------------
	DevBuf d_nodes, d_tris, d_shade_tris, d_materials, d_materials_raw, d_uint_tex, d_float_tex, d_tex_desc, d_sky,
		d_area, d_point, d_spot, d_dir, d_blue_noise;
	SceneView scene{};
	BvhBuildResult bvh;
	std::vector<BuildTriangle> build_tris;
	uint64_t flat_tri_count = 0;

	// ---- device geometry arena (geometry.cu): every mesh back to back + the instance table, so a refit never
	// touches the host: set_mesh uploads only the changed mesh, update() launches k_refit + k_flatten_shade ----
	DevBuf d_verts, d_indices, d_mesh_tris, d_instances, d_flat_inst, d_tri_order, d_parent_slot, d_arrivals;
	std::vector<uint32_t> mesh_vert_off, mesh_tri_off; // arena offsets per mesh
	std::vector<uint8_t> mesh_dirty;				   // host copy newer than the arena
	std::vector<uint8_t> inst_moved;				   // instance (or its mesh) changed since the last build: refits bound it whole
	std::vector<uint8_t> tri_moved;					   // the same per flattened triangle (host refit)
	DevBuf d_ref_boxes;
	bool arena_valid = false;
	bool device_geometry = true; // setting "refit" = device | host
	GeometryView geo{};
	struct Skin
	{
		DevBuf base_v, base_n, joints, weights, matrices, normals;
		size_t vertex_count = 0;
		uint32_t max_joint = 0;
		DevBuf pose_p, pose_n, morph_w; // morph targets: (n_targets + 1) poses of vertex_count float4 each
		size_t n_targets = 0;			// 0: no morph targets registered
		bool has_skin = false;
		bool device_newer = false; // the arena holds a pose the host copy of the mesh does not
	};
	std::vector<std::unique_ptr<Skin>> skins; // per mesh index, null when the mesh has no skin
	cudaEvent_t ev_geo_a = nullptr, ev_geo_b = nullptr;
	bool geo_timed = false, last_update_on_device = false, last_update_was_refit = false;
	uint64_t refits = 0, builds = 0;
	bool shade_ieee = false; // setting "shade_math" = fast | ieee
	bool wide8 = false;		 // setting "bvh" = 4 | 8 (compressed 8-wide layout, cwbvh.h)
	bool lbvh = false;		 // setting "builder" = sbvh (host, SAH + spatial splits) | lbvh | ploc (device, lbvh.h)
	bool ploc = false;		 // device builder: parallel locally-ordered clustering instead of the radix tree
	bool device_built = false; // the current tree was built on the device: the host has no copy of its topology
	bool lbvh_presplit = false; // setting "lbvh_presplit": early split clipping before the Morton sort (lbvh.h step 0); measured slower
	bool device_ref_boxes = false; // d_ref_boxes was written by the device builder
	int device_depth = 0;
	DevBuf d_lbvh_scratch;
	DevBuf d_cw_nodes, d_nodes16, d_prim_cache, d_occ_cache;
	// ---- two-level scene (TlInstance, device_types.h): setting "levels" = 1 (flatten) | 2 | auto (default) ----
	int levels_setting = 0;				   // 0 = auto (default): two levels when the flattened scene would exceed flatten_budget triangles
	uint64_t flatten_budget = 1ull << 26;  // setting "flatten_budget" (flattened triangles; 2^26 of them are ~15 GB of records and nodes)
	bool two_level = false;				   // the committed device scene is two-level
	struct GroupTree // one object-space tree over all meshes that share their list of placements (update_two_level)
	{
		std::vector<uint32_t> members;				   // mesh indices, ascending
		std::vector<uint32_t> tri_mesh_rank, tri_index; // group triangle -> (member rank, triangle of that mesh)
		BvhBuildResult bvh;
	};
	std::vector<GroupTree> group_trees; // kept across updates: a group is rebuilt only when its members or their geometry change
	std::vector<TlInstance> tl_table;	// host copies for the stage-level entry points
	std::vector<uint32_t> tl_inst_map;
	DevBuf d_tl_instances, d_tl_inst_map;
	int tl_depth = 0; // top-level depth + deepest group tree
	uint32_t tl_groups = 0;

	// ---- wavefront state -----------------------------------------------------------------------------
	// One wavefront carries up to `batch_spp` samples of every local pixel (BatchView); the planes hold local_pixels *
	// batch_spp work items.  A frame of spp samples is ceil(spp / batch_spp) wavefronts run one after the other on the
	// context's stream: 3 + 4 * bounces launches each, no host synchronisation.
	DevBuf d_debug;
	DevBuf d_O[2], d_D[2], d_T[2], d_hit, d_sO, d_sD, d_sE, d_sample_acc, d_acc, d_fb, d_counters, d_ext_seen, d_probe, d_frame,
		d_scratch_cursor;
	DevBuf d_sort_key, d_sort_hist, d_sort_base, d_sort_chunk, d_sort_grid;
	DevBuf d_sample_albedo, d_sample_normal, d_albedo, d_normal, d_aov_out; // setting "aov": depth-0 feature planes
	bool aov = false;
	float to_eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
	DevBuf d_display; // RGBA8 output of the tone-map pass (rfwb200_read_display), allocated on first use
	WavefrontView wf{};
	uint32_t counters_capacity = 0; // wavefronts per frame the counter rows are allocated for
	size_t items_capacity = 0;		// work items the wavefront planes are allocated for
	int spp_batch = 0;				// setting "spp_batch": samples per wavefront, 0 = as many as fit (<= MAX_BATCH_SPP, <= 2^24 items)
	uint32_t sort_bins_allocated = 0;
	bool sample_minor = true; // setting "sample_layout" = planes | pixel (default; BatchView::sample_minor): measured 12.63 -> 12.38 ms, profiles/r02/sweep11
	bool sort_grid_dirty = true; // the bin grid (k_sort_setup: from the root of the current tree) is computed once per scene change, not per frame

	// ---- display target of a sharded frame (DisplayTarget, device_types.h) and in-process device groups ----
	struct Display
	{
		void *base = nullptr; // [float4 image[w * h]] [tail: arrivals @ +0, consumed @ +128, sync arrivals @ +256, seen flags @ +512]
		size_t image_bytes = 0;
		bool owner = false, ipc = false;
		float4 *image = nullptr;
		uint32_t *arrivals = nullptr, *consumed = nullptr, *sync_arrivals = nullptr, *sync_seen = nullptr;
		uint32_t frames = 0; // sharded frames this context has folded into the image
		uint32_t syncs = 0;	 // k_shard_sync launches so far (every rank issues the same sequence)
		uint32_t stamp = 0;	 // wavefronts started so far
	} display;
	cudaStream_t copy_stream = nullptr;			 // rfwb200_read_framebuffer_async: the read-back runs beside the next frame's kernels
	cudaEvent_t ev_frame_done = nullptr, ev_copy_done = nullptr;
	bool copy_pending = false;
	DevBuf d_display_local;						 // [0] CTA counter of k_fold, [1] error word of the flow-control kernels
	std::vector<rfwb200_context *> peers;		 // ranks 1..n-1 of an in-process group; this context is rank 0 = the display rank
	std::vector<std::unique_ptr<DeviceWorker>> workers; // one per peer
	const rfwb200_context *build_donor = nullptr; // rank 0 of the group: its host-built tree is adopted instead of built again

	// ---- settings / state ------------------------------------------------------------------------------
	RenderSettings rs{2, 10.0f, 1e-5f, 1, 0, 8, nullptr, -1, 9, 1, 0, 9, 1, 5, 0, 0, 5, 0}; // bounces: packed nodes + unsorted connect rays; camera rays: fp32 nodes (DESIGN.md sweep) // smem_nodes 0: measured fastest on B200 (DESIGN.md "staging")
	int spp = 1;
	bool mode_pt = true;
	LaunchDims dims{};
	bool dims_valid = false;
	uint32_t sample_index = 0;
	uint32_t probe_x = 0, probe_y = 0;
	uint32_t last_spp = 0, last_first_sample = 0, last_batches = 0;
	uint64_t launches = 0;
	rfwb200_render_stats stats{};
	cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
	bool frame_in_flight = false;
	// optional per-stage timing (setting "timing"): events bracket every launch of the last frame
	bool timing = false;
	struct StageEvent
	{
		int category; // 0 primary, 1 secondary (trace depth 1), 2 deep (trace depth >= 2), 3 shade, 4 finalize, 5 re-ordering
		cudaEvent_t a, b;
	};
	std::vector<StageEvent> stage_events;
	size_t stage_events_used = 0;
	std::vector<DepthCounters> host_counters;
};

using Ctx = rfwb200_context;

namespace
{

int ensure_device(Ctx *c)
{
	CK(cudaSetDevice(c->device));
	return RFWB200_OK;
}

uint32_t local_tile_count(const ShardView &s)
{
	const uint32_t total = s.tiles_x * s.tiles_y;
	return total > s.rank ? (total - s.rank + s.world - 1) / s.world : 0;
}

void recompute_shard(Ctx *c)
{
	ShardView &s = c->shard;
	s.width = c->width, s.height = c->height;
	if (s.world == 0)
		s.world = 1, s.rank = 0;
	if (s.tile_w == 0)
		s.tile_w = 32, s.tile_h = 8;
	s.tiles_x = (s.width + s.tile_w - 1) / s.tile_w;
	s.tiles_y = (s.height + s.tile_h - 1) / s.tile_h;
	s.local_tiles = local_tile_count(s);
	s.local_pixels = s.local_tiles * s.tile_w * s.tile_h;
	s.inv_tile_pixels = 1.0f / float(s.tile_w * s.tile_h), s.inv_tiles_x = 1.0f / float(s.tiles_x);
	s.inv_blocks_per_row = 1.0f / float(s.tile_w >> 3);
}

size_t shard_stride_pixels(const ShardView &s)
{
	const uint32_t total = s.tiles_x * s.tiles_y;
	return size_t((total + s.world - 1) / s.world) * s.tile_w * s.tile_h;
}

// samples of a frame of `spp` samples that travel in one wavefront
uint32_t batch_spp_for(const Ctx *c, uint32_t spp)
{
	const uint32_t P = std::max<uint32_t>(c->shard.local_pixels, 32);
	uint32_t b = std::min<uint32_t>(spp, MAX_BATCH_SPP);
	if (c->spp_batch > 0)
		b = std::min<uint32_t>(b, uint32_t(c->spp_batch));
	b = std::min<uint32_t>(b, std::max<uint32_t>(1u, (1u << 24) / P)); // the path index is a 24-bit field of the state word
	return std::max<uint32_t>(b, 1u);
}

// per-pixel state: accumulator, framebuffer, hit caches
int alloc_wavefront(Ctx *c)
{
	const size_t P = std::max<size_t>(c->shard.local_pixels, 32);
	CK(c->d_acc.reserve(P * sizeof(float4)));
	const size_t fb_pixels = c->shard.world == 1 ? size_t(c->width) * c->height : shard_stride_pixels(c->shard);
	CK(c->d_fb.reserve(std::max(fb_pixels, P) * sizeof(float4)));
	CK(c->d_probe.reserve(sizeof(ProbeResult)));
	CK(c->d_frame.reserve(sizeof(FrameParams)));
	CK(c->d_sort_grid.reserve(sizeof(SortGrid)));
	CK(c->d_scratch_cursor.reserve(256));
	WavefrontView &w = c->wf;
	w.accumulator = c->d_acc.as<float4>();
	w.framebuffer = c->d_fb.as<float4>();
	w.probe = c->d_probe.as<ProbeResult>();
	CK(c->d_prim_cache.reserve(P * sizeof(uint32_t)));
	CK(cudaMemsetAsync(c->d_prim_cache.ptr, 0xff, P * sizeof(uint32_t), c->stream));
	w.prim_cache = c->d_prim_cache.as<uint32_t>();
	CK(c->d_occ_cache.reserve(2 * P * sizeof(uint32_t)));
	CK(cudaMemsetAsync(c->d_occ_cache.ptr, 0xff, 2 * P * sizeof(uint32_t), c->stream));
	w.occ_cache = c->d_occ_cache.as<uint32_t>();
	w.frame = c->d_frame.as<FrameParams>();
	w.grid = c->d_sort_grid.as<SortGrid>();
	c->sort_grid_dirty = true;
	return RFWB200_OK;
}

// per-work-item state of a wavefront of `bspp` samples per pixel, and the bins of the re-ordering pass
int ensure_wavefront(Ctx *c, uint32_t bspp)
{
	const size_t items = std::max<size_t>(size_t(c->shard.local_pixels) * bspp, 32);
	if (items > c->items_capacity)
	{
		CK(cudaStreamSynchronize(c->stream));
		const size_t plane = items * sizeof(float4);
		for (int i = 0; i < 2; i++)
		{
			CK(c->d_O[i].reserve(plane));
			CK(c->d_D[i].reserve(plane));
			CK(c->d_T[i].reserve(plane));
		}
		CK(c->d_hit.reserve(plane));
		CK(c->d_sO.reserve(plane));
		CK(c->d_sD.reserve(plane));
		CK(c->d_sE.reserve(plane));
		CK(c->d_sample_acc.reserve(plane));
		CK(c->d_sort_key.reserve(items * sizeof(uint2)));
		c->items_capacity = items;
	}
	if (c->aov)
	{
		const size_t P = std::max<size_t>(c->shard.local_pixels, 32);
		if (c->d_sample_albedo.bytes < items * sizeof(float4) || c->d_albedo.bytes < P * sizeof(float4))
		{
			CK(cudaStreamSynchronize(c->stream));
			CK(c->d_sample_albedo.reserve(items * sizeof(float4)));
			CK(c->d_sample_normal.reserve(items * sizeof(float4)));
			const bool fresh = c->d_albedo.bytes < P * sizeof(float4);
			CK(c->d_albedo.reserve(P * sizeof(float4)));
			CK(c->d_normal.reserve(P * sizeof(float4)));
			if (fresh)
			{
				CK(cudaMemsetAsync(c->d_albedo.ptr, 0, c->d_albedo.bytes, c->stream));
				CK(cudaMemsetAsync(c->d_normal.ptr, 0, c->d_normal.bytes, c->stream));
			}
		}
	}
	REQUIRE(3 * c->rs.sort_cell_bits + c->rs.sort_dir_bits <= 21, "sort_cell_bits = 6 needs sort_dir_bits = 3 (at most 2^21 bins)");
	const uint32_t bins = 1u << (3 * c->rs.sort_cell_bits + c->rs.sort_dir_bits);
	if (bins > c->sort_bins_allocated)
	{
		CK(cudaStreamSynchronize(c->stream));
		CK(c->d_sort_hist.reserve(size_t(bins) * sizeof(uint32_t)));
		CK(c->d_sort_base.reserve(size_t(bins) * sizeof(uint32_t)));
		CK(c->d_sort_chunk.reserve(size_t(std::max<uint32_t>(bins / SORT_CHUNK, 1)) * sizeof(uint32_t)));
		CK(cudaMemsetAsync(c->d_sort_hist.ptr, 0, c->d_sort_hist.bytes, c->stream)); // k_sort_scan leaves the counts at zero again
		c->sort_bins_allocated = bins;
	}
	WavefrontView &w = c->wf;
	for (int i = 0; i < 2; i++)
		w.O[i] = c->d_O[i].as<float4>(), w.D[i] = c->d_D[i].as<float4>(), w.T[i] = c->d_T[i].as<float4>();
	w.hit = c->d_hit.as<float4>();
	w.sO = c->d_sO.as<float4>(), w.sD = c->d_sD.as<float4>(), w.sE = c->d_sE.as<float4>();
	w.sample_acc = c->d_sample_acc.as<float4>();
	w.sample_albedo = c->aov ? c->d_sample_albedo.as<float4>() : nullptr, w.sample_normal = c->aov ? c->d_sample_normal.as<float4>() : nullptr;
	w.albedo_acc = c->aov ? c->d_albedo.as<float4>() : nullptr, w.normal_acc = c->aov ? c->d_normal.as<float4>() : nullptr;
	w.sort_key = c->d_sort_key.as<uint2>();
	w.sort_hist = c->d_sort_hist.as<uint32_t>(), w.sort_base = c->d_sort_base.as<uint32_t>(), w.sort_chunk = c->d_sort_chunk.as<uint32_t>();
	return RFWB200_OK;
}

int ensure_counters(Ctx *c, uint32_t batches)
{
	if (batches > c->counters_capacity)
	{
		CK(cudaStreamSynchronize(c->stream));
		CK(c->d_counters.reserve(size_t(batches) * MAX_DEPTH_SLOTS * sizeof(DepthCounters)));
		CK(c->d_ext_seen.reserve(size_t(batches) * MAX_DEPTH_SLOTS * MAX_BATCH_SPP * sizeof(uint32_t)));
		c->counters_capacity = batches;
	}
	c->wf.counters = c->d_counters.as<DepthCounters>();
	c->wf.ext_seen = c->d_ext_seen.as<uint32_t>();
	return RFWB200_OK;
}

// instance flattening: world-space triangles + repacked shading records
// (shade_out / det_eps_out are null on the device path: geometry.cu produces those records itself)
int flatten_scene(Ctx *c, std::vector<ShadeTri> *shade_out, std::vector<float> *det_eps_out)
{
	std::vector<ShadeTri> shade_dummy;
	std::vector<float> eps_dummy;
	const bool want_shade = shade_out != nullptr;
	std::vector<ShadeTri> &shade = want_shade ? *shade_out : shade_dummy;
	std::vector<float> &det_eps = det_eps_out ? *det_eps_out : eps_dummy;
	size_t total = 0;
	for (const HostInstance &in : c->instances)
		if (in.mesh >= 0 && size_t(in.mesh) < c->meshes.size())
			total += c->meshes[in.mesh].triangles.size();
	c->build_tris.resize(total);
	if (want_shade)
		shade.resize(total), det_eps.resize(total);
	size_t at = 0;
	for (size_t ii = 0; ii < c->instances.size(); ii++)
	{
		const HostInstance &in = c->instances[ii];
		if (in.mesh < 0 || size_t(in.mesh) >= c->meshes.size())
			continue;
		const HostMesh &m = c->meshes[in.mesh];
		const float *M = in.transform;
		const float det = M[0] * (M[5] * M[10] - M[9] * M[6]) - M[4] * (M[1] * M[10] - M[9] * M[2]) +
						  M[8] * (M[1] * M[6] - M[5] * M[2]);
		const float eps = 1e-6f * std::fabs(det);
		const size_t nt = m.triangles.size();
		const size_t nv = m.vertices.size() / 4;
		for (size_t t = 0; t < nt; t++, at++)
		{
			uint32_t vi[3];
			for (int k = 0; k < 3; k++)
				vi[k] = m.indices.empty() ? uint32_t(t * 3 + k) : m.indices[t * 3 + k];
			if (vi[0] >= nv || vi[1] >= nv || vi[2] >= nv)
				return set_error(RFWB200_ERR_INVALID, "mesh index out of range");
			BuildTriangle &bt = c->build_tris[at];
			mul_point(M, &m.vertices[4 * vi[0]], bt.v0);
			mul_point(M, &m.vertices[4 * vi[1]], bt.v1);
			mul_point(M, &m.vertices[4 * vi[2]], bt.v2);
			if (!want_shade)
				continue;
			const rfwb200_triangle &src = m.triangles[t];
			ShadeTri &st = shade[at];
			st.u0 = src.u0, st.u1 = src.u1, st.u2 = src.u2, st.light_tri_idx = src.light_tri_idx;
			st.v0 = src.v0, st.v1 = src.v1, st.v2 = src.v2, st.material = src.material;
			float n[3];
			mul_mat3(in.normal, src.vN0, n);
			st.n0x = n[0], st.n0y = n[1], st.n0z = n[2];
			mul_mat3(in.normal, src.vN1, n);
			st.n1x = n[0], st.n1y = n[1], st.n1z = n[2];
			mul_mat3(in.normal, src.vN2, n);
			st.n2x = n[0], st.n2y = n[1], st.n2z = n[2];
			const float gn[3] = {src.Nx, src.Ny, src.Nz};
			mul_mat3(in.normal, gn, n);
			const float il = 1.0f / std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
			st.Nx = n[0] * il, st.Ny = n[1] * il, st.Nz = n[2] * il;
			st.area = src.area, st.lod = src.LOD;
			st.inst_id = uint32_t(ii), st.prim_id = uint32_t(t);
			det_eps[at] = eps;
		}
	}
	c->flat_tri_count = total;
	return RFWB200_OK;
}

int upload_bvh(Ctx *c, const std::vector<float> &det_eps)
{
	const size_t n = c->bvh.tri_order.size();
	std::vector<TriRec> recs(std::max<size_t>(n, 1));
	memset(recs.data(), 0, recs.size() * sizeof(TriRec));
	for (size_t i = 0; i < n; i++)
	{
		const uint32_t src = c->bvh.tri_order[i];
		const BuildTriangle &t = c->build_tris[src];
		TriRec &r = recs[i];
		r.p0x = t.v0[0], r.p0y = t.v0[1], r.p0z = t.v0[2];
		r.e1x = t.v1[0] - t.v0[0], r.e1y = t.v1[1] - t.v0[1], r.e1z = t.v1[2] - t.v0[2];
		r.e2x = t.v2[0] - t.v0[0], r.e2y = t.v2[1] - t.v0[1], r.e2z = t.v2[2] - t.v0[2];
		r.shade_idx = src;
		r.det_eps = det_eps[src];
	}
	CK(c->d_nodes.reserve(c->bvh.nodes.size() * sizeof(BvhNode4)));
	CK(c->d_tris.reserve(recs.size() * sizeof(TriRec)));
	CK(cudaMemcpyAsync(c->d_nodes.ptr, c->bvh.nodes.data(), c->bvh.nodes.size() * sizeof(BvhNode4), cudaMemcpyHostToDevice,
					   c->stream));
	CK(cudaMemcpyAsync(c->d_tris.ptr, recs.data(), recs.size() * sizeof(TriRec), cudaMemcpyHostToDevice, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	c->scene.nodes = c->d_nodes.as<BvhNode4>();
	c->scene.tris = c->d_tris.as<TriRec>();
	c->scene.node_count = uint32_t(c->bvh.nodes.size());
	c->scene.tri_count = uint32_t(n);
	return RFWB200_OK;
}


// the 80-byte form of the nodes the wavefront trace kernel reads (device_types.h BvhNode4Packed), after any build / refit
int pack_nodes(Ctx *c)
{
	c->scene.nodes16 = nullptr;
	if (c->scene.cw_nodes || c->scene.node_count == 0)
		return RFWB200_OK;
	CK(c->d_nodes16.reserve(size_t(c->scene.node_count) * sizeof(BvhNode4Packed)));
	CK(launch_pack_nodes(c->scene.nodes, c->d_nodes16.as<BvhNode4Packed>(), c->scene.node_count, c->stream));
	c->scene.nodes16 = c->d_nodes16.as<uint4>();
	c->launches += 1;
	return RFWB200_OK;
}

// instances whose triangles no longer are where the builder saw them: refits bound those triangles whole, all others keep
// the builder's (spatially split) reference boxes
void mark_mesh_moved(Ctx *c, size_t mesh_index)
{
	c->inst_moved.resize(c->instances.size(), 0);
	for (size_t ii = 0; ii < c->instances.size(); ii++)
		if (c->instances[ii].mesh == int(mesh_index))
			c->inst_moved[ii] = 1;
}

// ---- device geometry path (geometry.cu) --------------------------------------------------------------------------------
int upload_mesh_to_arena(Ctx *c, size_t mi)
{
	const HostMesh &m = c->meshes[mi];
	const size_t nv = m.vertices.size() / 4, nt = m.triangles.size();
	if (nv)
		CK(cudaMemcpyAsync(c->d_verts.as<float4>() + c->mesh_vert_off[mi], m.vertices.data(), nv * sizeof(float4),
						   cudaMemcpyHostToDevice, c->stream));
	if (nt)
	{
		CK(cudaMemcpyAsync(static_cast<char *>(c->d_mesh_tris.ptr) + size_t(c->mesh_tri_off[mi]) * sizeof(rfwb200_triangle),
						   m.triangles.data(), nt * sizeof(rfwb200_triangle), cudaMemcpyHostToDevice, c->stream));
		uint32_t *dst = c->d_indices.as<uint32_t>() + size_t(c->mesh_tri_off[mi]) * 3;
		if (!m.indices.empty())
			CK(cudaMemcpyAsync(dst, m.indices.data(), nt * 3 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
		else
		{
			std::vector<uint32_t> seq(nt * 3);
			for (size_t i = 0; i < seq.size(); i++)
				seq[i] = uint32_t(i);
			CK(cudaMemcpyAsync(dst, seq.data(), seq.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
			CK(cudaStreamSynchronize(c->stream)); // seq dies here
		}
	}
	c->mesh_dirty[mi] = 0;
	return RFWB200_OK;
}

int upload_arena(Ctx *c)
{
	const size_t n = c->meshes.size();
	c->mesh_vert_off.assign(n, 0), c->mesh_tri_off.assign(n, 0), c->mesh_dirty.assign(n, 1);
	size_t nv = 0, nt = 0;
	for (size_t i = 0; i < n; i++)
	{
		c->mesh_vert_off[i] = uint32_t(nv), c->mesh_tri_off[i] = uint32_t(nt);
		nv += c->meshes[i].vertices.size() / 4, nt += c->meshes[i].triangles.size();
	}
	if (nv >= (1ull << 32) || nt * 3 >= (1ull << 32))
		return set_error(RFWB200_ERR_INVALID, "geometry arena exceeds 32-bit offsets");
	CK(c->d_verts.reserve(std::max<size_t>(nv, 1) * sizeof(float4)));
	CK(c->d_mesh_tris.reserve(std::max<size_t>(nt, 1) * sizeof(rfwb200_triangle)));
	CK(c->d_indices.reserve(std::max<size_t>(nt, 1) * 3 * sizeof(uint32_t)));
	for (size_t i = 0; i < n; i++)
		if (int r = upload_mesh_to_arena(c, i))
			return r;
	c->arena_valid = true;
	return RFWB200_OK;
}

int upload_instances(Ctx *c, bool with_flat_inst)
{
	std::vector<DeviceInstance> di(std::max<size_t>(c->instances.size(), 1));
	memset(di.data(), 0, di.size() * sizeof(DeviceInstance));
	std::vector<uint32_t> flat_inst;
	uint32_t flat = 0;
	for (size_t ii = 0; ii < c->instances.size(); ii++)
	{
		const HostInstance &in = c->instances[ii];
		DeviceInstance &d = di[ii];
		d.flat_off = flat;
		if (in.mesh < 0 || size_t(in.mesh) >= c->meshes.size())
			continue;
		memcpy(d.transform, in.transform, sizeof(d.transform));
		memcpy(d.normal, in.normal, sizeof(d.normal));
		const float *M = in.transform;
		const float det = M[0] * (M[5] * M[10] - M[9] * M[6]) - M[4] * (M[1] * M[10] - M[9] * M[2]) +
						  M[8] * (M[1] * M[6] - M[5] * M[2]);
		d.det_eps = 1e-6f * std::fabs(det);
		d.vert_off = c->mesh_vert_off[in.mesh], d.tri_off = c->mesh_tri_off[in.mesh];
		d.tri_count = uint32_t(c->meshes[in.mesh].triangles.size());
		d.moved = (ii < c->inst_moved.size() && c->inst_moved[ii]) ? 1u : 0u;
		if (with_flat_inst)
			flat_inst.insert(flat_inst.end(), d.tri_count, uint32_t(ii));
		flat += d.tri_count;
	}
	CK(c->d_instances.reserve(di.size() * sizeof(DeviceInstance)));
	CK(cudaMemcpyAsync(c->d_instances.ptr, di.data(), di.size() * sizeof(DeviceInstance), cudaMemcpyHostToDevice, c->stream));
	if (with_flat_inst)
	{
		CK(c->d_flat_inst.reserve(std::max<size_t>(flat_inst.size(), 1) * sizeof(uint32_t)));
		if (!flat_inst.empty())
			CK(cudaMemcpyAsync(c->d_flat_inst.ptr, flat_inst.data(), flat_inst.size() * sizeof(uint32_t), cudaMemcpyHostToDevice,
							   c->stream));
	}
	CK(cudaStreamSynchronize(c->stream)); // the staging vectors die here
	return RFWB200_OK;
}

// point the kernels' views at the 4-wide tree in d_nodes / d_tri_order / d_parent_slot (nn nodes, nr leaf references)
void bind_geometry_views(Ctx *c, size_t nn, size_t nr)
{
	GeometryView &g = c->geo;
	g = GeometryView{};
	g.verts = c->d_verts.as<float4>(), g.indices = c->d_indices.as<uint32_t>(), g.mesh_tris = c->d_mesh_tris.ptr;
	g.instances = c->d_instances.as<DeviceInstance>(), g.flat_inst = c->d_flat_inst.as<uint32_t>();
	g.flat_count = uint32_t(c->flat_tri_count);
	g.nodes = c->d_nodes.as<BvhNode4>(), g.tri_order = c->d_tri_order.as<uint32_t>();
	g.parent_slot = c->d_parent_slot.as<uint32_t>(), g.arrivals = c->d_arrivals.as<uint32_t>();
	g.ref_boxes = ((c->bvh.ref_boxes.size() == 6 * nr || c->device_ref_boxes) && nr > 0) ? c->d_ref_boxes.as<float>() : nullptr;
	g.node_count = uint32_t(nn), g.ref_count = uint32_t(nr);
	g.out_tris = c->d_tris.as<TriRec>(), g.out_shade = c->d_shade_tris.as<ShadeTri>();
	c->scene.nodes = c->d_nodes.as<BvhNode4>(), c->scene.tris = c->d_tris.as<TriRec>();
	c->scene.shade_tris = c->d_shade_tris.as<ShadeTri>();
	c->scene.node_count = uint32_t(nn), c->scene.tri_count = uint32_t(nr);
	c->scene.cw_nodes = nullptr, c->scene.cw_node_count = 0;
}

// nodes (topology + the builder's boxes), leaf order and the parent links the bottom-up refit climbs
int upload_topology(Ctx *c)
{
	if (c->bvh.wide8)
	{
		// compressed 8-wide layout: the nodes as the host builder quantised them; refits of this layout run on the host
		const size_t nn8 = c->bvh.cw_nodes.size(), nr8 = c->bvh.tri_order.size();
		CK(c->d_cw_nodes.reserve(nn8 * sizeof(CwNode)));
		CK(c->d_nodes.reserve(sizeof(BvhNode4)));
		CK(c->d_tris.reserve(std::max<size_t>(nr8, 1) * sizeof(TriRec)));
		CK(c->d_tri_order.reserve(std::max<size_t>(nr8, 1) * sizeof(uint32_t)));
		CK(c->d_shade_tris.reserve(std::max<size_t>(c->flat_tri_count, 1) * sizeof(ShadeTri)));
		CK(cudaMemcpyAsync(c->d_cw_nodes.ptr, c->bvh.cw_nodes.data(), nn8 * sizeof(CwNode), cudaMemcpyHostToDevice, c->stream));
		if (nr8)
			CK(cudaMemcpyAsync(c->d_tri_order.ptr, c->bvh.tri_order.data(), nr8 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
		CK(cudaStreamSynchronize(c->stream));
		GeometryView &g8 = c->geo;
		g8 = GeometryView{};
		g8.verts = c->d_verts.as<float4>(), g8.indices = c->d_indices.as<uint32_t>(), g8.mesh_tris = c->d_mesh_tris.ptr;
		g8.instances = c->d_instances.as<DeviceInstance>(), g8.flat_inst = c->d_flat_inst.as<uint32_t>();
		g8.flat_count = uint32_t(c->flat_tri_count);
		g8.tri_order = c->d_tri_order.as<uint32_t>();
		g8.node_count = 0, g8.ref_count = uint32_t(nr8);
		g8.out_tris = c->d_tris.as<TriRec>(), g8.out_shade = c->d_shade_tris.as<ShadeTri>();
		c->scene.nodes = c->d_nodes.as<BvhNode4>(), c->scene.node_count = 0; // unused by the kernels in this mode
		c->scene.cw_nodes = c->d_cw_nodes.as<uint4>(), c->scene.cw_node_count = uint32_t(nn8);
		c->scene.tris = c->d_tris.as<TriRec>(), c->scene.shade_tris = c->d_shade_tris.as<ShadeTri>();
		c->scene.tri_count = uint32_t(nr8);
		return RFWB200_OK;
	}
	c->scene.cw_nodes = nullptr, c->scene.cw_node_count = 0;
	const size_t nn = c->bvh.nodes.size(), nr = c->bvh.tri_order.size();
	if (nn >= (1u << 30))
		return set_error(RFWB200_ERR_INVALID, "BVH has more than 2^30 nodes");
	std::vector<uint32_t> parent_slot(std::max<size_t>(nn, 1), 0xffffffffu);
	for (size_t n = 0; n < nn; n++)
		for (int s = 0; s < c->bvh.nodes[n].pad[0]; s++)
			if (c->bvh.nodes[n].child[s] >= 0)
				parent_slot[c->bvh.nodes[n].child[s]] = uint32_t(n << 2) | uint32_t(s);
	CK(c->d_nodes.reserve(nn * sizeof(BvhNode4)));
	CK(c->d_tris.reserve(std::max<size_t>(nr, 1) * sizeof(TriRec)));
	CK(c->d_tri_order.reserve(std::max<size_t>(nr, 1) * sizeof(uint32_t)));
	CK(c->d_parent_slot.reserve(parent_slot.size() * sizeof(uint32_t)));
	CK(c->d_arrivals.reserve(parent_slot.size() * sizeof(uint32_t)));
	CK(c->d_shade_tris.reserve(std::max<size_t>(c->flat_tri_count, 1) * sizeof(ShadeTri)));
	CK(cudaMemcpyAsync(c->d_nodes.ptr, c->bvh.nodes.data(), nn * sizeof(BvhNode4), cudaMemcpyHostToDevice, c->stream));
	if (nr)
		CK(cudaMemcpyAsync(c->d_tri_order.ptr, c->bvh.tri_order.data(), nr * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
	CK(cudaMemcpyAsync(c->d_parent_slot.ptr, parent_slot.data(), parent_slot.size() * sizeof(uint32_t), cudaMemcpyHostToDevice,
					   c->stream));
	if (c->bvh.ref_boxes.size() == 6 * nr && nr > 0)
	{
		CK(c->d_ref_boxes.reserve(c->bvh.ref_boxes.size() * sizeof(float)));
		CK(cudaMemcpyAsync(c->d_ref_boxes.ptr, c->bvh.ref_boxes.data(), c->bvh.ref_boxes.size() * sizeof(float), cudaMemcpyHostToDevice,
						   c->stream));
	}
	CK(cudaStreamSynchronize(c->stream));
	bind_geometry_views(c, nn, nr);
	return RFWB200_OK;
}

// world-space intersection + shading records (and, for a refit, every box of the tree) from the arena
int device_generate(Ctx *c, bool refit_boxes, bool record_begin = true)
{
	if (!c->ev_geo_a)
	{
		CK(cudaEventCreate(&c->ev_geo_a));
		CK(cudaEventCreate(&c->ev_geo_b));
	}
	if (record_begin)
		CK(cudaEventRecord(c->ev_geo_a, c->stream));
	c->geo.write_boxes = refit_boxes ? 1 : 0;
	if (c->bvh.wide8)
		CK(launch_records(c->geo, c->stream));
	else
		CK(launch_refit(c->geo, c->stream));
	CK(launch_flatten_shade(c->geo, c->stream));
	if (int r = pack_nodes(c))
		return r;
	CK(cudaEventRecord(c->ev_geo_b, c->stream));
	c->launches += 2;
	c->geo_timed = true;
	return RFWB200_OK;
}

// a rebuild needs the current pose of device-skinned meshes on the host
int sync_skinned_meshes_to_host(Ctx *c)
{
	bool any = false;
	for (size_t mi = 0; mi < c->skins.size() && mi < c->meshes.size(); mi++)
	{
		Ctx::Skin *sk = c->skins[mi].get();
		// (the arena LAYOUT may already be stale — a new mesh was added — but this mesh's slice is still intact:
		// set_mesh on the mesh itself clears device_newer)
		if (!sk || !sk->device_newer || mi >= c->mesh_vert_off.size() || !c->d_verts.ptr)
			continue;
		HostMesh &hm = c->meshes[mi];
		CK(cudaMemcpyAsync(hm.vertices.data(), c->d_verts.as<float4>() + c->mesh_vert_off[mi], hm.vertices.size() * sizeof(float),
						   cudaMemcpyDeviceToHost, c->stream));
		CK(cudaMemcpyAsync(hm.triangles.data(),
						   static_cast<const char *>(c->d_mesh_tris.ptr) + size_t(c->mesh_tri_off[mi]) * sizeof(rfwb200_triangle),
						   hm.triangles.size() * sizeof(rfwb200_triangle), cudaMemcpyDeviceToHost, c->stream));
		sk->device_newer = false;
		any = true;
	}
	if (any)
		CK(cudaStreamSynchronize(c->stream));
	return RFWB200_OK;
}

int upload_frame_params(Ctx *c, const rfwb200_camera_view *view, uint32_t sample_base)
{
	FrameParams fp;
	for (int i = 0; i < 3; i++)
	{
		fp.pos[i] = view->pos[i], fp.p1[i] = view->p1[i];
		fp.right[i] = view->p2[i] - view->p1[i];
		fp.up[i] = view->p3[i] - view->p1[i];
	}
	fp.aperture = view->aperture, fp.spread_angle = view->spread_angle;
	fp.sample_base = sample_base;
	memcpy(fp.to_eye, c->to_eye, sizeof(fp.to_eye));
	fp.probe_pixel = (c->probe_x < c->width && c->probe_y < c->height) ? c->probe_y * c->width + c->probe_x : 0xffffffffu;
	CK(cudaMemcpyAsync(c->d_frame.ptr, &fp, sizeof(fp), cudaMemcpyHostToDevice, c->stream));
	return RFWB200_OK;
}

int ensure_dims(Ctx *c)
{
	if (c->dims_valid)
		return RFWB200_OK;
	CK(configure_launches(c->rs, c->scene.node_count, c->dims));
	c->dims_valid = true;
	return RFWB200_OK;
}

// brackets one launch with events when per-stage timing is on
struct StageTimer
{
	Ctx *c;
	cudaEvent_t b = nullptr;
	StageTimer(Ctx *ctx, int category) : c(ctx)
	{
		if (!c->timing)
			return;
		if (c->stage_events_used == c->stage_events.size())
		{
			Ctx::StageEvent e{category, nullptr, nullptr};
			if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess)
				return;
			c->stage_events.push_back(e);
		}
		Ctx::StageEvent &e = c->stage_events[c->stage_events_used++];
		e.category = category;
		cudaEventRecord(e.a, c->stream);
		b = e.b;
	}
	~StageTimer()
	{
		if (b)
			cudaEventRecord(b, c->stream);
	}
};

#define FORWARD(call)                                                                                                   \
	do                                                                                                                  \
	{                                                                                                                   \
		for (rfwb200_context * p_ : c->peers)                                                                           \
			if (int r_ = (call))                                                                                        \
				return r_;                                                                                              \
	} while (0)

// counters + two sets of flag rows (wavefronts alternate between them, so a rank that is one wavefront ahead never
// overwrites flags another rank has not read yet)
constexpr size_t DISPLAY_TAIL_BYTES = 512 + 2 * MAX_DEPTH_SLOTS * MAX_BATCH_SPP * sizeof(uint32_t);

void display_detach(Ctx *c)
{
	Ctx::Display &d = c->display;
	if (d.base)
	{
		cudaSetDevice(c->device);
		cudaStreamSynchronize(c->stream);
		if (d.owner)
			cudaFree(d.base);
		else if (d.ipc)
			cudaIpcCloseMemHandle(d.base);
	}
	d = Ctx::Display{};
}

// lays the pointers of a display allocation out for an image of this context's size
int display_bind(Ctx *c, void *base, bool owner, bool ipc)
{
	Ctx::Display &d = c->display;
	d.base = base, d.owner = owner, d.ipc = ipc;
	d.image_bytes = (size_t(c->width) * c->height * sizeof(float4) + 255) & ~size_t(255);
	d.image = static_cast<float4 *>(base);
	d.arrivals = reinterpret_cast<uint32_t *>(static_cast<char *>(base) + d.image_bytes);
	d.consumed = d.arrivals + 32, d.sync_arrivals = d.arrivals + 64, d.sync_seen = d.arrivals + 128;
	d.frames = 0, d.syncs = 0, d.stamp = 0;
	CK(c->d_display_local.reserve(256));
	CK(cudaMemsetAsync(c->d_display_local.ptr, 0, 256, c->stream));
	return RFWB200_OK;
}

DisplayTarget display_target(const Ctx *c)
{
	DisplayTarget t{};
	if (c->display.image)
		t.image = c->display.image, t.arrivals = c->display.arrivals, t.local_done = c->d_display_local.as<uint32_t>();
	return t;
}

int check_ready(Ctx *c)
{
	if (!c->initialised)
		return set_error(RFWB200_ERR_STATE, "rfwb200_init has not been called");
	if (c->geometry_dirty || c->scene.nodes == nullptr)
		return set_error(RFWB200_ERR_STATE, "scene changed: call rfwb200_update before rendering");
	return RFWB200_OK;
}

} // namespace

extern "C"
{

	const char *rfwb200_last_error(void) { return g_last_error.c_str(); }
	const char *rfwb200_version(void) { return "rfwb200 0.1 (sm_100a)"; }

	int rfwb200_create(int device, rfwb200_context **out)
	{
		REQUIRE(out != nullptr, "out is null");
		*out = nullptr;
		int count = 0;
		const cudaError_t e = cudaGetDeviceCount(&count);
		if (e != cudaSuccess || count == 0)
			return set_error(RFWB200_ERR_NO_DEVICE,
							 std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count == 0") +
								 " (this library has no CPU fallback)");
		REQUIRE(device >= 0 && device < count, "device ordinal out of range");
		std::unique_ptr<Ctx> c(new Ctx());
		c->device = device;
		CK(cudaSetDevice(device));
		cudaDeviceProp prop;
		CK(cudaGetDeviceProperties(&prop, device));
		if (prop.major < 10)
			return set_error(RFWB200_ERR_NO_DEVICE, std::string("device '") + prop.name +
														"' is not Blackwell (sm_100a code only, no fallback path)");
		CK(cudaEventCreate(&c->ev_begin));
		CK(cudaEventCreate(&c->ev_end));
		// blue-noise tables (context/blue_noise.h:8204-8219 layout), embedded in the library
		CK(c->d_blue_noise.reserve(rfwb200_bluenoise_table_size));
		CK(cudaMemcpy(c->d_blue_noise.ptr, rfwb200_bluenoise_table, rfwb200_bluenoise_table_size, cudaMemcpyHostToDevice));
		c->scene.blue_noise = c->d_blue_noise.as<uint8_t>();
		// a 1x1 black sky so a miss is defined before set_sky
		const float black[4] = {0, 0, 0, 0};
		CK(c->d_sky.reserve(sizeof(black)));
		CK(cudaMemcpy(c->d_sky.ptr, black, sizeof(black), cudaMemcpyHostToDevice));
		c->scene.sky = c->d_sky.as<float>(), c->scene.sky_w = 1, c->scene.sky_h = 1;
		c->shard.world = 1, c->shard.rank = 0, c->shard.tile_w = 32, c->shard.tile_h = 8;
		*out = c.release();
		return RFWB200_OK;
	}

	int rfwb200_cleanup(rfwb200_context *c)
	{
		if (!c)
			return RFWB200_OK;
		for (rfwb200_context *p : c->peers)
			rfwb200_cleanup(p);
		cudaSetDevice(c->device);
		cudaStreamSynchronize(c->stream);
		return RFWB200_OK;
	}

	int rfwb200_destroy(rfwb200_context *c)
	{
		if (!c)
			return RFWB200_OK;
		rfwb200_cleanup(c);
		c->workers.clear(); // joins the per-device host threads
		for (rfwb200_context *p : c->peers)
			rfwb200_destroy(p); // peers drop their mapping of the display image before its owner frees it
		c->peers.clear();
		cudaSetDevice(c->device);
		display_detach(c);
		if (c->copy_stream)
			cudaStreamDestroy(c->copy_stream), cudaEventDestroy(c->ev_frame_done), cudaEventDestroy(c->ev_copy_done);
		if (c->owns_stream && c->stream)
			cudaStreamDestroy(c->stream);
		if (c->ev_begin)
			cudaEventDestroy(c->ev_begin);
		if (c->ev_end)
			cudaEventDestroy(c->ev_end);
		for (auto &e : c->stage_events)
			cudaEventDestroy(e.a), cudaEventDestroy(e.b);
		if (c->ev_geo_a)
			cudaEventDestroy(c->ev_geo_a), cudaEventDestroy(c->ev_geo_b);
		delete c;
		return RFWB200_OK;
	}

	int rfwb200_init(rfwb200_context *c, uint32_t width, uint32_t height)
	{
		REQUIRE(c != nullptr, "context is null");
		REQUIRE(width > 0 && height > 0, "width and height must be positive");
		REQUIRE(uint64_t(width) * height < (1ull << 31), "more than 2^31 pixels");
		if (int r = ensure_device(c))
			return r;
		const uint32_t old_w = c->width, old_h = c->height;
		c->width = width, c->height = height;
		recompute_shard(c);
		// the 24-bit path index of the state word addresses the tile-padded local work items of this shard
		if (uint64_t(c->shard.local_tiles) * c->shard.tile_w * c->shard.tile_h > (1ull << 24))
		{
			c->width = old_w, c->height = old_h;
			recompute_shard(c);
			return set_error(RFWB200_ERR_INVALID, "more than 2^24 (tile-padded) pixels in this shard: path index no longer fits the state word; "
												  "shard the frame over more ranks (rfwb200_set_shard)");
		}
		if (int r = alloc_wavefront(c))
			return r;
		CK(cudaMemsetAsync(c->d_acc.ptr, 0, c->d_acc.bytes, c->stream));
		CK(cudaMemsetAsync(c->d_fb.ptr, 0, c->d_fb.bytes, c->stream));
		c->sample_index = 0;
		c->initialised = true;
		c->items_capacity = 0; // the wavefront planes are sized per frame (ensure_wavefront)
		if ((c->display.base && !c->display.owner) || (c->display.owner && (old_w != width || old_h != height)))
			display_detach(c); // a display image of another size: the caller attaches again
		if (!c->peers.empty())
		{
			// group: rank 0 owns the display image every rank folds into
			for (rfwb200_context *p : c->peers)
				display_detach(p);
			FORWARD(rfwb200_init(p_, width, height));
			if (int r = ensure_device(c))
				return r;
			if (int r = rfwb200_display_create(c, nullptr))
				return r;
			FORWARD(rfwb200_display_attach(p_, c));
			if (int r = ensure_device(c))
				return r;
		}
		return RFWB200_OK;
	}

	int rfwb200_set_shard(rfwb200_context *c, uint32_t rank, uint32_t world, uint32_t tile_w, uint32_t tile_h)
	{
		REQUIRE(c != nullptr, "context is null");
		REQUIRE(world >= 1 && rank < world, "rank must be < world");
		REQUIRE(tile_w >= 8 && tile_w % 8 == 0 && tile_h >= 4 && tile_h % 4 == 0, "tile must be a multiple of 8x4 pixels");
		REQUIRE(tile_w <= 1024 && tile_h <= 1024, "tile must be at most 1024x1024 pixels");
		REQUIRE(c->peers.empty() && c->build_donor == nullptr, "the ranks of an in-process group are assigned by rfwb200_create_group");
		c->shard.rank = rank, c->shard.world = world, c->shard.tile_w = tile_w, c->shard.tile_h = tile_h;
		if (c->initialised)
			return rfwb200_init(c, c->width, c->height);
		return RFWB200_OK;
	}

	int rfwb200_set_stream(rfwb200_context *c, void *cuda_stream)
	{
		REQUIRE(c != nullptr, "context is null");
		c->stream = static_cast<cudaStream_t>(cuda_stream);
		return RFWB200_OK;
	}

	int rfwb200_set_sky(rfwb200_context *c, const float *rgb, size_t width, size_t height)
	{
		REQUIRE(c && rgb && width > 0 && height > 0, "bad sky");
		FORWARD(rfwb200_set_sky(p_, rgb, width, height));
		if (int r = ensure_device(c))
			return r;
		std::vector<float> tmp(width * height * 4);
		for (size_t i = 0; i < width * height; i++)
			tmp[4 * i] = rgb[3 * i], tmp[4 * i + 1] = rgb[3 * i + 1], tmp[4 * i + 2] = rgb[3 * i + 2], tmp[4 * i + 3] = 0.f;
		CK(cudaStreamSynchronize(c->stream));
		CK(c->d_sky.reserve(tmp.size() * sizeof(float)));
		CK(cudaMemcpy(c->d_sky.ptr, tmp.data(), tmp.size() * sizeof(float), cudaMemcpyHostToDevice));
		c->scene.sky = c->d_sky.as<float>(), c->scene.sky_w = uint32_t(width), c->scene.sky_h = uint32_t(height);
		return RFWB200_OK;
	}

	static int resolve_and_upload_materials(rfwb200_context *c);

	int rfwb200_set_textures(rfwb200_context *c, const rfwb200_texture_data *tex, size_t count)
	{
		REQUIRE(c && (tex || count == 0), "bad textures");
		FORWARD(rfwb200_set_textures(p_, tex, count));
		if (int r = ensure_device(c))
			return r;
		// one linear uint array and one float4 array with per-texture offsets (CUDART/src/Context.cpp:201-268)
		size_t nu = 0, nf = 0;
		for (size_t i = 0; i < count; i++)
		{
			REQUIRE(tex[i].data != nullptr, "texture without data");
			(tex[i].type == RFWB200_TEX_UINT ? nu : nf) += tex[i].texel_count;
		}
		REQUIRE(nu < (1ull << 32) && nf < (1ull << 32), "texture pool exceeds 2^32 texels");
		std::vector<uint32_t> upool(std::max<size_t>(nu, 4), 0);
		std::vector<float> fpool(std::max<size_t>(nf, 4) * 4, 0.f);
		c->textures.resize(count);
		std::vector<uint32_t> desc(std::max<size_t>(count, 1) * 4, 0);
		size_t uo = 0, fo = 0;
		for (size_t i = 0; i < count; i++)
		{
			HostTexture t{tex[i].type, tex[i].width, tex[i].height, tex[i].texel_count, 0};
			if (tex[i].type == RFWB200_TEX_UINT)
			{
				t.addr = uint32_t(uo);
				memcpy(&upool[uo], tex[i].data, size_t(tex[i].texel_count) * 4);
				uo += tex[i].texel_count;
			}
			else
			{
				t.addr = uint32_t(fo);
				memcpy(&fpool[fo * 4], tex[i].data, size_t(tex[i].texel_count) * 16);
				fo += tex[i].texel_count;
			}
			c->textures[i] = t;
			desc[4 * i] = uint32_t(t.type), desc[4 * i + 1] = t.width, desc[4 * i + 2] = t.height, desc[4 * i + 3] = t.addr;
		}
		CK(cudaStreamSynchronize(c->stream));
		CK(c->d_uint_tex.reserve(upool.size() * 4));
		CK(c->d_float_tex.reserve(fpool.size() * 4));
		CK(c->d_tex_desc.reserve(desc.size() * 4));
		CK(cudaMemcpy(c->d_uint_tex.ptr, upool.data(), upool.size() * 4, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(c->d_float_tex.ptr, fpool.data(), fpool.size() * 4, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(c->d_tex_desc.ptr, desc.data(), desc.size() * 4, cudaMemcpyHostToDevice));
		c->scene.uint_texels = c->d_uint_tex.as<uint32_t>(), c->scene.uint_texel_count = uint32_t(upool.size());
		c->scene.float_texels = c->d_float_tex.as<float>(), c->scene.float_texel_count = uint32_t(fpool.size() / 4);
		if (!c->materials_raw.empty()) // materials uploaded earlier keep texture ids: point them at the new pool
		{
			bool all_known = true;
			for (const rfwb200_material_tex_ids &t : c->material_tex_ids)
				for (int k = 0; k < 11; k++)
					all_known &= t.texture[k] < int(count);
			if (all_known) // otherwise the caller is about to send matching materials (set_textures precedes set_materials)
				return resolve_and_upload_materials(c);
		}
		return RFWB200_OK;
	}

	int rfwb200_set_materials(rfwb200_context *c, const rfwb200_material *mats, const rfwb200_material_tex_ids *ids,
							  size_t count)
	{
		REQUIRE(c && (mats || count == 0), "bad materials");
		FORWARD(rfwb200_set_materials(p_, mats, ids, count));
		if (int r = ensure_device(c))
			return r;
		c->materials_raw.assign(mats, mats + count);
		if (ids)
			c->material_tex_ids.assign(ids, ids + count);
		else
			c->material_tex_ids.clear();
		return resolve_and_upload_materials(c);
	}

	// texaddr fields of the uploaded materials <- texel offsets of the current texture pool (CUDART/src/Context.cpp:167-191);
	// runs again when the textures change, so materials never point at the offsets of an older pool
	static int resolve_and_upload_materials(rfwb200_context *c)
	{
		const size_t count = c->materials_raw.size();
		c->materials = c->materials_raw;
		const rfwb200_material_tex_ids *ids = c->material_tex_ids.empty() ? nullptr : c->material_tex_ids.data();
		for (size_t i = 0; i < count && ids; i++)
		{
			// CUDART/src/Context.cpp:167-191 (slot 8 = ROUGHNESS1 has no descriptor)
			rfwb200_material &m = c->materials[i];
			rfwb200_map_desc *slots[11] = {&m.tex0, &m.tex1, &m.tex2, &m.nmap0, &m.nmap1, &m.nmap2,
										   &m.smap, &m.rmap, nullptr, &m.cmap,	&m.amap};
			for (int k = 0; k < 11; k++)
			{
				const int id = ids[i].texture[k];
				if (id == -1 || !slots[k])
					continue;
				REQUIRE(id >= 0 && size_t(id) < c->textures.size(), "material references a texture that was not set");
				slots[k]->texaddr = c->textures[id].addr;
			}
		}
		CK(cudaStreamSynchronize(c->stream));
		const size_t bytes = std::max<size_t>(count, 1) * sizeof(rfwb200_material);
		CK(c->d_materials.reserve(bytes));
		CK(c->d_materials_raw.reserve(bytes));
		if (count)
		{
			CK(cudaMemcpy(c->d_materials.ptr, c->materials.data(), count * sizeof(rfwb200_material), cudaMemcpyHostToDevice));
			CK(cudaMemcpy(c->d_materials_raw.ptr, c->materials_raw.data(), count * sizeof(rfwb200_material),
						  cudaMemcpyHostToDevice));
		}
		c->scene.materials = c->d_materials.ptr, c->scene.material_count = uint32_t(count);
		return RFWB200_OK;
	}

	int rfwb200_set_mesh(rfwb200_context *c, size_t index, const rfwb200_mesh *mesh)
	{
		REQUIRE(c && mesh, "bad mesh");
		REQUIRE(mesh->vertices && mesh->triangles, "mesh needs vertices and triangles");
		REQUIRE(index < (1u << 24), "mesh index too large");
		FORWARD(rfwb200_set_mesh(p_, index, mesh));
		if (index >= c->meshes.size())
			c->meshes.resize(index + 1), c->topology_dirty = true;
		HostMesh &m = c->meshes[index];
		// same vertex and triangle count => refit (EmbreeRT/src/Mesh.cpp:21-36, top_level_bvh.cpp:26)
		if (m.vertices.size() != mesh->vertex_count * 4 || m.triangles.size() != mesh->triangle_count ||
			m.indices.size() != (mesh->indices ? mesh->triangle_count * 3 : 0))
			c->topology_dirty = true;
		m.vertices.assign(mesh->vertices, mesh->vertices + mesh->vertex_count * 4);
		if (mesh->indices)
			m.indices.assign(mesh->indices, mesh->indices + mesh->triangle_count * 3);
		else
			m.indices.clear();
		m.triangles.assign(mesh->triangles, mesh->triangles + mesh->triangle_count);
		m.version++;
		c->geometry_dirty = true;
		if (c->mesh_dirty.size() < c->meshes.size())
			c->mesh_dirty.resize(c->meshes.size(), 1);
		c->mesh_dirty[index] = 1;
		if (c->topology_dirty)
			c->arena_valid = false;
		if (index < c->skins.size() && c->skins[index])
		{
			if (c->skins[index]->vertex_count != mesh->vertex_count)
				c->skins[index].reset(); // a different mesh: the skin no longer applies
			else
				c->skins[index]->device_newer = false;
		}
		return RFWB200_OK;
	}

	int rfwb200_set_instance(rfwb200_context *c, size_t i, size_t mesh_index, const float transform[16],
							 const float normal_matrix[9])
	{
		REQUIRE(c && transform && normal_matrix, "bad instance");
		REQUIRE(mesh_index < c->meshes.size(), "instance references a mesh that was not set");
		REQUIRE(i <= c->instances.size() + (1u << 20), "instance index too large");
		FORWARD(rfwb200_set_instance(p_, i, mesh_index, transform, normal_matrix));
		if (i >= c->instances.size())
			c->instances.resize(i + 1), c->topology_dirty = true;
		HostInstance &in = c->instances[i];
		if (in.mesh != int(mesh_index))
			c->topology_dirty = true;
		in.mesh = int(mesh_index);
		memcpy(in.transform, transform, sizeof(float) * 16);
		memcpy(in.normal, normal_matrix, sizeof(float) * 9);
		c->inst_moved.resize(c->instances.size(), 0);
		c->inst_moved[i] = 1;
		c->geometry_dirty = true;
		return RFWB200_OK;
	}

	int rfwb200_set_lights(rfwb200_context *c, rfwb200_light_count n, const rfwb200_area_light *a,
						   const rfwb200_point_light *p, const rfwb200_spot_light *s, const rfwb200_directional_light *d)
	{
		REQUIRE(c != nullptr, "context is null");
		REQUIRE((n.area == 0 || a) && (n.point == 0 || p) && (n.spot == 0 || s) && (n.directional == 0 || d),
				"light count without array");
		FORWARD(rfwb200_set_lights(p_, n, a, p, s, d));
		if (int r = ensure_device(c))
			return r;
		CK(cudaStreamSynchronize(c->stream));
		CK(c->d_area.reserve(std::max<size_t>(n.area, 1) * sizeof(rfwb200_area_light)));
		CK(c->d_point.reserve(std::max<size_t>(n.point, 1) * sizeof(rfwb200_point_light)));
		CK(c->d_spot.reserve(std::max<size_t>(n.spot, 1) * sizeof(rfwb200_spot_light)));
		CK(c->d_dir.reserve(std::max<size_t>(n.directional, 1) * sizeof(rfwb200_directional_light)));
		if (n.area)
			CK(cudaMemcpy(c->d_area.ptr, a, n.area * sizeof(*a), cudaMemcpyHostToDevice));
		if (n.point)
			CK(cudaMemcpy(c->d_point.ptr, p, n.point * sizeof(*p), cudaMemcpyHostToDevice));
		if (n.spot)
			CK(cudaMemcpy(c->d_spot.ptr, s, n.spot * sizeof(*s), cudaMemcpyHostToDevice));
		if (n.directional)
			CK(cudaMemcpy(c->d_dir.ptr, d, n.directional * sizeof(*d), cudaMemcpyHostToDevice));
		c->scene.area_lights = c->d_area.ptr, c->scene.point_lights = c->d_point.ptr;
		c->scene.spot_lights = c->d_spot.ptr, c->scene.dir_lights = c->d_dir.ptr;
		c->scene.lights = LightCounts{n.area, n.point, n.spot, n.directional};
		return RFWB200_OK;
	}

	static int update_one(rfwb200_context *c);

	int rfwb200_update(rfwb200_context *c)
	{
		REQUIRE(c != nullptr, "context is null");
		// rank 0 of a group builds (or refits) on the host once; the other ranks adopt its tree and only run the uploads
		// and the device kernels on their own GPU — every device traverses the SAME tree, so the frame does not depend on
		// the number of devices
		if (int r = update_one(c))
			return r;
		FORWARD(update_one(p_));
		return ensure_device(c);
	}

	static int update_two_level(rfwb200_context *c);

	static int update_one(rfwb200_context *c)
	{
		if (int r = ensure_device(c))
			return r;
		if (!c->geometry_dirty && c->scene.nodes)
			return RFWB200_OK;
		{
			uint64_t flattened = 0;
			for (const HostInstance &in : c->instances)
				if (in.mesh >= 0 && size_t(in.mesh) < c->meshes.size())
					flattened += c->meshes[in.mesh].triangles.size();
			if (c->levels_setting == 2 || (c->levels_setting == 0 && flattened > c->flatten_budget))
				return update_two_level(c);
			if (c->two_level) // back to the flattened form: nothing of the two-level scene is reused
			{
				c->two_level = false, c->scene.tl_instances = nullptr, c->scene.tl_instance_count = 0, c->scene.tl_inst_map = nullptr;
				c->group_trees.clear(), c->tl_table.clear(), c->tl_inst_map.clear();
				c->bvh = BvhBuildResult();
				c->arena_valid = false, c->topology_dirty = true, c->device_built = false;
				c->built_layout.clear();
			}
		}
		// topology unchanged since the last build => refit (bvh_tree.cpp:104-114)
		std::vector<std::pair<int, size_t>> layout;
		size_t total = 0;
		for (const HostInstance &in : c->instances)
		{
			const bool ok = in.mesh >= 0 && size_t(in.mesh) < c->meshes.size();
			layout.emplace_back(in.mesh, ok ? c->meshes[in.mesh].triangles.size() : 0);
			total += layout.back().second;
		}
		// a triangle's material index is used unchecked by the shade kernel: validate the meshes that changed
		c->mesh_dirty.resize(c->meshes.size(), 1);
		for (size_t mi = 0; mi < c->meshes.size(); mi++)
			if (c->mesh_dirty[mi])
				for (const rfwb200_triangle &t : c->meshes[mi].triangles)
					if (t.material >= c->materials_raw.size())
						return set_error(RFWB200_ERR_INVALID, "mesh " + std::to_string(mi) + " references material " + std::to_string(t.material) +
																  " but only " + std::to_string(c->materials_raw.size()) + " materials were set");
		const bool can_refit = !c->topology_dirty && layout == c->built_layout && (!c->bvh.nodes.empty() || !c->bvh.cw_nodes.empty() || c->device_built) && c->built_tri_count == total;
		const bool device = c->device_geometry && total > 0;
		c->mesh_dirty.resize(c->meshes.size(), 1);
		if (can_refit)
			for (size_t mi = 0; mi < c->meshes.size(); mi++)
				if (c->mesh_dirty[mi])
					mark_mesh_moved(c, mi);
		if (device && can_refit && c->arena_valid && !c->wide8)
		{
			// Device refit: only the meshes that changed and the 128-B instance records cross PCIe; the boxes, the
			// intersection records and the shading records are regenerated by two launches on the render stream.
			for (size_t mi = 0; mi < c->meshes.size(); mi++)
				if (c->mesh_dirty[mi])
					if (int r = upload_mesh_to_arena(c, mi))
						return r;
			if (int r = upload_instances(c, false))
				return r;
			if (int r = device_generate(c, true))
				return r;
			c->refits++, c->last_update_on_device = true, c->last_update_was_refit = true;
			c->geometry_dirty = false, c->sort_grid_dirty = true;
			return RFWB200_OK;
		}
		CK(cudaStreamSynchronize(c->stream));
		if (int r = sync_skinned_meshes_to_host(c))
			return r;
		if (c->lbvh && device && !c->wide8)
		{
			// Device build (lbvh.h): the host only re-uploads the arena; Morton sort, radix tree, boxes and the 4-wide
			// collapse run as kernels, then the same record / pack kernels as after a host build.
			REQUIRE(total < (1u << 30), "more than 2^30 triangles");
			c->flat_tri_count = total;
			c->bvh = BvhBuildResult(); // no host copy of this tree
			c->inst_moved.assign(c->instances.size(), 0);
			if (int r = upload_arena(c))
				return r;
			if (int r = upload_instances(c, true))
				return r;
			const size_t ref_cap = c->lbvh_presplit ? total + total / 2 + 1024 : total; // early split clipping: up to 50 % more references
			const size_t nn_cap = ref_cap;													  // a 4-wide node has at least two children
			CK(c->d_nodes.reserve(nn_cap * sizeof(BvhNode4)));
			CK(c->d_tris.reserve(ref_cap * sizeof(TriRec)));
			CK(c->d_tri_order.reserve(ref_cap * sizeof(uint32_t)));
			CK(c->d_parent_slot.reserve(nn_cap * sizeof(uint32_t)));
			CK(c->d_arrivals.reserve(nn_cap * sizeof(uint32_t)));
			CK(c->d_shade_tris.reserve(total * sizeof(ShadeTri)));
			CK(c->d_ref_boxes.reserve(ref_cap * 6 * sizeof(float)));
			CK(c->d_lbvh_scratch.reserve(lbvh_scratch_bytes(total, ref_cap)));
			c->device_ref_boxes = false;
			bind_geometry_views(c, 0, total);
			if (!c->ev_geo_a)
			{
				CK(cudaEventCreate(&c->ev_geo_a));
				CK(cudaEventCreate(&c->ev_geo_b));
			}
			const auto tb0 = std::chrono::steady_clock::now();
			CK(cudaEventRecord(c->ev_geo_a, c->stream));
			uint32_t nn = 0, nrefs = 0;
			int depth = 0, launches = 0;
			CK(lbvh_build(c->geo, (c->lbvh_presplit ? 1 : 0) | (c->ploc ? 2 : 0), c->d_lbvh_scratch.ptr, c->d_lbvh_scratch.bytes, ref_cap, c->d_nodes.as<BvhNode4>(),
						  c->d_tri_order.as<uint32_t>(), c->d_parent_slot.as<uint32_t>(), c->d_ref_boxes.as<float>(), &nn, &nrefs, &depth,
						  &launches, c->stream));
			c->launches += uint64_t(launches);
			c->device_ref_boxes = true;
			if (3 * depth + 1 > TRAVERSAL_STACK)
				return set_error(RFWB200_ERR_INVALID, "LBVH deeper than the traversal stack allows (use builder=sbvh)");
			bind_geometry_views(c, nn, nrefs);
			if (int r = device_generate(c, false, false))
				return r;
			CK(cudaStreamSynchronize(c->stream));
			c->bvh.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tb0).count();
			c->bvh.depth = depth;
			c->device_built = true, c->device_depth = depth;
			c->built_layout = layout, c->built_tri_count = total;
			c->builds++, c->last_update_on_device = true, c->last_update_was_refit = false;
			c->geometry_dirty = false, c->topology_dirty = false, c->sort_grid_dirty = true;
			c->dims_valid = false;
			return RFWB200_OK;
		}
		REQUIRE(!(c->lbvh && !c->device_geometry), "builder=lbvh needs the device geometry path (refit=device)");
		c->device_ref_boxes = false;
		REQUIRE(!(can_refit && c->device_built && c->bvh.nodes.empty()), "a device-built tree can only be refitted on the device (refit=device)");
		c->device_built = false;
		std::vector<ShadeTri> shade;
		std::vector<float> det_eps;
		if (int r = flatten_scene(c, device ? nullptr : &shade, device ? nullptr : &det_eps))
			return r;
		const auto t0 = std::chrono::steady_clock::now();
		const rfwb200_context *donor = c->build_donor;
		if (donor && !donor->device_built && donor->built_layout == layout && donor->bvh.wide8 == c->wide8 &&
			donor->built_tri_count == c->build_tris.size() && (!donor->bvh.nodes.empty() || !donor->bvh.cw_nodes.empty()))
		{
			c->bvh = donor->bvh; // topology + boxes as rank 0 built / refitted them
			c->built_tri_count = c->build_tris.size();
			if (!can_refit)
				c->inst_moved.assign(c->instances.size(), 0);
			(can_refit ? c->refits : c->builds)++;
		}
		else if (can_refit && c->bvh.wide8 == c->wide8)
		{
			c->tri_moved.clear();
			c->inst_moved.resize(c->instances.size(), 0);
			for (size_t ii = 0; ii < c->instances.size(); ii++)
				c->tri_moved.insert(c->tri_moved.end(), layout[ii].second, c->inst_moved[ii]);
			if (c->wide8)
				refit_cwbvh(c->build_tris.data(), c->build_tris.size(), c->bvh, c->tri_moved.data());
			else
				refit_bvh4(c->build_tris.data(), c->build_tris.size(), c->bvh, c->tri_moved.data());
			c->refits++;
		}
		else
		{
			const int threads = int(std::max(1u, std::thread::hardware_concurrency()));
			if (c->wide8)
				build_cwbvh(c->build_tris.data(), c->build_tris.size(), threads, c->bvh, c->spatial_splits);
			else
				build_bvh4(c->build_tris.data(), c->build_tris.size(), threads, c->bvh, c->spatial_splits);
			c->built_tri_count = c->build_tris.size();
			c->inst_moved.assign(c->instances.size(), 0);
			c->builds++;
		}
		c->bvh.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		if (c->wide8 ? (c->bvh.depth > CW_MAX_DEPTH) : (3 * c->bvh.depth + 1 > TRAVERSAL_STACK))
			return set_error(RFWB200_ERR_INVALID, "BVH deeper than the traversal stack allows");
		c->built_layout = layout;
		c->last_update_on_device = false, c->last_update_was_refit = can_refit;
		c->geo_timed = false;
		if (device)
		{
			// the host only builds the topology; records are produced on the device from the arena (the builder's
			// clipped boxes of spatially split references are kept: write_boxes = 0)
			if (int r = upload_arena(c))
				return r;
			if (int r = upload_instances(c, true))
				return r;
			if (int r = upload_topology(c))
				return r;
			if (int r = device_generate(c, false))
				return r;
		}
		else
		{
			REQUIRE(!c->wide8, "bvh=8 needs refit=device and a non-empty scene");
			c->arena_valid = false;
			std::fill(c->mesh_dirty.begin(), c->mesh_dirty.end(), uint8_t(0)); // consumed by the host flatten above
			c->scene.cw_nodes = nullptr, c->scene.cw_node_count = 0;
			if (int r = upload_bvh(c, det_eps))
				return r;
			CK(c->d_shade_tris.reserve(std::max<size_t>(shade.size(), 1) * sizeof(ShadeTri)));
			if (!shade.empty())
				CK(cudaMemcpy(c->d_shade_tris.ptr, shade.data(), shade.size() * sizeof(ShadeTri), cudaMemcpyHostToDevice));
			c->scene.shade_tris = c->d_shade_tris.as<ShadeTri>();
		}
		if (!device)
			if (int r = pack_nodes(c))
				return r;
		c->geometry_dirty = false, c->topology_dirty = false, c->sort_grid_dirty = true;
		c->dims_valid = false; // node count may have changed the staged prefix
		return RFWB200_OK;
	}

	// Two-level commit (setting "levels"): object-space trees below a top-level tree over instances — what the reference keeps
	// (CUDART/src/Context.cpp:270-311,394-456), and the fallback for scenes whose flattened form would not fit.
	//
	// The reference builds one tree per MESH, so a model of 394 meshes instanced 38 times is 14,972 overlapping instance boxes
	// a ray has to enter one after the other.  Here the unit below the top level is a GROUP: all meshes that are instanced
	// with exactly the same list of (transform, normal matrix) pairs share one object space, so their triangles go into ONE
	// tree and every entry of the list becomes ONE top-level instance (the 394 x 38 model: one tree of 262 k triangles, 38
	// instances; a model loaded once under a common transform: one tree, one instance — the speed of the flattened form at
	// the memory of the unique meshes).  A hit names (top-level instance, mesh triangle); the (instance, primitive) pair the
	// caller knows is looked up in a small table (tl_inst_map).  Group trees are cached by their member meshes: moving an
	// instance rebuilds the top level only, on the host (microseconds per thousand instances).
	static int update_two_level(rfwb200_context *c)
	{
		REQUIRE(!c->wide8, "bvh=8 is a flattened layout: use levels=1");
		for (const auto &sk : c->skins)
			REQUIRE(!sk, "device skinning / morph targets write the flattened scene's arena: use levels=1 for animated meshes");
		CK(cudaStreamSynchronize(c->stream));
		const size_t nm = c->meshes.size();
		c->mesh_dirty.resize(nm, 1);
		for (size_t mi = 0; mi < nm; mi++)
			if (c->mesh_dirty[mi])
				for (const rfwb200_triangle &t : c->meshes[mi].triangles)
					if (t.material >= c->materials_raw.size())
						return set_error(RFWB200_ERR_INVALID, "mesh " + std::to_string(mi) + " references material " + std::to_string(t.material) +
																  " but only " + std::to_string(c->materials_raw.size()) + " materials were set");
		const auto t0 = std::chrono::steady_clock::now();
		const int threads = int(std::max(1u, std::thread::hardware_concurrency()));

		// ---- groups: meshes with identical sorted lists of (transform, normal matrix) -----------------------------------------
		std::vector<std::vector<Placement>> placed;
		std::vector<std::vector<uint32_t>> groups; // member meshes, ascending
		{
			std::vector<uint8_t> has_tris(nm, 0);
			for (size_t mi = 0; mi < nm; mi++)
				has_tris[mi] = c->meshes[mi].triangles.empty() ? 0 : 1;
			group_by_placement(c->instances, has_tris, placed, groups);
		}

		// ---- one tree per group, cached by its member list -----------------------------------------------------------------------
		std::vector<Ctx::GroupTree> trees(groups.size());
		std::vector<BuildTriangle> bt;
		for (size_t g = 0; g < groups.size(); g++)
		{
			bool reuse = false;
			for (Ctx::GroupTree &old : c->group_trees)
				if (old.members == groups[g])
				{
					reuse = true;
					for (uint32_t mi : groups[g])
						reuse = reuse && !c->mesh_dirty[mi];
					if (reuse)
						trees[g] = std::move(old), old.members.clear();
					break;
				}
			if (reuse)
				continue;
			Ctx::GroupTree &gt = trees[g];
			gt.members = groups[g];
			size_t total = 0;
			for (uint32_t mi : groups[g])
				total += c->meshes[mi].triangles.size();
			bt.resize(total);
			gt.tri_mesh_rank.resize(total), gt.tri_index.resize(total);
			size_t at = 0;
			for (size_t r = 0; r < groups[g].size(); r++)
			{
				const HostMesh &m = c->meshes[groups[g][r]];
				const size_t nt = m.triangles.size(), nv = m.vertices.size() / 4;
				for (size_t t = 0; t < nt; t++, at++)
				{
					for (int k = 0; k < 3; k++)
					{
						const uint32_t vi = m.indices.empty() ? uint32_t(t * 3 + k) : m.indices[t * 3 + k];
						if (vi >= nv)
							return set_error(RFWB200_ERR_INVALID, "mesh index out of range");
						float *dst = k == 0 ? bt[at].v0 : (k == 1 ? bt[at].v1 : bt[at].v2);
						dst[0] = m.vertices[4 * vi], dst[1] = m.vertices[4 * vi + 1], dst[2] = m.vertices[4 * vi + 2];
					}
					gt.tri_mesh_rank[at] = uint32_t(r), gt.tri_index[at] = uint32_t(t);
				}
			}
			build_bvh4(bt.data(), total, threads, gt.bvh, c->spatial_splits);
			c->builds++;
		}
		c->group_trees = std::move(trees);
		std::fill(c->mesh_dirty.begin(), c->mesh_dirty.end(), uint8_t(0));

		// ---- layout: [top-level nodes][group 0 nodes][group 1 nodes]...; records per group reference; shading records per mesh triangle
		std::vector<uint32_t> tri_base(nm, 0);
		size_t ntri = 0;
		for (size_t mi = 0; mi < nm; mi++)
			tri_base[mi] = uint32_t(ntri), ntri += c->meshes[mi].triangles.size();
		std::vector<float> boxes;
		std::vector<TlInstance> table;
		std::vector<uint32_t> inst_map;
		std::vector<uint32_t> table_group;
		int deepest = 0;
		for (size_t g = 0; g < groups.size(); g++)
		{
			const Ctx::GroupTree &gt = c->group_trees[g];
			deepest = std::max(deepest, gt.bvh.depth);
			float b[6] = {3.0e38f, 3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f};
			const BvhNode4 &root = gt.bvh.nodes[0];
			for (int s = 0; s < root.pad[0]; s++)
			{
				b[0] = std::min(b[0], root.minx[s]), b[1] = std::min(b[1], root.miny[s]), b[2] = std::min(b[2], root.minz[s]);
				b[3] = std::max(b[3], root.maxx[s]), b[4] = std::max(b[4], root.maxy[s]), b[5] = std::max(b[5], root.maxz[s]);
			}
			const std::vector<Placement> &list = placed[groups[g][0]];
			for (size_t j = 0; j < list.size(); j++)
			{
				const float *M = list[j].m;
				float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
				for (int corner = 0; corner < 8; corner++)
				{
					const float p[3] = {b[(corner & 1) ? 3 : 0], b[(corner & 2) ? 4 : 1], b[(corner & 4) ? 5 : 2]};
					float w[3];
					mul_point(M, p, w);
					for (int a = 0; a < 3; a++)
						lo[a] = std::min(lo[a], w[a]), hi[a] = std::max(hi[a], w[a]);
				}
				// a float step of slack per coordinate: the object-space walk rounds differently from a world-space box test
				for (int a = 0; a < 3; a++)
					boxes.push_back(lo[a] - (4e-7f * std::max(std::fabs(lo[a]), std::fabs(hi[a])) + 1e-30f));
				for (int a = 0; a < 3; a++)
					boxes.push_back(hi[a] + (4e-7f * std::max(std::fabs(lo[a]), std::fabs(hi[a])) + 1e-30f));
				TlInstance ti;
				memset(&ti, 0, sizeof(ti));
				// inverse of the affine transform (column-major M), in double, rounded once
				const double a = M[0], bb = M[4], cc = M[8], d = M[1], e = M[5], f = M[9], gg = M[2], h = M[6], k = M[10];
				const double det = a * (e * k - f * h) - bb * (d * k - f * gg) + cc * (d * h - e * gg);
				REQUIRE(det != 0.0, "instance " + std::to_string(list[j].instance) + " has a singular transform");
				const double id = 1.0 / det;
				const double r[3][3] = {{(e * k - f * h) * id, (cc * h - bb * k) * id, (bb * f - cc * e) * id},
										{(f * gg - d * k) * id, (a * k - cc * gg) * id, (cc * d - a * f) * id},
										{(d * h - e * gg) * id, (bb * gg - a * h) * id, (a * e - bb * d) * id}};
				const double tx = M[12], ty = M[13], tz = M[14];
				for (int row = 0; row < 3; row++)
				{
					ti.inv[4 * row + 0] = float(r[row][0]), ti.inv[4 * row + 1] = float(r[row][1]), ti.inv[4 * row + 2] = float(r[row][2]);
					ti.inv[4 * row + 3] = float(-(r[row][0] * tx + r[row][1] * ty + r[row][2] * tz));
				}
				memcpy(ti.normal, M + 16, sizeof(ti.normal));
				ti.pad[0] = uint32_t(inst_map.size()); // first entry of this instance's (mesh rank -> caller's instance) row
				for (uint32_t mi : groups[g])
					inst_map.push_back(placed[mi][j].instance);
				table.push_back(ti);
				table_group.push_back(uint32_t(g));
			}
		}
		std::vector<BvhNode4> nodes;
		const int top_depth = build_tlas4(boxes.data(), table.size(), nodes);
		const size_t top_nodes = nodes.size();
		std::vector<uint32_t> node_base(groups.size(), 0), ref_base(groups.size(), 0);
		size_t nn = top_nodes, nr = 0;
		for (size_t g = 0; g < groups.size(); g++)
		{
			node_base[g] = uint32_t(nn), ref_base[g] = uint32_t(nr);
			nn += c->group_trees[g].bvh.nodes.size(), nr += c->group_trees[g].bvh.tri_order.size();
		}
		REQUIRE(nn < (1ull << 30) && nr < (1ull << 29) && ntri < (1ull << 31), "two-level scene exceeds the 32-bit node / reference encoding");
		REQUIRE(table.size() < (1ull << 29), "too many instances");
		if (3 * (top_depth + deepest) + 3 > TRAVERSAL_STACK)
			return set_error(RFWB200_ERR_INVALID, "two-level scene deeper than the traversal stack allows");
		for (size_t i = 0; i < table.size(); i++)
			table[i].blas_root = node_base[table_group[i]];
		nodes.resize(nn);
		std::vector<TriRec> recs(std::max<size_t>(nr, 1));
		std::vector<ShadeTri> shade(std::max<size_t>(ntri, 1));
		memset(recs.data(), 0, recs.size() * sizeof(TriRec));
		memset(shade.data(), 0, shade.size() * sizeof(ShadeTri));
		for (size_t g = 0; g < groups.size(); g++)
		{
			const Ctx::GroupTree &gt = c->group_trees[g];
			const BvhBuildResult &b = gt.bvh;
			for (size_t k = 0; k < b.nodes.size(); k++)
			{
				BvhNode4 n = b.nodes[k];
				for (int s = 0; s < n.pad[0]; s++)
				{
					if (n.child[s] >= 0)
						n.child[s] += int32_t(node_base[g]);
					else
					{
						const uint32_t v = uint32_t(~n.child[s]);
						n.child[s] = ~int32_t((((v >> 2) + ref_base[g]) << 2) | (v & 3u));
					}
				}
				nodes[node_base[g] + k] = n;
			}
			for (size_t i = 0; i < b.tri_order.size(); i++)
			{
				const uint32_t gt_tri = b.tri_order[i];
				const uint32_t mi = groups[g][gt.tri_mesh_rank[gt_tri]], t = gt.tri_index[gt_tri];
				const HostMesh &m = c->meshes[mi];
				uint32_t vi[3];
				for (int k = 0; k < 3; k++)
					vi[k] = m.indices.empty() ? uint32_t(t * 3 + k) : m.indices[t * 3 + k];
				const float *v0 = &m.vertices[4 * vi[0]], *v1 = &m.vertices[4 * vi[1]], *v2 = &m.vertices[4 * vi[2]];
				TriRec &r = recs[ref_base[g] + i];
				r.p0x = v0[0], r.p0y = v0[1], r.p0z = v0[2];
				r.e1x = v1[0] - v0[0], r.e1y = v1[1] - v0[1], r.e1z = v1[2] - v0[2];
				r.e2x = v2[0] - v0[0], r.e2y = v2[1] - v0[1], r.e2z = v2[2] - v0[2];
				r.shade_idx = tri_base[mi] + t;
				r.det_eps = 1e-6f; // T_EPSILON, tested in object space like the reference (CUDAIntersect.h:61-63)
			}
			for (size_t rank = 0; rank < groups[g].size(); rank++)
			{
				const uint32_t mi = groups[g][rank];
				const HostMesh &m = c->meshes[mi];
				for (size_t t = 0; t < m.triangles.size(); t++)
				{
					const rfwb200_triangle &src = m.triangles[t];
					ShadeTri &st = shade[tri_base[mi] + t];
					st.u0 = src.u0, st.u1 = src.u1, st.u2 = src.u2, st.light_tri_idx = src.light_tri_idx;
					st.v0 = src.v0, st.v1 = src.v1, st.v2 = src.v2, st.material = src.material;
					st.n0x = src.vN0[0], st.n0y = src.vN0[1], st.n0z = src.vN0[2];
					st.n1x = src.vN1[0], st.n1y = src.vN1[1], st.n1z = src.vN1[2];
					st.n2x = src.vN2[0], st.n2y = src.vN2[1], st.n2z = src.vN2[2];
					st.Nx = src.Nx, st.Ny = src.Ny, st.Nz = src.Nz; // the instance's normal matrix is applied when the hit is shaded
					st.area = src.area, st.lod = src.LOD;
					st.inst_id = uint32_t(rank); // column of the instance's row in tl_inst_map
					st.prim_id = uint32_t(t);
				}
			}
		}
		if (table.empty())
			table.emplace_back(), memset(&table[0], 0, sizeof(TlInstance));
		if (inst_map.empty())
			inst_map.push_back(0);
		CK(c->d_nodes.reserve(nodes.size() * sizeof(BvhNode4)));
		CK(c->d_tris.reserve(recs.size() * sizeof(TriRec)));
		CK(c->d_shade_tris.reserve(shade.size() * sizeof(ShadeTri)));
		CK(c->d_tl_instances.reserve(table.size() * sizeof(TlInstance)));
		CK(c->d_tl_inst_map.reserve(inst_map.size() * sizeof(uint32_t)));
		CK(cudaMemcpyAsync(c->d_nodes.ptr, nodes.data(), nodes.size() * sizeof(BvhNode4), cudaMemcpyHostToDevice, c->stream));
		CK(cudaMemcpyAsync(c->d_tris.ptr, recs.data(), recs.size() * sizeof(TriRec), cudaMemcpyHostToDevice, c->stream));
		CK(cudaMemcpyAsync(c->d_shade_tris.ptr, shade.data(), shade.size() * sizeof(ShadeTri), cudaMemcpyHostToDevice, c->stream));
		CK(cudaMemcpyAsync(c->d_tl_instances.ptr, table.data(), table.size() * sizeof(TlInstance), cudaMemcpyHostToDevice, c->stream));
		CK(cudaMemcpyAsync(c->d_tl_inst_map.ptr, inst_map.data(), inst_map.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
		CK(cudaStreamSynchronize(c->stream)); // the staging vectors die here
		c->tl_table = std::move(table), c->tl_inst_map = std::move(inst_map);
		c->scene.nodes = c->d_nodes.as<BvhNode4>(), c->scene.tris = c->d_tris.as<TriRec>(), c->scene.shade_tris = c->d_shade_tris.as<ShadeTri>();
		c->scene.node_count = uint32_t(nn), c->scene.tri_count = uint32_t(nr);
		c->scene.cw_nodes = nullptr, c->scene.cw_node_count = 0;
		if (int r = pack_nodes(c)) // the 80-byte form of every tree (top level and groups alike) for the wavefront kernel
			return r;
		c->scene.tl_instances = c->d_tl_instances.as<TlInstance>(), c->scene.tl_instance_count = uint32_t(c->tl_table.size());
		c->scene.tl_inst_map = c->d_tl_inst_map.as<uint32_t>();
		c->flat_tri_count = ntri;
		c->two_level = true, c->tl_depth = top_depth + deepest, c->tl_groups = uint32_t(groups.size());
		c->bvh = BvhBuildResult();
		c->bvh.depth = c->tl_depth;
		c->bvh.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		c->device_built = true; // get_bvh_info reads the counts from the device scene
		c->arena_valid = false, c->built_layout.clear(), c->built_tri_count = 0;
		c->inst_moved.assign(c->instances.size(), 0);
		c->last_update_on_device = false, c->last_update_was_refit = false, c->geo_timed = false;
		c->geometry_dirty = false, c->topology_dirty = false, c->sort_grid_dirty = true;
		c->dims_valid = false;
		return RFWB200_OK;
	}

	/* ---- device skinning (extension; replaces the CPU skinning in front of set_mesh, gltf/mesh.cpp:18-48) ---- */
	int rfwb200_set_mesh_skin(rfwb200_context *c, size_t mesh_index, const float *base_vertices, const float *base_normals,
							  const uint32_t *joints, const float *weights, size_t vertex_count)
	{
		REQUIRE(c && base_vertices && base_normals && joints && weights, "bad skin");
		REQUIRE(mesh_index < c->meshes.size(), "skin references a mesh that was not set");
		FORWARD(rfwb200_set_mesh_skin(p_, mesh_index, base_vertices, base_normals, joints, weights, vertex_count));
		REQUIRE(c->meshes[mesh_index].vertices.size() == vertex_count * 4, "skin and mesh vertex counts differ");
		if (int r = ensure_device(c))
			return r;
		if (c->skins.size() <= mesh_index)
			c->skins.resize(mesh_index + 1);
		auto sk = std::make_unique<Ctx::Skin>();
		sk->vertex_count = vertex_count;
		uint32_t max_joint = 0;
		for (size_t i = 0; i < vertex_count * 4; i++)
			max_joint = std::max(max_joint, joints[i]);
		sk->max_joint = max_joint;
		const size_t b16 = std::max<size_t>(vertex_count, 1) * 16;
		CK(sk->base_v.reserve(b16));
		CK(sk->base_n.reserve(b16));
		CK(sk->joints.reserve(b16));
		CK(sk->weights.reserve(b16));
		CK(sk->normals.reserve(b16));
		CK(cudaMemcpy(sk->base_v.ptr, base_vertices, vertex_count * 16, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(sk->base_n.ptr, base_normals, vertex_count * 16, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(sk->joints.ptr, joints, vertex_count * 16, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(sk->weights.ptr, weights, vertex_count * 16, cudaMemcpyHostToDevice));
		sk->has_skin = true;
		c->skins[mesh_index] = std::move(sk);
		return RFWB200_OK;
	}

	int rfwb200_set_mesh_morph_targets(rfwb200_context *c, size_t mesh_index, const float *pose_positions, const float *pose_normals,
									   size_t target_count, size_t vertex_count)
	{
		REQUIRE(c && pose_positions && pose_normals && target_count >= 1, "bad morph targets");
		REQUIRE(mesh_index < c->meshes.size(), "morph targets reference a mesh that was not set");
		FORWARD(rfwb200_set_mesh_morph_targets(p_, mesh_index, pose_positions, pose_normals, target_count, vertex_count));
		REQUIRE(c->meshes[mesh_index].vertices.size() == vertex_count * 4, "morph target and mesh vertex counts differ");
		if (int r = ensure_device(c))
			return r;
		if (c->skins.size() <= mesh_index)
			c->skins.resize(mesh_index + 1);
		auto sk = std::make_unique<Ctx::Skin>();
		sk->vertex_count = vertex_count, sk->n_targets = target_count;
		const size_t bytes = (target_count + 1) * vertex_count * 16;
		CK(sk->pose_p.reserve(bytes));
		CK(sk->pose_n.reserve(bytes));
		CK(sk->morph_w.reserve(target_count * sizeof(float)));
		CK(sk->normals.reserve(std::max<size_t>(vertex_count, 1) * 16));
		CK(cudaMemcpy(sk->pose_p.ptr, pose_positions, bytes, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(sk->pose_n.ptr, pose_normals, bytes, cudaMemcpyHostToDevice));
		c->skins[mesh_index] = std::move(sk);
		return RFWB200_OK;
	}

	int rfwb200_set_mesh_morph_weights(rfwb200_context *c, size_t mesh_index, const float *weights, size_t weight_count)
	{
		REQUIRE(c && weights, "bad morph weights");
		REQUIRE(mesh_index < c->skins.size() && c->skins[mesh_index] && c->skins[mesh_index]->n_targets > 0,
				"mesh has no morph targets (rfwb200_set_mesh_morph_targets)");
		FORWARD(rfwb200_set_mesh_morph_weights(p_, mesh_index, weights, weight_count));
		if (int r = ensure_device(c))
			return r;
		Ctx::Skin &sk = *c->skins[mesh_index];
		REQUIRE(weight_count == sk.n_targets, "one weight per morph target (gltf/mesh.cpp:128)");
		if (!c->device_geometry)
			return set_error(RFWB200_ERR_STATE, "device morphing needs setting refit=device");
		if (!c->arena_valid || c->topology_dirty)
			return set_error(RFWB200_ERR_STATE, "call rfwb200_update once after set_mesh before posing a mesh");
		if (c->mesh_dirty[mesh_index])
			if (int r = upload_mesh_to_arena(c, mesh_index))
				return r;
		const HostMesh &hm = c->meshes[mesh_index];
		CK(cudaMemcpyAsync(sk.morph_w.ptr, weights, weight_count * sizeof(float), cudaMemcpyHostToDevice, c->stream));
		MorphView v{};
		v.pose_positions = sk.pose_p.as<float4>(), v.pose_normals = sk.pose_n.as<float4>();
		v.weights = sk.morph_w.as<float>(), v.n_weights = uint32_t(weight_count);
		v.indices = c->d_indices.as<uint32_t>() + size_t(c->mesh_tri_off[mesh_index]) * 3;
		v.out_vertices = c->d_verts.as<float4>() + c->mesh_vert_off[mesh_index];
		v.out_normals = sk.normals.as<float4>();
		v.mesh_tris = static_cast<char *>(c->d_mesh_tris.ptr) + size_t(c->mesh_tri_off[mesh_index]) * sizeof(rfwb200_triangle);
		v.vertex_count = uint32_t(sk.vertex_count), v.triangle_count = uint32_t(hm.triangles.size());
		CK(launch_morph(v, c->stream));
		c->launches += 2;
		sk.device_newer = true;
		mark_mesh_moved(c, mesh_index);
		c->geometry_dirty = true;
		return RFWB200_OK;
	}

	int rfwb200_set_mesh_pose(rfwb200_context *c, size_t mesh_index, const float *joint_matrices, size_t joint_count)
	{
		REQUIRE(c && joint_matrices && joint_count > 0, "bad pose");
		REQUIRE(mesh_index < c->skins.size() && c->skins[mesh_index] && c->skins[mesh_index]->has_skin, "mesh has no skin (rfwb200_set_mesh_skin)");
		FORWARD(rfwb200_set_mesh_pose(p_, mesh_index, joint_matrices, joint_count));
		if (int r = ensure_device(c))
			return r;
		Ctx::Skin &sk = *c->skins[mesh_index];
		REQUIRE(sk.max_joint < joint_count, "a vertex references a joint beyond joint_count");
		if (!c->device_geometry)
			return set_error(RFWB200_ERR_STATE, "device skinning needs setting refit=device");
		if (!c->arena_valid || c->topology_dirty)
			return set_error(RFWB200_ERR_STATE, "call rfwb200_update once after set_mesh before posing a mesh");
		if (c->mesh_dirty[mesh_index]) // a newer host copy (uv, materials, ...) goes first, the pose on top of it
			if (int r = upload_mesh_to_arena(c, mesh_index))
				return r;
		const HostMesh &hm = c->meshes[mesh_index];
		CK(sk.matrices.reserve(joint_count * 64));
		CK(cudaMemcpyAsync(sk.matrices.ptr, joint_matrices, joint_count * 64, cudaMemcpyHostToDevice, c->stream));
		SkinView v{};
		v.base_vertices = sk.base_v.as<float4>(), v.base_normals = sk.base_n.as<float4>();
		v.joints = sk.joints.as<uint4>(), v.weights = sk.weights.as<float4>();
		v.joint_matrices = sk.matrices.as<float>();
		v.indices = c->d_indices.as<uint32_t>() + size_t(c->mesh_tri_off[mesh_index]) * 3;
		v.out_vertices = c->d_verts.as<float4>() + c->mesh_vert_off[mesh_index];
		v.out_normals = sk.normals.as<float4>();
		v.mesh_tris = static_cast<char *>(c->d_mesh_tris.ptr) + size_t(c->mesh_tri_off[mesh_index]) * sizeof(rfwb200_triangle);
		v.vertex_count = uint32_t(sk.vertex_count), v.triangle_count = uint32_t(hm.triangles.size());
		CK(launch_skin(v, c->stream));
		c->launches += 2;
		sk.device_newer = true;
		mark_mesh_moved(c, mesh_index);
		c->geometry_dirty = true;
		return RFWB200_OK;
	}

	int rfwb200_get_geometry_stats(rfwb200_context *c, rfwb200_geometry_stats *out)
	{
		REQUIRE(c && out, "bad arguments");
		memset(out, 0, sizeof(*out));
		out->on_device = c->last_update_on_device, out->was_refit = c->last_update_was_refit;
		out->host_ms = c->last_update_on_device ? 0.0f : float(c->bvh.build_ms);
		out->refits = c->refits, out->builds = c->builds;
		if (c->geo_timed)
		{
			CK(cudaEventSynchronize(c->ev_geo_b));
			CK(cudaEventElapsedTime(&out->device_ms, c->ev_geo_a, c->ev_geo_b));
		}
		return RFWB200_OK;
	}

	/* test hook: 0 = BvhNode4[], 1 = TriRec[], 2 = ShadeTri[] of the committed scene */
	int rfwb200_debug_read_scene(rfwb200_context *c, int which, void *host, size_t capacity_bytes, size_t *bytes_out)
	{
		REQUIRE(c && bytes_out, "bad arguments");
		REQUIRE(which >= 0 && which <= 2, "which must be 0 (nodes), 1 (triangles) or 2 (shading triangles)");
		if (int r = check_ready(c))
			return r;
		const void *src = which == 0 ? (c->scene.cw_nodes ? (const void *)c->scene.cw_nodes : (const void *)c->scene.nodes)
						  : which == 1 ? (const void *)c->scene.tris
									   : (const void *)c->scene.shade_tris;
		const size_t bytes = which == 0	  ? (c->scene.cw_nodes ? size_t(c->scene.cw_node_count) * sizeof(CwNode) : size_t(c->scene.node_count) * sizeof(BvhNode4))
							 : which == 1 ? size_t(c->scene.tri_count) * sizeof(TriRec)
										  : size_t(c->flat_tri_count) * sizeof(ShadeTri);
		*bytes_out = bytes;
		if (!host)
			return RFWB200_OK;
		REQUIRE(capacity_bytes >= bytes, "buffer too small");
		CK(cudaStreamSynchronize(c->stream));
		if (bytes)
			CK(cudaMemcpy(host, src, bytes, cudaMemcpyDeviceToHost));
		return RFWB200_OK;
	}

	int rfwb200_set_setting(rfwb200_context *c, const char *key, const char *value)
	{
		REQUIRE(c && key && value, "bad setting");
		FORWARD(rfwb200_set_setting(p_, key, value));
		const std::string k = key, v = value;
		if (k == "spp")
		{
			const int n = atoi(v.c_str());
			REQUIRE(n >= 1 && n <= 65536, "spp must be in [1, 65536]");
			c->spp = n;
		}
		else if (k == "mode")
		{
			REQUIRE(v == "pt" || v == "embree", "mode must be 'pt' or 'embree'");
			c->mode_pt = (v == "pt");
		}
		else if (k == "max_path_length")
		{
			const int n = atoi(v.c_str());
			REQUIRE(n >= 0 && n < MAX_DEPTH_SLOTS, "max_path_length must be in [0, 7]");
			c->rs.max_path_length = n;
		}
		else if (k == "clamp")
			c->rs.clamp_value = float(atof(v.c_str()));
		else if (k == "survival_scale")
			c->rs.survival_scale = (v == "on" || v == "1") ? 1 : 0;
		else if (k == "smem_nodes")
		{
			const int n = atoi(v.c_str());
			REQUIRE(n >= 0 && n <= 1700, "smem_nodes must be in [0, 1700] (227 KB of shared memory)");
			c->rs.smem_nodes = n;
			c->dims_valid = false;
		}
		else if (k == "sample_lanes")
		{
			// Round 1 ran the samples of a frame as separate 1-spp wavefronts on up to four streams; all samples now travel
			// in one wavefront (spp_batch), so the setting is accepted and has no effect.
			const int n = atoi(v.c_str());
			REQUIRE(n >= 1 && n <= 4, "sample_lanes must be in [1, 4]");
		}
		else if (k == "spp_batch")
		{
			const int n = atoi(v.c_str());
			REQUIRE(n >= 0 && n <= MAX_BATCH_SPP, "spp_batch must be in [0, 64] (0 = as many samples per wavefront as fit)");
			c->spp_batch = n;
		}
		else if (k == "sort")
		{
			REQUIRE(v == "on" || v == "off" || v == "1" || v == "0", "sort must be on or off");
			c->rs.sort_mode = (v == "on" || v == "1") ? 1 : 0;
		}
		else if (k == "sample_layout")
		{
			REQUIRE(v == "planes" || v == "pixel", "sample_layout must be 'planes' (a warp = one sample of 32 pixels) or 'pixel' (a warp = all samples of 32 / spp pixels)");
			c->sample_minor = (v == "pixel");
		}
		else if (k == "sort_cell_bits")
		{
			const int n = atoi(v.c_str());
			REQUIRE(n >= 3 && n <= 6, "sort_cell_bits must be in [3, 6] (grid cells per axis = 2^bits)");
			c->rs.sort_cell_bits = n, c->sort_grid_dirty = true;
		}
		else if (k == "sort_dir_bits")
		{
			const int n = atoi(v.c_str());
			REQUIRE(n == 3 || n == 5 || n == 6, "sort_dir_bits must be 3 (octant), 5 (octant x dominant axis) or 6 (8 x 8 octahedral map)");
			c->rs.sort_dir_bits = n;
		}
		else if (k == "shade_loop")
		{
			REQUIRE(v == "cursor" || v == "static", "shade_loop must be 'cursor' (32 paths per warp from a device cursor) or 'static' (grid strides, next job prefetched)");
			c->rs.shade_static = (v == "static") ? 1 : 0;
		}
		else if (k == "sort_major")
		{
			REQUIRE(v == "cell" || v == "octant", "sort_major must be 'cell' or 'octant'");
			c->rs.sort_dir_major = (v == "octant") ? 1 : 0;
		}
		else if (k == "spatial_splits")
		{
			c->spatial_splits = (v == "on" || v == "1");
			c->geometry_dirty = c->topology_dirty = true; // takes effect at the next update()
		}
		else if (k == "fetch_threshold")
		{
			const int n = atoi(v.c_str());
			REQUIRE(n >= 1 && n <= 32, "fetch_threshold must be in [1, 32]");
			c->rs.fetch_threshold = n;
		}
		else if (k == "aov")
		{
			REQUIRE(v == "on" || v == "off" || v == "1" || v == "0", "aov must be on or off");
			c->aov = (v == "on" || v == "1"); // takes effect with the next Reset frame
		}
		else if (k == "fetch_chunk")
		{
			const int n = atoi(v.c_str());
			REQUIRE(n == 0 || (n >= 32 && n <= (1 << 20)), "fetch_chunk must be 0 (one shared front) or in [32, 2^20]");
			c->rs.fetch_chunk = n;
		}
		else if (k == "timing")
			c->timing = (v == "on" || v == "1");
		else if (k == "trace_variant")
		{
			const int n = atoi(v.c_str());
			REQUIRE(n >= 0 && n <= 13, "trace_variant must be in [0, 13]");
			c->rs.trace_variant = c->rs.primary_variant = n; // one value for both kinds of launch; primary_variant overrides
		}
		else if (k == "primary_cache")
			c->rs.primary_cache = (v == "on" || v == "1") ? 1 : 0;
		else if (k == "shadow_cache")
			c->rs.shadow_cache = (v == "pixel" || v == "2") ? 2 : ((v == "on" || v == "lane" || v == "1") ? 1 : 0);
		else if (k == "primary_variant")
		{
			const int n = atoi(v.c_str());
			REQUIRE(n >= 0 && n <= 13, "primary_variant must be in [0, 13]");
			c->rs.primary_variant = n;
		}
		else if (k == "refit")
		{
			REQUIRE(v == "device" || v == "host", "refit must be 'device' or 'host'");
			const bool dev = (v == "device");
			if (dev != c->device_geometry)
				c->device_geometry = dev, c->arena_valid = false, c->geometry_dirty = c->topology_dirty = true;
		}
		else if (k == "builder")
		{
			REQUIRE(v == "sbvh" || v == "lbvh" || v == "ploc", "builder must be 'sbvh' (host), 'lbvh' or 'ploc' (device)");
			const bool lb = (v == "lbvh" || v == "ploc"), pl = (v == "ploc");
			if (lb != c->lbvh || pl != c->ploc)
				c->lbvh = lb, c->ploc = pl, c->arena_valid = false, c->geometry_dirty = c->topology_dirty = true;
		}
		else if (k == "lbvh_presplit")
		{
			const bool on = (v == "on" || v == "1");
			if (on != c->lbvh_presplit)
			{
				c->lbvh_presplit = on;
				if (c->lbvh)
					c->arena_valid = false, c->geometry_dirty = c->topology_dirty = true;
			}
		}
		else if (k == "bvh")
		{
			REQUIRE(v == "4" || v == "8", "bvh must be 4 or 8");
			const bool w8 = (v == "8");
			if (w8 != c->wide8)
				c->wide8 = w8, c->arena_valid = false, c->geometry_dirty = c->topology_dirty = true;
		}
		else if (k == "shade_math")
		{
			REQUIRE(v == "fast" || v == "ieee", "shade_math must be 'fast' or 'ieee'");
			c->shade_ieee = (v == "ieee");
		}
		else if (k == "levels")
		{
			REQUIRE(v == "1" || v == "2" || v == "auto", "levels must be 1 (instances flattened into one tree), 2 (top level over per-mesh trees) or auto");
			const int n = v == "auto" ? 0 : atoi(v.c_str());
			if (n != c->levels_setting)
				c->levels_setting = n, c->geometry_dirty = c->topology_dirty = true;
		}
		else if (k == "flatten_budget")
		{
			const long long n = atoll(v.c_str());
			REQUIRE(n >= 0, "flatten_budget is a number of flattened triangles");
			c->flatten_budget = uint64_t(n);
			if (c->levels_setting == 0)
				c->geometry_dirty = c->topology_dirty = true;
		}
		else if (k == "threads")
		{
		}
		else
			return set_error(RFWB200_ERR_INVALID, "unknown setting '" + k + "'");
		return RFWB200_OK;
	}

	int rfwb200_get_settings(const rfwb200_context *c, char *buf, size_t buf_size)
	{
		REQUIRE(c && buf && buf_size > 0, "bad buffer");
		const std::string s = "spp=" + std::to_string(c->spp) + "\nmode=pt|embree\nmax_path_length=" +
							  std::to_string(c->rs.max_path_length) + "\nclamp=" + std::to_string(c->rs.clamp_value) +
							  "\nsurvival_scale=on|off\nsmem_nodes=" + std::to_string(c->rs.smem_nodes) + "\nspp_batch=" +
							  std::to_string(c->spp_batch) + "\nsort=on|off\nsort_cell_bits=" + std::to_string(c->rs.sort_cell_bits) +
							  "\nsort_major=cell|octant\naov=on|off\nfetch_chunk=" + std::to_string(c->rs.fetch_chunk) + "\nfetch_threshold=" + std::to_string(c->rs.fetch_threshold) +
							  "\ntrace_variant=" + std::to_string(c->rs.trace_variant) + "\nprimary_variant=" +
							  std::to_string(c->rs.primary_variant) + "\nbvh=4|8\nbuilder=sbvh|lbvh|ploc\nspatial_splits=on|off\nrefit=device|host" +
							  "\nshade_math=fast|ieee\nshade_loop=" + (c->rs.shade_static ? "static" : "cursor") + "\ntiming=on|off\nlevels=1|2|auto\nflatten_budget=" + std::to_string(c->flatten_budget) + "\nlevels_in_use=" + (c->two_level ? "2" : "1") + "\ntop_level_instances=" + std::to_string(c->two_level ? c->tl_table.size() : 0) +
							  "\ninstance_groups=" + std::to_string(c->two_level ? c->tl_groups : 0) + "\n";
		snprintf(buf, buf_size, "%s", s.c_str());
		return RFWB200_OK;
	}

	static int render_one(rfwb200_context *c, const rfwb200_camera_view *view, int status);

	int rfwb200_render_frame(rfwb200_context *c, const rfwb200_camera_view *view, int status)
	{
		REQUIRE(c && view, "bad arguments");
		if (c->peers.empty())
			return render_one(c, view, status);
		// in-process group: every further device's frame is enqueued by its own host thread while this thread enqueues
		// rank 0's; all ranks write their tiles into rank 0's display image, and rank 0's stream then waits for the
		// arrivals, so whatever the caller enqueues next on this context sees the complete frame
		const rfwb200_camera_view v = *view;
		for (size_t i = 0; i < c->peers.size(); i++)
		{
			rfwb200_context *p = c->peers[i];
			c->workers[i]->submit([p, v, status](std::string &msg) {
				const int r = render_one(p, &v, status);
				if (r)
					msg = g_last_error;
				return r;
			});
		}
		int rc = render_one(c, view, status);
		std::string first_msg = rc ? g_last_error : std::string();
		for (size_t i = 0; i < c->peers.size(); i++)
		{
			std::string msg;
			const int r = c->workers[i]->wait(msg);
			if (r && !rc)
				rc = r, first_msg = "device " + std::to_string(c->peers[i]->device) + ": " + msg;
		}
		if (rc)
			return set_error(rc, first_msg);
		if (c->mode_pt)
			return rfwb200_display_wait(c);
		return RFWB200_OK;
	}

	static int render_one(rfwb200_context *c, const rfwb200_camera_view *view, int status)
	{
		if (int r = ensure_device(c))
			return r;
		if (int r = check_ready(c))
			return r;
		if (int r = ensure_dims(c))
			return r;
		const uint32_t spp = c->mode_pt ? uint32_t(c->spp) : 1u;
		const uint32_t bspp = batch_spp_for(c, spp), batches = (spp + bspp - 1) / bspp;
		if (int r = ensure_counters(c, batches))
			return r;
		if (c->mode_pt)
			if (int r = ensure_wavefront(c, bspp))
				return r;
		cudaStream_t st = c->stream;
		CK(cudaEventRecord(c->ev_begin, st));
		if (status == RFWB200_RESET)
		{
			CK(cudaMemsetAsync(c->d_acc.ptr, 0, size_t(c->shard.local_pixels) * sizeof(float4), st));
			if (c->aov && c->d_albedo.ptr)
			{
				CK(cudaMemsetAsync(c->d_albedo.ptr, 0, size_t(c->shard.local_pixels) * sizeof(float4), st));
				CK(cudaMemsetAsync(c->d_normal.ptr, 0, size_t(c->shard.local_pixels) * sizeof(float4), st));
			}
			c->sample_index = 0;
		}
		CK(cudaMemsetAsync(c->d_counters.ptr, 0, size_t(batches) * MAX_DEPTH_SLOTS * sizeof(DepthCounters), st));
		CK(cudaMemsetAsync(c->d_ext_seen.ptr, 0, size_t(batches) * MAX_DEPTH_SLOTS * MAX_BATCH_SPP * sizeof(uint32_t), st));
		const bool to_display = c->mode_pt && c->display.image != nullptr;
		// the hit caches remember positions in the flattened record array; a two-level scene has none
		RenderSettings rs_frame = c->rs;
		if (c->two_level)
			rs_frame.primary_cache = 0, rs_frame.shadow_cache = 0;
		if (c->sample_index == 0)
		{
			const ProbeResult none{0, 0, 0.f, 0};
			CK(cudaMemcpyAsync(c->d_probe.ptr, &none, sizeof(none), cudaMemcpyHostToDevice, st));
		}
		if (int r = upload_frame_params(c, view, c->sample_index))
			return r;
		c->last_first_sample = c->sample_index;
		c->stage_events_used = 0;
		if (c->mode_pt)
		{
			const uint32_t maxd = uint32_t(c->rs.max_path_length);
			const bool sort = c->rs.sort_mode != 0 && maxd > 0;
			auto shade = c->shade_ieee ? launch_shade_ieee : launch_shade;
			const bool shard_sync = to_display && c->shard.world > 1;
			if (sort && c->sort_grid_dirty)
			{
				CK(launch_sort_setup(c->scene, c->wf, rs_frame, st)); // grid of the bins from the root of the current tree
				c->launches += 1;
				c->sort_grid_dirty = false;
			}
			for (uint32_t b = 0; b < batches; b++)
			{
				BatchView bv;
				bv.spp = std::min(bspp, spp - b * bspp), bv.first_sample = b * bspp, bv.index = b;
				bv.items = c->shard.local_pixels * bv.spp, bv.inv_spp = 1.0f / float(bv.spp);
				bv.sample_minor = c->sample_minor ? 1u : 0u;
				ShardSync sync{};
				if (shard_sync)
				{
					sync.stamp = ++c->display.stamp;
					sync.seen = c->display.sync_seen + size_t(sync.stamp & 1u) * MAX_DEPTH_SLOTS * MAX_BATCH_SPP;
					sync.arrivals = c->display.sync_arrivals;
				}
				// ranks of a sharded frame merge their "sample s still has extension rays" flags behind every shade launch
				// whose connect rays are traced (k_shard_sync): one arrival + one short wait per bounce
				auto merge_flags = [&](uint32_t depth) -> int {
					if (!shard_sync || depth >= maxd)
						return RFWB200_OK;
					c->display.syncs++;
					// every rank issues the same sequence of syncs (maxd per wavefront); the lag of maxd - 1 syncs means "all ranks
					// have arrived at the first sync of the wavefront before this one's last", see k_shard_sync
					const uint32_t lag = maxd - 1u;
					const uint32_t lagged = c->display.syncs > lag ? (c->display.syncs - lag) * c->shard.world : 0u;
					CK(launch_shard_sync(sync, depth, c->display.syncs * c->shard.world, lagged, bv.spp,
										 c->wf.ext_seen + size_t(b * MAX_DEPTH_SLOTS + depth) * MAX_BATCH_SPP, c->d_display_local.as<uint32_t>() + 1, st));
					c->launches += 1;
					return RFWB200_OK;
				};
				{
					StageTimer t(c, 0);
					CK(launch_primary(c->scene, c->shard, c->wf, rs_frame, bv, c->dims, st));
				}
				{
					StageTimer t(c, 3);
					CK(shade(c->scene, c->shard, c->wf, rs_frame, bv, 0, 0, 1, sync, c->dims, st));
				}
				c->launches += 2;
				if (int r = merge_flags(0))
					return r;
				for (uint32_t d = 1; d <= maxd; d++)
				{
					// re-ordering on: shade appends to planes [1], the sort moves them into planes [0], trace and shade read [0];
					// off: the planes alternate per depth (Kernels.cu:578-584)
					const uint32_t in = sort ? 0u : (d & 1u), out = sort ? 1u : (in ^ 1u);
					if (sort)
					{
						StageTimer t(c, 5);
						CK(launch_sort(c->wf, rs_frame, bv, d, c->dims, st));
						c->launches += 2;
					}
					{
						StageTimer t(c, d == 1 ? 1 : 2);
						CK(launch_trace(c->scene, c->shard, c->wf, rs_frame, bv, d, in, c->dims, st));
					}
					{
						StageTimer t(c, 3);
						CK(shade(c->scene, c->shard, c->wf, rs_frame, bv, d, in, out, sync, c->dims, st));
					}
					c->launches += 2;
					if (int r = merge_flags(d))
						return r;
				}
				const bool last = b + 1 == batches;
				if (last && c->copy_pending) // an asynchronous read-back of the previous frame: the image is overwritten only after it
				{
					CK(cudaStreamWaitEvent(st, c->ev_copy_done, 0));
					c->copy_pending = false;
				}
				if (last && to_display && c->display.owner)
				{
					// everything that read the previous frame on this stream (or the asynchronous read-back) has run: the other
					// ranks may overwrite the image.  Placed in front of the fold, not of the frame, so that all ranks trace the
					// next frame while the previous one is still being read.
					CK(launch_display_release(c->display.consumed, c->display.frames, st));
					c->launches += 1;
				}
				if (last && to_display && !c->display.owner && c->display.frames > 0)
				{
					// the display rank must have released the previous frame before this rank's tiles overwrite it
					CK(launch_display_spin(c->display.consumed, c->display.frames, c->d_display_local.as<uint32_t>() + 1, st));
					c->launches += 1;
				}
				{
					StageTimer t(c, 4);
					// blit_buffer, Kernels.cu:181-203
					CK(launch_fold(c->shard, c->wf, bv, 1.0f / float(c->sample_index + spp), last ? 1 : 0,
								   (last && to_display) ? display_target(c) : DisplayTarget{}, st));
				}
				c->launches += 1;
				if (last && to_display)
					c->display.frames++;
			}
			c->sample_index += spp;
		}
		else
		{
			// EmbreeRT renders one un-accumulated sample per call (Context.cpp:104-300)
			StageTimer t(c, 0);
			CK(launch_emode(c->scene, c->shard, c->wf, rs_frame, c->d_materials_raw.ptr, c->d_tex_desc.as<uint32_t>(),
							uint32_t(c->textures.size()), c->dims, st));
			c->launches += 1;
			c->sample_index += 1;
		}
		CK(cudaEventRecord(c->ev_end, st));
		c->last_spp = spp, c->last_batches = batches;
		c->frame_in_flight = true;
		return RFWB200_OK;
	}

	// the display rank of a sharded frame (rank 0 of an in-process group, or the owner of an exported display image) presents the
	// assembled frame exactly like a single-device context
	static bool presents_display(const rfwb200_context *c) { return c && c->display.owner && c->display.image && c->mode_pt && c->shard.world > 1; }

	void *rfwb200_device_framebuffer(rfwb200_context *c)
	{
		if (presents_display(c))
			return c->display.image;
		return c ? c->d_fb.ptr : nullptr;
	}

	size_t rfwb200_local_pixel_count(const rfwb200_context *c)
	{
		if (!c)
			return 0;
		return (c->shard.world == 1 || presents_display(c)) ? size_t(c->width) * c->height : size_t(c->shard.local_pixels);
	}

	size_t rfwb200_shard_stride(const rfwb200_context *c) { return c ? shard_stride_pixels(c->shard) : 0; }

	int rfwb200_read_framebuffer(rfwb200_context *c, float *host_rgba, size_t capacity_pixels)
	{
		REQUIRE(c && host_rgba, "bad arguments");
		if (int r = ensure_device(c))
			return r;
		const size_t n = rfwb200_local_pixel_count(c);
		REQUIRE(capacity_pixels >= n, "host buffer too small");
		CK(cudaMemcpyAsync(host_rgba, rfwb200_device_framebuffer(c), n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
		return RFWB200_OK;
	}

	int rfwb200_read_framebuffer_async(rfwb200_context *c, float *pinned_host_rgba, size_t capacity_pixels)
	{
		REQUIRE(c && pinned_host_rgba, "bad arguments");
		if (int r = ensure_device(c))
			return r;
		const size_t n = rfwb200_local_pixel_count(c);
		REQUIRE(capacity_pixels >= n, "host buffer too small");
		if (!c->copy_stream)
		{
			CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
			CK(cudaEventCreateWithFlags(&c->ev_frame_done, cudaEventDisableTiming));
			CK(cudaEventCreateWithFlags(&c->ev_copy_done, cudaEventDisableTiming));
		}
		CK(cudaEventRecord(c->ev_frame_done, c->stream));
		CK(cudaStreamWaitEvent(c->copy_stream, c->ev_frame_done, 0));
		CK(cudaMemcpyAsync(pinned_host_rgba, rfwb200_device_framebuffer(c), n * sizeof(float4), cudaMemcpyDeviceToHost, c->copy_stream));
		CK(cudaEventRecord(c->ev_copy_done, c->copy_stream));
		c->copy_pending = true;
		return RFWB200_OK;
	}

	int rfwb200_read_wait(rfwb200_context *c)
	{
		REQUIRE(c != nullptr, "context is null");
		if (int r = ensure_device(c))
			return r;
		if (c->ev_copy_done)
			CK(cudaEventSynchronize(c->ev_copy_done));
		return RFWB200_OK;
	}

	int rfwb200_set_aov_transform(rfwb200_context *c, const float m[9])
	{
		REQUIRE(c && m, "bad arguments");
		FORWARD(rfwb200_set_aov_transform(p_, m));
		memcpy(c->to_eye, m, sizeof(c->to_eye));
		return RFWB200_OK;
	}

	int rfwb200_read_aov(rfwb200_context *c, int which, float *host_rgba, size_t capacity_pixels)
	{
		REQUIRE(c && host_rgba && (which == 0 || which == 1), "bad arguments (which: 0 = albedo, 1 = normal)");
		if (int r = ensure_device(c))
			return r;
		REQUIRE(c->aov && c->d_albedo.ptr && c->sample_index > 0, "no feature planes: set \"aov\" = on and render a Reset frame first");
		const size_t n = c->shard.world == 1 ? size_t(c->width) * c->height : size_t(c->shard.local_pixels);
		REQUIRE(capacity_pixels >= n, "host buffer too small");
		CK(c->d_aov_out.reserve(std::max<size_t>(n, c->shard.local_pixels) * sizeof(float4)));
		CK(launch_aov_finalize(c->shard, (which == 0 ? c->d_albedo : c->d_normal).as<float4>(), 1.0f / float(c->sample_index), c->d_aov_out.as<float4>(), c->stream));
		c->launches += 1;
		CK(cudaMemcpyAsync(host_rgba, c->d_aov_out.ptr, n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
		return RFWB200_OK;
	}

	int rfwb200_tone_map(rfwb200_context *c, float contrast, float brightness, const void *device_rgba32f, void *device_rgba8,
						 size_t pixels)
	{
		REQUIRE(c != nullptr, "context is null");
		if (int r = ensure_device(c))
			return r;
		const float4 *src = static_cast<const float4 *>(device_rgba32f);
		if (!src)
		{
			src = static_cast<const float4 *>(rfwb200_device_framebuffer(c));
			pixels = rfwb200_local_pixel_count(c);
		}
		REQUIRE(src != nullptr, "no framebuffer: call init first");
		REQUIRE(pixels <= 0xffffffffull, "too many pixels");
		uint32_t *dst = static_cast<uint32_t *>(device_rgba8);
		if (!dst)
		{
			CK(c->d_display.reserve(std::max<size_t>(pixels, 1) * sizeof(uint32_t)));
			dst = c->d_display.as<uint32_t>();
		}
		CK(launch_tone_map(src, dst, uint32_t(pixels), contrast, brightness, c->stream));
		c->launches += 1;
		return RFWB200_OK;
	}

	void *rfwb200_device_display(rfwb200_context *c) { return c ? c->d_display.ptr : nullptr; }

	int rfwb200_read_display(rfwb200_context *c, float contrast, float brightness, uint8_t *host_rgba8, size_t capacity_pixels)
	{
		REQUIRE(c && host_rgba8, "bad arguments");
		const size_t n = rfwb200_local_pixel_count(c);
		REQUIRE(capacity_pixels >= n, "host buffer too small");
		if (int r = rfwb200_tone_map(c, contrast, brightness, nullptr, nullptr, 0))
			return r;
		CK(cudaMemcpyAsync(host_rgba8, c->d_display.ptr, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
		return RFWB200_OK;
	}

	int rfwb200_assemble_shards(rfwb200_context *c, const void *gathered, void *image_out)
	{
		REQUIRE(c && gathered && image_out, "bad arguments");
		if (int r = ensure_device(c))
			return r;
		CK(launch_assemble(c->shard, static_cast<const float4 *>(gathered), shard_stride_pixels(c->shard),
						   static_cast<float4 *>(image_out), c->stream));
		c->launches += 1;
		return RFWB200_OK;
	}

	// ---- display target of a sharded frame ------------------------------------------------------------------------
	int rfwb200_display_create(rfwb200_context *c, void **image_out)
	{
		REQUIRE(c != nullptr, "context is null");
		if (int r = ensure_device(c))
			return r;
		if (!c->initialised)
			return set_error(RFWB200_ERR_STATE, "rfwb200_init has not been called");
		display_detach(c);
		void *base = nullptr;
		const size_t image_bytes = (size_t(c->width) * c->height * sizeof(float4) + 255) & ~size_t(255);
		CK(cudaMalloc(&base, image_bytes + DISPLAY_TAIL_BYTES));
		CK(cudaMemset(base, 0, image_bytes + DISPLAY_TAIL_BYTES));
		if (int r = display_bind(c, base, true, false))
			return r;
		if (image_out)
			*image_out = c->display.image;
		return RFWB200_OK;
	}

	int rfwb200_display_export(rfwb200_context *c, unsigned char handle_out[64])
	{
		REQUIRE(c && handle_out, "bad arguments");
		REQUIRE(c->display.owner, "this context does not own a display image (rfwb200_display_create)");
		if (int r = ensure_device(c))
			return r;
		static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
		cudaIpcMemHandle_t h;
		CK(cudaIpcGetMemHandle(&h, c->display.base));
		memcpy(handle_out, &h, 64);
		return RFWB200_OK;
	}

	int rfwb200_display_import(rfwb200_context *c, const unsigned char handle[64])
	{
		REQUIRE(c && handle, "bad arguments");
		if (int r = ensure_device(c))
			return r;
		if (!c->initialised)
			return set_error(RFWB200_ERR_STATE, "rfwb200_init has not been called");
		display_detach(c);
		cudaIpcMemHandle_t h;
		memcpy(&h, handle, 64);
		void *base = nullptr;
		CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
		return display_bind(c, base, false, true);
	}

	int rfwb200_display_attach(rfwb200_context *c, rfwb200_context *display_rank)
	{
		REQUIRE(c && display_rank && c != display_rank, "bad arguments");
		REQUIRE(display_rank->display.owner, "the display rank owns no display image (rfwb200_display_create)");
		REQUIRE(c->initialised && c->width == display_rank->width && c->height == display_rank->height,
				"both contexts must be initialised with the same frame size");
		if (int r = ensure_device(c))
			return r;
		display_detach(c);
		if (c->device != display_rank->device)
		{
			int can = 0;
			CK(cudaDeviceCanAccessPeer(&can, c->device, display_rank->device));
			REQUIRE(can, "no peer access between the two devices");
			const cudaError_t e = cudaDeviceEnablePeerAccess(display_rank->device, 0);
			if (e == cudaErrorPeerAccessAlreadyEnabled)
				cudaGetLastError();
			else
				CK(e);
		}
		return display_bind(c, display_rank->display.base, false, false);
	}

	int rfwb200_display_wait(rfwb200_context *c)
	{
		REQUIRE(c != nullptr, "context is null");
		REQUIRE(c->display.owner, "only the display rank waits for a frame");
		if (int r = ensure_device(c))
			return r;
		CK(launch_display_spin(c->display.arrivals, c->display.frames * c->shard.world, c->d_display_local.as<uint32_t>() + 1, c->stream));
		c->launches += 1;
		return RFWB200_OK;
	}

	void *rfwb200_display_image(rfwb200_context *c) { return c ? c->display.image : nullptr; }
	// test hook: blocking copy from a device pointer valid on this context's device, ordered after its stream
	int rfwb200_debug_read_device(rfwb200_context *c, const void *device_ptr, void *host, size_t bytes)
	{
		REQUIRE(c && device_ptr && host, "bad arguments");
		if (int r = ensure_device(c))
			return r;
		CK(cudaMemcpyAsync(host, device_ptr, bytes, cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
		return RFWB200_OK;
	}

	// ---- in-process device group -----------------------------------------------------------------------------------
	int rfwb200_create_group(const int *devices, size_t count, rfwb200_context **out)
	{
		REQUIRE(devices && out && count >= 1 && count <= 64, "bad device list");
		*out = nullptr;
		rfwb200_context *master = nullptr;
		if (int r = rfwb200_create(devices[0], &master))
			return r;
		for (size_t i = 1; i < count; i++)
		{
			rfwb200_context *p = nullptr;
			int r = rfwb200_create(devices[i], &p);
			if (!r)
			{
				cudaSetDevice(p->device);
				const cudaError_t e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
				if (e != cudaSuccess)
					r = set_error(RFWB200_ERR_CUDA, cudaGetErrorString(e));
			}
			if (r)
			{
				const std::string msg = g_last_error;
				if (p)
					rfwb200_destroy(p);
				rfwb200_destroy(master);
				return set_error(r, msg);
			}
			p->owns_stream = true;
			p->build_donor = master;
			p->shard.rank = uint32_t(i), p->shard.world = uint32_t(count);
			master->peers.push_back(p);
			master->workers.emplace_back(new DeviceWorker());
		}
		master->shard.rank = 0, master->shard.world = uint32_t(count);
		*out = master;
		return RFWB200_OK;
	}

	int rfwb200_synchronize(rfwb200_context *c)
	{
		REQUIRE(c != nullptr, "context is null");
		FORWARD(rfwb200_synchronize(p_));
		if (int r = ensure_device(c))
			return r;
		CK(cudaStreamSynchronize(c->stream));
		if (c->d_display_local.ptr && c->display.image)
		{
			uint32_t words[2] = {0, 0};
			CK(cudaMemcpy(words, c->d_display_local.ptr, sizeof(words), cudaMemcpyDeviceToHost));
			if (words[1])
			{
				CK(cudaMemset(c->d_display_local.as<uint32_t>() + 1, 0, 4));
				return set_error(RFWB200_ERR_STATE, "display flow control timed out: a rank of the sharded frame did not arrive / release within 4 s");
			}
		}
		return RFWB200_OK;
	}

	int rfwb200_set_probe_index(rfwb200_context *c, uint32_t x, uint32_t y)
	{
		REQUIRE(c != nullptr, "context is null");
		FORWARD(rfwb200_set_probe_index(p_, x, y));
		c->probe_x = x, c->probe_y = y;
		return RFWB200_OK;
	}

	int rfwb200_get_probe_results(rfwb200_context *c, uint32_t *inst, uint32_t *prim, float *dist)
	{
		REQUIRE(c && inst && prim && dist, "bad arguments");
		if (int r = ensure_device(c))
			return r;
		ProbeResult p{0, 0, 0.f, 0};
		if (c->d_probe.ptr)
		{
			CK(cudaMemcpyAsync(&p, c->d_probe.ptr, sizeof(p), cudaMemcpyDeviceToHost, c->stream));
			CK(cudaStreamSynchronize(c->stream));
		}
		*inst = uint32_t(p.inst), *prim = uint32_t(p.prim), *dist = p.dist;
		// in a group the probed pixel belongs to one rank's shard; the others report the cleared record
		for (rfwb200_context *q : c->peers)
		{
			uint32_t qi = 0, qp = 0;
			float qd = 0.f;
			if (int r = rfwb200_get_probe_results(q, &qi, &qp, &qd))
				return r;
			if (qd > 0.f && *dist == 0.f)
				*inst = qi, *prim = qp, *dist = qd;
		}
		return ensure_device(c);
	}

	static int fetch_counters(Ctx *c)
	{
		const size_t n = size_t(c->last_batches) * MAX_DEPTH_SLOTS;
		c->host_counters.assign(n, DepthCounters{});
		if (n == 0)
			return RFWB200_OK;
		CK(cudaMemcpyAsync(c->host_counters.data(), c->d_counters.ptr, n * sizeof(DepthCounters), cudaMemcpyDeviceToHost,
						   c->stream));
		CK(cudaStreamSynchronize(c->stream));
		return RFWB200_OK;
	}

	static int frame_counters_one(rfwb200_context *c, rfwb200_frame_counters *out);

	int rfwb200_get_frame_counters(rfwb200_context *c, rfwb200_frame_counters *out)
	{
		REQUIRE(c && out, "bad arguments");
		if (int r = frame_counters_one(c, out))
			return r;
		for (rfwb200_context *q : c->peers) // a group reports the whole frame
		{
			rfwb200_frame_counters o;
			if (int r = frame_counters_one(q, &o))
				return r;
			out->n_gen += o.n_gen, out->n_ext += o.n_ext, out->n_shade += o.n_shade, out->n_ext_out += o.n_ext_out;
			out->n_nee += o.n_nee, out->n_acc += o.n_acc, out->pixels += o.pixels;
		}
		return ensure_device(c);
	}

	static int frame_counters_one(rfwb200_context *c, rfwb200_frame_counters *out)
	{
		if (int r = ensure_device(c))
			return r;
		memset(out, 0, sizeof(*out));
		if (int r = fetch_counters(c))
			return r;
		// live pixels of this shard
		uint64_t live = 0;
		{
			const ShardView &s = c->shard;
			for (uint32_t lt = 0; lt < s.local_tiles; lt++)
			{
				const uint32_t gt = lt * s.world + s.rank;
				const uint32_t ty = gt / s.tiles_x, tx = gt % s.tiles_x;
				const uint32_t w = std::min(s.tile_w, s.width - std::min(s.width, tx * s.tile_w));
				const uint32_t h = std::min(s.tile_h, s.height - std::min(s.height, ty * s.tile_h));
				live += uint64_t(w) * h;
			}
		}
		out->pixels = live;
		out->samples = c->last_spp;
		if (!c->mode_pt)
		{
			out->n_gen = out->n_ext = live;
			return RFWB200_OK;
		}
		const uint32_t maxd = uint32_t(c->rs.max_path_length);
		out->n_gen = out->n_ext = out->n_shade = live * c->last_spp;
		for (uint32_t s = 0; s < c->last_batches; s++)
		{
			const DepthCounters *dc = &c->host_counters[size_t(s) * MAX_DEPTH_SLOTS];
			for (uint32_t d = 0; d <= maxd; d++)
			{
				out->n_ext_out += dc[d].ext;
				out->n_acc += dc[d].acc;
				if (d >= 1)
				{
					out->n_ext += dc[d - 1].ext;
					out->n_shade += dc[d - 1].ext;
					out->n_nee += dc[d].shadow_traced;
				}
			}
		}
		return RFWB200_OK;
	}

	int rfwb200_get_stats(rfwb200_context *c, rfwb200_render_stats *out)
	{
		REQUIRE(c && out, "bad arguments");
		if (int r = ensure_device(c))
			return r;
		memset(out, 0, sizeof(*out));
		if (!c->frame_in_flight)
			return RFWB200_OK;
		CK(cudaEventSynchronize(c->ev_end));
		float ms = 0;
		CK(cudaEventElapsedTime(&ms, c->ev_begin, c->ev_end));
		out->render_time = ms;
		// per-stage device times of the last frame (reference fields, context.h:50-72). Connect rays are
		// traced inside the same launch as the extension rays of their bounce, so shadow_time stays 0 and
		// their cost is part of secondary_time / deep_time.
		for (size_t i = 0; i < c->stage_events_used; i++)
		{
			const Ctx::StageEvent &e = c->stage_events[i];
			float t = 0;
			if (cudaEventElapsedTime(&t, e.a, e.b) != cudaSuccess)
				continue;
			switch (e.category)
			{
			case 0: out->primary_time += t; break;
			case 1: out->secondary_time += t; break;
			case 2: out->deep_time += t; break;
			case 3: out->shade_time += t; break;
			case 5: out->animation_time += t; break; // no field of the reference's RenderStats fits the re-ordering pass; its
													 // device time is reported in the one field this backend otherwise leaves 0
			default: out->finalize_time += t; break;
			}
		}
		rfwb200_frame_counters fc;
		if (int r = rfwb200_get_frame_counters(c, &fc))
			return r;
		out->primary_count = uint32_t(fc.n_gen);
		if (c->mode_pt && c->last_spp)
		{
			uint64_t sec = 0, deep = 0, shadow = 0;
			std::vector<const rfwb200_context *> ranks(1, c);
			ranks.insert(ranks.end(), c->peers.begin(), c->peers.end());
			for (const rfwb200_context *q : ranks) // host_counters of every rank were fetched by get_frame_counters above
				for (uint32_t s = 0; s < q->last_batches && size_t(s + 1) * MAX_DEPTH_SLOTS <= q->host_counters.size(); s++)
				{
					const DepthCounters *dc = &q->host_counters[size_t(s) * MAX_DEPTH_SLOTS];
					for (uint32_t d = 1; d <= uint32_t(c->rs.max_path_length); d++)
					{
						(d == 1 ? sec : deep) += dc[d - 1].ext;
						shadow += dc[d].shadow_traced;
					}
				}
			out->secondary_count = uint32_t(sec), out->deep_count = uint32_t(deep), out->shadow_count = uint32_t(shadow);
		}
		return RFWB200_OK;
	}

	uint64_t rfwb200_launch_count(const rfwb200_context *c)
	{
		if (!c)
			return 0;
		uint64_t n = c->launches;
		for (const rfwb200_context *q : c->peers)
			n += q->launches;
		return n;
	}

	int rfwb200_get_bvh_info(const rfwb200_context *c, uint64_t *nodes, uint64_t *triangles, float *sah_cost, float *build_ms)
	{
		REQUIRE(c != nullptr, "context is null");
		if (nodes)
			*nodes = c->device_built ? c->scene.node_count : (c->bvh.wide8 ? c->bvh.cw_nodes.size() : c->bvh.nodes.size());
		if (triangles)
			*triangles = c->device_built ? c->scene.tri_count : c->bvh.tri_order.size();
		if (sah_cost)
			*sah_cost = c->bvh.sah_cost;
		if (build_ms)
			*build_ms = float(c->bvh.build_ms);
		return RFWB200_OK;
	}

	// ---- stage-level entry points -----------------------------------------------------------------

	int rfwb200_trace_closest(rfwb200_context *c, const float *origins, const float *directions, size_t n, float t_min,
							  rfwb200_hit *hits_out)
	{
		REQUIRE(c && origins && directions && hits_out, "bad arguments");
		REQUIRE(n < (1ull << 31), "too many rays");
		if (int r = ensure_device(c))
			return r;
		if (c->geometry_dirty || !c->scene.nodes)
			return set_error(RFWB200_ERR_STATE, "call rfwb200_update first");
		if (int r = ensure_dims(c))
			return r;
		if (n == 0)
			return RFWB200_OK;
		CK(c->d_scratch_cursor.reserve(256));
		DevBuf dO, dD, dH, dI;
		CK(dO.reserve(n * 16));
		CK(dD.reserve(n * 16));
		CK(dH.reserve(n * 16));
		if (c->two_level)
			CK(dI.reserve(n * 4));
		cudaStream_t st = c->stream;
		CK(cudaMemcpyAsync(dO.ptr, origins, n * 16, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(dD.ptr, directions, n * 16, cudaMemcpyHostToDevice, st));
		CK(cudaMemsetAsync(c->d_scratch_cursor.ptr, 0, 4, st));
		CK(launch_trace_closest(c->scene, c->rs, dO.as<float4>(), dD.as<float4>(), uint32_t(n), t_min, dH.as<float4>(),
								c->d_scratch_cursor.as<uint32_t>(), c->dims, st, c->two_level ? dI.as<uint32_t>() : nullptr));
		c->launches += 1;
		std::vector<float> raw(n * 4);
		std::vector<uint32_t> hit_inst(c->two_level ? n : 0);
		CK(cudaMemcpyAsync(raw.data(), dH.ptr, n * 16, cudaMemcpyDeviceToHost, st));
		if (c->two_level)
			CK(cudaMemcpyAsync(hit_inst.data(), dI.ptr, n * 4, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		// translate shading-record indices back to (instance, primitive)
		std::vector<ShadeTri> shade(c->flat_tri_count);
		if (c->flat_tri_count)
			CK(cudaMemcpy(shade.data(), c->d_shade_tris.ptr, shade.size() * sizeof(ShadeTri), cudaMemcpyDeviceToHost));
		for (size_t i = 0; i < n; i++)
		{
			uint32_t tri;
			memcpy(&tri, &raw[4 * i + 3], 4);
			hits_out[i].t = raw[4 * i], hits_out[i].u = raw[4 * i + 1], hits_out[i].v = raw[4 * i + 2];
			if (tri == 0xffffffffu || tri >= shade.size())
				hits_out[i].inst_id = -1, hits_out[i].prim_id = -1, hits_out[i].u = 0, hits_out[i].v = 0;
			else
			{
				hits_out[i].prim_id = int32_t(shade[tri].prim_id);
				if (!c->two_level)
					hits_out[i].inst_id = int32_t(shade[tri].inst_id);
				else // (top-level instance, member rank of the triangle's mesh) -> the caller's instance
					hits_out[i].inst_id = hit_inst[i] < c->tl_table.size() ? int32_t(c->tl_inst_map[c->tl_table[hit_inst[i]].pad[0] + shade[tri].inst_id]) : -1;
			}
		}
		return RFWB200_OK;
	}

	int rfwb200_trace_occluded(rfwb200_context *c, const float *origins, const float *directions, const float *t_max,
							   size_t n, float t_min, uint8_t *occluded_out)
	{
		REQUIRE(c && origins && directions && t_max && occluded_out, "bad arguments");
		REQUIRE(n < (1ull << 31), "too many rays");
		if (int r = ensure_device(c))
			return r;
		if (c->geometry_dirty || !c->scene.nodes)
			return set_error(RFWB200_ERR_STATE, "call rfwb200_update first");
		if (int r = ensure_dims(c))
			return r;
		if (n == 0)
			return RFWB200_OK;
		std::vector<float> dt(n * 4);
		for (size_t i = 0; i < n; i++)
			dt[4 * i] = directions[4 * i], dt[4 * i + 1] = directions[4 * i + 1], dt[4 * i + 2] = directions[4 * i + 2],
					dt[4 * i + 3] = t_max[i];
		CK(c->d_scratch_cursor.reserve(256));
		DevBuf dO, dD, dR;
		CK(dO.reserve(n * 16));
		CK(dD.reserve(n * 16));
		CK(dR.reserve(n));
		cudaStream_t st = c->stream;
		CK(cudaMemcpyAsync(dO.ptr, origins, n * 16, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(dD.ptr, dt.data(), n * 16, cudaMemcpyHostToDevice, st));
		CK(cudaMemsetAsync(c->d_scratch_cursor.ptr, 0, 4, st));
		CK(launch_trace_occluded(c->scene, c->rs, dO.as<float4>(), dD.as<float4>(), uint32_t(n), t_min, dR.as<uint8_t>(),
								 c->d_scratch_cursor.as<uint32_t>(), c->dims, st));
		c->launches += 1;
		CK(cudaMemcpyAsync(occluded_out, dR.ptr, n, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		return RFWB200_OK;
	}

	int rfwb200_generate_primary(rfwb200_context *c, const rfwb200_camera_view *view, uint32_t sample_index,
								 float *origins_out, float *directions_out, size_t capacity_rays)
	{
		REQUIRE(c && view && origins_out && directions_out, "bad arguments");
		if (int r = ensure_device(c))
			return r;
		if (!c->initialised)
			return set_error(RFWB200_ERR_STATE, "rfwb200_init has not been called");
		REQUIRE(c->shard.world == 1, "generate_primary is a single-shard debug entry point");
		const size_t P = size_t(c->width) * c->height;
		REQUIRE(capacity_rays >= P, "host buffers too small");
		if (int r = upload_frame_params(c, view, sample_index))
			return r;
		DevBuf dO, dD;
		CK(dO.reserve(P * 16));
		CK(dD.reserve(P * 16));
		cudaStream_t st = c->stream;
		CK(cudaMemsetAsync(dO.ptr, 0, P * 16, st));
		CK(cudaMemsetAsync(dD.ptr, 0, P * 16, st));
		CK(launch_generate_only(c->scene, c->shard, c->wf, sample_index, c->mode_pt ? 0 : 1, dO.as<float4>(), dD.as<float4>(), st));
		c->launches += 1;
		CK(cudaMemcpyAsync(origins_out, dO.ptr, P * 16, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(directions_out, dD.ptr, P * 16, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		return RFWB200_OK;
	}

	// Debug/test hook: copy one wavefront plane of the last frame to the host.  which: 0/1 = O[0]/O[1], 2/3 = D[0]/D[1],
	// 4/5 = T[0]/T[1], 6 = hit, 7/8/9 = connect O/D/E, 10 = accumulator; n float4 elements.
	int rfwb200_debug_read_plane(rfwb200_context *c, int which, float *host, size_t n)
	{
		REQUIRE(c && host, "bad arguments");
		if (int r = ensure_device(c))
			return r;
		const DevBuf *planes[11] = {&c->d_O[0], &c->d_O[1], &c->d_D[0], &c->d_D[1], &c->d_T[0], &c->d_T[1],
									&c->d_hit,	&c->d_sO,	&c->d_sD,	&c->d_sE,	&c->d_sample_acc};
		REQUIRE(which >= 0 && which < 11, "bad plane");
		REQUIRE(n * 16 <= planes[which]->bytes, "plane smaller than requested");
		CK(cudaMemcpyAsync(host, planes[which]->ptr, n * 16, cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
		return RFWB200_OK;
	}
	// Debug hook: per-warp {start ns, end ns, rays, smid} of the trace launch at `depth` (sample 0 of the next frames).
	int rfwb200_debug_trace_timeline(rfwb200_context *c, int depth, unsigned long long *host, size_t max_warps, size_t *n_warps)
	{
		REQUIRE(c != nullptr, "context is null");
		if (int r = ensure_device(c))
			return r;
		if (int r = ensure_dims(c))
			return r;
		const size_t warps = size_t(c->dims.trace_grid) * (c->dims.trace_block / 32);
		if (n_warps)
			*n_warps = warps;
		if (!host)
		{
			CK(c->d_debug.reserve(warps * 32));
			CK(cudaMemset(c->d_debug.ptr, 0, warps * 32));
			c->rs.debug = static_cast<unsigned long long *>(c->d_debug.ptr), c->rs.debug_depth = depth;
			return RFWB200_OK;
		}
		REQUIRE(max_warps >= warps && c->d_debug.ptr, "buffer too small or timeline not armed");
		CK(cudaStreamSynchronize(c->stream));
		CK(cudaMemcpy(host, c->d_debug.ptr, warps * 32, cudaMemcpyDeviceToHost));
		c->rs.debug = nullptr, c->rs.debug_depth = -1;
		return RFWB200_OK;
	}
	int rfwb200_debug_read_counters(rfwb200_context *c, uint32_t *out8_per_depth, size_t depth_slots)
	{
		REQUIRE(c && out8_per_depth, "bad arguments");
		if (int r = ensure_device(c))
			return r;
		REQUIRE(depth_slots <= size_t(c->counters_capacity) * MAX_DEPTH_SLOTS, "more slots than allocated");
		CK(cudaMemcpyAsync(out8_per_depth, c->d_counters.ptr, depth_slots * sizeof(DepthCounters), cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
		return RFWB200_OK;
	}

	// Host-only self check of the BVH builder (no GPU involved): builds the 4-wide BVH over `n_tris` world-space
	// triangles (float[9] each) and walks it on the CPU for `n_rays` rays with the kernels' node semantics.
	// CPU walk of a flattened 4-wide BVH with the node semantics of the kernels (shared by the self checks below)
	static int walk_bvh4_host(const BvhBuildResult &bvh, const std::vector<BuildTriangle> &bt, const float *origins3, const float *dirs3,
							  size_t n_rays, float *t_out, int32_t *tri_out, uint32_t *visits_out)
	{
		REQUIRE(3 * bvh.depth + 2 <= TRAVERSAL_STACK, "BVH deeper than the traversal stack allows");
		for (size_t r = 0; r < n_rays; r++)
		{
			const float *o = origins3 + 3 * r, *d = dirs3 + 3 * r;
			float id[3], ood[3];
			for (int a = 0; a < 3; a++)
			{
				const float dd = std::fabs(d[a]) > 1e-30f ? d[a] : std::copysign(1e-30f, d[a]);
				id[a] = 1.0f / dd, ood[a] = o[a] * id[a];
			}
			float tmax = 1e34f;
			int32_t best = -1;
			int stack[TRAVERSAL_STACK], sp = 0, cur = 0;
			uint32_t node_visits = 0, tri_tests = 0;
			for (;;)
			{
				if (cur >= 0)
				{
					node_visits++;
					const BvhNode4 &n = bvh.nodes[cur];
					int hit[4], nh = 0;
					float key[4];
					for (int k = 0; k < 4; k++)
					{
						const float lo[3] = {n.minx[k], n.miny[k], n.minz[k]}, hi[3] = {n.maxx[k], n.maxy[k], n.maxz[k]};
						float tn = -3e38f, tf = 3e38f;
						bool nan = false;
						for (int a = 0; a < 3; a++)
						{
							const float t1 = lo[a] * id[a] - ood[a], t2 = hi[a] * id[a] - ood[a];
							nan |= std::isnan(t1) || std::isnan(t2);
							tn = std::max(tn, std::min(t1, t2)), tf = std::min(tf, std::max(t1, t2));
						}
						if (!nan && tf >= tn && tn < tmax && tf >= 1e-5f)
							hit[nh] = k, key[nh] = tn, nh++;
					}
					for (int i = 0; i < nh; i++) // far to near onto the stack
						for (int j = i + 1; j < nh; j++)
							if (key[j] > key[i])
								std::swap(key[i], key[j]), std::swap(hit[i], hit[j]);
					if (nh == 0)
					{
						if (sp == 0)
							break;
						cur = stack[--sp];
						continue;
					}
					for (int i = 0; i + 1 < nh; i++)
						stack[sp++] = n.child[hit[i]];
					cur = n.child[hit[nh - 1]];
				}
				else
				{
					const uint32_t v = uint32_t(~cur), first = v >> 2, cnt = (v & 3u) + 1u;
					for (uint32_t i = first; i < first + cnt && i < bvh.tri_order.size(); i++)
					{
						tri_tests++;
						const BuildTriangle &t = bt[bvh.tri_order[i]];
						const float e1[3] = {t.v1[0] - t.v0[0], t.v1[1] - t.v0[1], t.v1[2] - t.v0[2]};
						const float e2[3] = {t.v2[0] - t.v0[0], t.v2[1] - t.v0[1], t.v2[2] - t.v0[2]};
						const float h[3] = {d[1] * e2[2] - e2[1] * d[2], d[2] * e2[0] - e2[2] * d[0], d[0] * e2[1] - e2[0] * d[1]};
						const float det = e1[0] * h[0] + e1[1] * h[1] + e1[2] * h[2];
						if (det > -1e-12f && det < 1e-12f)
							continue;
						const float f = 1.0f / det;
						const float s[3] = {o[0] - t.v0[0], o[1] - t.v0[1], o[2] - t.v0[2]};
						const float u = f * (s[0] * h[0] + s[1] * h[1] + s[2] * h[2]);
						if (u < 0.0f || u > 1.0f)
							continue;
						const float q[3] = {s[1] * e1[2] - e1[1] * s[2], s[2] * e1[0] - e1[2] * s[0], s[0] * e1[1] - e1[0] * s[1]};
						const float vv = f * (d[0] * q[0] + d[1] * q[1] + d[2] * q[2]);
						if (vv < 0.0f || u + vv > 1.0f)
							continue;
						const float tt = f * (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]);
						if (tt > 1e-5f && tt < tmax)
							tmax = tt, best = int32_t(bvh.tri_order[i]);
					}
					if (sp == 0)
						break;
					cur = stack[--sp];
				}
			}
			t_out[r] = tmax, tri_out[r] = best;
			if (visits_out)
				visits_out[2 * r] = node_visits, visits_out[2 * r + 1] = tri_tests;
		}
		return RFWB200_OK;
	}

	int rfwb200_host_bvh_check(const float *tris9, size_t n_tris, int spatial_splits, const float *origins3,
							   const float *dirs3, size_t n_rays, float *t_out, int32_t *tri_out, uint64_t *nodes_out,
							   uint64_t *refs_out, int32_t *depth_out, float *sah_out, uint32_t *visits_out)
	{
		REQUIRE(tris9 && origins3 && dirs3 && t_out && tri_out, "bad arguments");
		std::vector<BuildTriangle> bt(n_tris);
		for (size_t i = 0; i < n_tris; i++)
			memcpy(&bt[i], tris9 + 9 * i, 9 * sizeof(float));
		BvhBuildResult bvh;
		build_bvh4(bt.data(), n_tris, int(std::max(1u, std::thread::hardware_concurrency())), bvh, spatial_splits != 0);
		if (nodes_out)
			*nodes_out = bvh.nodes.size();
		if (refs_out)
			*refs_out = bvh.tri_order.size();
		if (depth_out)
			*depth_out = bvh.depth;
		if (sah_out)
			*sah_out = bvh.sah_cost;
		return walk_bvh4_host(bvh, bt, origins3, dirs3, n_rays, t_out, tri_out, visits_out);
	}

	// Self check of the grouping rule of two-level scenes without a GPU: which meshes share a tree, how many top-level instances
	// the scene has, and which caller instance stands behind (top-level instance, member) — the table the kernels read.
	int rfwb200_host_group_check(const int32_t *mesh_of_instance, const float *transforms16, const float *normals9, size_t n_instances,
								 size_t n_meshes, int32_t *group_of_mesh_out, uint64_t *groups_out, uint64_t *top_level_instances_out,
								 uint32_t *inst_map_out, size_t inst_map_capacity, uint64_t *inst_map_size_out)
	{
		REQUIRE((mesh_of_instance && transforms16 && normals9) || n_instances == 0, "bad arguments");
		REQUIRE(group_of_mesh_out && groups_out && top_level_instances_out && inst_map_size_out, "bad arguments");
		std::vector<HostInstance> inst(n_instances);
		for (size_t i = 0; i < n_instances; i++)
		{
			inst[i].mesh = mesh_of_instance[i];
			memcpy(inst[i].transform, transforms16 + 16 * i, sizeof(float) * 16);
			memcpy(inst[i].normal, normals9 + 9 * i, sizeof(float) * 9);
		}
		std::vector<std::vector<Placement>> placed;
		std::vector<std::vector<uint32_t>> groups;
		group_by_placement(inst, std::vector<uint8_t>(n_meshes, 1), placed, groups);
		for (size_t m = 0; m < n_meshes; m++)
			group_of_mesh_out[m] = -1;
		uint64_t tl = 0, map_size = 0;
		for (size_t g = 0; g < groups.size(); g++)
		{
			for (uint32_t m : groups[g])
				group_of_mesh_out[m] = int32_t(g);
			const size_t k = placed[groups[g][0]].size();
			for (size_t j = 0; j < k; j++, tl++)
				for (uint32_t m : groups[g])
				{
					if (inst_map_out && map_size < inst_map_capacity)
						inst_map_out[map_size] = placed[m][j].instance;
					map_size++;
				}
		}
		*groups_out = groups.size(), *top_level_instances_out = tl, *inst_map_size_out = map_size;
		return RFWB200_OK;
	}

	// Self check of the top-level builder of two-level scenes without a GPU: every box is named by exactly one leaf slot, every
	// node's slot boxes contain what is below them, and for every ray the set of boxes reached through the tree equals the
	// set a loop over all boxes finds with the same slab test.
	int rfwb200_host_tlas_check(const float *boxes6, size_t n_boxes, const float *origins3, const float *dirs3, size_t n_rays,
								uint64_t *nodes_out, int32_t *depth_out, uint64_t *structure_errors_out, uint64_t *ray_mismatches_out,
								uint64_t *boxes_hit_out)
	{
		REQUIRE((boxes6 || n_boxes == 0) && (n_rays == 0 || (origins3 && dirs3)) && structure_errors_out && ray_mismatches_out, "bad arguments");
		std::vector<BvhNode4> nodes;
		const int depth = build_tlas4(boxes6, n_boxes, nodes);
		if (nodes_out)
			*nodes_out = nodes.size();
		if (depth_out)
			*depth_out = depth;
		uint64_t bad = 0;
		std::vector<uint32_t> named(n_boxes, 0);
		std::vector<uint8_t> visited(nodes.size(), 0);
		// bounds of a subtree, checked against the slot box that covers it
		std::function<void(uint32_t, float *)> bound = [&](uint32_t ni, float *out6) {
			for (int a = 0; a < 3; a++)
				out6[a] = 3.0e38f, out6[3 + a] = -3.0e38f;
			if (ni >= nodes.size() || visited[ni]++)
			{
				bad++;
				return;
			}
			const BvhNode4 &n = nodes[ni];
			if (n.pad[0] < (n_boxes ? 1 : 0) || n.pad[0] > 4)
				bad++;
			for (int s = 0; s < n.pad[0] && s < 4; s++)
			{
				float sub[6];
				if (n.child[s] >= 0)
					bound(uint32_t(n.child[s]), sub);
				else
				{
					const uint32_t item = uint32_t(~n.child[s]) >> 2;
					if (item >= n_boxes || (uint32_t(~n.child[s]) & 3u) != 0u)
					{
						bad++;
						continue;
					}
					named[item]++;
					memcpy(sub, boxes6 + 6 * item, sizeof(sub));
				}
				const float slot[6] = {n.minx[s], n.miny[s], n.minz[s], n.maxx[s], n.maxy[s], n.maxz[s]};
				for (int a = 0; a < 3; a++)
				{
					if (!(slot[a] <= sub[a]) || !(slot[3 + a] >= sub[3 + a]))
						bad++;
					out6[a] = std::min(out6[a], slot[a]), out6[3 + a] = std::max(out6[3 + a], slot[3 + a]);
				}
			}
		};
		float all[6];
		bound(0, all);
		for (size_t i = 0; i < n_boxes; i++)
			if (named[i] != 1)
				bad++;
		*structure_errors_out = bad;
		auto slab_hit = [](const float *lo, const float *hi, const float *o, const float *id) {
			float tn = 0.0f, tf = 3.0e38f;
			for (int a = 0; a < 3; a++)
			{
				const float t1 = (lo[a] - o[a]) * id[a], t2 = (hi[a] - o[a]) * id[a];
				tn = std::max(tn, std::min(t1, t2)), tf = std::min(tf, std::max(t1, t2));
			}
			return tn <= tf;
		};
		uint64_t mismatches = 0, total_hit = 0;
		std::vector<uint32_t> via_tree, direct, stack;
		for (size_t r = 0; r < n_rays; r++)
		{
			const float *o = origins3 + 3 * r, *d = dirs3 + 3 * r;
			float id[3];
			for (int a = 0; a < 3; a++)
				id[a] = 1.0f / (std::fabs(d[a]) > 1e-30f ? d[a] : std::copysign(1e-30f, d[a]));
			via_tree.clear(), direct.clear(), stack.assign(1, 0u);
			while (!stack.empty())
			{
				const BvhNode4 &n = nodes[stack.back()];
				stack.pop_back();
				for (int s = 0; s < n.pad[0] && s < 4; s++)
				{
					const float lo[3] = {n.minx[s], n.miny[s], n.minz[s]}, hi[3] = {n.maxx[s], n.maxy[s], n.maxz[s]};
					if (!slab_hit(lo, hi, o, id))
						continue;
					if (n.child[s] >= 0)
						stack.push_back(uint32_t(n.child[s]));
					else if (slab_hit(boxes6 + 6 * (uint32_t(~n.child[s]) >> 2), boxes6 + 6 * (uint32_t(~n.child[s]) >> 2) + 3, o, id))
						via_tree.push_back(uint32_t(~n.child[s]) >> 2);
				}
			}
			for (size_t i = 0; i < n_boxes; i++)
				if (slab_hit(boxes6 + 6 * i, boxes6 + 6 * i + 3, o, id))
					direct.push_back(uint32_t(i));
			std::sort(via_tree.begin(), via_tree.end());
			if (via_tree != direct)
				mismatches++;
			total_hit += direct.size();
		}
		*ray_mismatches_out = mismatches;
		if (boxes_hit_out)
			*boxes_hit_out = total_hit;
		return RFWB200_OK;
	}

	struct SerialAlloc1 // atomicAdd(counter, 1)
	{
		uint32_t *next;
		uint32_t operator()() const { return (*next)++; }
	};
	struct SerialAlloc // what atomicAdd on the level counter does, for one "thread" at a time
	{
		uint32_t *next;
		LB_HD uint32_t operator()(uint32_t k) const
		{
			const uint32_t r = *next;
			*next += k;
			return r;
		}
	};

	// Self check of the GPU builder's algorithm without a GPU: the per-element functions of lbvh.h — the very code the
	// kernels in geometry.cu run one element per thread — executed in loops on the host (std::sort in place of the radix
	// sort), then walked like any other 4-wide tree.
	int rfwb200_host_lbvh_check(const float *tris9, size_t n_tris, int presplit_and_flags, const float *origins3, const float *dirs3, size_t n_rays,
								float *t_out, int32_t *tri_out, uint64_t *nodes_out, uint64_t *refs_out, int32_t *depth_out, uint32_t *visits_out)
	{
		const int presplit = presplit_and_flags & 1;
		const bool ploc = (presplit_and_flags & 2) != 0; // builder=ploc: clustering instead of the radix tree
		REQUIRE(tris9 && origins3 && dirs3 && t_out && tri_out, "bad arguments");
		REQUIRE(n_tris >= 1 && n_tris < (1u << 29), "need 1 .. 2^29 triangles");
		std::vector<BuildTriangle> bt(n_tris);
		for (size_t i = 0; i < n_tris; i++)
			memcpy(&bt[i], tris9 + 9 * i, 9 * sizeof(float));
		auto pad_box = [](LbvhBox &b) {
			for (int a = 0; a < 3; a++)
			{
				const float m = std::max(std::fabs(b.lo[a]), std::fabs(b.hi[a])), pad = std::max(1e-5f, m * 2.4e-7f);
				b.lo[a] -= pad, b.hi[a] += pad;
			}
		};
		// k_lbvh_bounds
		std::vector<LbvhBox> full(n_tris);
		float slo[3] = {3e38f, 3e38f, 3e38f}, shi[3] = {-3e38f, -3e38f, -3e38f};
		for (size_t i = 0; i < n_tris; i++)
		{
			LbvhBox &b = full[i];
			for (int a = 0; a < 3; a++)
			{
				b.lo[a] = std::min(std::min(bt[i].v0[a], bt[i].v1[a]), bt[i].v2[a]);
				b.hi[a] = std::max(std::max(bt[i].v0[a], bt[i].v1[a]), bt[i].v2[a]);
				const float c = 0.5f * (b.lo[a] + b.hi[a]);
				slo[a] = std::min(slo[a], c), shi[a] = std::max(shi[a], c);
			}
			b.pad0 = b.pad1 = 0;
		}
		// k_lbvh_counts / k_lbvh_emit: one reference per piece
		float cell = presplit ? std::max(std::max(shi[0] - slo[0], shi[1] - slo[1]), shi[2] - slo[2]) * (1.0f / 64.0f) : 0.0f;
		const size_t ref_cap = n_tris + n_tris / 2 + 1024; // the device orchestrator's capacity rule (context.cpp, lbvh_build)
		for (int attempt = 0; presplit && attempt < 8 && cell > 0.0f; attempt++, cell *= 2.0f)
		{
			size_t total = 0;
			for (size_t i = 0; i < n_tris; i++)
				total += size_t(lb_piece_count(full[i], cell));
			if (total <= ref_cap)
				break;
			if (attempt == 7)
				cell = 0.0f; // gives up: one reference per triangle
		}
		std::vector<LbvhBox> boxes;
		std::vector<uint32_t> ref_tri;
		for (size_t i = 0; i < n_tris; i++)
		{
			const int cnt = presplit ? lb_piece_count(full[i], cell) : 1;
			for (int j = 0; j < cnt; j++)
			{
				LbvhBox b = full[i];
				if (cnt > 1)
				{
					int axis;
					float lo, hi;
					lb_piece_slab(full[i], cnt, j, axis, lo, hi);
					b = lb_clip_to_slab(bt[i].v0, bt[i].v1, bt[i].v2, full[i], axis, lo, hi);
				}
				pad_box(b);
				boxes.push_back(b), ref_tri.push_back(uint32_t(i));
			}
		}
		const int n = int(boxes.size());
		if (refs_out)
			*refs_out = uint64_t(n);
		std::vector<LbvhBox> leaf_box(n), inner_box(n);
		float inv[3];
		for (int a = 0; a < 3; a++)
			inv[a] = shi[a] > slo[a] ? 1.0f / (shi[a] - slo[a]) : 0.0f;
		std::vector<uint64_t> keys(n);
		for (int i = 0; i < n; i++)
			keys[i] = (uint64_t(lb_morton30(0.5f * (boxes[i].lo[0] + boxes[i].hi[0]), 0.5f * (boxes[i].lo[1] + boxes[i].hi[1]),
											 0.5f * (boxes[i].lo[2] + boxes[i].hi[2]), slo, inv))
					   << 32) |
					  uint64_t(i);
		std::sort(keys.begin(), keys.end());
		BvhBuildResult bvh;
		bvh.tri_order.resize(n);
		for (int i = 0; i < n; i++)
		{
			const uint32_t r = uint32_t(keys[i] & 0xffffffffull);
			bvh.tri_order[i] = ref_tri[r], leaf_box[i] = boxes[r];
		}
		std::vector<int32_t> left(n), right(n), first(n), last(n), parent_inner(n), parent_leaf(n);
		std::vector<uint32_t> arrivals(n, 0u);
		Lbvh2View t{};
		t.keys = keys.data(), t.n = n;
		t.left = left.data(), t.right = right.data(), t.first = first.data(), t.last = last.data();
		t.parent_inner = parent_inner.data(), t.parent_leaf = parent_leaf.data();
		t.inner_box = inner_box.data(), t.leaf_box = leaf_box.data(), t.arrivals = arrivals.data();
		std::vector<int32_t> counts(n, 0), position(n, 0);
		int ploc_rounds = 0;
		if (ploc && n > 1)
		{
			// k_ploc_nearest / k_ploc_merge / compaction, one "thread" after the other (geometry.cu ploc rounds)
			std::vector<int32_t> id_a(n), id_b(n), nn(n);
			std::vector<LbvhBox> box_a(leaf_box), box_b(n);
			std::vector<uint32_t> keep(n);
			for (int i = 0; i < n; i++)
				id_a[i] = ~i;
			uint32_t made = 0;
			int c = n;
			while (c > 1)
			{
				PlocRound pr{c, id_a.data(), box_a.data(), nn.data(), id_b.data(), box_b.data(), keep.data()};
				for (int i = 0; i < c; i++)
					ploc_nearest(pr, i, PLOC_RADIUS);
				for (int i = 0; i < c; i++)
					ploc_merge(pr, t, counts.data(), i, SerialAlloc1{&made});
				int c2 = 0;
				for (int i = 0; i < c; i++)
					if (keep[i])
						id_a[c2] = id_b[i], box_a[c2] = box_b[i], c2++;
				REQUIRE(c2 < c, "a clustering round merged nothing");
				c = c2, ploc_rounds++;
			}
			REQUIRE(made == uint32_t(n - 1), "clustering did not produce n - 1 inner nodes");
			for (int i = 0; i < n; i++)
				position[i] = ploc_leaf_position(t, counts.data(), i);
			for (int k = 0; k < n - 1; k++)
				ploc_node_range(t, counts.data(), position.data(), k);
			// permute the per-reference arrays to depth-first order and relabel the leaf children (k_ploc_permute)
			std::vector<LbvhBox> lb2(n);
			std::vector<uint32_t> to2(n);
			std::vector<int32_t> pl2(n);
			for (int i = 0; i < n; i++)
				lb2[position[i]] = leaf_box[i], to2[position[i]] = bvh.tri_order[i], pl2[position[i]] = parent_leaf[i];
			for (int k = 0; k < n - 1; k++)
			{
				if (left[k] < 0)
					left[k] = ~position[~left[k]];
				if (right[k] < 0)
					right[k] = ~position[~right[k]];
			}
			leaf_box.swap(lb2), bvh.tri_order.swap(to2), parent_leaf.swap(pl2);
			t.leaf_box = leaf_box.data(), t.parent_leaf = parent_leaf.data();
			std::vector<uint8_t> seen(n, 0);
			for (int i = 0; i < n; i++)
				seen[position[i]]++;
			for (int i = 0; i < n; i++)
				REQUIRE(seen[i] == 1, "depth-first positions are not a permutation");
			std::fill(arrivals.begin(), arrivals.begin() + (n - 1), 2u);
		}
		for (int i = 0; i < n - 1 && !ploc; i++)
			lb_build_inner(t, i);
		for (int i = 0; i < n && n > 1 && !ploc; i++) // k_lbvh_boxes, one "thread" after the other
		{
			int cur = parent_leaf[i];
			while (cur >= 0)
			{
				if (arrivals[cur]++ == 0u)
					break;
				inner_box[cur] = lb_union(lb_box(t, left[cur]), lb_box(t, right[cur]));
				cur = parent_inner[cur];
			}
		}
		bvh.nodes.resize(n);
		bvh.node_parent.resize(n);
		std::vector<LbvhPending> qa(n), qb(n);
		qa[0] = LbvhPending{n == 1 ? ~0 : 0, 0xffffffffu};
		uint32_t level_base = 0, count = 1;
		int levels = 0;
		while (count > 0)
		{
			uint32_t next = 0;
			for (uint32_t i = 0; i < count; i++)
				lb_collapse_node(t, qa[i], level_base + i, level_base + count, bvh.nodes.data(), bvh.node_parent.data(), qb.data(),
								 SerialAlloc{&next});
			level_base += count, count = next, levels++;
			std::swap(qa, qb);
			REQUIRE(levels <= 64, "radix tree deeper than 64 levels");
		}
		bvh.nodes.resize(level_base);
		bvh.depth = levels;
		if (nodes_out)
			*nodes_out = level_base;
		if (depth_out)
			*depth_out = levels;
		// every inner node must have been completed by exactly two arrivals
		for (int i = 0; i < n - 1; i++)
			REQUIRE(arrivals[i] == 2u, "bottom-up pass did not reach every inner node");
		return walk_bvh4_host(bvh, bt, origins3, dirs3, n_rays, t_out, tri_out, visits_out);
	}

	// Same self check for the compressed 8-wide layout (cwbvh.h): the walk below is the kernels' traversal loop
	// (cw_intersect_children is the very function k_wavefront_trace uses), so the builder, the quantisation and the
	// group / octant bit logic are validated without a GPU.  `refit_jitter` > 0 additionally moves every vertex by a
	// seeded pseudo-random offset of that size and refits on the host before tracing (the moved triangles are the ones
	// traced), exercising refit_cwbvh.
	int rfwb200_host_cwbvh_check(const float *tris9, size_t n_tris, int spatial_splits, float refit_jitter, const float *origins3,
								 const float *dirs3, size_t n_rays, float *t_out, int32_t *tri_out, uint64_t *nodes_out,
								 uint64_t *refs_out, int32_t *depth_out, float *sah_out, uint32_t *visits_out, float *tris9_out)
	{
		REQUIRE(tris9 && origins3 && dirs3 && t_out && tri_out, "bad arguments");
		std::vector<BuildTriangle> bt(n_tris);
		for (size_t i = 0; i < n_tris; i++)
			memcpy(&bt[i], tris9 + 9 * i, 9 * sizeof(float));
		BvhBuildResult bvh;
		build_cwbvh(bt.data(), n_tris, int(std::max(1u, std::thread::hardware_concurrency())), bvh, spatial_splits != 0);
		if (refit_jitter > 0.0f)
		{
			uint32_t seed = 12345u;
			for (size_t i = 0; i < n_tris; i++)
				for (int k = 0; k < 9; k++)
				{
					seed = seed * 1664525u + 1013904223u;
					(&bt[i].v0[0])[k] += refit_jitter * (float(seed >> 8) * (1.0f / 8388608.0f) - 1.0f);
				}
			refit_cwbvh(bt.data(), n_tris, bvh);
		}
		if (tris9_out)
			for (size_t i = 0; i < n_tris; i++)
				memcpy(tris9_out + 9 * i, &bt[i], 9 * sizeof(float));
		if (nodes_out)
			*nodes_out = bvh.cw_nodes.size();
		if (refs_out)
			*refs_out = bvh.tri_order.size();
		if (depth_out)
			*depth_out = bvh.depth;
		if (sah_out)
			*sah_out = bvh.sah_cost;
		REQUIRE(bvh.depth <= CW_MAX_DEPTH, "BVH deeper than the traversal stack allows");
		for (size_t r = 0; r < n_rays; r++)
		{
			const float *o = origins3 + 3 * r, *d = dirs3 + 3 * r;
			CwRay ray;
			ray.ox = o[0], ray.oy = o[1], ray.oz = o[2];
			float id[3];
			for (int a = 0; a < 3; a++)
			{
				const float dd = std::fabs(d[a]) > 1e-30f ? d[a] : std::copysign(1e-30f, d[a]);
				id[a] = 1.0f / dd;
			}
			ray.idx = id[0], ray.idy = id[1], ray.idz = id[2];
			ray.octinv = cw_octinv(id[0], id[1], id[2]);
			const float tmin = 1e-5f;
			float tmax = 1e34f;
			int32_t best = -1;
			uint32_t stack[2 * CW_STACK];
			int sp = 0;
			uint32_t ng_x = 0, ng_y = 0x80000000u, tg_x = 0, tg_y = 0; // node group (base, hits|imask), triangle group (base, bits)
			uint32_t node_visits = 0, tri_tests = 0;
			for (;;)
			{
				if (ng_y > 0x00ffffffu)
				{
					const uint32_t hits = ng_y;
					const uint32_t bit = 31u - uint32_t(__builtin_clz(hits));
					ng_y &= ~(1u << bit);
					if (ng_y > 0x00ffffffu)
					{
						REQUIRE(sp < CW_STACK, "traversal stack overflow");
						stack[2 * sp] = ng_x, stack[2 * sp + 1] = ng_y, sp++;
					}
					const uint32_t slot = (bit - 24u) ^ ray.octinv;
					const uint32_t rel = uint32_t(__builtin_popcount(hits & ~(0xffffffffu << slot) & 0xffu));
					const uint32_t ni = ng_x + rel;
					REQUIRE(ni < bvh.cw_nodes.size(), "child index out of range");
					const uint32_t *w = reinterpret_cast<const uint32_t *>(&bvh.cw_nodes[ni]);
					node_visits++;
					const uint32_t hm = cw_intersect_children(w, ray, tmin, tmax);
					ng_x = w[4], tg_x = w[5];
					ng_y = (hm & 0xff000000u) | (w[3] >> 24), tg_y = hm & 0x00ffffffu;
				}
				else
					tg_x = ng_x, tg_y = ng_y, ng_x = ng_y = 0;
				while (tg_y)
				{
					const uint32_t bit = 31u - uint32_t(__builtin_clz(tg_y));
					tg_y &= ~(1u << bit);
					const uint32_t i = tg_x + bit;
					REQUIRE(i < bvh.tri_order.size(), "triangle index out of range");
					tri_tests++;
					const BuildTriangle &t = bt[bvh.tri_order[i]];
					const float e1[3] = {t.v1[0] - t.v0[0], t.v1[1] - t.v0[1], t.v1[2] - t.v0[2]};
					const float e2[3] = {t.v2[0] - t.v0[0], t.v2[1] - t.v0[1], t.v2[2] - t.v0[2]};
					const float h[3] = {d[1] * e2[2] - e2[1] * d[2], d[2] * e2[0] - e2[2] * d[0], d[0] * e2[1] - e2[0] * d[1]};
					const float det = e1[0] * h[0] + e1[1] * h[1] + e1[2] * h[2];
					if (det > -1e-12f && det < 1e-12f)
						continue;
					const float f = 1.0f / det;
					const float s[3] = {o[0] - t.v0[0], o[1] - t.v0[1], o[2] - t.v0[2]};
					const float u = f * (s[0] * h[0] + s[1] * h[1] + s[2] * h[2]);
					if (u < 0.0f || u > 1.0f)
						continue;
					const float q[3] = {s[1] * e1[2] - e1[1] * s[2], s[2] * e1[0] - e1[2] * s[0], s[0] * e1[1] - e1[0] * s[1]};
					const float vv = f * (d[0] * q[0] + d[1] * q[1] + d[2] * q[2]);
					if (vv < 0.0f || u + vv > 1.0f)
						continue;
					const float tt = f * (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]);
					if (tt > tmin && tt < tmax)
						tmax = tt, best = int32_t(bvh.tri_order[i]);
				}
				if (ng_y <= 0x00ffffffu)
				{
					if (sp == 0)
						break;
					sp--, ng_x = stack[2 * sp], ng_y = stack[2 * sp + 1];
				}
			}
			t_out[r] = tmax, tri_out[r] = best;
			if (visits_out)
				visits_out[2 * r] = node_visits, visits_out[2 * r + 1] = tri_tests;
		}
		return RFWB200_OK;
	}

} // extern "C"
