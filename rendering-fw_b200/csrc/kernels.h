// kernels.h — host-callable launchers of the sm_100a wavefront kernels (kernels.cu)
#pragma once
#include "device_types.h"

#include <cuda_runtime.h>
#include <stddef.h>

namespace rfwb200
{

struct LaunchDims
{
	int trace_grid = 0, trace_block = 256;
	int shade_grid = 0, shade_block = 128;
	int move_grid = 0;
	size_t trace_smem = 0; // dynamic shared memory of the trace kernels (staged BVH prefix)
	size_t trace_smem16 = 0; // the same for the packed-node variant (80 B per staged node)
	int trace_grid_staged16 = 0;
};

// queries occupancy for the current device and smem_nodes setting, sets function attributes
cudaError_t shade_occupancy(int block, int *per_sm); // defined next to k_shade
cudaError_t configure_launches(const RenderSettings &rs, uint32_t node_count, LaunchDims &dims);

// --- PT-mode stages.  One wavefront (BatchView: up to MAX_BATCH_SPP samples of every local pixel) =
//     primary, shade(0), [sort, trace(d), shade(d)] for d = 1..max, fold ---
cudaError_t launch_primary(const SceneView &sc, const ShardView &sh, const WavefrontView &wf, const RenderSettings &rs,
						   const BatchView &bv, const LaunchDims &dims, cudaStream_t stream);
// in_buf: ray planes the paths are read from; out_buf: planes the extension rays are appended to
cudaError_t launch_shade(const SceneView &sc, const ShardView &sh, const WavefrontView &wf, const RenderSettings &rs,
						 const BatchView &bv, uint32_t depth, uint32_t in_buf, uint32_t out_buf, const ShardSync &sync, const LaunchDims &dims,
						 cudaStream_t stream);
// the same kernel built without -use_fast_math (setting "shade_math" = "ieee")
cudaError_t launch_shade_ieee(const SceneView &sc, const ShardView &sh, const WavefrontView &wf, const RenderSettings &rs,
							  const BatchView &bv, uint32_t depth, uint32_t in_buf, uint32_t out_buf, const ShardSync &sync, const LaunchDims &dims,
							  cudaStream_t stream);
cudaError_t launch_trace(const SceneView &sc, const ShardView &sh, const WavefrontView &wf, const RenderSettings &rs,
						 const BatchView &bv, uint32_t depth, uint32_t in_buf, const LaunchDims &dims, cudaStream_t stream);
// accumulator += samples of the wavefront; write_fb: framebuffer = accumulator * scale (Kernels.cu:181-203)
// dt.image != null: the finalised pixels also go, row-major, into the display rank's image, and the last CTA adds one arrival
cudaError_t launch_fold(const ShardView &sh, const WavefrontView &wf, const BatchView &bv, float scale, int write_fb, const DisplayTarget &dt,
						cudaStream_t stream);
// one-thread flow-control kernels of the display image: wait until *counter >= need (gives up after 4 s and sets *err),
// publish a counter value
cudaError_t launch_display_spin(const uint32_t *counter, uint32_t need, uint32_t *err, cudaStream_t stream);
cudaError_t launch_display_release(uint32_t *consumed, uint32_t value, cudaStream_t stream);
// sharded frame, between shade(depth) and trace(depth + 1): arrive; wait for `need_full` arrivals and merge the ranks' ext-seen
// flags if one of this rank's `spp` flags is clear, else only for `need_lagged`
cudaError_t launch_shard_sync(const ShardSync &sync, uint32_t depth, uint32_t need_full, uint32_t need_lagged, uint32_t spp,
							  uint32_t *local_seen_row, uint32_t *err, cudaStream_t stream);
// re-ordering of the extension rays shade(depth - 1) appended (planes [1] -> planes [0]): bin scan + move
cudaError_t launch_sort_setup(const SceneView &sc, const WavefrontView &wf, const RenderSettings &rs, cudaStream_t stream);
cudaError_t launch_sort(const WavefrontView &wf, const RenderSettings &rs, const BatchView &bv, uint32_t depth, const LaunchDims &dims,
						cudaStream_t stream);
// feature plane (albedo / normal sums of the depth-0 vertices) * scale -> out (row-major for a single shard)
cudaError_t launch_aov_finalize(const ShardView &sh, const float4 *acc, float scale, float4 *out, cudaStream_t stream);
// display pass (assets/shaders/tone-map.frag): ACES fit of the finalised framebuffer, packed to RGBA8
cudaError_t launch_tone_map(const float4 *framebuffer, uint32_t *rgba8_out, uint32_t n, float contrast, float brightness,
							cudaStream_t stream);

// --- E-mode (image model of the reference's EmbreeRT backend), one fused kernel ---
cudaError_t launch_emode(const SceneView &sc, const ShardView &sh, const WavefrontView &wf, const RenderSettings &rs,
						 const void *raw_materials, const uint32_t *tex_desc /* 5 uints per texture */,
						 uint32_t tex_count, const LaunchDims &dims, cudaStream_t stream);

// --- stage-level entry points on caller rays ---
cudaError_t launch_trace_closest(const SceneView &sc, const RenderSettings &rs, const float4 *origins,
								 const float4 *directions, uint32_t n, float t_min, float4 *hits_out, uint32_t *cursor,
								 const LaunchDims &dims, cudaStream_t stream, uint32_t *inst_out = nullptr /* two-level scenes */);
cudaError_t launch_trace_occluded(const SceneView &sc, const RenderSettings &rs, const float4 *origins,
								  const float4 *directions_tmax, uint32_t n, float t_min, uint8_t *occluded_out,
								  uint32_t *cursor, const LaunchDims &dims, cudaStream_t stream);
cudaError_t launch_generate_only(const SceneView &sc, const ShardView &sh, const WavefrontView &wf, uint32_t sample_index,
								 int emode, float4 *origins_out, float4 *directions_out, cudaStream_t stream);

// de-tile `world` gathered shards (each `stride` float4) into a row-major image
cudaError_t launch_assemble(const ShardView &sh, const float4 *gathered, size_t stride, float4 *image,
							cudaStream_t stream);

} // namespace rfwb200
