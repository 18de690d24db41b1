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
	size_t trace_smem = 0; // dynamic shared memory of the trace kernels (staged BVH prefix)
};

// queries occupancy for the current device and smem_nodes setting, sets function attributes
cudaError_t shade_occupancy(int block, int *per_sm); // defined next to k_shade
cudaError_t configure_launches(const RenderSettings &rs, uint32_t node_count, LaunchDims &dims);

// --- PT-mode stages (one sample = primary, shade(0), [trace(d), shade(d)] for d = 1..max) ---
cudaError_t launch_primary(const SceneView &sc, const ShardView &sh, const WavefrontView &wf, const RenderSettings &rs,
						   uint32_t sample_in_frame, const LaunchDims &dims, cudaStream_t stream);
cudaError_t launch_shade(const SceneView &sc, const ShardView &sh, const WavefrontView &wf, const RenderSettings &rs,
						 uint32_t sample_in_frame, uint32_t depth, const LaunchDims &dims, cudaStream_t stream);
// the same kernel built without -use_fast_math (setting "shade_math" = "ieee")
cudaError_t launch_shade_ieee(const SceneView &sc, const ShardView &sh, const WavefrontView &wf, const RenderSettings &rs,
							  uint32_t sample_in_frame, uint32_t depth, const LaunchDims &dims, cudaStream_t stream);
cudaError_t launch_trace(const SceneView &sc, const ShardView &sh, const WavefrontView &wf, const RenderSettings &rs,
						 uint32_t sample_in_frame, uint32_t depth, const LaunchDims &dims, cudaStream_t stream);
cudaError_t launch_finalize(const ShardView &sh, const WavefrontView &wf, float scale, cudaStream_t stream);
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
								 const LaunchDims &dims, cudaStream_t stream);
cudaError_t launch_trace_occluded(const SceneView &sc, const RenderSettings &rs, const float4 *origins,
								  const float4 *directions_tmax, uint32_t n, float t_min, uint8_t *occluded_out,
								  uint32_t *cursor, const LaunchDims &dims, cudaStream_t stream);
cudaError_t launch_generate_only(const SceneView &sc, const ShardView &sh, const WavefrontView &wf, uint32_t sample_index,
								 int emode, float4 *origins_out, float4 *directions_out, cudaStream_t stream);

// de-tile `world` gathered shards (each `stride` float4) into a row-major image
cudaError_t launch_assemble(const ShardView &sh, const float4 *gathered, size_t stride, float4 *image,
							cudaStream_t stream);

} // namespace rfwb200
