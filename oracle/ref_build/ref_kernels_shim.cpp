// ref_kernels_shim.cpp — TEST INFRASTRUCTURE.  The reference's own CUDA wavefront path tracer,
// RFW/backends/CUDART/src/Kernels.cu (generatePrimaryRay, intersect_rays for the three stages, shade_rays, blit_buffer and
// every set* symbol upload), compiled for the HOST from the reference tree and run one CUDA thread after the other.  The
// kernels use no shared memory, barriers or warp intrinsics, so thread-by-thread execution is what a GPU computes up to
// the order of atomics.  The Makefile compiles a transformed copy, kept in a temporary directory for the duration of the compile, in which the six
// `kernel<<<grid, block>>>(args)` launches read `RFW_LAUNCH(kernel, grid, block, args)` (the only syntax a host compiler
// cannot parse); shim_inc/cuda_kernels_standin.h supplies the CUDA names the file uses.  What is ours here: the C
// interface below and the bounce loop of CUDAContext::render_frame (CUDART/src/Context.cpp:83-159), which cannot be
// compiled (GL interop, the Rust BVH crate) and is restated call for call.
//
// tests/test_ref_pin.py uses this to pin the oracle's restatement of the whole PT pipeline — blue-noise camera rays,
// two-level MBVH traversal, shade_rays control flow, NEE, connect, bounce loop — on a scene that keeps the documented
// deviations D1-D6 (oracle/rfw_oracle.cpp header) out of play.
#include <cuda_kernels_standin.h>

#include "kernels_host.inc" // temporary: Kernels.cu with the launch syntax rewritten (Makefile)

#include <vector>

#define REF_API extern "C" __attribute__((visibility("default")))

struct RefMesh
{
	const float *vertices4;
	const unsigned *indices3; // may be null
	const void *triangles160;
	const void *mbvh_nodes;
	const unsigned *prim_indices;
};
struct RefInstance
{
	int mesh;
	float transform[16], inverse[16], normal[16]; // column-major mat4 each (normal = mat4 of the normal matrix)
};
struct RefScene
{
	int n_meshes;
	const RefMesh *meshes;
	int n_instances;
	const RefInstance *instances;
	const void *tlas_nodes;
	const unsigned *tlas_prims;
	const void *materials192;
	const unsigned *uint_texels;
	const float *float_texels4;
	const float *sky3;
	unsigned sky_w, sky_h;
	unsigned n_area, n_point, n_spot, n_dir;
	const void *area, *point, *spot, *dir;
	const unsigned *blue_noise; // 5 x 65536 uints (createBlueNoiseBuffer)
};

// probe pixel of the next render (RenderContext::set_probe_index) and what shade_rays stored for it at path length 0
// (Kernels.cu:626-631), read back as CUDAContext::render_frame does for the first sample (Context.cpp:103-109)
static unsigned g_probe_index = 0xffffffffu;
static int g_probed_instance = -1, g_probed_prim = -1;
static float g_probed_distance = -1.0f;
REF_API void rfwref_set_probe_index(unsigned index)
{
	g_probe_index = index;
	g_probed_instance = g_probed_prim = -1, g_probed_distance = -1.0f;
}
REF_API void rfwref_get_probe_results(int *instance, int *prim, float *distance)
{
	*instance = g_probed_instance, *prim = g_probed_prim, *distance = g_probed_distance;
}

static mat4 load_mat4(const float *m)
{
	mat4 r;
	memcpy(&r, m, 64);
	return r;
}

// Renders samples [first, first + count) of a w x h frame from a cleared accumulator; accumulator_out receives the
// raw accumulator (vec4 per pixel), primary_* (optional, w*h vec4 each) the buffers after the Primary stage of the first
// sample, counters_out (optional, 3 uints per sample and depth slot of 8) the queue sizes the host loop saw.
REF_API int rfwref_cudart_render(const RefScene *sc, const float *view14, unsigned w, unsigned h, unsigned first, unsigned count,
								 float clamp_value, float *accumulator_out, float *primary_origins, float *primary_directions,
								 float *primary_states, unsigned *counters_out)
{
	const unsigned N = w * h;
	std::vector<InstanceBVHDescriptor> inst(sc->n_instances);
	for (int i = 0; i < sc->n_instances; i++)
	{
		const RefInstance &in = sc->instances[i];
		const RefMesh &m = sc->meshes[in.mesh];
		InstanceBVHDescriptor &d = inst[i];
		d.mbvh = static_cast<const bvh::MBVHNode *>(m.mbvh_nodes);
		d.bvh_indices = m.prim_indices;
		d.vertices = reinterpret_cast<const vec4 *>(m.vertices4);
		d.indices = reinterpret_cast<const uvec3 *>(m.indices3);
		d.triangles = static_cast<const DeviceTriangle *>(m.triangles160);
		d.instance_transform = load_mat4(in.transform);
		d.inverse_transform = load_mat4(in.inverse);
		const mat4 nm = load_mat4(in.normal);
		for (int c = 0; c < 3; c++)
			d.normal_transform.c[c] = nm[c];
		d.bvh = nullptr;
	}
	CameraView cam;
	cam.pos = vec3(view14[0], view14[1], view14[2]), cam.p1 = vec3(view14[3], view14[4], view14[5]);
	cam.p2 = vec3(view14[6], view14[7], view14[8]), cam.p3 = vec3(view14[9], view14[10], view14[11]);
	cam.aperture = view14[12], cam.spreadAngle = view14[13];
	std::vector<Counters> counter_store(1); // on the heap, like the mapped host copy CUDAContext reads (m_Counters)
	Counters &cnt = counter_store[0];
	memset(&cnt, 0, sizeof(cnt));
	cnt.probeIdx = g_probe_index; // counters->probeIdx = probe.x + probe.y * width (Context.cpp:70)
	std::vector<vec4> acc(N, vec4(0.0f)), states(2 * N), origins(2 * N), directions(2 * N), throughputs(2 * N);
	std::vector<PotentialContribution> connect(N);

	// the uploads of CUDAContext (Context.cpp:36-60, 167-268, 394-456), through the reference's own set* functions
	setTopLevelMBVH(const_cast<bvh::MBVHNode *>(static_cast<const bvh::MBVHNode *>(sc->tlas_nodes)));
	setTopPrimIndices(const_cast<uint *>(sc->tlas_prims));
	setInstances(inst.data());
	setCameraView(&cam);
	setCounters(&cnt);
	setAccumulator(acc.data());
	setStride(N);
	setPathStates(states.data()), setPathOrigins(origins.data()), setPathDirections(directions.data());
	setPathThroughputs(throughputs.data());
	setPotentialContributions(connect.data());
	setMaterials(const_cast<DeviceMaterial *>(static_cast<const DeviceMaterial *>(sc->materials192)));
	setFloatTextures(const_cast<vec4 *>(reinterpret_cast<const vec4 *>(sc->float_texels4)));
	setUintTextures(const_cast<uint *>(sc->uint_texels));
	setSkybox(const_cast<vec3 *>(reinterpret_cast<const vec3 *>(sc->sky3)));
	setSkyDimensions(sc->sky_w, sc->sky_h);
	setGeometryEpsilon(1e-5f);
	setBlueNoiseBuffer(const_cast<uint *>(sc->blue_noise));
	setScreenDimensions(w, h);
	LightCount lc;
	lc.areaLightCount = sc->n_area, lc.pointLightCount = sc->n_point, lc.spotLightCount = sc->n_spot, lc.directionalLightCount = sc->n_dir;
	setLightCount(lc);
	setAreaLights(const_cast<DeviceAreaLight *>(static_cast<const DeviceAreaLight *>(sc->area)));
	setPointLights(const_cast<DevicePointLight *>(static_cast<const DevicePointLight *>(sc->point)));
	setSpotLights(const_cast<DeviceSpotLight *>(static_cast<const DeviceSpotLight *>(sc->spot)));
	setDirectionalLights(const_cast<DeviceDirectionalLight *>(static_cast<const DeviceDirectionalLight *>(sc->dir)));
	setClampValue(clamp_value);

	for (unsigned s = first; s < first + count; s++)
	{
		// CUDAContext::render_frame, Context.cpp:83-159 (timers, stats and the GL blit left out)
		cnt.samplesTaken = s; // what `counters->samplesTaken = m_SampleIndex` of the previous frame left (Context.cpp:156)
		unsigned pathLength = 0;
		const unsigned pathCount = N;
		InitCountersForExtend(pathCount, s);
		intersectRays(Primary, pathLength, w, h);
		if (s == first)
		{
			if (primary_origins)
				memcpy(primary_origins, origins.data(), size_t(N) * 16);
			if (primary_directions)
				memcpy(primary_directions, directions.data(), size_t(N) * 16);
			if (primary_states)
				memcpy(primary_states, states.data(), size_t(N) * 16);
		}
		shadeRays(pathLength, pathCount);
		unsigned activePaths = cnt.extensionRays;
		if (s == 0) // Context.cpp:103-109 (m_SampleIndex == 0)
			g_probed_instance = cnt.probedInstanceId, g_probed_prim = cnt.probedPrimId, g_probed_distance = cnt.probedDistance;
		unsigned *rec = counters_out ? counters_out + size_t(s - first) * 8 * 3 : nullptr;
		if (rec)
			rec[0] = cnt.extensionRays, rec[1] = cnt.shadowRays, rec[2] = pathCount;
		while (activePaths > 0 && pathLength < MAX_PATH_LENGTH)
		{
			pathLength = pathLength + 1;
			if (cnt.shadowRays > 0)
				intersectRays(Shadow, pathLength, cnt.shadowRays);
			InitCountersSubsequent();
			intersectRays(Secondary, pathLength, activePaths);
			shadeRays(pathLength, activePaths);
			if (rec && pathLength < 8)
				rec[3 * pathLength] = cnt.extensionRays, rec[3 * pathLength + 1] = cnt.shadowRays, rec[3 * pathLength + 2] = activePaths;
			activePaths = cnt.extensionRays;
		}
	}
	memcpy(accumulator_out, acc.data(), size_t(N) * 16);
	return 0;
}

// finalize: the reference's own blit_buffer (Kernels.cu:181-203) from an accumulator into an RGBA32F image through the surface
// stand-in, with the scale CUDAContext::render_frame passes (1 / m_SampleIndex after the increment, Context.cpp:149-152)
REF_API int rfwref_cudart_blit(const float *accumulator_in, unsigned w, unsigned h, unsigned samples, float *image_out)
{
	std::vector<vec4> acc(size_t(w) * h);
	memcpy(acc.data(), accumulator_in, acc.size() * 16);
	setAccumulator(acc.data());
	output.pixels = image_out, output.width = w;
	blitBuffer(w, h, samples);
	output.pixels = nullptr;
	return 0;
}
