// adapter_driver.cpp — plays rfw::system for the check build: dlopen("B200RT.so"), resolve the two factory symbols
// by name (system.cpp:134-157), then drive the plugin purely through the reference's abstract class
// rfw::RenderContext: a small Cornell-style scene, synchronize order of system.cpp:247-433, render_frame, read back.
// Prints the mean radiance so a test can compare it with the same scene rendered through the C ABI directly.
#include <rfw/context/context.h>

#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <vector>

const float rfw::Camera::DEFAULT_BRIGHTNESS = 0.0f;
const float rfw::Camera::DEFAULT_CONTRAST = 0.0f;
const glm::vec3 rfw::Camera::DEFAULT_POSITION = glm::vec3(0.0f);
const glm::vec3 rfw::Camera::DEFAULT_DIRECTION = glm::vec3(0.0f, 0.0f, 1.0f);

// Stand-ins for the three GL entry points the plugin resolves from its host process (the driver is linked -rdynamic):
// with `--gl` the driver hands the plugin a texture id like rfw::system does, and checks that render_frame uploaded the
// frame into "the texture" (EmbreeRT/src/Context.cpp:289-297 is the behaviour being mirrored).
static unsigned g_bound = 0, g_uploads = 0, g_upload_w = 0, g_upload_h = 0, g_upload_tex = 0;
static std::vector<float> g_texture;
extern "C" __attribute__((visibility("default"))) void glBindTexture(unsigned target, unsigned tex) { g_bound = tex; }
extern "C" __attribute__((visibility("default"))) void glTexSubImage2D(unsigned target, int level, int x, int y, int w, int h, unsigned format,
																   unsigned type, const void *data)
{
	if (target == 0x0DE1 && level == 0 && x == 0 && y == 0 && format == 0x1908 && type == 0x1406 && data)
	{
		g_texture.assign(static_cast<const float *>(data), static_cast<const float *>(data) + size_t(w) * h * 4);
		g_uploads++, g_upload_w = unsigned(w), g_upload_h = unsigned(h), g_upload_tex = g_bound;
	}
}
extern "C" __attribute__((visibility("default"))) unsigned glGetError() { return 0; }

typedef rfw::RenderContext *(*CreateFn)();
typedef void (*DestroyFn)(rfw::RenderContext *);
typedef void (*ReadFn)(rfw::RenderContext *, float *);

static rfw::Triangle make_tri(glm::vec3 a, glm::vec3 b, glm::vec3 c, uint mat)
{
	rfw::Triangle t;
	memset(&t, 0, sizeof(t));
	const glm::vec3 n = normalize(cross(b - a, c - a));
	t.vN0 = t.vN1 = t.vN2 = n;
	t.Nx = n.x, t.Ny = n.y, t.Nz = n.z;
	t.vertex0 = a, t.vertex1 = b, t.vertex2 = c;
	t.material = mat;
	t.area = 0.5f * length(cross(b - a, c - a));
	return t;
}

int main(int argc, char **argv)
{
	const char *lib = argc > 1 ? argv[1] : "./B200RT.so";
	const bool with_gl = argc > 2 && strcmp(argv[2], "--gl") == 0;
	void *h = dlopen(lib, RTLD_NOW);
	if (!h)
	{
		printf("dlopen failed: %s\n", dlerror());
		return 2;
	}
	CreateFn create = (CreateFn)dlsym(h, "createRenderContext");
	DestroyFn destroy = (DestroyFn)dlsym(h, "destroyRenderContext");
	ReadFn read_pixels = (ReadFn)dlsym(h, "b200rt_read_pixels");
	if (!create || !destroy || !read_pixels)
	{
		printf("missing factory symbols\n");
		return 2;
	}
	rfw::RenderContext *ctx = nullptr;
	try
	{
		ctx = create();
	}
	catch (const std::exception &e)
	{
		printf("create threw: %s\n", e.what()); // expected on a box without a B200: no CPU fallback
		return 3;
	}
	const uint W = 128, H = 96;
	try
	{
		printf("targets %zu\n", ctx->get_supported_targets().size());
		GLuint tex = with_gl ? 7 : 0;
		ctx->init(&tex, W, H);
		ctx->set_sky({glm::vec3(0.2f, 0.3f, 0.4f)}, 1, 1);
		ctx->set_textures({});
		std::vector<rfw::DeviceMaterial> mats(2);
		memset(mats.data(), 0, mats.size() * sizeof(rfw::DeviceMaterial));
		auto set_color = [&](int i, float r, float g, float b, float rough) {
			rfw::Material &m = reinterpret_cast<rfw::Material &>(mats[i]);
			m.diffuse_r = half(r), m.diffuse_g = half(g), m.diffuse_b = half(b);
			m.parameters.x = uint(rough * 255.0f) << 24;
			m.parameters.z = uint(0.5f * 255.0f) << 24; // eta 1.0 stored as eta*0.5
		};
		set_color(0, 0.7f, 0.7f, 0.7f, 1.0f);
		set_color(1, 20.f, 20.f, 20.f, 1.0f);
		std::vector<rfw::MaterialTexIds> ids(2);
		ctx->set_materials(mats, ids);
		// floor + back wall (material 0) and a ceiling light (material 1)
		std::vector<glm::vec4> verts;
		std::vector<rfw::Triangle> tris;
		auto quad = [&](glm::vec3 a, glm::vec3 b, glm::vec3 c, glm::vec3 d, uint mat) {
			for (glm::vec3 p : {a, b, c, a, c, d})
				verts.push_back(glm::vec4(p, 1.0f));
			tris.push_back(make_tri(a, b, c, mat));
			tris.push_back(make_tri(a, c, d, mat));
		};
		quad({-2, 0, 0}, {-2, 0, 6}, {2, 0, 6}, {2, 0, 0}, 0);
		quad({-2, 0, 6}, {-2, 4, 6}, {2, 4, 6}, {2, 0, 6}, 0);
		quad({-0.5f, 3.9f, 2.5f}, {0.5f, 3.9f, 2.5f}, {0.5f, 3.9f, 3.5f}, {-0.5f, 3.9f, 3.5f}, 1);
		rfw::DeviceAreaLight lights[2];
		for (int i = 0; i < 2; i++)
		{
			rfw::Triangle &t = tris[4 + i];
			t.lightTriIdx = i;
			memset(&lights[i], 0, sizeof(lights[i]));
			const glm::vec3 c = (t.vertex0 + t.vertex1 + t.vertex2) * (1.0f / 3.0f);
			lights[i].pos_energy = glm::vec4(c, length(glm::vec3(20.f)));
			lights[i].normal_area = glm::vec4(t.Nx, t.Ny, t.Nz, t.area);
			lights[i].radiance = glm::vec4(20.f, 20.f, 20.f, 0.f);
			lights[i].vertex0_triIdx = glm::vec4(t.vertex0, 0.f);
			lights[i].vertex1_instIdx = glm::vec4(t.vertex1, 0.f);
			lights[i].vertex2 = glm::vec4(t.vertex2, 0.f);
		}
		rfw::Mesh mesh;
		mesh.vertices = verts.data(), mesh.triangles = tris.data();
		mesh.vertexCount = verts.size(), mesh.triangleCount = tris.size();
		ctx->set_mesh(0, mesh);
		ctx->set_instance(0, 0, glm::mat4(1.0f), glm::mat3(1.0f));
		rfw::LightCount lc{2, 0, 0, 0};
		ctx->set_lights(lc, lights, nullptr, nullptr, nullptr);
		ctx->update();
		ctx->set_setting(rfw::RenderSetting("spp", "4"));
		rfw::Camera cam;
		cam.position = glm::vec3(0.f, 2.f, -4.f);
		cam.direction = glm::vec3(0.f, 0.f, 1.f);
		cam.aspectRatio = float(W) / float(H);
		cam.pixelCount = glm::ivec2(W, H);
		ctx->set_probe_index(glm::uvec2(W / 2, H - 4));
		ctx->render_frame(cam, rfw::Reset);
		ctx->render_frame(cam, rfw::Converge);
		std::vector<float> px(size_t(W) * H * 4);
		read_pixels(ctx, px.data());
		double sum = 0;
		for (size_t i = 0; i < size_t(W) * H; i++)
			sum += px[4 * i] + px[4 * i + 1] + px[4 * i + 2];
		unsigned inst = 0, prim = 0;
		float dist = 0;
		ctx->get_probe_results(&inst, &prim, &dist);
		const rfw::RenderStats st = ctx->get_stats();
		printf("mean %.6f probe %u %u %.4f primary_count %u\n", sum / (3.0 * W * H), inst, prim, dist, st.primaryCount);
		if (with_gl)
		{
			const bool same = g_texture.size() == px.size() && memcmp(g_texture.data(), px.data(), px.size() * sizeof(float)) == 0;
			printf("gl uploads %u tex %u size %ux%u identical_to_read_pixels %d\n", g_uploads, g_upload_tex, g_upload_w, g_upload_h, same ? 1 : 0);
			if (g_uploads != 2 || g_upload_tex != 7 || !same)
			{
				destroy(ctx);
				return 5;
			}
		}
	}
	catch (const std::exception &e)
	{
		printf("plugin threw: %s\n", e.what());
		destroy(ctx);
		return 4;
	}
	destroy(ctx);
	dlclose(h);
	printf("adapter ok\n");
	return 0;
}
