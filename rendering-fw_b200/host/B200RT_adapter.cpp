// B200RT_adapter.cpp — the reference-side binding: a rfw::RenderContext plugin that forwards to the C ABI of
// librfwb200.so.  Build it inside the reference tree like any backend (RFW/backends/<X>/CMakeLists.txt pattern:
// shared library with PREFIX "" named B200RT, linked against RenderContext + librfwb200); rfw::system loads it with
// load_render_api("B200RT") (RFW/system/src/rfw/system.cpp:107-158) and drives it unchanged.
//
// Interface implemented: rfw::RenderContext, RFW/system/context/rfw/context/context.h:74-111.
// Factory symbols: RFW/system/context/rfw/context/export.h:11-15.
// Errors: C ABI codes become std::runtime_error, the reference's convention (CUDART/src/CheckCUDA.h:7-20).
#include <rfw/context/context.h>

#include "../../include/rfwb200.h"

#include <dlfcn.h>

#include <cmath>
#include <cstdlib>
#include <sstream>
#include <stdexcept>
#include <string>

static_assert(sizeof(rfw::Triangle) == sizeof(rfwb200_triangle), "Triangle wire format");
static_assert(sizeof(rfw::DeviceMaterial) == sizeof(rfwb200_material), "DeviceMaterial wire format");
static_assert(sizeof(rfw::MaterialTexIds) == sizeof(rfwb200_material_tex_ids), "MaterialTexIds wire format");
static_assert(sizeof(rfw::TextureData) == sizeof(rfwb200_texture_data), "TextureData wire format");
static_assert(sizeof(rfw::Mesh) == sizeof(rfwb200_mesh), "Mesh wire format");
static_assert(sizeof(rfw::DeviceAreaLight) == sizeof(rfwb200_area_light), "AreaLight wire format");
static_assert(sizeof(rfw::DevicePointLight) == sizeof(rfwb200_point_light), "PointLight wire format");
static_assert(sizeof(rfw::DeviceSpotLight) == sizeof(rfwb200_spot_light), "SpotLight wire format");
static_assert(sizeof(rfw::DeviceDirectionalLight) == sizeof(rfwb200_directional_light), "DirectionalLight wire format");
static_assert(sizeof(rfw::CameraView) == sizeof(rfwb200_camera_view), "CameraView wire format");
static_assert(sizeof(rfw::RenderStats) == sizeof(rfwb200_render_stats), "RenderStats wire format");
static_assert(sizeof(glm::mat4) == 64 && sizeof(glm::mat3) == 36 && sizeof(glm::vec3) == 12, "glm layouts");

namespace
{
void check(int rc, const char *what)
{
	if (rc != RFWB200_OK)
		throw std::runtime_error(std::string("B200RT: ") + what + ": " + rfwb200_last_error());
}

// The three GL entry points the texture target needs, resolved from the host process at run time: rfw::system created
// the GL context and the texture (RFW/system/src/rfw/system.cpp:204-211), so libGL is already loaded there.  The plugin
// itself links no GL library (a headless host has none); without these symbols only RenderTarget::BUFFER is advertised.
struct GLApi
{
	void (*BindTexture)(unsigned, unsigned) = nullptr;
	void (*TexSubImage2D)(unsigned, int, int, int, int, int, unsigned, unsigned, const void *) = nullptr;
	unsigned (*GetError)() = nullptr;
	bool ok() const { return BindTexture && TexSubImage2D; }
	static GLApi load()
	{
		GLApi g;
		g.BindTexture = reinterpret_cast<decltype(g.BindTexture)>(dlsym(RTLD_DEFAULT, "glBindTexture"));
		g.TexSubImage2D = reinterpret_cast<decltype(g.TexSubImage2D)>(dlsym(RTLD_DEFAULT, "glTexSubImage2D"));
		g.GetError = reinterpret_cast<decltype(g.GetError)>(dlsym(RTLD_DEFAULT, "glGetError"));
		return g;
	}
};

class B200Context final : public rfw::RenderContext
{
  public:
	B200Context() : m_GL(GLApi::load())
	{
		// RFWB200_DEVICES="0,1,2,3": one context that owns several GPUs of the box and shards every frame over them
		// (rfwb200_create_group); RFWB200_DEVICE=n / nothing: one GPU
		std::vector<int> devices;
		if (const char *list = std::getenv("RFWB200_DEVICES"))
		{
			std::stringstream ss(list);
			std::string tok;
			while (std::getline(ss, tok, ','))
				if (!tok.empty())
					devices.push_back(std::atoi(tok.c_str()));
		}
		if (devices.size() > 1)
			check(rfwb200_create_group(devices.data(), devices.size(), &m_Ctx), "create_group");
		else
		{
			const char *dev = std::getenv("RFWB200_DEVICE");
			check(rfwb200_create(devices.size() == 1 ? devices[0] : (dev ? std::atoi(dev) : 0), &m_Ctx), "create");
		}
	}
	~B200Context() override { cleanup(); }

	[[nodiscard]] std::vector<rfw::RenderTarget> get_supported_targets() const override
	{
		if (m_GL.ok())
			return {rfw::RenderTarget::BUFFER, rfw::RenderTarget::OPENGL_TEXTURE};
		return {rfw::RenderTarget::BUFFER};
	}

	// The frame is rendered into a linear-HDR RGBA32F device buffer.  With a GL texture target the finished frame is
	// copied into the texture at the end of render_frame, RGBA / FLOAT rows in framebuffer order, as EmbreeRT fills its
	// target (EmbreeRT/src/Context.cpp:289-297: PBO + glTexSubImage2D) — rfw::system only draws that texture.
	void init(GLuint *glTextureID, uint width, uint height) override
	{
		m_Texture = glTextureID ? *glTextureID : 0;
		if (m_Texture != 0 && !m_GL.ok())
			throw std::runtime_error("B200RT: init(GLuint*) needs glBindTexture / glTexSubImage2D in the host process (no GL library is loaded)");
		m_Width = width, m_Height = height;
		check(rfwb200_init(m_Ctx, width, height), "init");
		if (m_Texture != 0)
			m_Staging.assign(size_t(width) * height * 4, 0.0f);
	}

	void cleanup() override
	{
		if (m_Ctx) // idempotent: the system calls cleanup() and then destroyRenderContext() (system.cpp:164-165)
			rfwb200_destroy(m_Ctx);
		m_Ctx = nullptr;
	}

	void render_frame(const rfw::Camera &camera, rfw::RenderStatus status) override
	{
		const rfwb200_camera_view view = get_view(camera);
		check(rfwb200_render_frame(m_Ctx, &view, status == rfw::Reset ? RFWB200_RESET : RFWB200_CONVERGE), "render_frame");
		if (m_Texture != 0)
		{
			constexpr unsigned GL_TEXTURE_2D_ = 0x0DE1, GL_RGBA_ = 0x1908, GL_FLOAT_ = 0x1406;
			read_pixels(m_Staging.data());
			m_GL.BindTexture(GL_TEXTURE_2D_, m_Texture);
			m_GL.TexSubImage2D(GL_TEXTURE_2D_, 0, 0, 0, int(m_Width), int(m_Height), GL_RGBA_, GL_FLOAT_, m_Staging.data());
			m_GL.BindTexture(GL_TEXTURE_2D_, 0);
			if (m_GL.GetError)
				if (const unsigned e = m_GL.GetError())
					throw std::runtime_error("B200RT: glTexSubImage2D failed with GL error " + std::to_string(e));
		}
	}

	void set_materials(const std::vector<rfw::DeviceMaterial> &materials, const std::vector<rfw::MaterialTexIds> &texDescriptors) override
	{
		check(rfwb200_set_materials(m_Ctx, reinterpret_cast<const rfwb200_material *>(materials.data()),
									reinterpret_cast<const rfwb200_material_tex_ids *>(texDescriptors.data()), materials.size()),
			  "set_materials");
	}
	void set_textures(const std::vector<rfw::TextureData> &textures) override
	{
		check(rfwb200_set_textures(m_Ctx, reinterpret_cast<const rfwb200_texture_data *>(textures.data()), textures.size()), "set_textures");
	}
	void set_mesh(size_t index, const rfw::Mesh &mesh) override
	{
		check(rfwb200_set_mesh(m_Ctx, index, reinterpret_cast<const rfwb200_mesh *>(&mesh)), "set_mesh");
	}
	void set_instance(size_t i, size_t meshIdx, const mat4 &transform, const mat3 &inverse_transform) override
	{
		check(rfwb200_set_instance(m_Ctx, i, meshIdx, reinterpret_cast<const float *>(&transform), reinterpret_cast<const float *>(&inverse_transform)),
			  "set_instance");
	}
	void set_sky(const std::vector<glm::vec3> &pixels, size_t width, size_t height) override
	{
		check(rfwb200_set_sky(m_Ctx, reinterpret_cast<const float *>(pixels.data()), width, height), "set_sky");
	}
	void set_lights(rfw::LightCount lightCount, const rfw::DeviceAreaLight *areaLights, const rfw::DevicePointLight *pointLights,
					const rfw::DeviceSpotLight *spotLights, const rfw::DeviceDirectionalLight *directionalLights) override
	{
		const rfwb200_light_count n{lightCount.areaLightCount, lightCount.pointLightCount, lightCount.spotLightCount, lightCount.directionalLightCount};
		check(rfwb200_set_lights(m_Ctx, n, reinterpret_cast<const rfwb200_area_light *>(areaLights),
								 reinterpret_cast<const rfwb200_point_light *>(pointLights), reinterpret_cast<const rfwb200_spot_light *>(spotLights),
								 reinterpret_cast<const rfwb200_directional_light *>(directionalLights)),
			  "set_lights");
	}
	void get_probe_results(unsigned int *instanceIndex, unsigned int *primitiveIndex, float *distance) const override
	{
		check(rfwb200_get_probe_results(m_Ctx, instanceIndex, primitiveIndex, distance), "get_probe_results");
	}
	rfw::AvailableRenderSettings get_settings() const override
	{
		rfw::AvailableRenderSettings s;
		s.settingKeys = {"spp", "mode", "max_path_length"};
		s.settingValues = {{"1", "2", "4", "8", "16"}, {"pt", "embree"}, {"0", "1", "2", "3", "4"}};
		return s;
	}
	void set_setting(const rfw::RenderSetting &setting) override
	{
		check(rfwb200_set_setting(m_Ctx, setting.name.c_str(), setting.value.c_str()), "set_setting");
	}
	void update() override { check(rfwb200_update(m_Ctx), "update"); }
	void set_probe_index(glm::uvec2 probePos) override { check(rfwb200_set_probe_index(m_Ctx, probePos.x, probePos.y), "set_probe_index"); }
	rfw::RenderStats get_stats() const override
	{
		rfw::RenderStats stats;
		check(rfwb200_get_stats(m_Ctx, reinterpret_cast<rfwb200_render_stats *>(&stats)), "get_stats");
		return stats;
	}

	// extension used by headless hosts: copy the finished frame (width*height RGBA32F) to host memory
	void read_pixels(float *rgba) const { check(rfwb200_read_framebuffer(m_Ctx, rgba, size_t(m_Width) * m_Height), "read_pixels"); }

  private:
	// rfw::Camera::get_view (context/Camera.cpp:74-88,109-115) on the camera's public fields, so the plugin does not
	// have to link the RenderContext static library
	static rfwb200_camera_view get_view(const rfw::Camera &c)
	{
		const glm::vec3 z = c.direction;
		const glm::vec3 x = normalize(cross(z, glm::vec3(0.0f, 1.0f, 0.0f)));
		const glm::vec3 y = cross(x, z);
		rfwb200_camera_view v;
		const float spread = (c.FOV * 3.14159265358979323846f / 180) / float(c.pixelCount.y);
		const float screenSize = std::tan(c.FOV / 2.0f / (180.0f / 3.14159265358979323846f));
		const glm::vec3 centre = c.position + c.focalDistance * z;
		const glm::vec3 p1 = centre - screenSize * x * c.focalDistance * c.aspectRatio + screenSize * c.focalDistance * y;
		const glm::vec3 p2 = centre + screenSize * x * c.focalDistance * c.aspectRatio + screenSize * c.focalDistance * y;
		const glm::vec3 p3 = centre - screenSize * x * c.focalDistance * c.aspectRatio - screenSize * c.focalDistance * y;
		for (int i = 0; i < 3; i++)
			v.pos[i] = c.position[i], v.p1[i] = p1[i], v.p2[i] = p2[i], v.p3[i] = p3[i];
		v.aperture = c.aperture, v.spread_angle = spread;
		return v;
	}

	rfwb200_context *m_Ctx = nullptr;
	GLApi m_GL;
	std::vector<float> m_Staging;
	GLuint m_Texture = 0;
	uint m_Width = 0, m_Height = 0;
};
} // namespace

extern "C" __attribute__((visibility("default"))) rfw::RenderContext *createRenderContext() { return new B200Context(); }
extern "C" __attribute__((visibility("default"))) void destroyRenderContext(rfw::RenderContext *ptr)
{
	ptr->cleanup();
	delete ptr;
}
// headless helper for hosts without a GL context (not part of the reference interface)
extern "C" __attribute__((visibility("default"))) void b200rt_read_pixels(rfw::RenderContext *ptr, float *rgba)
{
	static_cast<B200Context *>(ptr)->read_pixels(rgba);
}
