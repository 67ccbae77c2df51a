// headless.cpp -- the reference's frame loop (src/main.cpp:95-131, 277-290, 319-322) without SDL2/ImGui,
// written against include/tracer.hpp exactly as main.cpp is written against the reference Tracer:
// build the vectors, set options / scene_data, clear + update_scene when something changed,
// render(ticks, pixels) every frame, save_ppm at the end.
//
//   g++ -std=c++17 -Iinclude examples/headless.cpp -Lsimple_raytracer_b200 -lsrt_b200 \
//       -Wl,-rpath,$PWD/simple_raytracer_b200 -o build/headless
//   build/headless out.ppm [width height frames]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tracer.hpp"

using namespace srt_facade;

static srt_float3 f3(float x, float y, float z) { return srt_float3{x, y, z, 0.f}; }

static Material make_material(srt_float3 color, float smoothness = 0.f, float metallic = 0.f, float specular = 0.f,
                              float transmittance = 0.f, float ior = 1.f, srt_float3 emission = {0, 0, 0, 0},
                              float strength = 0.f) {  // Material ctor, reference include/material.hpp:23-37
	Material m;
	std::memset(&m, 0, sizeof m);
	m.color = color, m.smoothness = smoothness, m.metallic = metallic, m.specular = specular;
	m.transmittance = transmittance, m.refraction_index = ior, m.emission = emission, m.emission_strength = strength;
	return m;
}
static Shape make_sphere(int material, srt_float3 c, float r) {  // Shape(cl_int, const Sphere&), shape.hpp:96-100
	Shape s;
	std::memset(&s, 0, sizeof s);
	s.type = SRT_SHAPE_SPHERE, s.material = material, s.shape.sphere.position = c, s.shape.sphere.radius = r;
	return s;
}
static Shape make_plane(int material, srt_float3 p, srt_float3 n) {
	Shape s;
	std::memset(&s, 0, sizeof s);
	s.type = SRT_SHAPE_PLANE, s.material = material, s.shape.plane.position = p, s.shape.plane.normal = n;
	return s;
}

int main(int argc, char **argv) {
	const char *out = argc > 1 ? argv[1] : "out.ppm";
	const int width = argc > 3 ? std::atoi(argv[2]) : 480, height = argc > 3 ? std::atoi(argv[3]) : 270;
	const int frames = argc > 4 ? std::atoi(argv[4]) : 8;

	// a tiny procedural sky (the reference decodes assets/skybox.png with stb_image)
	const int sw = 64, sh = 32;
	std::vector<float> sky(sw * sh * 4);
	for (int y = 0; y < sh; ++y)
		for (int x = 0; x < sw; ++x) {
			float v = (y + 0.5f) / sh;
			float *t = &sky[4 * (y * sw + x)];
			t[0] = 0.25f + 0.3f * v, t[1] = 0.35f + 0.35f * v, t[2] = 0.5f + 0.45f * v, t[3] = 1.f;
		}

	std::vector<Shape> shapes;
	std::vector<Triangle> triangles;
	std::vector<Material> materials;
	materials.push_back(make_material(f3(0.8f, 0.8f, 0.8f)));
	materials.push_back(make_material(f3(1, 1, 1), 1.f, 0.f, 0.f, 1.f, 1.5f));
	materials.push_back(make_material(f3(0.25f, 0.4f, 0.95f), 0.95f, 1.f));
	materials.push_back(make_material(f3(1, 0.2f, 0.15f), 0, 0, 0, 0, 1, f3(1, 0.15f, 0.1f), 5.f));
	shapes.push_back(make_plane(0, f3(0, -2, 0), f3(0, 1, 0)));
	shapes.push_back(make_sphere(0, f3(-3.2f, 0, -3), 2.f));
	shapes.push_back(make_sphere(1, f3(0.6f, -0.8f, -0.5f), 1.2f));
	shapes.push_back(make_sphere(2, f3(3.4f, -0.6f, -2.6f), 1.4f));
	shapes.push_back(make_sphere(3, f3(-0.4f, -1.3f, -4.6f), 0.7f));

	Tracer tracer(width, height, sky.data(), sw, sh);
	tracer.options.num_samples = 2;   // src/main.cpp:116-118
	tracer.options.num_bounces = 10;
	tracer.options.show_normals = false;
	tracer.scene_data.sun_focus = 25.0f;  // src/main.cpp:120-126
	tracer.scene_data.sun_color = f3(1.f, 1.f, 0xd3 / 255.f);
	tracer.scene_data.sun_intensity = 1.0f;
	const float inv = 1.0f / std::sqrt(2.0f);
	tracer.scene_data.sun_direction = f3(inv, -inv, 0.f);

	std::vector<uint8_t> pixels(static_cast<size_t>(width) * height * 4);
	uint32_t time_not_moved = 1;
	for (int tick = 0; tick < frames; ++tick) {
		if (time_not_moved == 1) {  // src/main.cpp:277-280
			tracer.clear_canvas();
			tracer.update_scene(shapes, triangles, materials);
		}
		auto &options = tracer.options;  // src/main.cpp:283-288
		options.aspect_ratio = static_cast<float>(width) / height;
		options.fov_scale = 1.0f;  // tan(90 deg / 2)
		std::memset(options.camera_to_world, 0, sizeof options.camera_to_world);
		options.camera_to_world[0].x = options.camera_to_world[1].y = options.camera_to_world[2].z = 1.f;
		options.camera_to_world[3] = srt_float4{0.f, 0.5f, 5.5f, 1.f};  // translate(position), helper.hpp:21-26
		options.time = 1000003u + tick;
		options.tick = tick;
		tracer.render(time_not_moved, pixels);  // src/main.cpp:290
		time_not_moved++;
	}
	if (srt_save_ppm(out, pixels.data(), width, height) != SRT_OK) {  // `P` key, src/main.cpp:319-322
		std::fprintf(stderr, "cannot write %s\n", out);
		return 1;
	}
	std::printf("wrote %s (%dx%d, %d frames accumulated)\n", out, width, height, frames);
	return 0;
}
