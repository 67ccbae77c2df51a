// headless.cpp -- the reference's frame loop (src/main.cpp:95-131, 277-290, 319-322) without SDL2/ImGui,
// written against include/tracer.hpp + include/scene.hpp exactly as main.cpp is written against the reference's
// Tracer / shape.hpp / parser.hpp: build the vectors with the reference's constructors, Box::create_triangle first
// (main.cpp:102), optionally "Add model" from a mesh file (interface.cpp:281-301), set options / scene_data,
// clear + update_scene when something changed, render(ticks, pixels) every frame, save_ppm at the end.
//
//   g++ -std=c++17 -Iinclude examples/headless.cpp -Lsimple_raytracer_b200 -lsrt_b200
//       -Wl,-rpath,$PWD/simple_raytracer_b200 -o build/headless
//   build/headless out.ppm [width height frames [mesh.stl|mesh.obj]]
// With SRT_DUMP_SCENE=<prefix> in the environment the three scene vectors are written to <prefix>.shapes /
// .triangles / .materials as raw bytes and the program stops before touching the GPU (used by the CPU tests).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "tracer.hpp"

using namespace srt_facade;

int main(int argc, char **argv) {
	const char *out = argc > 1 ? argv[1] : "out.ppm";
	const int width = argc > 3 ? std::atoi(argv[2]) : 480, height = argc > 3 ? std::atoi(argv[3]) : 270;
	const int frames = argc > 4 ? std::atoi(argv[4]) : 8;
	const std::string mesh = argc > 5 ? argv[5] : "";

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
	Box::create_triangle(triangles);  // src/main.cpp:102: the cube's 12 triangles are always triangles [0, 12)
	materials.push_back(Material(Color(0.8f, 0.8f, 0.8f)));
	materials.push_back(Material(color::white, 1.f, 0.f, 0.f, 1.f, 1.5f));
	materials.push_back(Material(Color(0.25f, 0.4f, 0.95f), 0.95f, 1.f));
	materials.push_back(Material(Color(1.f, 0.2f, 0.15f), 0.f, 0.f, 0.f, 0.f, 1.f, Color(1.f, 0.15f, 0.1f), 5.f));
	shapes.push_back({0, Plane({0, -2, 0}, {0, 1, 0})});
	shapes.push_back({0, Sphere({-3.2f, 0, -3}, 2.f)});
	shapes.push_back({1, Sphere({0.6f, -0.8f, -0.5f}, 1.2f)});
	shapes.push_back({2, Sphere({3.4f, -0.6f, -2.6f}, 1.4f)});
	shapes.push_back({3, Sphere({-0.4f, -1.3f, -4.6f}, 0.7f)});
	shapes.push_back({2, Box::model({1.5f, -1.f, 1.f}, vec3(2.f))});  // "Add box", src/interface.cpp:162
	if (!mesh.empty()) {  // "Add model", src/interface.cpp:281-301
		const bool obj = mesh.size() > 4 && mesh.compare(mesh.size() - 4, 4, ".obj") == 0;
		const auto indices = obj ? load_obj_model(mesh, triangles) : load_stl_model(mesh, triangles);
		if (!indices.has_value()) {
			std::fprintf(stderr, "Inexistant file %s\n", mesh.c_str());
			return 2;
		}
		shapes.push_back({0, Model(triangles, indices->first, indices->second)});
	}

	if (const char *prefix = std::getenv("SRT_DUMP_SCENE")) {
		auto dump = [&](const char *ext, const void *p, size_t bytes) {
			FILE *f = std::fopen((std::string(prefix) + ext).c_str(), "wb");
			if (!f || std::fwrite(p, 1, bytes, f) != bytes) std::exit(3);
			std::fclose(f);
		};
		dump(".shapes", shapes.data(), shapes.size() * sizeof(Shape));
		dump(".triangles", triangles.data(), triangles.size() * sizeof(Triangle));
		dump(".materials", materials.data(), materials.size() * sizeof(Material));
		return 0;
	}

	// With SRT_SKYBOX_PNG=<file> the tracer decodes the sky box itself, as the reference's constructor does with
	// "assets/skybox.png" (tracer.cpp:42-52); otherwise the procedural texels above are handed over.
	const char *sky_png = std::getenv("SRT_SKYBOX_PNG");
	Tracer tracer = sky_png ? Tracer(width, height, std::string(sky_png)) : Tracer(width, height, sky.data(), sw, sh);
	tracer.options.num_samples = 2;   // src/main.cpp:116-118
	tracer.options.num_bounces = 10;
	tracer.options.show_normals = false;
	tracer.scene_data.sun_focus = 25.0f;  // src/main.cpp:120-126
	tracer.scene_data.sun_color = to_record(color::from_hex(0xffffd3));
	tracer.scene_data.sun_intensity = 1.0f;
	tracer.scene_data.sun_direction = to_record(normalize(vec3(1.0f, -1.0f, 0.0f)));
	Camera camera{{0.0f, 0.5f, 5.5f}, 0.0f, 0.0f};

	std::vector<uint8_t> pixels(static_cast<size_t>(width) * height * 4);
	tracer.pin_output(pixels);  // optional: `pixels` outlives every render() call below
	uint32_t time_not_moved = 1;
	for (int tick = 0; tick < frames; ++tick) {
		if (time_not_moved == 1) {  // src/main.cpp:277-280
			tracer.clear_canvas();
			tracer.update_scene(shapes, triangles, materials);
		}
		auto &options = tracer.options;  // src/main.cpp:283-288
		options.aspect_ratio = static_cast<float>(width) / height;
		options.fov_scale = 1.0f;  // tan(90 deg / 2)
		camera.camera_matrix(options.camera_to_world);  // helper.hpp:21-26
		options.time = 1000003u + tick;
		options.tick = tick;
		tracer.render(time_not_moved, pixels);  // src/main.cpp:290
		time_not_moved++;
	}
	tracer.unpin_output();
	save_ppm(out, pixels, width, height);  // `P` key, src/main.cpp:319-322
	std::printf("wrote %s (%dx%d, %d frames accumulated, %zu shapes, %zu triangles)\n", out, width, height, frames,
	            shapes.size(), triangles.size());
	return 0;
}
