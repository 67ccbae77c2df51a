// tracer.hpp -- header-only C++ facade with the reference's `Tracer` surface over the C ABI (srt.h).
//
// The reference class (reference include/tracer.hpp:26-88, src/tracer.cpp:11-116) is what main.cpp and
// interface.cpp talk to: a constructor, public mutable `options` / `scene_data`, `update_scene`,
// `clear_canvas`, `render(ticks_stopped, output)`.  This header keeps exactly those names, argument
// meanings and the frame protocol (src/main.cpp:277-290) so that callers compile unchanged; the
// boost.compute members are replaced by one srt_tracer handle.  Errors surface as std::runtime_error
// where the reference throws boost::compute::opencl_error.
//
// Neither glm nor the OpenCL `cl_*` typedefs are required: `Shape` / `Triangle` / `Material` (scene.hpp) derive from
// the plain-C records of srt.h, byte-identical to the reference's structs (static_asserted in srt_api.cu), and carry
// the reference's constructors, `Model` / `Box` helpers and the `load_*_model` / `save_ppm` functions.
// A build that still has glm keeps using its own shape.hpp/material.hpp types and passes
// reinterpret_cast pointers -- the layouts are the same.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "scene.hpp"
#include "srt.h"

namespace srt_facade {

class Tracer {
  public:
	using RenderData = srt_render_data;  // reference tracer.hpp:48-67
	using SceneData = srt_scene_data;    // reference tracer.hpp:69-80

	RenderData options{};
	SceneData scene_data{};

	// Tracer::Tracer(width, height), tracer.cpp:11-68.  The reference reads "assets/skybox.png" itself;
	// here the decoded RGBA-f32 texels (row 0 = bottom, stb's pow(x/255, 2.2) applied) are passed in.
	Tracer(int width, int height, const float *skybox_rgba, int sky_w, int sky_h, int device = -1) {
		options.width = width;
		options.height = height;
		options.num_samples = 4;  // RenderData(int, int), tracer.hpp:61-66
		options.num_bounces = 10;
		if (srt_create(width, height, skybox_rgba, sky_w, sky_h, device, &handle_) != SRT_OK)
			throw std::runtime_error(std::string("srt_create: ") + srt_last_error(nullptr));
	}
	// Tracer::Tracer(width, height) as the reference writes it: the constructor decodes the sky-box PNG itself
	// (tracer.cpp:42-45 opens "assets/skybox.png"; the path is an argument here, with that default).
	Tracer(int width, int height, const std::string &skybox_png = "assets/skybox.png", int device = -1) {
		options.width = width;
		options.height = height;
		options.num_samples = 4;
		options.num_bounces = 10;
		float *sky = nullptr;
		int sw = 0, sh = 0;
		if (srt_load_skybox_png(skybox_png.c_str(), &sky, &sw, &sh) != SRT_OK)
			throw std::runtime_error("Tracer: cannot read the sky box " + skybox_png);
		const int rc = srt_create(width, height, sky, sw, sh, device, &handle_);
		srt_free(sky);
		if (rc != SRT_OK) throw std::runtime_error(std::string("srt_create: ") + srt_last_error(nullptr));
	}
	~Tracer() { srt_destroy(handle_); }
	Tracer(const Tracer &) = delete;
	Tracer &operator=(const Tracer &) = delete;

	// tracer.cpp:70-96
	void update_scene(const std::vector<Shape> &shapes, const std::vector<Triangle> &triangles,
	                  const std::vector<Material> &materials) {
		scene_data.num_shapes = static_cast<int32_t>(shapes.size());
		check(srt_upload_scene(handle_, shapes.data(), shapes.size(), triangles.data(), triangles.size(), materials.data(),
		                       materials.size(), &scene_data));
	}

	// tracer.cpp:98-101
	void clear_canvas() { check(srt_clear(handle_)); }

	// tracer.cpp:103-116: render kernel, average kernel, blocking read-back of width*height*4 ARGB bytes
	void render(uint32_t ticks_stopped, std::vector<uint8_t> &output) {
		if (output.size() < static_cast<size_t>(options.width) * options.height * 4)
			throw std::runtime_error("Tracer::render: output must hold width*height*4 bytes (src/main.cpp:128)");
		check(srt_render_frame(handle_, &options, ticks_stopped, output.data()));
	}

	// Optional (srt_pin_output): page-lock the `pixels` vector that is handed to render() every frame
	// (src/main.cpp:128,290) so the read-back lands in it directly.  The vector must not be resized or destroyed
	// before unpin_output() or this Tracer's destructor.
	void pin_output(std::vector<uint8_t> &output) { check(srt_pin_output(handle_, output.data(), output.size())); }
	void unpin_output() { check(srt_unpin_output(handle_)); }

	srt_tracer *handle() { return handle_; }

  private:
	void check(int rc) {
		if (rc != SRT_OK) throw std::runtime_error(srt_last_error(handle_));
	}
	srt_tracer *handle_ = nullptr;
};

}  // namespace srt_facade
