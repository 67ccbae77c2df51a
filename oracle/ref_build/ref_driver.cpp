// ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Runs the reference's OWN kernel source on the CPU.  /root/reference/src/render.cl is compiled by g++
// from where it lies: oracle/Makefile pipes it through rewrite_cl.py (vector literals -> brace lists, the
// one syntactic difference C++ cannot absorb) into the #include below; cl_shim.hpp supplies the OpenCL C
// language surface.  This file adds what an OpenCL runtime would: the NDRange loop that sets
// get_global_id() and calls the kernels `render` (render.cl:483-523) and `average` (:525-535) once per
// work-item, exactly as src/tracer.cpp:103-115 enqueues them (global size width x height, resp.
// width*height).  It plays the role north_star gives to "render.cl on a CPU OpenCL device".
//
// Built with -ffp-contract=off: the compiler may not fuse the kernel's own a*b+c expressions; the only
// fused operations are inside the builtins of oracle_math.h.
#ifdef _OPENMP
#include <omp.h>
#endif

#include "cl_shim.hpp"  // after every system header: it defines OpenCL keywords as macros

#ifndef REF_KERNEL_SOURCE
#error "REF_KERNEL_SOURCE must name the (rewritten) reference kernel stream, see oracle/Makefile"
#endif

namespace refcl {
thread_local size_t g_global_id[2];

#include REF_KERNEL_SOURCE

// the records must have the layout the host structs have (include/shape.hpp, material.hpp, tracer.hpp:48-80)
static_assert(sizeof(Material) == 64 && offsetof(Material, color) == 32 && offsetof(Material, emission) == 48, "Material");
static_assert(sizeof(Sphere) == 32 && offsetof(Sphere, radius) == 16, "Sphere");
static_assert(sizeof(Plane) == 32 && sizeof(Vertex) == 32 && sizeof(Triangle) == 96, "Plane/Vertex/Triangle");
static_assert(sizeof(Model) == 112 && offsetof(Model, bounding_min) == 16 && offsetof(Model, transform) == 48, "Model");
static_assert(sizeof(Shape) == 128 && offsetof(Shape, shape) == 16, "Shape");
static_assert(sizeof(RenderData) == 112 && offsetof(RenderData, show_normals) == 24 &&
                  offsetof(RenderData, camera_to_world) == 32 && offsetof(RenderData, time) == 96, "RenderData");
static_assert(sizeof(SceneData) == 96 && offsetof(SceneData, horizon_color) == 16 && offsetof(SceneData, sun_direction) == 80,
              "SceneData");
}  // namespace refcl

static int pick_threads(int threads) {
	if (threads > 0) return threads;
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

extern "C" {

// One launch of kernel `render` over the work-items [x0,x1) x [y0,y1) of the width x height NDRange.
// Rows with (y / band_h) % band_n != band_i are skipped when band_n > 1 (same convention as oracle_render).
void ref_render(const void *render_data, const void *scene_data, float *canvas, const void *shapes,
                const void *triangles, const void *materials, const float *sky, int sky_w, int sky_h, int x0, int y0,
                int x1, int y1, int band_h, int band_i, int band_n, int threads) {
	using namespace refcl;
	const RenderData data = *static_cast<const RenderData *>(render_data);
	const SceneData scene = *static_cast<const SceneData *>(scene_data);
	const image2d_desc image = {sky, sky_w, sky_h};
	threads = pick_threads(threads);
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
	for (int gy = y0; gy < y1; gy++) {
		if (band_n > 1 && (gy / band_h) % band_n != band_i) continue;
		for (int gx = x0; gx < x1; gx++) {
			g_global_id[0] = (size_t)gx;
			g_global_id[1] = (size_t)gy;
			render(data, scene, reinterpret_cast<float3 *>(canvas), static_cast<const Shape *>(shapes),
			       static_cast<const Triangle *>(triangles), static_cast<const Material *>(materials), &image, 0);
		}
	}
}

// Kernel `average` over n work-items.
void ref_average(uint32_t num_steps, const float *canvas, uint8_t *output, size_t n) {
	using namespace refcl;
	for (size_t id = 0; id < n; id++) {
		g_global_id[0] = id;
		g_global_id[1] = 0;
		average(num_steps, reinterpret_cast<const float3 *>(canvas), reinterpret_cast<uchar4 *>(output));
	}
}

int ref_max_threads(void) { return pick_threads(0); }

}  // extern "C"
