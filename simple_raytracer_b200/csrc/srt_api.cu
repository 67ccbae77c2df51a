// srt_api.cu -- host side of the C ABI declared in include/srt.h.
//
// Replaces the boost.compute layer of reference src/tracer.cpp:1-116 (context / queue / buffers /
// image / arg binding) with a CUDA stream, device buffers and the sm_100a kernels of
// render_kernels.cuh.  No CPU fallback: every entry point either runs on the GPU or fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/srt.h"
#include "bvh_build.hpp"
#include "render_kernels.cuh"

static_assert(sizeof(srt_material) == 64 && offsetof(srt_material, color) == 32 && offsetof(srt_material, emission) == 48, "material");
static_assert(sizeof(srt_sphere) == 32 && offsetof(srt_sphere, radius) == 16, "sphere");
static_assert(sizeof(srt_plane) == 32 && offsetof(srt_plane, normal) == 16, "plane");
static_assert(sizeof(srt_triangle) == 96, "triangle");
static_assert(sizeof(srt_model) == 112 && offsetof(srt_model, bounding_min) == 16 && offsetof(srt_model, bounding_max) == 32 &&
              offsetof(srt_model, transform) == 48, "model");
static_assert(sizeof(srt_shape) == 128 && offsetof(srt_shape, shape) == 16, "shape");
static_assert(sizeof(srt_render_data) == 112 && offsetof(srt_render_data, show_normals) == 24 &&
              offsetof(srt_render_data, camera_to_world) == 32 && offsetof(srt_render_data, time) == 96 &&
              offsetof(srt_render_data, tick) == 100, "render_data");
static_assert(sizeof(srt_scene_data) == 96 && offsetof(srt_scene_data, horizon_color) == 16 &&
              offsetof(srt_scene_data, sun_color) == 64 && offsetof(srt_scene_data, sun_direction) == 80, "scene_data");
static_assert(sizeof(srt_counters) == sizeof(srt::Counters), "counters");

namespace {

std::string g_create_error;

template <typename T>
struct DevBuf {  // grow-only device buffer (rebuild_if_too_small, reference tracer.cpp:5-9)
	T *ptr = nullptr;
	size_t cap = 0;
	cudaError_t reserve(size_t n) {
		if (n <= cap) return cudaSuccess;
		if (ptr) cudaFree(ptr);
		ptr = nullptr;
		cap = 0;
		cudaError_t e = cudaMalloc(&ptr, std::max<size_t>(n, 1) * sizeof(T));
		if (e == cudaSuccess) cap = n;
		return e;
	}
	void release() {
		if (ptr) cudaFree(ptr);
		ptr = nullptr;
		cap = 0;
	}
};

}  // namespace

struct srt_tracer {
	int device = 0;
	int width = 0, height = 0;
	int sm_count = 0;
	cudaStream_t stream = nullptr;
	std::string error;

	float4 *canvas = nullptr;  // float3 with 16-byte stride (tracer.cpp:39)
	uchar4 *output = nullptr;  // ARGB8 (tracer.cpp:40)
	uint8_t *pinned_out = nullptr;   // staging buffer of every read-back into memory the library does not own
	void *registered_out = nullptr;  // srt_pin_output: a caller buffer page-locked at the CALLER's request
	uint8_t *registered_dev = nullptr;  // the same memory as the device sees it (null: not mapped)
	size_t registered_bytes = 0;
	cudaEvent_t chunk_ev[4] = {nullptr, nullptr, nullptr, nullptr};  // read-back pipeline (read_back)
	float4 *sky = nullptr;
	int sky_w = 0, sky_h = 0;
	unsigned long long *cursor = nullptr;  // 64 bits: lanes may step past the last item without ever wrapping
	srt::Counters *counters = nullptr;

	DevBuf<int4> shape_hdr;
	DevBuf<float4> shape_a, shape_b, model_xf, materials;
	DevBuf<float4> tri_aos, tri_hot, tri_n;
	DevBuf<float2> tri_flt;  // 5 per SoA triangle (40 B filter records) + 16 B of padding
	DevBuf<float4> tri_uv;   // 5 per SoA triangle (80 B two-strip filter records)
	DevBuf<float> model_k;   // per shape slot
	DevBuf<float4> scratch;  // one float4 per (pixel, sample) of a launch
	DevBuf<srt::ModelSpan> spans;
	// optional BVH (srt_set_accel): a labelled extension outside the parity path
	int accel = SRT_ACCEL_NONE;
	bool bvh_ready = false;
	int bvh_depth = 0;
	DevBuf<float4> bvh_nodes;
	DevBuf<int> bvh_order, bvh_root;
	std::vector<int4> host_hdr;  // the shape headers of the uploaded scene (the BVH builder walks the models)
	size_t n_shapes = 0, n_materials = 0, n_soa_tris = 0;
	bool has_models = false, has_big_models = false;
	srt_scene_data scene_data{};
	bool have_scene = false;

	int band_h = 1, band_i = 0, band_n = 1;
	int uv_max_tris = srt::UV_MAX_TRIS;  // srt_set_sweep_filter
	int schedule = SRT_SCHEDULE_AUTO;    // srt_set_schedule
	int render_grid[2][5][2] = {};  // [counted][mode][wavefront]
	srt::ShapeTable shape_table{};  // the first CONST_SHAPES shape records, passed as a kernel parameter

	int frame_pipeline = SRT_FRAME_AUTO;  // srt_set_frame_pipeline
	// set for the duration of one srt_render_frame into the pinned caller vector: the launch's epilogue is
	// frame_epilogue_kernel (accumulate + average + store to the host) instead of accumulate_kernel
	uchar4 *epilogue_host = nullptr;
	uint32_t epilogue_steps = 0;

	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing;  // kernel launches since last query
	uint64_t timed_launches = 0;                               // reference launches they cover (batches count each)
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> event_pool;
};

namespace {

int fail(srt_tracer *t, int code, const char *fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	if (t) t->error = buf;
	else g_create_error = buf;
	return code;
}

#define SRT_CUDA(t, call)                                                                             \
	do {                                                                                              \
		cudaError_t e__ = (call);                                                                     \
		if (e__ != cudaSuccess)                                                                       \
			return fail((t), SRT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
	} while (0)

#define SRT_BIND(t)                        \
	do {                                   \
		if (!(t)) return SRT_ERR_INVALID;  \
		SRT_CUDA((t), cudaSetDevice((t)->device)); \
	} while (0)

srt::DevScene dev_scene(const srt_tracer *t) {
	srt::DevScene s{};
	s.num_shapes = (int)t->n_shapes;
	s.has_models = t->has_models ? 1 : 0;
	s.shape_hdr = t->shape_hdr.ptr;
	s.shape_a = t->shape_a.ptr;
	s.shape_b = t->shape_b.ptr;
	s.tri_hot = t->tri_hot.ptr;
	s.tri_flt = t->tri_flt.ptr;
	s.tri_uv = t->tri_uv.ptr;
	s.model_k = t->model_k.ptr;
	s.tri_n = t->tri_n.ptr;
	s.model_xf = t->model_xf.ptr;
	s.materials = t->materials.ptr;
	s.sky = t->sky;
	const bool bvh = t->accel == SRT_ACCEL_BVH && t->bvh_ready;
	s.bvh_nodes = bvh ? t->bvh_nodes.ptr : nullptr;
	s.bvh_order = bvh ? t->bvh_order.ptr : nullptr;
	s.bvh_root = bvh ? t->bvh_root.ptr : nullptr;
	s.sky_w = t->sky_w;
	s.sky_h = t->sky_h;
	const srt_scene_data &sd = t->scene_data;
	s.sun_focus = sd.sun_focus;
	s.sun_intensity = sd.sun_intensity;
	s.sun_color[0] = sd.sun_color.x, s.sun_color[1] = sd.sun_color.y, s.sun_color[2] = sd.sun_color.z;
	s.sun_dir[0] = sd.sun_direction.x, s.sun_dir[1] = sd.sun_direction.y, s.sun_dir[2] = sd.sun_direction.z;
	return s;
}

int make_params(srt_tracer *t, const srt_render_data *rd, srt::RenderParams &p) {
	if (!rd) return fail(t, SRT_ERR_INVALID, "render data is null");
	if (rd->width != t->width || rd->height != t->height)
		return fail(t, SRT_ERR_INVALID, "RenderData is %dx%d but the tracer was created %dx%d", rd->width, rd->height,
		            t->width, t->height);
	if (rd->num_samples < 1 || rd->num_bounces < 0)
		return fail(t, SRT_ERR_INVALID, "num_samples must be >= 1 and num_bounces >= 0");
	if (!t->have_scene) return fail(t, SRT_ERR_INVALID, "no scene uploaded");
	p.width = rd->width;
	p.height = rd->height;
	p.num_samples = rd->num_samples;
	p.num_bounces = rd->num_bounces;
	p.aspect_ratio = rd->aspect_ratio;
	p.fov_scale = rd->fov_scale;
	p.show_normals = rd->show_normals ? 1 : 0;
	memcpy(p.c2w, rd->camera_to_world, sizeof p.c2w);
	p.num_launches = 1;
	p.times[0] = rd->time;
	p.inv_ns = (rd->num_samples & (rd->num_samples - 1)) == 0 ? 1.0f / (float)rd->num_samples : 0.0f;
	p.uv_max_tris = t->uv_max_tris;
	p.band_h = t->band_h;
	p.band_i = t->band_i;
	p.band_n = t->band_n;
	int rows = rd->height;
	if (t->band_n > 1) {
		rows = 0;
		for (int y = 0; y < rd->height; ++y)
			if ((y / t->band_h) % t->band_n == t->band_i) ++rows;
	}
	p.my_rows = rows;
	p.total_pixels = (unsigned int)rows * (unsigned int)rd->width;
	if ((unsigned long long)p.total_pixels * (unsigned long long)rd->num_samples > srt::MAX_ITEMS)
		return fail(t, SRT_ERR_INVALID, "width*height*num_samples exceeds %llu work items per launch", srt::MAX_ITEMS);
	p.items_per_launch = p.total_pixels * (unsigned int)rd->num_samples;
	p.total_items = p.items_per_launch;
#if SRT_FASTDIV
	const uint32_t divisors[3] = {p.items_per_launch, (uint32_t)rd->num_samples, (uint32_t)rd->width};
	for (int k = 0; k < 3; ++k) {
		const srt::FastDiv fd = srt::fast_div_make(divisors[k]);
		p.dv_m[k] = fd.m, p.dv_s[k] = fd.s;
	}
#endif
	return SRT_OK;
}

template <bool COUNT, int MODE, bool WF>
int launch_render_impl(srt_tracer *t, const srt::RenderParams &p);

// The wavefront schedule queues a hit until 32 are there to shade: worth it once a launch keeps every thread busy for
// many items (+1.6 % on BASELINE config 2), not for a launch of a few items per thread (config 1: -6 %).  Its hit record
// keeps the bounce count in 8 bits and the shape index in 24; anything else runs the plain schedule.
template <bool COUNT, int MODE>
int launch_render_impl(srt_tracer *t, const srt::RenderParams &p) {
	const bool fits = p.num_bounces <= 256 && t->n_shapes < ((size_t)1 << 24);
	const bool wf = SRT_WAVEFRONT && MODE != srt::MODE_BIG_MODELS && fits && t->schedule != SRT_SCHEDULE_PLAIN &&
	                (p.total_items >= (4u << 20) || t->schedule == SRT_SCHEDULE_WAVEFRONT);
	if (MODE != srt::MODE_BIG_MODELS && wf) return launch_render_impl<COUNT, MODE, MODE != srt::MODE_BIG_MODELS>(t, p);
	return launch_render_impl<COUNT, MODE, false>(t, p);
}

template <bool COUNT, int MODE, bool WF>
int launch_render_impl(srt_tracer *t, const srt::RenderParams &p) {
	auto kernel = srt::render_kernel<COUNT, MODE, WF>;
	constexpr bool MODELS = MODE != srt::MODE_ANALYTIC && MODE != srt::MODE_ANALYTIC_CONST;
	const int smem = MODE == srt::MODE_BIG_MODELS ? srt::RENDER_SMEM_BYTES + srt::BIG_SKYQ_BYTES
	                 : WF                          ? srt::wavefront_smem_bytes(MODELS)
	                                               : srt::QUEUE_SMEM_BYTES;
	int &grid = t->render_grid[COUNT ? 1 : 0][MODE][WF ? 1 : 0];
	if (grid == 0) {
		int per_sm = 0;
		if (smem) SRT_CUDA(t, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
		SRT_CUDA(t, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, srt::RENDER_THREADS, smem));
		grid = std::max(per_sm, 1) * t->sm_count;  // persistent: one full wave, work pulled from the cursor
	}
	SRT_CUDA(t, t->scratch.reserve(p.total_items));
	SRT_CUDA(t, cudaMemsetAsync(t->cursor, 0, sizeof(unsigned long long), t->stream));
	std::pair<cudaEvent_t, cudaEvent_t> ev;
	if (!t->event_pool.empty()) {
		ev = t->event_pool.back();
		t->event_pool.pop_back();
	} else {
		SRT_CUDA(t, cudaEventCreate(&ev.first));
		SRT_CUDA(t, cudaEventCreate(&ev.second));
	}
	const srt::DevScene sc = dev_scene(t);
	cudaError_t le = cudaEventRecord(ev.first, t->stream);
	kernel<<<grid, srt::RENDER_THREADS, smem, t->stream>>>(p, sc, t->shape_table, t->scratch.ptr, t->cursor, t->counters);
	if (t->epilogue_host)
		srt::frame_epilogue_kernel<<<((p.total_pixels + 3) / 4 + 255) / 256, 256, 0, t->stream>>>(p, t->scratch.ptr, t->canvas, t->output,
		                                                                                        t->epilogue_host, t->epilogue_steps);
	else
		srt::accumulate_kernel<<<(p.total_pixels + 255) / 256, 256, 0, t->stream>>>(p, t->scratch.ptr, t->canvas);
	if (le == cudaSuccess) le = cudaEventRecord(ev.second, t->stream);
	if (le == cudaSuccess) le = cudaGetLastError();
	if (le != cudaSuccess) {
		t->event_pool.push_back(ev);  // the pair goes back to the pool on every path
		return fail(t, SRT_ERR_CUDA, "render launch failed: %s", cudaGetErrorString(le));
	}
	if (t->timing.size() >= 4096) {  // nobody is reading the timings: recycle
		for (auto &e : t->timing) t->event_pool.push_back(e);
		t->timing.clear();
		t->timed_launches = 0;
	}
	t->timing.push_back(ev);
	t->timed_launches += (uint64_t)p.num_launches;
	return SRT_OK;
}

template <bool COUNT>
int launch_params(srt_tracer *t, const srt::RenderParams &p) {
	if (!t->has_models && t->n_shapes <= (size_t)srt::CONST_SHAPES) return launch_render_impl<COUNT, srt::MODE_ANALYTIC_CONST>(t, p);
	if (!t->has_models) return launch_render_impl<COUNT, srt::MODE_ANALYTIC>(t, p);
	if (t->accel == SRT_ACCEL_BVH && t->bvh_ready && t->has_big_models) return launch_render_impl<COUNT, srt::MODE_BVH>(t, p);
	if (!t->has_big_models) return launch_render_impl<COUNT, srt::MODE_SMALL_MODELS>(t, p);
	return launch_render_impl<COUNT, srt::MODE_BIG_MODELS>(t, p);
}

template <bool COUNT>
int launch_render(srt_tracer *t, const srt_render_data *rd) {
	srt::RenderParams p{};
	if (int rc = make_params(t, rd, p)) return rc;
	if (p.num_bounces == 0 || p.total_items == 0) return SRT_OK;  // render.cl:403: zero bounces add zero radiance
	return launch_params<COUNT>(t, p);
}

// (Re)build the optional hierarchy over the uploaded scene: one per model of more than INLINE_MODEL_TRIS triangles,
// over the world-space triangles exactly as the kernel intersects them (tri_hot, read back from the device).
int build_bvh(srt_tracer *t) {
	t->bvh_ready = false;
	const size_t n_shapes = t->host_hdr.size();
	std::vector<int> roots(n_shapes, -1);
	std::vector<srt_bvh::Node> nodes;
	std::vector<int32_t> order;
	int depth = 0;
	try {
		std::vector<float> hot(12 * t->n_soa_tris);
		if (t->n_soa_tris)
			SRT_CUDA(t, cudaMemcpy(hot.data(), t->tri_hot.ptr, hot.size() * sizeof(float), cudaMemcpyDeviceToHost));
		for (size_t i = 0; i < n_shapes; ++i) {
			const int4 h = t->host_hdr[i];
			if (h.x != SRT_SHAPE_MODEL || h.w <= srt::INLINE_MODEL_TRIS) continue;
			roots[i] = (int)nodes.size();
			depth = std::max(depth, srt_bvh::build(hot.data(), h.z, h.w, nodes, order));
		}
	} catch (const std::exception &) {
		return fail(t, SRT_ERR_INVALID, "out of host memory while building the BVH");
	}
	if (depth > srt::BVH_STACK) return fail(t, SRT_ERR_INVALID, "BVH depth %d exceeds the traversal stack (%d)", depth, srt::BVH_STACK);
	SRT_CUDA(t, t->bvh_nodes.reserve(4 * nodes.size()));
	SRT_CUDA(t, t->bvh_order.reserve(order.size()));
	SRT_CUDA(t, t->bvh_root.reserve(n_shapes));
	if (!nodes.empty()) SRT_CUDA(t, cudaMemcpy(t->bvh_nodes.ptr, nodes.data(), nodes.size() * sizeof(srt_bvh::Node), cudaMemcpyHostToDevice));
	if (!order.empty()) SRT_CUDA(t, cudaMemcpy(t->bvh_order.ptr, order.data(), order.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
	if (n_shapes) SRT_CUDA(t, cudaMemcpy(t->bvh_root.ptr, roots.data(), n_shapes * sizeof(int), cudaMemcpyHostToDevice));
	t->bvh_depth = depth;
	t->bvh_ready = true;
	return SRT_OK;
}

// true when two launches differ in nothing the kernel reads except `time` (`tick` is unused, render.cl:91)
bool same_but_time(const srt_render_data &a, const srt_render_data &b) {
	return a.width == b.width && a.height == b.height && a.num_samples == b.num_samples && a.num_bounces == b.num_bounces &&
	       memcmp(&a.aspect_ratio, &b.aspect_ratio, 4) == 0 && memcmp(&a.fov_scale, &b.fov_scale, 4) == 0 &&
	       (a.show_normals != 0) == (b.show_normals != 0) && memcmp(a.camera_to_world, b.camera_to_world, 64) == 0;
}

}  // namespace

extern "C" {

int srt_abi_version(void) { return SRT_ABI_VERSION; }

const char *srt_last_error(const srt_tracer *t) { return t ? t->error.c_str() : g_create_error.c_str(); }

int srt_create(int width, int height, const float *skybox_rgba, int sky_w, int sky_h, int device, srt_tracer **out) {
	if (!out) return fail(nullptr, SRT_ERR_INVALID, "out is null");
	*out = nullptr;
	if (width <= 0 || height <= 0 || (long long)width * height > 0x7fffffffLL)
		return fail(nullptr, SRT_ERR_INVALID, "bad image size %dx%d", width, height);
	if (!skybox_rgba || sky_w <= 0 || sky_h <= 0) return fail(nullptr, SRT_ERR_INVALID, "sky box is required (reference tracer.cpp:42-52)");
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
		cudaGetLastError();
		return fail(nullptr, SRT_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
	}
	if (device < 0) {
		if (cudaGetDevice(&device) != cudaSuccess) device = 0;
	}
	if (device >= ndev) return fail(nullptr, SRT_ERR_INVALID, "device %d out of range (%d devices)", device, ndev);
	srt_tracer *t = new (std::nothrow) srt_tracer();
	if (!t) return fail(nullptr, SRT_ERR_INVALID, "out of host memory");
	t->device = device;
	t->width = width;
	t->height = height;
	t->sky_w = sky_w;
	t->sky_h = sky_h;
#define CREATE_CUDA(call)                                                                               \
	do {                                                                                                \
		cudaError_t e__ = (call);                                                                       \
		if (e__ != cudaSuccess) {                                                                       \
			fail(nullptr, SRT_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));               \
			srt_destroy(t);                                                                             \
			return SRT_ERR_CUDA;                                                                        \
		}                                                                                               \
	} while (0)
	CREATE_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop{};
	CREATE_CUDA(cudaGetDeviceProperties(&prop, device));
	t->sm_count = prop.multiProcessorCount;
	CREATE_CUDA(cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking));
	const size_t n = (size_t)width * height;
	CREATE_CUDA(cudaMalloc(&t->canvas, n * sizeof(float4)));
	CREATE_CUDA(cudaMalloc(&t->output, n * sizeof(uchar4)));
	CREATE_CUDA(cudaMallocHost(&t->pinned_out, n * 4));
	CREATE_CUDA(cudaMalloc(&t->cursor, sizeof(unsigned long long)));
	for (auto &e : t->chunk_ev) CREATE_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	CREATE_CUDA(cudaMalloc(&t->counters, sizeof(srt::Counters)));
	CREATE_CUDA(cudaMalloc(&t->sky, (size_t)sky_w * sky_h * sizeof(float4)));
	CREATE_CUDA(cudaMemcpyAsync(t->sky, skybox_rgba, (size_t)sky_w * sky_h * sizeof(float4), cudaMemcpyHostToDevice, t->stream));
	CREATE_CUDA(cudaMemsetAsync(t->canvas, 0, n * sizeof(float4), t->stream));
	CREATE_CUDA(cudaMemsetAsync(t->counters, 0, sizeof(srt::Counters), t->stream));
	CREATE_CUDA(cudaStreamSynchronize(t->stream));
#undef CREATE_CUDA
	*out = t;
	return SRT_OK;
}

int srt_destroy(srt_tracer *t) {
	if (!t) return SRT_OK;
	cudaSetDevice(t->device);
	if (t->stream) cudaStreamSynchronize(t->stream);
	for (auto &e : t->timing) t->event_pool.push_back(e);
	for (auto &e : t->event_pool) {
		cudaEventDestroy(e.first);
		cudaEventDestroy(e.second);
	}
	if (t->registered_out) cudaHostUnregister(t->registered_out);
	for (auto &e : t->chunk_ev)
		if (e) cudaEventDestroy(e);
	cudaGetLastError();
	cudaFree(t->canvas);
	cudaFree(t->output);
	cudaFreeHost(t->pinned_out);
	cudaFree(t->cursor);
	cudaFree(t->counters);
	cudaFree(t->sky);
	t->shape_hdr.release(), t->shape_a.release(), t->shape_b.release(), t->model_xf.release(), t->materials.release();
	t->bvh_nodes.release(), t->bvh_order.release(), t->bvh_root.release();
	t->tri_aos.release(), t->tri_hot.release(), t->tri_flt.release(), t->tri_uv.release(), t->model_k.release(), t->tri_n.release(), t->spans.release(), t->scratch.release();
	if (t->stream) cudaStreamDestroy(t->stream);
	delete t;
	return SRT_OK;
}

int srt_upload_scene(srt_tracer *t, const srt_shape *shapes, size_t n_shapes, const srt_triangle *triangles,
                     size_t n_triangles, const srt_material *materials, size_t n_materials,
                     const srt_scene_data *scene_data) {
	SRT_BIND(t);
	if ((n_shapes && !shapes) || (n_triangles && !triangles) || (n_materials && !materials) || !scene_data)
		return fail(t, SRT_ERR_INVALID, "null array with non-zero count");
	if (n_shapes > 0x7fffffffu || n_triangles > 0x7fffffffu) return fail(t, SRT_ERR_INVALID, "scene too large");

	// host-side repack of the (small) shape list into the SoA device records; validation of what the
	// reference leaves unchecked (render.cl:412 material index, :325 triangle range)
	std::vector<int4> hdr(n_shapes);
	std::vector<float4> a(n_shapes), b(n_shapes), xf(4 * n_shapes);
	std::vector<srt::ModelSpan> spans;
	size_t soa = 0;
	for (size_t i = 0; i < n_shapes; ++i) {
		const srt_shape &s = shapes[i];
		if (s.material < 0 || (size_t)s.material >= n_materials)
			return fail(t, SRT_ERR_INVALID, "shape %zu: material %d out of range (%zu materials)", i, s.material, n_materials);
		hdr[i] = make_int4(s.type, s.material, 0, 0);
		a[i] = b[i] = make_float4(0, 0, 0, 0);
		for (int c = 0; c < 4; ++c) xf[4 * i + c] = make_float4(0, 0, 0, 0);
		if (s.type == SRT_SHAPE_SPHERE) {
			a[i] = make_float4(s.shape.sphere.position.x, s.shape.sphere.position.y, s.shape.sphere.position.z, s.shape.sphere.radius);
		} else if (s.type == SRT_SHAPE_PLANE) {
			a[i] = make_float4(s.shape.plane.position.x, s.shape.plane.position.y, s.shape.plane.position.z, 0);
			b[i] = make_float4(s.shape.plane.normal.x, s.shape.plane.normal.y, s.shape.plane.normal.z, 0);
		} else if (s.type == SRT_SHAPE_MODEL) {
			const srt_model &m = s.shape.model;
			if ((size_t)m.triangle_index + m.num_triangles > n_triangles)
				return fail(t, SRT_ERR_INVALID, "shape %zu: triangles [%u,+%u) exceed the %zu uploaded", i, m.triangle_index, m.num_triangles, n_triangles);
			a[i] = make_float4(m.bounding_min.x, m.bounding_min.y, m.bounding_min.z, 0);
			b[i] = make_float4(m.bounding_max.x, m.bounding_max.y, m.bounding_max.z, 0);
			for (int c = 0; c < 4; ++c) xf[4 * i + c] = make_float4(m.transform[c].x, m.transform[c].y, m.transform[c].z, m.transform[c].w);
			if (m.num_triangles > (unsigned)srt::MAX_SWEEP_TRIS)
				return fail(t, SRT_ERR_INVALID, "shape %zu: %u triangles in one model (limit %d)", i, m.num_triangles, srt::MAX_SWEEP_TRIS);
			soa = (soa + 1) & ~(size_t)1;  // even start: the 40-byte filter records of a model begin 16-byte aligned
			hdr[i].z = (int)soa;
			hdr[i].w = (int)m.num_triangles;
			if (m.num_triangles) spans.push_back(srt::ModelSpan{(int)i, (int)m.triangle_index, (int)soa, (int)m.num_triangles});
			soa += m.num_triangles;
			if (soa > 0x7fffffffu) return fail(t, SRT_ERR_INVALID, "too many model triangles");
		} else {
			return fail(t, SRT_ERR_INVALID, "shape %zu: unknown type %d", i, s.type);
		}
	}

	// From here on the device buffers are being replaced: until the last copy has landed the handle has NO scene, so
	// a failure half-way leaves it refusing to render ("no scene uploaded") instead of launching on freed buffers.
	t->have_scene = false;
	t->n_shapes = 0;
	SRT_CUDA(t, t->shape_hdr.reserve(n_shapes));
	SRT_CUDA(t, t->shape_a.reserve(n_shapes));
	SRT_CUDA(t, t->shape_b.reserve(n_shapes));
	SRT_CUDA(t, t->model_xf.reserve(4 * n_shapes));
	SRT_CUDA(t, t->materials.reserve(4 * n_materials));
	SRT_CUDA(t, t->tri_aos.reserve(6 * n_triangles));
	SRT_CUDA(t, t->tri_hot.reserve(3 * soa));
	SRT_CUDA(t, t->tri_flt.reserve(5 * soa + 2));
	SRT_CUDA(t, t->tri_uv.reserve(5 * soa + 1));
	SRT_CUDA(t, t->model_k.reserve(n_shapes));
	SRT_CUDA(t, t->tri_n.reserve(3 * soa));
	SRT_CUDA(t, t->spans.reserve(spans.size()));
	cudaStream_t st = t->stream;
	// the host vectors above die at return, and the caller may reuse its arrays: these copies are
	// from pageable memory, which cudaMemcpyAsync stages before returning
	if (n_shapes) {
		SRT_CUDA(t, cudaMemcpyAsync(t->shape_hdr.ptr, hdr.data(), n_shapes * sizeof(int4), cudaMemcpyHostToDevice, st));
		SRT_CUDA(t, cudaMemcpyAsync(t->shape_a.ptr, a.data(), n_shapes * sizeof(float4), cudaMemcpyHostToDevice, st));
		SRT_CUDA(t, cudaMemcpyAsync(t->shape_b.ptr, b.data(), n_shapes * sizeof(float4), cudaMemcpyHostToDevice, st));
		SRT_CUDA(t, cudaMemcpyAsync(t->model_xf.ptr, xf.data(), 4 * n_shapes * sizeof(float4), cudaMemcpyHostToDevice, st));
	}
	if (n_materials) {
		SRT_CUDA(t, cudaMemcpyAsync(t->materials.ptr, materials, n_materials * sizeof(srt_material), cudaMemcpyHostToDevice, st));
		srt::prepare_materials_kernel<<<((int)n_materials + 127) / 128, 128, 0, st>>>(t->materials.ptr, (int)n_materials);
		SRT_CUDA(t, cudaGetLastError());
	}
	if (n_triangles)
		SRT_CUDA(t, cudaMemcpyAsync(t->tri_aos.ptr, triangles, n_triangles * sizeof(srt_triangle), cudaMemcpyHostToDevice, st));
	if (!spans.empty()) {
		SRT_CUDA(t, cudaMemcpyAsync(t->spans.ptr, spans.data(), spans.size() * sizeof(srt::ModelSpan), cudaMemcpyHostToDevice, st));
		const int total = (int)soa;
		SRT_CUDA(t, cudaMemsetAsync(t->model_k.ptr, 0, n_shapes * sizeof(float), st));
		SRT_CUDA(t, cudaMemsetAsync(t->tri_flt.ptr, 0, (5 * soa + 2) * sizeof(float2), st));  // alignment gaps and the tail pad
		SRT_CUDA(t, cudaMemsetAsync(t->tri_uv.ptr, 0, (5 * soa + 1) * sizeof(float4), st));
		srt::prepare_triangles_kernel<<<(total + 255) / 256, 256, 0, st>>>(t->tri_aos.ptr, t->spans.ptr, (int)spans.size(), total,
		                                                                  t->model_xf.ptr, t->tri_hot.ptr, t->tri_flt.ptr,
		                                                                  t->tri_uv.ptr, t->model_k.ptr, t->tri_n.ptr);
		SRT_CUDA(t, cudaGetLastError());
	}
	SRT_CUDA(t, cudaStreamSynchronize(st));  // copy-in semantics, like the blocking writes of tracer.cpp:76-86
	t->n_shapes = n_shapes;
	t->n_materials = n_materials;
	t->n_soa_tris = soa;
	t->has_big_models = std::any_of(hdr.begin(), hdr.end(), [](const int4 &h) { return h.x == SRT_SHAPE_MODEL && h.w > srt::INLINE_MODEL_TRIS; });
	t->has_models = soa > 0 || std::any_of(hdr.begin(), hdr.end(), [](const int4 &h) { return h.x == SRT_SHAPE_MODEL; });
	t->scene_data = *scene_data;
	t->scene_data.num_shapes = (int)n_shapes;  // tracer.cpp:94
	t->host_hdr = hdr;
	t->shape_table = srt::ShapeTable{};
	for (size_t i = 0; i < std::min<size_t>(n_shapes, srt::CONST_SHAPES); ++i)
		t->shape_table.hdr[i] = hdr[i], t->shape_table.a[i] = a[i], t->shape_table.b[i] = b[i];
#if SRT_PAIR_SCAN
	if (!t->has_models && n_shapes <= (size_t)srt::CONST_SHAPES) {  // the ops of scan_pairs: neighbours of one type, two at a time
		srt::ShapeTable &tab = t->shape_table;
		int n_ops = 0;
		for (size_t i = 0; i < n_shapes; ++n_ops) {
			const bool pair = i + 1 < n_shapes && hdr[i + 1].x == hdr[i].x;
			const size_t j = pair ? i + 1 : i;  // a single shape: the B half repeats A and is ignored
			tab.op_hdr[n_ops] = make_int4(hdr[i].x, (int)i, pair ? (int)j : -1, 0);
			if (hdr[i].x == SRT_SHAPE_SPHERE) {
				// -r*r: ONE correctly rounded single-precision product, the operation cfma_(-r, r, .) starts with (this
				// translation unit's host code is compiled with -ffp-contract=off, x86-64 SSE arithmetic)
				const volatile float ra = a[i].w, rb = a[j].w;
				const volatile float nra = -ra * ra, nrb = -rb * rb;
				tab.op[n_ops][0] = make_float4(a[i].x, a[j].x, a[i].y, a[j].y);
				tab.op[n_ops][1] = make_float4(a[i].z, a[j].z, nra, nrb);
				tab.op[n_ops][2] = make_float4(0, 0, 0, 0);
			} else {
				tab.op[n_ops][0] = make_float4(a[i].x, a[j].x, a[i].y, a[j].y);
				tab.op[n_ops][1] = make_float4(a[i].z, a[j].z, b[i].x, b[j].x);
				tab.op[n_ops][2] = make_float4(b[i].y, b[j].y, b[i].z, b[j].z);
			}
			i = j + 1;
		}
		tab.n_ops = n_ops;
	}
#endif
	t->bvh_ready = false;
	t->have_scene = true;
	if (t->accel == SRT_ACCEL_BVH) return build_bvh(t);
	return SRT_OK;
}

int srt_set_accel(srt_tracer *t, int accel) {
	SRT_BIND(t);
	if (accel != SRT_ACCEL_NONE && accel != SRT_ACCEL_BVH) return fail(t, SRT_ERR_INVALID, "unknown acceleration mode %d", accel);
	t->accel = accel;
	if (accel == SRT_ACCEL_BVH && t->have_scene && !t->bvh_ready) {
		SRT_CUDA(t, cudaStreamSynchronize(t->stream));
		return build_bvh(t);
	}
	return SRT_OK;
}

int srt_clear(srt_tracer *t) {
	SRT_BIND(t);
	SRT_CUDA(t, cudaMemsetAsync(t->canvas, 0, (size_t)t->width * t->height * sizeof(float4), t->stream));
	return SRT_OK;
}

int srt_render(srt_tracer *t, const srt_render_data *rd) {
	SRT_BIND(t);
	return launch_render<false>(t, rd);
}

// launches of `p`'s shape one kernel may cover: MAX_BATCH, a 4 GiB budget of per-sample scratch, 2^32 work items
static size_t batch_cap(const srt::RenderParams &p) {
	const size_t scratch_budget = (size_t)4 << 30;
	const size_t fit = std::max<size_t>(1, scratch_budget / ((size_t)p.items_per_launch * sizeof(float4)));
	return std::min<size_t>({(size_t)srt::MAX_BATCH, fit, (size_t)(srt::MAX_ITEMS / p.items_per_launch)});
}

int srt_reserve_batch(srt_tracer *t, const srt_render_data *rd, size_t n) {
	SRT_BIND(t);
	srt::RenderParams p{};
	if (int rc = make_params(t, rd, p)) return rc;
	if (p.total_items == 0 || n == 0) return SRT_OK;
	SRT_CUDA(t, t->scratch.reserve((size_t)p.items_per_launch * std::min(n, batch_cap(p))));
	return SRT_OK;
}

int srt_render_batch(srt_tracer *t, const srt_render_data *rds, size_t n) {
	SRT_BIND(t);
	if (n && !rds) return fail(t, SRT_ERR_INVALID, "render data is null");
	size_t i = 0;
	while (i < n) {
		srt::RenderParams p{};
		if (int rc = make_params(t, &rds[i], p)) return rc;
		size_t run = 1;
		if (p.num_bounces > 0 && p.total_items > 0) {
			const size_t cap = batch_cap(p);
			while (i + run < n && run < cap && same_but_time(rds[i], rds[i + run])) {
				p.times[run] = rds[i + run].time;
				++run;
			}
			p.num_launches = (int)run;
			p.total_items = p.items_per_launch * (unsigned int)run;
			if (int rc = launch_params<false>(t, p)) return rc;
		}
		i += run;
	}
	return SRT_OK;
}

int srt_render_counted(srt_tracer *t, const srt_render_data *rd, srt_counters *counters) {
	SRT_BIND(t);
	if (!counters) return fail(t, SRT_ERR_INVALID, "counters is null");
	SRT_CUDA(t, cudaMemsetAsync(t->counters, 0, sizeof(srt::Counters), t->stream));
	if (int rc = launch_render<true>(t, rd)) return rc;
	srt::Counters c{};
	SRT_CUDA(t, cudaMemcpyAsync(&c, t->counters, sizeof c, cudaMemcpyDeviceToHost, t->stream));
	SRT_CUDA(t, cudaStreamSynchronize(t->stream));
	counters->samples += c.samples;
	counters->bounces += c.bounces;
	counters->tri_tests += c.tri_tests;
	counters->aabb_pass += c.aabb_pass;
	counters->hits += c.hits;
	counters->sky += c.sky;
	return SRT_OK;
}

int srt_resolve_device(srt_tracer *t, uint32_t num_steps) {
	return srt_resolve_device_range(t, num_steps, 0, t ? (size_t)t->width * t->height : 0);
}

int srt_resolve_device_range(srt_tracer *t, uint32_t num_steps, size_t first_pixel, size_t count) {
	SRT_BIND(t);
	const size_t total = (size_t)t->width * t->height;
	if (first_pixel > total || count > total - first_pixel) return fail(t, SRT_ERR_INVALID, "pixel range out of bounds");
	if (count == 0) return SRT_OK;
	const int n = (int)count;
	srt::average_kernel<<<(n + 255) / 256, 256, 0, t->stream>>>(num_steps, t->canvas + first_pixel, t->output + first_pixel, n);
	SRT_CUDA(t, cudaGetLastError());
	return SRT_OK;
}

// The blocking read of the ARGB8 image, tracer.cpp:115.  The library never page-locks memory it does not own: the
// copy goes through the handle's pinned staging buffer, in four chunks so that the host memcpy of chunk k overlaps
// the DMA of chunk k + 1.  A caller who keeps ONE output buffer alive across frames (src/main.cpp:128) can opt in
// to a direct copy with srt_pin_output.
static int read_back(srt_tracer *t, uint8_t *dst) {
	const size_t bytes = (size_t)t->width * t->height * 4;
	const uint8_t *src = reinterpret_cast<const uint8_t *>(t->output);
	if (t->registered_out && dst >= (uint8_t *)t->registered_out &&
	    dst + bytes <= (uint8_t *)t->registered_out + t->registered_bytes) {
		SRT_CUDA(t, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, t->stream));
		SRT_CUDA(t, cudaStreamSynchronize(t->stream));
		return SRT_OK;
	}
	const int chunks = bytes >= (1u << 20) ? 4 : 1;
	const size_t step = ((bytes + chunks - 1) / chunks + 63) & ~(size_t)63;
	for (int k = 0; k < chunks; ++k) {
		const size_t off = std::min(bytes, k * step), len = std::min(bytes - off, step);
		if (len) SRT_CUDA(t, cudaMemcpyAsync(t->pinned_out + off, src + off, len, cudaMemcpyDeviceToHost, t->stream));
		SRT_CUDA(t, cudaEventRecord(t->chunk_ev[k], t->stream));
	}
	for (int k = 0; k < chunks; ++k) {
		const size_t off = std::min(bytes, k * step), len = std::min(bytes - off, step);
		SRT_CUDA(t, cudaEventSynchronize(t->chunk_ev[k]));
		memcpy(dst + off, t->pinned_out + off, len);
	}
	return SRT_OK;
}

int srt_read_output(srt_tracer *t, uint8_t *argb_out) {
	SRT_BIND(t);
	if (!argb_out) return fail(t, SRT_ERR_INVALID, "output is null");
	return read_back(t, argb_out);
}

int srt_resolve(srt_tracer *t, uint32_t num_steps, uint8_t *argb_out) {
	SRT_BIND(t);
	if (!argb_out) return fail(t, SRT_ERR_INVALID, "output is null");
	if (int rc = srt_resolve_device(t, num_steps)) return rc;
	return read_back(t, argb_out);
}

int srt_pin_output(srt_tracer *t, void *buffer, size_t bytes) {
	SRT_BIND(t);
	if (!buffer || bytes == 0) return fail(t, SRT_ERR_INVALID, "srt_pin_output: null buffer");
	if (int rc = srt_unpin_output(t)) return rc;
	cudaError_t e = cudaHostRegister(buffer, bytes, cudaHostRegisterMapped);
	if (e != cudaSuccess) {
		cudaGetLastError();  // not sticky: the staging path keeps working
		return fail(t, SRT_ERR_CUDA, "cudaHostRegister failed: %s", cudaGetErrorString(e));
	}
	t->registered_out = buffer;
	t->registered_bytes = bytes;
	void *dev = nullptr;  // mapped into the device's address space: srt_render_frame's epilogue stores the image into it
	if (cudaHostGetDevicePointer(&dev, buffer, 0) != cudaSuccess) {
		cudaGetLastError();
		dev = nullptr;  // copies still work
	}
	t->registered_dev = static_cast<uint8_t *>(dev);
	return SRT_OK;
}

int srt_unpin_output(srt_tracer *t) {
	SRT_BIND(t);
	if (!t->registered_out) return SRT_OK;
	SRT_CUDA(t, cudaStreamSynchronize(t->stream));
	cudaError_t e = cudaHostUnregister(t->registered_out);
	t->registered_out = nullptr;
	t->registered_dev = nullptr;
	t->registered_bytes = 0;
	if (e != cudaSuccess) {
		cudaGetLastError();
		return fail(t, SRT_ERR_CUDA, "cudaHostUnregister failed: %s (was the buffer freed while pinned?)", cudaGetErrorString(e));
	}
	return SRT_OK;
}

int srt_render_frame(srt_tracer *t, const srt_render_data *rd, uint32_t ticks_stopped, uint8_t *argb_out) {
	SRT_BIND(t);
	if (!argb_out) return fail(t, SRT_ERR_INVALID, "output is null");
	srt::RenderParams p{};
	if (int rc = make_params(t, rd, p)) return rc;
	// launches restricted to row bands (tile sharding) and launches without work (render.cl:403: zero bounces add zero
	// radiance) take the separate steps
	const bool full_frame = t->band_n <= 1 && p.num_bounces > 0 && p.total_items > 0;
	// a full frame into the page-locked caller vector: one epilogue kernel accumulates, resolves and stores the image
	// into the caller's memory directly (no separate `average` launch, no copy-engine transfer after it)
	const bool direct = t->registered_out && argb_out >= (uint8_t *)t->registered_out &&
	                    argb_out + (size_t)t->width * t->height * 4 <= (uint8_t *)t->registered_out + t->registered_bytes;
	if (t->frame_pipeline != SRT_FRAME_SEPARATE && full_frame && direct && t->registered_dev && ((uintptr_t)argb_out & 15) == 0) {
		t->epilogue_host = reinterpret_cast<uchar4 *>(t->registered_dev + (argb_out - (uint8_t *)t->registered_out));
		t->epilogue_steps = ticks_stopped;
		const int rc = launch_params<false>(t, p);
		t->epilogue_host = nullptr;
		if (rc) return rc;
		SRT_CUDA(t, cudaStreamSynchronize(t->stream));
		return SRT_OK;
	}
	if (int rc = srt_render(t, rd)) return rc;
	return srt_resolve(t, ticks_stopped, argb_out);
}

int srt_set_frame_pipeline(srt_tracer *t, int mode) {
	if (!t) return SRT_ERR_INVALID;
	if (mode != SRT_FRAME_AUTO && mode != SRT_FRAME_SEPARATE)
		return fail(t, SRT_ERR_INVALID, "unknown frame pipeline %d", mode);
	t->frame_pipeline = mode;
	return SRT_OK;
}

int srt_set_row_bands(srt_tracer *t, int band_height, int band_index, int band_count) {
	if (!t) return SRT_ERR_INVALID;
	if (band_count <= 1) {
		t->band_h = 1, t->band_i = 0, t->band_n = 1;
		return SRT_OK;
	}
	if (band_height < 1 || band_index < 0 || band_index >= band_count) return fail(t, SRT_ERR_INVALID, "bad row bands");
	t->band_h = band_height, t->band_i = band_index, t->band_n = band_count;
	return SRT_OK;
}

int srt_set_schedule(srt_tracer *t, int schedule) {
	if (!t) return SRT_ERR_INVALID;
	if (schedule != SRT_SCHEDULE_AUTO && schedule != SRT_SCHEDULE_PLAIN && schedule != SRT_SCHEDULE_WAVEFRONT)
		return fail(t, SRT_ERR_INVALID, "unknown schedule %d", schedule);
	t->schedule = schedule;
	return SRT_OK;
}

int srt_set_sweep_filter(srt_tracer *t, int mode) {
	if (!t) return SRT_ERR_INVALID;
	if (mode == SRT_FILTER_AUTO) t->uv_max_tris = srt::UV_MAX_TRIS;
	else if (mode == SRT_FILTER_ONE_STRIP) t->uv_max_tris = -1;
	else if (mode == SRT_FILTER_TWO_STRIP) t->uv_max_tris = 0x7fffffff;
	else return fail(t, SRT_ERR_INVALID, "unknown sweep filter %d", mode);
	return SRT_OK;
}

int srt_read_canvas(srt_tracer *t, float *rgba_out) {
	SRT_BIND(t);
	if (!rgba_out) return fail(t, SRT_ERR_INVALID, "output is null");
	SRT_CUDA(t, cudaMemcpyAsync(rgba_out, t->canvas, (size_t)t->width * t->height * sizeof(float4), cudaMemcpyDeviceToHost, t->stream));
	SRT_CUDA(t, cudaStreamSynchronize(t->stream));
	return SRT_OK;
}

int srt_write_canvas(srt_tracer *t, const float *rgba_in) {
	SRT_BIND(t);
	if (!rgba_in) return fail(t, SRT_ERR_INVALID, "input is null");
	SRT_CUDA(t, cudaMemcpyAsync(t->canvas, rgba_in, (size_t)t->width * t->height * sizeof(float4), cudaMemcpyHostToDevice, t->stream));
	SRT_CUDA(t, cudaStreamSynchronize(t->stream));
	return SRT_OK;
}

int srt_canvas_device_ptr(srt_tracer *t, void **ptr, size_t *bytes) {
	if (!t || !ptr) return SRT_ERR_INVALID;
	*ptr = t->canvas;
	if (bytes) *bytes = (size_t)t->width * t->height * sizeof(float4);
	return SRT_OK;
}

int srt_output_device_ptr(srt_tracer *t, void **ptr, size_t *bytes) {
	if (!t || !ptr) return SRT_ERR_INVALID;
	*ptr = t->output;
	if (bytes) *bytes = (size_t)t->width * t->height * 4;
	return SRT_OK;
}

int srt_stream(srt_tracer *t, void **cuda_stream) {
	if (!t || !cuda_stream) return SRT_ERR_INVALID;
	*cuda_stream = t->stream;
	return SRT_OK;
}

int srt_synchronize(srt_tracer *t) {
	SRT_BIND(t);
	SRT_CUDA(t, cudaStreamSynchronize(t->stream));
	return SRT_OK;
}

int srt_debug_primary(srt_tracer *t, const srt_render_data *rd, int32_t *shape_idx, float *t_out) {
	SRT_BIND(t);
	if (!shape_idx || !t_out) return fail(t, SRT_ERR_INVALID, "output is null");
	srt::RenderParams p{};
	if (int rc = make_params(t, rd, p)) return rc;
	const int n = p.width * p.height;
	int *d_idx = nullptr;
	float *d_t = nullptr;
	SRT_CUDA(t, cudaMalloc(&d_idx, n * sizeof(int)));
	if (cudaMalloc(&d_t, n * sizeof(float)) != cudaSuccess) {
		cudaFree(d_idx);
		return fail(t, SRT_ERR_CUDA, "cudaMalloc failed");
	}
	if (t->accel == SRT_ACCEL_BVH && t->bvh_ready)
		srt::primary_kernel<true><<<(n + 255) / 256, 256, 0, t->stream>>>(p, dev_scene(t), d_idx, d_t);
	else
		srt::primary_kernel<false><<<(n + 255) / 256, 256, 0, t->stream>>>(p, dev_scene(t), d_idx, d_t);
	cudaError_t e = cudaGetLastError();
	if (e == cudaSuccess) e = cudaMemcpyAsync(shape_idx, d_idx, n * sizeof(int), cudaMemcpyDeviceToHost, t->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(t_out, d_t, n * sizeof(float), cudaMemcpyDeviceToHost, t->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(t->stream);
	cudaFree(d_idx);
	cudaFree(d_t);
	if (e != cudaSuccess) return fail(t, SRT_ERR_CUDA, "debug_primary failed: %s", cudaGetErrorString(e));
	return SRT_OK;
}

int srt_debug_math(srt_tracer *t, int op, const float *x, const float *y, float *out, size_t n) {
	SRT_BIND(t);
	if (!x || !y || !out) return fail(t, SRT_ERR_INVALID, "null array");
	if (n == 0) return SRT_OK;
	float *d = nullptr;
	SRT_CUDA(t, cudaMalloc(&d, 3 * n * sizeof(float)));
	cudaError_t e = cudaMemcpyAsync(d, x, n * sizeof(float), cudaMemcpyHostToDevice, t->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(d + n, y, n * sizeof(float), cudaMemcpyHostToDevice, t->stream);
	if (e == cudaSuccess) {
		srt::math_kernel<<<(unsigned)((n + 255) / 256), 256, 0, t->stream>>>(op, d, d + n, d + 2 * n, n);
		e = cudaGetLastError();
	}
	if (e == cudaSuccess) e = cudaMemcpyAsync(out, d + 2 * n, n * sizeof(float), cudaMemcpyDeviceToHost, t->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(t->stream);
	cudaFree(d);
	if (e != cudaSuccess) return fail(t, SRT_ERR_CUDA, "debug_math failed: %s", cudaGetErrorString(e));
	return SRT_OK;
}

int srt_measure_fp32_peak(srt_tracer *t, double *tflops, double *sm_clock_mhz_est) {
	SRT_BIND(t);
	if (!tflops) return fail(t, SRT_ERR_INVALID, "tflops is null");
	float *d = nullptr;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	auto cleanup = [&]() {
		if (e0) cudaEventDestroy(e0);
		if (e1) cudaEventDestroy(e1);
		cudaFree(d);
	};
	if (cudaMalloc(&d, sizeof(float)) != cudaSuccess || cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
		cleanup();
		cudaGetLastError();
		return fail(t, SRT_ERR_CUDA, "peak probe: allocation failed");
	}
	const int blocks = t->sm_count * 8, threads = 256, iters = 1 << 14;
	double best = 0.0;
	for (int rep = 0; rep < 6; ++rep) {
		cudaEventRecord(e0, t->stream);
		srt::fma_peak_kernel<<<blocks, threads, 0, t->stream>>>(d, iters, 1.0000001f, 1e-9f);
		cudaEventRecord(e1, t->stream);
		cudaError_t e = cudaStreamSynchronize(t->stream);
		if (e != cudaSuccess) {
			cleanup();
			return fail(t, SRT_ERR_CUDA, "peak kernel failed: %s", cudaGetErrorString(e));
		}
		float ms = 0.f;
		cudaEventElapsedTime(&ms, e0, e1);
		double fl = 2.0 * 16.0 * (double)iters * (double)blocks * threads;
		if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
	}
	cleanup();
	*tflops = best;
	if (sm_clock_mhz_est) *sm_clock_mhz_est = best * 1e12 / (2.0 * 128.0 * t->sm_count) / 1e6;
	return SRT_OK;
}

int srt_render_time_ms(srt_tracer *t, double *total_ms, uint64_t *launches) {
	SRT_BIND(t);
	SRT_CUDA(t, cudaStreamSynchronize(t->stream));
	double sum = 0.0;
	for (auto &e : t->timing) {
		float ms = 0.f;
		SRT_CUDA(t, cudaEventElapsedTime(&ms, e.first, e.second));
		sum += ms;
		t->event_pool.push_back(e);
	}
	if (total_ms) *total_ms = sum;
	if (launches) *launches = t->timed_launches;
	t->timing.clear();
	t->timed_launches = 0;
	return SRT_OK;
}

}  // extern "C"
