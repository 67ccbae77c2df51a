/*
 * srt.h -- C ABI of the B200 path-tracing library (libsrt_b200.so).
 *
 * This is the drop-in boundary for the ONE hot path of davawen/Simple-Raytracer: the
 * `render` + `average` OpenCL kernels (reference src/render.cl:483-535) as driven by the
 * boost.compute host layer `Tracer` (reference include/tracer.hpp:26-88, src/tracer.cpp:1-116).
 * Every entry point names the reference interface it replaces.  Plain pointers and sizes
 * only; all functions return 0 on success or a non-zero srt_status, with the text available
 * from srt_last_error().  A handle is single-threaded (like Tracer); distinct handles, one
 * per GPU, may be driven from distinct threads.
 *
 * The scene / argument records are byte-identical to the reference's host structs
 * (include/shape.hpp:15-111, include/material.hpp:10-37, include/tracer.hpp:48-80), which in
 * turn mirror the device structs of src/render.cl:17-105 (OpenCL float3 == 16 bytes).
 */
#ifndef SRT_H
#define SRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRT_ABI_VERSION 1

typedef struct { float x, y, z, _w; } srt_float3; /* cl_float3: 16 bytes, 16-aligned; _w is padding */
typedef struct { float x, y, z, w; } srt_float4;

/* reference include/material.hpp:10-37 == src/render.cl:17-27 (64 B) */
typedef struct {
	float smoothness, metallic, specular, emission_strength, transmittance, refraction_index;
	float _pad[2];
	srt_float3 color, emission;
} srt_material;

typedef struct { srt_float3 position; float radius; float _pad[3]; } srt_sphere;  /* shape.hpp:15-20, 32 B */
typedef struct { srt_float3 position, normal; } srt_plane;                        /* shape.hpp:22-27, 32 B */
typedef struct { srt_float3 normal, pos; } srt_vertex;                            /* shape.hpp:30-33 */
typedef struct { srt_vertex vertices[3]; } srt_triangle;                          /* shape.hpp:29-44, 96 B */
typedef struct {                                                                  /* shape.hpp:47-69, 112 B */
	uint32_t triangle_index, num_triangles;
	uint32_t _pad[2];
	srt_float3 bounding_min, bounding_max;
	srt_float4 transform[4]; /* glm::mat4, column-major: transform[c] is column c */
} srt_model;

enum { SRT_SHAPE_SPHERE = 0, SRT_SHAPE_PLANE = 1, SRT_SHAPE_MODEL = 2 };          /* shape.hpp:79-83 */

typedef struct {                                                                  /* shape.hpp:85-111, 128 B */
	int32_t type, material;
	int32_t _pad[2];
	union { srt_sphere sphere; srt_plane plane; srt_model model; } shape;
} srt_shape;

/* Tracer::RenderData, tracer.hpp:48-67 == `RenderData` kernel argument 0, render.cl:79-92 (112 B) */
typedef struct {
	int32_t width, height, num_samples, num_bounces;
	float aspect_ratio, fov_scale;
	uint8_t show_normals;
	uint8_t _pad[7];
	srt_float4 camera_to_world[4]; /* column-major */
	uint32_t time, tick;
	uint32_t _pad2[2];
} srt_render_data;

/* Tracer::SceneData, tracer.hpp:69-80 == `SceneData` kernel argument 1, render.cl:94-105 (96 B) */
typedef struct {
	int32_t num_shapes; /* overwritten with n_shapes by srt_upload_scene, as tracer.cpp:94 does */
	float sun_focus, sun_intensity;
	int32_t _pad;
	srt_float3 horizon_color, zenith_color, ground_color, sun_color, sun_direction;
} srt_scene_data;

/* Algorithmic work counters of an instrumented launch (no reference counterpart). */
typedef struct {
	uint64_t samples;   /* camera paths started                 (render.cl:495) */
	uint64_t bounces;   /* closest_intersection calls           (render.cl:404) */
	uint64_t tri_tests; /* ray x triangle tests                 (render.cl:324-331) */
	uint64_t aabb_pass; /* model AABB tests that passed         (render.cl:319) */
	uint64_t hits;      /* bounces that hit a shape             (render.cl:406) */
	uint64_t sky;       /* paths that escaped to the sky box    (render.cl:463-466) */
} srt_counters;

typedef enum {
	SRT_OK = 0,
	SRT_ERR_INVALID = 1, /* bad argument (null pointer, bad size, index out of range) */
	SRT_ERR_CUDA = 2,    /* a CUDA runtime call failed; see srt_last_error */
	SRT_ERR_NO_DEVICE = 3
} srt_status;

typedef struct srt_tracer srt_tracer;

/* Replaces Tracer::Tracer(width,height), tracer.cpp:11-68: picks the device, allocates the
 * float3 (16 B stride) canvas and the ARGB8 output, uploads the sky box.  skybox_rgba is
 * sky_w*sky_h RGBA float32 texels, row 0 = v 0, i.e. what stbi_loadf_from_file(...,4) returns
 * after stbi_set_flip_vertically_on_load(1) (tracer.cpp:42-52).  The canvas is zero-filled
 * (the reference leaves it uninitialised until the first clear_canvas).  device < 0 keeps the
 * current CUDA device. */
int srt_create(int width, int height, const float *skybox_rgba, int sky_w, int sky_h, int device,
               srt_tracer **out);

/* Replaces Tracer::update_scene, tracer.cpp:70-96.  Copy-in semantics: the host arrays may be
 * modified as soon as the call returns.  Records are the unchanged 128/96/64/96-byte structs;
 * the device-side SoA conversion (pre-transformed triangle edges) happens inside.  Validates
 * what the reference leaves unchecked: material < n_materials and
 * triangle_index + num_triangles <= n_triangles for every shape. */
int srt_upload_scene(srt_tracer *t, const srt_shape *shapes, size_t n_shapes,
                     const srt_triangle *triangles, size_t n_triangles,
                     const srt_material *materials, size_t n_materials,
                     const srt_scene_data *scene_data);

/* Replaces Tracer::clear_canvas, tracer.cpp:98-101 (asynchronous, stream-ordered). */
int srt_clear(srt_tracer *t);

/* Replaces the `render` kernel launch of Tracer::render, tracer.cpp:103-108:
 * canvas[id] += mean over num_samples paths (render.cl:483-523).  Asynchronous. */
int srt_render(srt_tracer *t, const srt_render_data *rd);

/* n consecutive srt_render calls in one: canvas += mean(launch 0), += mean(launch 1), ... -- the same additions in
 * the same order, so the canvas is bit-identical to n separate calls.  Runs of launches that differ only in `time`
 * (progressive accumulation with a fixed camera, src/main.cpp:283-290) are executed by ONE persistent kernel over
 * the combined item space, which removes the ragged end of every launch but the last; anything else falls back to
 * one kernel per launch.  No reference counterpart (the reference renders one launch per displayed frame). */
int srt_render_batch(srt_tracer *t, const srt_render_data *rds, size_t n);
/* Capacity hint (like std::vector::reserve): allocate now the per-sample scratch a batch of n launches shaped like
 * *rd will need, so that the first srt_render_batch does not pay for the allocation. */
int srt_reserve_batch(srt_tracer *t, const srt_render_data *rd, size_t n);

/* Replaces the `average` launch + blocking read-back of Tracer::render, tracer.cpp:110-115:
 * argb_out receives width*height*4 bytes in A,R,G,B order (render.cl:525-535).  Synchronises. */
int srt_resolve(srt_tracer *t, uint32_t num_steps, uint8_t *argb_out);

/* Optional: page-lock ONE caller-owned output buffer so that srt_resolve / srt_render_frame / srt_read_output copy
 * straight into it instead of through the handle's pinned staging buffer (saves a host memcpy of width*height*4
 * bytes per frame).  The reference's caller allocates its `pixels` vector once and hands it in every frame
 * (src/main.cpp:128,290), which is the case this is for.  The library never pins memory on its own: the CALLER
 * guarantees that [buffer, buffer+bytes) stays allocated until srt_unpin_output, a second srt_pin_output, or
 * srt_destroy -- freeing it while pinned is undefined behaviour (CUDA's rule for cudaHostRegister).  Reads into
 * any other address keep using the staging path. */
int srt_pin_output(srt_tracer *t, void *buffer, size_t bytes);
int srt_unpin_output(srt_tracer *t);

/* Tracer::render(ticks_stopped, output), tracer.cpp:103-116, in one call.  The result -- canvas and image -- is what
 * srt_render followed by srt_resolve produces, bit for bit.  When argb_out lies (16-byte aligned) inside the vector
 * the caller page-locked with srt_pin_output and the launch covers the full frame, the steps after the render kernel --
 * the per-pixel sum of the samples, `average`, the read-back -- run as ONE epilogue kernel that stores the ARGB8 image
 * straight into the caller's memory (16-byte stores over PCIe at the copy engine's rate; no separate `average` launch,
 * no copy-engine transfer).  Otherwise, and always with SRT_FRAME_SEPARATE (parity tests compare the two), the
 * separate steps run: render kernel + accumulate, average, copy through the handle's pinned staging buffer. */
int srt_render_frame(srt_tracer *t, const srt_render_data *rd, uint32_t ticks_stopped, uint8_t *argb_out);
enum { SRT_FRAME_AUTO = 0, SRT_FRAME_SEPARATE = 1 };
int srt_set_frame_pipeline(srt_tracer *t, int mode);

/* Restrict rendering to interleaved row bands (tile sharding across GPUs): only rows y with
 * (y / band_height) % band_count == band_index are traced; global pixel ids are preserved so
 * every sample is bit-identical to a full-frame launch.  band_count <= 1 renders all rows. */
int srt_set_row_bands(srt_tracer *t, int band_height, int band_index, int band_count);

/* OPTIONAL acceleration structure -- a labelled extension OUTSIDE the parity-graded path.  The
 * reference brute-forces every triangle of a model whose box the ray enters (render.cl:324) and lists a BVH first among
 * its future plans (README.md:41).  SRT_ACCEL_NONE (the default) is that brute-force path, bit-exact with render.cl.
 * SRT_ACCEL_BVH builds, at this call and at every later srt_upload_scene, a bounding-volume hierarchy over each model
 * of more than 32 triangles and traverses it instead: same exact test on the same operands, closest hit, equal t to
 * the lowest triangle index -- but only triangles whose boxes the ray enters are tested, so results are equal to the
 * brute-force path only up to its rounding-noise hits on far-away edge-on triangles (tolerance-tested, not bit-tested). */
enum { SRT_ACCEL_NONE = 0, SRT_ACCEL_BVH = 1 };
int srt_set_accel(srt_tracer *t, int accel);

/* Harness / test entry points (no reference counterpart). */
/* Which conservative filter the dense triangle sweep runs before the exact test (results are bit-identical either
 * way; only the speed differs): AUTO = two-strip (u and v ranges) for models of at most 20000 triangles, one-strip
 * (u range) above; the other two force one kind for every model (parity tests run both, tuning). */
enum { SRT_FILTER_AUTO = 0, SRT_FILTER_ONE_STRIP = 1, SRT_FILTER_TWO_STRIP = 2 };
int srt_set_sweep_filter(srt_tracer *t, int mode);
/* How scenes without large models schedule a warp's work (results are bit-identical either way): PLAIN shades a hit in
 * the lane and trip that found it, WAVEFRONT passes hits through a per-warp ring and shades them 32 at a time (faster
 * on long launches, slower on launches of a few items per thread); AUTO picks by launch size.  WAVEFRONT is honoured
 * when the launch fits its records (at most 256 bounces, fewer than 2^24 shapes).  Parity tests run both; tuning. */
enum { SRT_SCHEDULE_AUTO = 0, SRT_SCHEDULE_PLAIN = 1, SRT_SCHEDULE_WAVEFRONT = 2 };
int srt_set_schedule(srt_tracer *t, int schedule);
int srt_read_canvas(srt_tracer *t, float *rgba_out);               /* width*height*4 floats; synchronises */
int srt_write_canvas(srt_tracer *t, const float *rgba_in);         /* restore an accumulation (checkpoint/resume) */
int srt_canvas_device_ptr(srt_tracer *t, void **ptr, size_t *bytes); /* for NCCL reduce of per-GPU canvases */
int srt_output_device_ptr(srt_tracer *t, void **ptr, size_t *bytes); /* ARGB8 buffer, for gathers */
int srt_resolve_device(srt_tracer *t, uint32_t num_steps);         /* `average` without the read-back */
/* `average` over pixels [first_pixel, first_pixel + count) only: after a reduce-scatter of per-GPU canvases every
 * GPU resolves the slice it owns (sample sharding). */
int srt_resolve_device_range(srt_tracer *t, uint32_t num_steps, size_t first_pixel, size_t count);
int srt_read_output(srt_tracer *t, uint8_t *argb_out);             /* read-back of the ARGB8 buffer alone; synchronises */
int srt_stream(srt_tracer *t, void **cuda_stream);
int srt_synchronize(srt_tracer *t);
/* Shape index (-1 = miss) and distance of every pixel's sample-0 camera ray (parity gate). */
int srt_debug_primary(srt_tracer *t, const srt_render_data *rd, int32_t *shape_idx, float *t_out);
/* Same as srt_render, with the work counters of that launch added into *counters (slower). */
int srt_render_counted(srt_tracer *t, const srt_render_data *rd, srt_counters *counters);
/* Device math self-test: op 0 log, 1 cos, 2 atan2pi(x,y), 3 pow(x,y), 4 sqrt, 5 schlick(mu=x,cos=y); 6 / 7 = the two
 * halves of the packed-FP32x2 log of the pair {x, y}, 8 / 9 = of the packed cos (must equal ops 0 / 1 on x and on y);
 * 10 = normalize's range-tested 1 / sqrt(x), 11 = the two correctly rounded intrinsics it stands in for; 12 / 13 = the
 * halves of the packed sqrt of {x, y}; 14 / 15 = EXHAUSTIVE checks of ops 10 and 12 / 13 against the intrinsics: with
 * n = 2^20, thread i compares them on the 4096 bit patterns from i * 4096 and returns the number of mismatches. */
int srt_debug_math(srt_tracer *t, int op, const float *x, const float *y, float *out, size_t n);
/* FP32 FMA-chain micro-benchmark on the handle's device: achieved TFLOP/s. */
int srt_measure_fp32_peak(srt_tracer *t, double *tflops, double *sm_clock_mhz_est);
/* Total duration in ms of the render launches since the last call (CUDA events on the handle's stream) and how many
 * reference launches they covered (a batch counts each of its launches). */
int srt_render_time_ms(srt_tracer *t, double *total_ms, uint64_t *launches);

int srt_destroy(srt_tracer *t);
const char *srt_last_error(const srt_tracer *t); /* t may be NULL: error of the last failed srt_create */
int srt_abi_version(void);

/* Mesh / image I/O next to the path (reference include/parser.hpp:14-28, src/parser.cpp).
 * Loaders append to a growable triangle array owned by the library; free with srt_free. */
int srt_load_stl(const char *path, srt_triangle **triangles, size_t *count);
int srt_load_obj(const char *path, srt_triangle **triangles, size_t *count);
int srt_save_ppm(const char *path, const uint8_t *argb, int width, int height);
void srt_free(void *p);
/* The sky-box image as Tracer::Tracer prepares it (reference src/tracer.cpp:42-52: stbi_set_flip_vertically_on_load(1),
 * stbi_loadf_from_file(..., 4)): an 8-bit PNG decoded to width*height RGBA float32 texels, memory row 0 = image bottom,
 * colour = (float)pow(x / 255.0f, 2.2f), alpha = x / 255.0f (lib/stb_image.h:1868-1874).  The array is what srt_create
 * takes; free it with srt_free.  8-bit grey / grey+alpha / RGB / RGBA / palette PNGs, non-interlaced. */
int srt_load_skybox_png(const char *path, float **rgba, int *width, int *height);
/* Model::compute_bounding_box, reference src/shape.cpp:45-58 */
int srt_model_bounds(const srt_triangle *triangles, size_t n_triangles, srt_model *model);

#ifdef __cplusplus
}
#endif
#endif
