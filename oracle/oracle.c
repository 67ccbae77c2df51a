/*
 * oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement of the reference's path-tracing hot path,
 * /root/reference/src/render.cl:1-535, following its control flow function by function (AoS
 * records, per-ray vertex transform kept as at render.cl:326-328 so that it costs what the
 * reference costs).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (simple_raytracer_b200/csrc) never does.
 *
 * PARITY PINNED AGAINST THE REFERENCE SOURCE: oracle/_ref/libref_render_cl.so is render.cl itself, compiled
 * by g++ from /root/reference (oracle/ref_build/: an OpenCL-C language shim, one mechanical rewrite of vector
 * literals, an NDRange loop).  tests/test_ref_parity.py requires this file to reproduce that build BIT FOR
 * BIT (canvases, ARGB8 images, primary-hit ids) on all BASELINE configs and on random scenes, and the committed
 * golden fixtures (tests/golden/) are outputs of that build.  This restatement exists because the kernel
 * exposes neither the primary-hit shape index / t nor the work counters the roofline needs.
 * What remains this repository's own documented choice -- because it is third-party to the reference and no
 * OpenCL runtime exists in this image -- is the arithmetic INSIDE the OpenCL builtins (oracle_math.h, shared
 * with the _ref build) and the decision not to contract render.cl's own a*b+c expressions (om_cfma;
 * -DORACLE_CONTRACT=1 builds the fused variant used to measure the sensitivity).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle_math.h"

/* ---- device ABI records, render.cl:5-105 (OpenCL float3 = 16 bytes, 16-aligned) ---------- */
typedef struct { float x, y, z, w; } cl_f3; /* .w is padding */
typedef struct { float x, y, z, w; } cl_f4;

typedef struct { /* render.cl:17-27 */
	float smoothness, metallic, specular, emission_strength, transmittance, refraction_index;
	float _pad[2];
	cl_f3 color, emission;
} Material;
typedef struct { cl_f3 position; float radius; float _pad[3]; } Sphere;   /* :29-32 */
typedef struct { cl_f3 position, normal; } Plane;                          /* :34-37 */
typedef struct { cl_f3 normal, pos; } Vertex;                              /* :39-42 */
typedef struct { Vertex v[3]; } Triangle;                                  /* :44-53 */
typedef struct { /* :55-61 */
	uint32_t triangle_index, num_triangles;
	uint32_t _pad[2];
	cl_f3 bounding_min, bounding_max;
	cl_f4 transform[4];
} Model;
enum { SHAPE_SPHERE = 0, SHAPE_PLANE = 1, SHAPE_MODEL = 2 };               /* :63-67 */
typedef struct { /* :69-77 */
	int32_t type, material;
	int32_t _pad[2];
	union { Sphere sphere; Plane plane; Model model; } shape;
} Shape;
typedef struct { /* :79-92 */
	int32_t width, height, num_samples, num_bounces;
	float aspect_ratio, fov_scale;
	char show_normals;
	char _pad[7];
	cl_f4 camera_to_world[4];
	uint32_t time, tick;
	uint32_t _pad2[2];
} RenderData;
typedef struct { /* :94-105 */
	int32_t num_shapes;
	float sun_focus, sun_intensity;
	int32_t _pad;
	cl_f3 horizon_color, zenith_color, ground_color, sun_color, sun_direction;
} SceneData;

_Static_assert(sizeof(Material) == 64 && offsetof(Material, color) == 32 && offsetof(Material, emission) == 48, "Material");
_Static_assert(sizeof(Sphere) == 32 && offsetof(Sphere, radius) == 16, "Sphere");
_Static_assert(sizeof(Plane) == 32 && offsetof(Plane, normal) == 16, "Plane");
_Static_assert(sizeof(Triangle) == 96, "Triangle");
_Static_assert(sizeof(Model) == 112 && offsetof(Model, bounding_min) == 16 && offsetof(Model, transform) == 48, "Model");
_Static_assert(sizeof(Shape) == 128 && offsetof(Shape, shape) == 16, "Shape");
_Static_assert(sizeof(RenderData) == 112 && offsetof(RenderData, show_normals) == 24 &&
               offsetof(RenderData, camera_to_world) == 32 && offsetof(RenderData, time) == 96, "RenderData");
_Static_assert(sizeof(SceneData) == 96 && offsetof(SceneData, horizon_color) == 16 &&
               offsetof(SceneData, sun_direction) == 80, "SceneData");

/* Algorithmic work counters. Not part of render.cl; used for the roofline figure. */
typedef struct {
	uint64_t samples;      /* camera paths started                        (render.cl:495) */
	uint64_t bounces;      /* closest_intersection calls                  (:404)          */
	uint64_t tri_tests;    /* ray x triangle tests                        (:324-331)      */
	uint64_t aabb_pass;    /* model AABB tests that passed                (:319)          */
	uint64_t hits;         /* bounces that hit something                  (:406)          */
	uint64_t sky;          /* paths that escaped to the sky box           (:463-466)      */
} OracleCounters;

typedef struct { v3 origin, direction; } Ray;                              /* :5-8 */
typedef struct { v3 position, normal; int front; } Intersection;          /* :10-15 */

typedef struct {
	const SceneData *data;
	const Shape *shapes;
	const Triangle *triangles;
	const Material *materials;
	const float *sky; /* RGBA f32, row 0 = v 0 */
	int sky_w, sky_h;
} Scene;

static inline v3 f3(cl_f3 a) { return v3_make(a.x, a.y, a.z); }

/* render.cl:114-120 -- row k = ((m0.k*x + m1.k*y) + m2.k*z) + m3.k*w, left to right */
static inline v3 matrix_by_vector3(const cl_f4 *m, v3 v, float w) {
	return v3_make(om_cfma(m[3].x, w, om_cfma(m[2].x, v.z, om_cfma(m[1].x, v.y, m[0].x * v.x))),
	               om_cfma(m[3].y, w, om_cfma(m[2].y, v.z, om_cfma(m[1].y, v.y, m[0].y * v.x))),
	               om_cfma(m[3].z, w, om_cfma(m[2].z, v.z, om_cfma(m[1].z, v.y, m[0].z * v.x))));
}
/* render.cl:135-137 */
static inline v3 transform_mat(const cl_f4 *m, v3 p, int translate) {
	return matrix_by_vector3(m, p, translate ? 1.0f : 0.0f);
}
/* render.cl:139-141:  v - 2*dot(v,n)*n */
static inline v3 reflect(v3 v, v3 n) {
	float k = 2.0f * v3_dot(v, n);
	return v3_make(om_cfma(-k, n.x, v.x), om_cfma(-k, n.y, v.y), om_cfma(-k, n.z, v.z));
}

/* render.cl:165-167: written out by the source as x*x + y*y + z*z, not a dot() call */
static inline float length_squared(v3 v) { return om_cfma(v.z, v.z, om_cfma(v.y, v.y, v.x * v.x)); }

/* render.cl:143-148.  (float)UINT_MAX == 2^32, so the division is an exact scaling. */
static inline uint32_t rng_hash(uint32_t *seed) {
	*seed = *seed * 747796405u + 2891336453u;
	uint32_t result = ((*seed >> ((*seed >> 28) + 4u)) ^ *seed) * 277803737u;
	result = (result >> 22) ^ result;
	return result;
}
static inline float random_float(uint32_t *seed) {
	return (float)rng_hash(seed) / 4294967296.0f;
}
/* render.cl:150-154; theta's draw precedes rho's */
static inline float random_float_normal(uint32_t *seed) {
	float theta = 6.28318530717958647692f * random_float(seed);
	float rho = om_sqrt(-2.0f * om_log(random_float(seed)));
	return rho * om_cos(theta);
}
/* render.cl:156-158; components drawn left to right */
static inline v3 random_direction(uint32_t *seed) {
	float x = random_float_normal(seed);
	float y = random_float_normal(seed);
	float z = random_float_normal(seed);
	return v3_normalize(v3_make(x, y, z));
}
/* render.cl:160-163 */
static inline v3 random_direction_hemisphere(v3 normal, uint32_t *seed) {
	v3 dir = random_direction(seed);
	return v3_scale(dir, om_sign(v3_dot(normal, dir)));
}
/* render.cl:173-178, evaluated in double as the 1.0 literals demand; pown(x,5) = ((((x*x)*x)*x)*x) */
static inline float shlick_reflectance(float mu, float cos_theta) {
	float r0 = (float)((1.0 - (double)mu) / (1.0 + (double)mu));
	r0 = r0 * r0;
	double c = 1.0 - (double)cos_theta;
	double c5 = (((c * c) * c) * c) * c;
	return (float)((double)r0 + (1.0 - (double)r0) * c5);
}

/* render.cl:180-204 */
static inline int intersect_sphere(const Sphere *sphere, const Ray *ray, float *t) {
	v3 rayToCenter = v3_sub(f3(sphere->position), ray->origin);
	float b = v3_dot(rayToCenter, ray->direction);
	float c = om_cfma(-sphere->radius, sphere->radius, v3_dot(rayToCenter, rayToCenter));
	float disc = om_cfma(b, b, -c);
	if (disc < 0.0f) return 0;
	float sq = om_sqrt(disc);
	*t = b - sq;
	if (*t < 0.0f) {
		*t = b + sq;
		if (*t < 0.0f) return 0;
	}
	return 1;
}
/* render.cl:206-221 */
static inline int intersect_plane(const Plane *plane, const Ray *ray, float *t) {
	v3 n = f3(plane->normal);
	float denom = v3_dot(n, ray->direction);
	if (__builtin_fabsf(denom) == 0.0f) return 0;
	float tmp = v3_dot(n, v3_sub(f3(plane->position), ray->origin)) / denom;
	if (tmp < 0.0f) return 0;
	*t = tmp;
	return 1;
}
/* render.cl:223-241; weights returned rotated (w2,w0,w1) */
static inline v3 barycentric_weights(const v3 pos[3], v3 p) {
	v3 v0 = v3_sub(pos[1], pos[0]);
	v3 v1 = v3_sub(pos[2], pos[0]);
	v3 v2 = v3_sub(p, pos[0]);
	float d00 = v3_dot(v0, v0);
	float d01 = v3_dot(v0, v1);
	float d11 = v3_dot(v1, v1);
	float d20 = v3_dot(v2, v0);
	float d21 = v3_dot(v2, v1);
	float denom = om_cfma(d00, d11, -(d01 * d01));
	float w0 = om_cfma(d11, d20, -(d01 * d21)) / denom;
	float w1 = om_cfma(d00, d21, -(d01 * d20)) / denom;
	float w2 = (1.0f - w0) - w1;
	return v3_make(w2, w0, w1);
}
/* render.cl:243-275 (Moller-Trumbore, no culling) */
static inline int intersect_triangle(const v3 pos[3], const Ray *ray, float *t) {
	v3 edge1 = v3_sub(pos[1], pos[0]);
	v3 edge2 = v3_sub(pos[2], pos[0]);
	v3 h = v3_cross(ray->direction, edge2);
	float a = v3_dot(edge1, h);
	if (a == 0.0f) return 0;
	float f = 1.0f / a;
	v3 s = v3_sub(ray->origin, pos[0]);
	float u = f * v3_dot(s, h);
	if (u < 0.0f || u > 1.0f) return 0;
	v3 q = v3_cross(s, edge1);
	float v = f * v3_dot(ray->direction, q);
	if (v < 0.0f || u + v > 1.0f) return 0;
	*t = f * v3_dot(edge2, q);
	return *t > 0.0f;
}
/* render.cl:279-290 */
static inline int intersection_aabb(v3 bmin, v3 bmax, const Ray *ray, v3 inv_dir, float tmax) {
	float tmin = 0.0f;
	const float bn[3] = {bmin.x, bmin.y, bmin.z}, bx[3] = {bmax.x, bmax.y, bmax.z};
	const float o[3] = {ray->origin.x, ray->origin.y, ray->origin.z};
	const float id[3] = {inv_dir.x, inv_dir.y, inv_dir.z};
	for (int d = 0; d < 3; d++) {
		float t1 = (bn[d] - o[d]) * id[d];
		float t2 = (bx[d] - o[d]) * id[d];
		tmin = om_max(tmin, om_min(t1, t2));
		tmax = om_min(tmax, om_max(t1, t2));
	}
	return tmin < tmax;
}

/* render.cl:293-378.  Returns the material index; *shape_out / *t_out are debug outputs the
 * OpenCL kernel does not have (primary-hit parity gate). */
static int closest_intersection(const Scene *scene, const Ray *ray, Intersection *rayhit,
                                int *shape_out, float *t_out, OracleCounters *cnt) {
	int closest = -1, closest_shape = -1;
	float tmin = INFINITY;
	v3 inv_dir = v3_make(1.0f / ray->direction.x, 1.0f / ray->direction.y, 1.0f / ray->direction.z);
	cnt->bounces++;

	for (int i = 0; i < scene->data->num_shapes; i++) {
		const Shape *shape = &scene->shapes[i];
		if (shape->type == SHAPE_SPHERE) {
			const Sphere *sphere = &shape->shape.sphere;
			float t_i;
			if (intersect_sphere(sphere, ray, &t_i) && t_i < tmin) {
				tmin = t_i;
				closest = shape->material;
				closest_shape = i;
				rayhit->position = v3_cfma(ray->direction, tmin, ray->origin);
				float r = sphere->radius;
				v3 d = v3_sub(rayhit->position, f3(sphere->position));
				rayhit->normal = v3_make(d.x / r, d.y / r, d.z / r);
			}
		} else if (shape->type == SHAPE_MODEL) {
			const Model *model = &shape->shape.model;
			if (!intersection_aabb(f3(model->bounding_min), f3(model->bounding_max), ray, inv_dir, tmin))
				continue;
			cnt->aabb_pass++;
			cnt->tri_tests += model->num_triangles;
			for (size_t k = 0; k < model->num_triangles; k++) {
				const Triangle *tri = &scene->triangles[model->triangle_index + k];
				v3 pos[3];
				for (int j = 0; j <= 2; j++) pos[j] = transform_mat(model->transform, f3(tri->v[j].pos), 1);
				float t_i;
				if (intersect_triangle(pos, ray, &t_i) && t_i < tmin) {
					tmin = t_i;
					closest = shape->material;
					closest_shape = i;
					rayhit->position = v3_cfma(ray->direction, tmin, ray->origin);
					v3 w = barycentric_weights(pos, rayhit->position);
					v3 n0 = f3(tri->v[0].normal), n1 = f3(tri->v[1].normal), n2 = f3(tri->v[2].normal);
					/* n0*w.x + n1*w.y + n2*w.z */
					v3 n = v3_make(om_cfma(n2.x, w.z, om_cfma(n1.x, w.y, n0.x * w.x)),
					               om_cfma(n2.y, w.z, om_cfma(n1.y, w.y, n0.y * w.x)),
					               om_cfma(n2.z, w.z, om_cfma(n1.z, w.y, n0.z * w.x)));
					n = transform_mat(model->transform, n, 0);
					rayhit->normal = v3_normalize(n);
				}
			}
		} else if (shape->type == SHAPE_PLANE) {
			const Plane *plane = &shape->shape.plane;
			float t_i;
			if (intersect_plane(plane, ray, &t_i) && t_i < tmin) {
				tmin = t_i;
				closest = shape->material;
				closest_shape = i;
				rayhit->normal = f3(plane->normal);
				rayhit->position = v3_cfma(ray->direction, tmin, ray->origin);
			}
		}
	}
	if (shape_out) *shape_out = closest_shape;
	if (t_out) *t_out = tmin;
	if (closest_shape < 0) return -1; /* render.cl:369-375 on a miss has no observable effect */

	rayhit->front = v3_dot(rayhit->normal, ray->direction) < 0.0f;
	if (!rayhit->front) rayhit->normal = v3_scale(rayhit->normal, -1.0f);
	else rayhit->normal = v3_scale(rayhit->normal, 1.0f);
	return closest;
}

/* read_imagef(...).xyz, render.cl:393: the builtin is defined in oracle_math.h */
static inline v3 sky_fetch(const Scene *scene, float u, float v) {
	float r[4];
	om_read_imagef_linear_clamp(scene->sky, scene->sky_w, scene->sky_h, u, v, r);
	return v3_make(r[0], r[1], r[2]);
}
/* render.cl:380-394 */
static v3 sky_box(const Ray *ray, const Scene *scene) {
	const SceneData *d = scene->data;
	float sd = om_max(v3_dot(ray->direction, v3_neg(f3(d->sun_direction))), 0.0f);
	float pw = om_pow(sd, d->sun_focus);
	v3 sun = v3_scale(v3_scale(f3(d->sun_color), pw), d->sun_intensity);
	float u = om_cfma(om_atan2pi(ray->direction.z, ray->direction.x), 0.5f, 0.5f);
	float v = om_cfma(ray->direction.y, 0.5f, 0.5f);
	return v3_add(sky_fetch(scene, u, v), sun);
}

/* render.cl:396-471 */
static v3 trace(const RenderData *render, const Scene *scene, const Ray *camray, uint32_t seed,
                OracleCounters *cnt) {
	v3 color = v3_make(0.f, 0.f, 0.f);
	v3 mask = v3_make(1.f, 1.f, 1.f);
	Ray ray = *camray;
	Intersection rayhit;
	memset(&rayhit, 0, sizeof rayhit);

	for (int i = 0; i < render->num_bounces; i++) {
		int material_index = closest_intersection(scene, &ray, &rayhit, NULL, NULL, cnt);
		if (material_index >= 0) {
			cnt->hits++;
			if (render->show_normals) {
				color = v3_make(om_cfma(rayhit.normal.x, 0.5f, 0.5f), om_cfma(rayhit.normal.y, 0.5f, 0.5f),
				                om_cfma(rayhit.normal.z, 0.5f, 0.5f));
				break;
			}
			const Material *material = &scene->materials[material_index];
			/* color += mask * emission * emission_strength */
			v3 e = v3_scale(v3_mul(mask, f3(material->emission)), material->emission_strength);
			color = v3_add(color, e);
			if (i == render->num_bounces - 1) break;

			ray.origin = rayhit.position;
			v3 random_dir = v3_normalize(v3_add(rayhit.normal, random_direction_hemisphere(rayhit.normal, &seed)));
			v3 reflected_dir = reflect(ray.direction, rayhit.normal);

			int is_metallic = material->metallic > random_float(&seed);
			int is_specular = material->specular > random_float(&seed);
			v3 rough_dir = v3_mix(random_dir, reflected_dir, material->smoothness);
			int is_transparent = material->transmittance > random_float(&seed);

			if (!is_transparent) {
				ray.direction = v3_mix(random_dir, rough_dir, (is_metallic || is_specular) ? 1.0f : 0.0f);
				v3 one = v3_make(1.0f, 1.0f, 1.0f);
				mask = v3_mul(mask, v3_mix(f3(material->color), one, is_specular ? 1.0f : 0.0f));
			} else {
				v3 in_dir = reflect(rough_dir, rayhit.normal);
				float mu = rayhit.front ? 1.0f / material->refraction_index : material->refraction_index;
				float cos_theta = om_min(1.0f, v3_dot(in_dir, v3_neg(rayhit.normal)));
				float sin_theta = om_sqrt(om_cfma(-cos_theta, cos_theta, 1.0f));
				int transparency_reflected = mu * sin_theta > 1.0f ||
				                             shlick_reflectance(mu, cos_theta) > random_float(&seed);
				if (transparency_reflected) {
					ray.direction = rough_dir;
				} else {
					v3 out_perp = v3_scale(v3_cfma(rayhit.normal, cos_theta, in_dir), mu);
					float k = -om_sqrt(__builtin_fabsf(1.0f - length_squared(out_perp)));
					ray.direction = v3_cfma(rayhit.normal, k, out_perp);
					mask = v3_mul(mask, f3(material->color));
				}
			}
			ray.direction = v3_normalize(ray.direction);
			/* origin += normal * sign(dot(normal, dir)) * 0.001 */
			float sg = om_sign(v3_dot(rayhit.normal, ray.direction)) * 0.001f;
			ray.origin = v3_cfma(rayhit.normal, sg, ray.origin);
		} else {
			cnt->sky++;
			mask = v3_mul(mask, sky_box(&ray, scene));
			color = v3_add(color, mask);
			break;
		}
	}
	return color;
}

/* render.cl:496 */
static inline uint32_t sample_seed(uint32_t sample, uint32_t id, uint32_t num_samples, uint32_t time) {
	return (sample + id * num_samples) * time * 5304u;
}
/* render.cl:498-516 */
static inline Ray camera_ray(const RenderData *data, int gx, int gy, uint32_t *seed) {
	float u0 = random_float(seed);
	float u1 = random_float(seed);
	float ndc_x = ((float)gx + u0) / (float)data->width;
	float ndc_y = ((float)gy + u1) / (float)data->height;
	float sx = (om_cfma(2.0f, ndc_x, -1.0f) * data->aspect_ratio) * data->fov_scale;
	float sy = om_cfma(-2.0f, ndc_y, 1.0f) * data->fov_scale;
	Ray ray;
	ray.origin = v3_make(data->camera_to_world[3].x, data->camera_to_world[3].y, data->camera_to_world[3].z);
	ray.direction = v3_normalize(matrix_by_vector3(data->camera_to_world, v3_make(sx, sy, -1.0f), 0.0f));
	return ray;
}

/* kernel `render`, render.cl:483-523, over the pixel window [x0,x1) x [y0,y1) of the
 * width x height image (global ids preserved).  canvas: width*height*4 floats (float3 stride
 * 16 B, src/tracer.cpp:39); .w untouched.  rows with (y / band_h) % band_n != band_i are skipped
 * when band_n > 1 (tile-sharding tests). */
void oracle_render(const RenderData *data, const SceneData *scene_data, float *canvas,
                   const Shape *shapes, const Triangle *triangles, const Material *materials,
                   const float *sky, int sky_w, int sky_h, int x0, int y0, int x1, int y1,
                   int band_h, int band_i, int band_n, int threads, OracleCounters *counters) {
	Scene scene = {scene_data, shapes, triangles, materials, sky, sky_w, sky_h};
	OracleCounters total;
	memset(&total, 0, sizeof total);
	if (threads <= 0) {
#ifdef _OPENMP
		threads = omp_get_max_threads();
#else
		threads = 1;
#endif
	}
#pragma omp parallel num_threads(threads)
	{
		OracleCounters cnt;
		memset(&cnt, 0, sizeof cnt);
#pragma omp for schedule(dynamic, 1)
		for (int gy = y0; gy < y1; gy++) {
			if (band_n > 1 && (gy / band_h) % band_n != band_i) continue;
			for (int gx = x0; gx < x1; gx++) {
				uint32_t id = (uint32_t)gx + (uint32_t)gy * (uint32_t)data->width;
				v3 color = v3_make(0.f, 0.f, 0.f);
				for (int sample = 0; sample < data->num_samples; sample++) {
					uint32_t seed = sample_seed((uint32_t)sample, id, (uint32_t)data->num_samples, data->time);
					Ray ray = camera_ray(data, gx, gy, &seed);
					cnt.samples++;
					color = v3_add(color, trace(data, &scene, &ray, seed, &cnt));
				}
				float ns = (float)data->num_samples;
				canvas[4 * (size_t)id + 0] += color.x / ns;
				canvas[4 * (size_t)id + 1] += color.y / ns;
				canvas[4 * (size_t)id + 2] += color.z / ns;
			}
		}
#pragma omp critical
		{
			total.samples += cnt.samples; total.bounces += cnt.bounces; total.tri_tests += cnt.tri_tests;
			total.aabb_pass += cnt.aabb_pass; total.hits += cnt.hits; total.sky += cnt.sky;
		}
	}
	if (counters) {
		counters->samples += total.samples; counters->bounces += total.bounces;
		counters->tri_tests += total.tri_tests; counters->aabb_pass += total.aabb_pass;
		counters->hits += total.hits; counters->sky += total.sky;
	}
}

/* Debug view the OpenCL kernel lacks: shape index (-1 = sky) and t of the sample-0 camera ray. */
void oracle_primary(const RenderData *data, const SceneData *scene_data, const Shape *shapes,
                    const Triangle *triangles, int32_t *shape_idx, float *t_out, int threads) {
	Scene scene = {scene_data, shapes, triangles, NULL, NULL, 0, 0};
	if (threads <= 0) {
#ifdef _OPENMP
		threads = omp_get_max_threads();
#else
		threads = 1;
#endif
	}
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
	for (int gy = 0; gy < data->height; gy++) {
		OracleCounters cnt;
		memset(&cnt, 0, sizeof cnt);
		for (int gx = 0; gx < data->width; gx++) {
			uint32_t id = (uint32_t)gx + (uint32_t)gy * (uint32_t)data->width;
			uint32_t seed = sample_seed(0u, id, (uint32_t)data->num_samples, data->time);
			Ray ray = camera_ray(data, gx, gy, &seed);
			Intersection hit;
			int shape = -1;
			float t = INFINITY;
			closest_intersection(&scene, &ray, &hit, &shape, &t, &cnt);
			shape_idx[id] = shape;
			t_out[id] = t;
		}
	}
}

/* render.cl:473-481 */
static inline float aces1(float x) {
	const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
	float num = x * om_cfma(x, a, b);
	float den = om_cfma(x, om_cfma(x, c, d), e);
	return om_clamp(num / den, 0.0f, 1.0f); /* NaN -> 0 (hazard ix) */
}
/* kernel `average`, render.cl:525-535: canvas/num_steps -> aces -> sqrt -> ARGB8 (truncating) */
void oracle_average(uint32_t num_steps, const float *canvas, uint8_t *output, size_t n) {
	float steps = (float)num_steps;
	for (size_t id = 0; id < n; id++) {
		output[4 * id + 0] = 255;
		for (int c = 0; c < 3; c++) {
			float v = om_sqrt(aces1(canvas[4 * id + c] / steps));
			output[4 * id + 1 + c] = (uint8_t)(int)(v * 255.0f);
		}
	}
}

/* ---- small entry points for unit tests -------------------------------------------------- */
uint32_t oracle_seed(uint32_t sample, uint32_t id, uint32_t num_samples, uint32_t time) {
	return sample_seed(sample, id, num_samples, time);
}
/* advances *seed; writes the raw hash and the float */
float oracle_random_float(uint32_t *seed, uint32_t *hash_out) {
	uint32_t s = *seed;
	uint32_t h = rng_hash(&s);
	if (hash_out) *hash_out = h;
	return random_float(seed);
}
/* op: 0 log, 1 cos, 2 atan2pi(x, y[i]) , 3 pow(x, y[i]), 4 sqrt, 5 schlick(x=mu, y=cos) */
void oracle_math(int op, const float *x, const float *y, float *out, size_t n) {
	for (size_t i = 0; i < n; i++) {
		switch (op) {
		case 0: out[i] = om_log(x[i]); break;
		case 1: out[i] = om_cos(x[i]); break;
		case 2: out[i] = om_atan2pi(x[i], y[i]); break;
		case 3: out[i] = om_pow(x[i], y[i]); break;
		case 4: out[i] = om_sqrt(x[i]); break;
		case 5: out[i] = shlick_reflectance(x[i], y[i]); break;
		default: out[i] = 0.0f;
		}
	}
}
/* closed-form intersection probes: kind 0 sphere(a=center,b.x=radius) 1 plane(a=pos,b=normal)
 * 2 triangle(a,b,c = positions) 3 aabb(a=min,b=max,c.x=tmax). returns hit flag, *t */
int oracle_intersect(int kind, const float *o, const float *d, const float *a, const float *b,
                     const float *c, float *t) {
	Ray ray = {v3_make(o[0], o[1], o[2]), v3_make(d[0], d[1], d[2])};
	*t = INFINITY;
	if (kind == 0) {
		Sphere s;
		memset(&s, 0, sizeof s);
		s.position.x = a[0]; s.position.y = a[1]; s.position.z = a[2]; s.radius = b[0];
		return intersect_sphere(&s, &ray, t);
	} else if (kind == 1) {
		Plane p;
		memset(&p, 0, sizeof p);
		p.position.x = a[0]; p.position.y = a[1]; p.position.z = a[2];
		p.normal.x = b[0]; p.normal.y = b[1]; p.normal.z = b[2];
		return intersect_plane(&p, &ray, t);
	} else if (kind == 2) {
		v3 pos[3] = {v3_make(a[0], a[1], a[2]), v3_make(b[0], b[1], b[2]), v3_make(c[0], c[1], c[2])};
		return intersect_triangle(pos, &ray, t);
	} else {
		v3 inv = v3_make(1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z);
		return intersection_aabb(v3_make(a[0], a[1], a[2]), v3_make(b[0], b[1], b[2]), &ray, inv, c[0]);
	}
}
int oracle_max_threads(void) {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}
