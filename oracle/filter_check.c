/*
 * filter_check.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Randomised CPU check of the claim the CUDA sweep filter rests on ("Triangle filter"):
 *
 *     the filter rejects a (ray, triangle) pair  ==>  the reference's own u test (render.cl:250-261) rejects it.
 *
 * The filter's arithmetic is restated here operation for operation (same fma chains, same constants as
 * simple_raytracer_b200/csrc/render_kernels.cuh: prepare_triangles_kernel's record, tri_filter_sweep); the
 * reference side is oracle_math.h's dot / cross exactly as oracle.c and the _ref build use them.  The harshest
 * admissible margins are used: K of the triangle itself (the kernel takes the maximum over the model) and R of the
 * ray itself (the kernel takes the maximum over the parked rays) -- larger margins only reject less.
 *
 * Pairs are drawn to sit where the claim is at risk: the ray is aimed at a point whose barycentric u is within a
 * few ulps .. 1e-3 of 0 or 1, or whose direction is within 1e-7 .. 1e-2 of the triangle's plane (det ~ 0), for
 * triangles of size 1e-4 .. 1e3 at distances up to 1e6 from the origin, plus uniformly random pairs.
 *
 * filter_check_uv does the same for the two-strip filter of small models (tri_filter_sweep_uv: u AND v ranges) against
 * the reference's u, v, u + v and t tests together, with extra pairs aimed at v ~ 0 and v ~ 1.
 *
 *   filter_check(n_pairs, seed, margin_scale, out[4]):  out = {pairs, filter rejects, violations, reference rejects}
 *   margin_scale = 1 is the kernel's margin; 0 removes it (the test uses that to show the margins matter).
 */
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle_math.h"

static inline uint64_t splitmix(uint64_t *s) {
	uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
	return z ^ (z >> 31);
}
static inline double urand(uint64_t *s) { return (double)(splitmix(s) >> 11) * (1.0 / 9007199254740992.0); }
static inline double srand1(uint64_t *s) { return 2.0 * urand(s) - 1.0; }
static inline double logu(uint64_t *s, double lo, double hi) { return lo * __builtin_exp(urand(s) * __builtin_log(hi / lo)); }

/* reference u decision, render.cl:250-261, on pre-transformed (v0, e1, e2): 1 = rejected (or det == 0, :253) */
static int reference_rejects_u(v3 v0, v3 e1, v3 e2, v3 o, v3 d) {
	v3 h = v3_cross(d, e2);
	float a = v3_dot(e1, h);
	if (a == 0.0f) return 1;
	float f = 1.0f / a;
	v3 s = v3_sub(o, v0);
	float u = f * v3_dot(s, h);
	return (u < 0.0f || u > 1.0f) ? 1 : 0; /* NaN: not rejected here */
}

/* the sweep filter: record as prepare_triangles_kernel builds it, test as tri_filter_sweep evaluates it */
static int filter_rejects(v3 v0, v3 e1, v3 e2, v3 o, v3 d, float margin_scale) {
	const float U = 5.9604644775390625e-8f;
	v3 np = v3_cross(e2, e1), m = v3_cross(e2, v0);
	float n1e1 = __builtin_fabsf(e1.x) + __builtin_fabsf(e1.y) + __builtin_fabsf(e1.z);
	float n1e2 = __builtin_fabsf(e2.x) + __builtin_fabsf(e2.y) + __builtin_fabsf(e2.z);
	float n1v0 = __builtin_fabsf(v0.x) + __builtin_fabsf(v0.y) + __builtin_fabsf(v0.z);
	float g = 48.0f * margin_scale * U * n1e2;
	float k = n1v0 + 3.0f * n1e1;
	float r = __builtin_fabsf(o.x) + __builtin_fabsf(o.y) + __builtin_fabsf(o.z) + k;
	float M = om_fma(g, r, 2e-6f * margin_scale);
	v3 c = v3_cross(o, d);
	float det = om_fma(d.z, np.z, om_fma(d.y, np.y, d.x * np.x));
	float t = om_fma(d.z, m.z, om_fma(d.y, m.y, d.x * m.x));
	float su = om_fma(e2.z, c.z, om_fma(e2.y, c.y, om_fma(e2.x, c.x, -t)));
	float diff = om_fma(-det, 0.500001f, su);  /* one rounding (FFMA), as in tri_filter_sweep */
	float w = om_fma(__builtin_fabsf(det), 0.500001f, M);
	return (__builtin_fabsf(diff) > w) ? 1 : 0;
}

/* reference decision through the u AND the v / u+v range tests, render.cl:250-268: 1 = the pair cannot become a hit */
static int reference_rejects_uv(v3 v0, v3 e1, v3 e2, v3 o, v3 d) {
	v3 h = v3_cross(d, e2);
	float a = v3_dot(e1, h);
	if (a == 0.0f) return 1;
	float f = 1.0f / a;
	v3 s = v3_sub(o, v0);
	float u = f * v3_dot(s, h);
	if (u < 0.0f || u > 1.0f) return 1;
	v3 q = v3_cross(s, e1);
	float v = f * v3_dot(d, q);
	if (v < 0.0f || u + v > 1.0f) return 1;
	float t = f * v3_dot(e2, q);
	return (t > 0.0f && t < __builtin_inff()) ? 0 : 1; /* a NaN / infinite / non-positive t is no hit either (:270) */
}

/* the two-strip sweep filter for small models: record as prepare_triangles_kernel builds it (n', -m, -m1, e2, e1, one
 * common margin scale), test as tri_filter_sweep_uv evaluates it.  With m1 = e1 x v0:  v det = d . m1 - e1 . c. */
static int filter_rejects_uv(v3 v0, v3 e1, v3 e2, v3 o, v3 d, float margin_scale) {
	const float U = 5.9604644775390625e-8f;
	v3 np = v3_cross(e2, e1), m = v3_cross(e2, v0), m1 = v3_cross(e1, v0);
	float n1e1 = __builtin_fabsf(e1.x) + __builtin_fabsf(e1.y) + __builtin_fabsf(e1.z);
	float n1e2 = __builtin_fabsf(e2.x) + __builtin_fabsf(e2.y) + __builtin_fabsf(e2.z);
	float n1v0 = __builtin_fabsf(v0.x) + __builtin_fabsf(v0.y) + __builtin_fabsf(v0.z);
	float emax = n1e1 > n1e2 ? n1e1 : n1e2;
	float g = 48.0f * margin_scale * U * emax;
	float k = n1v0 + 3.0f * emax;
	float r = __builtin_fabsf(o.x) + __builtin_fabsf(o.y) + __builtin_fabsf(o.z) + k;
	float M = om_fma(g, r, 2e-6f * margin_scale);
	v3 c = v3_cross(o, d);
	float det = om_fma(d.z, np.z, om_fma(d.y, np.y, d.x * np.x));
	float nt = om_fma(d.z, -m.z, om_fma(d.y, -m.y, d.x * -m.x));    /* -t  = d . (-m)  */
	float nt1 = om_fma(d.z, -m1.z, om_fma(d.y, -m1.y, d.x * -m1.x)); /* -t1 = d . (-m1) */
	float su = om_fma(c.z, e2.z, om_fma(c.y, e2.y, om_fma(c.x, e2.x, nt)));   /*  u det */
	float sv = om_fma(c.z, e1.z, om_fma(c.y, e1.y, om_fma(c.x, e1.x, nt1))); /* -v det */
	float du = om_fma(-det, 0.500001f, su), dv = om_fma(det, 0.500001f, sv);
	float w = om_fma(__builtin_fabsf(det), 0.500001f, M);
	return (__builtin_fabsf(du) > w || __builtin_fabsf(dv) > w) ? 1 : 0;
}

static inline v3 f3d(double x, double y, double z) { return v3_make((float)x, (float)y, (float)z); }

static void filter_check_impl(uint64_t n_pairs, uint64_t seed, float margin_scale, uint64_t out[4], int uv) {
	uint64_t pairs = 0, rejects = 0, violations = 0, ref_rejects = 0;
#pragma omp parallel reduction(+ : pairs, rejects, violations, ref_rejects)
	{
		int tid = 0, nth = 1;
#ifdef _OPENMP
		tid = omp_get_thread_num();
		nth = omp_get_num_threads();
#endif
		uint64_t s = seed * 0x2545f4914f6cdd1dull + (uint64_t)tid * 0x9e3779b97f4a7c15ull + 12345;
		for (uint64_t i = (uint64_t)tid; i < n_pairs; i += (uint64_t)nth) {
			/* triangle: size 1e-4 .. 1e3, centre up to 1e6 from the origin, arbitrary shape (sometimes a sliver) */
			double size = logu(&s, 1e-4, 1e3), off = (splitmix(&s) & 3) ? logu(&s, 1e-2, 1e6) : 0.0;
			double cx = off * srand1(&s), cy = off * srand1(&s), cz = off * srand1(&s);
			double ax = size * srand1(&s), ay = size * srand1(&s), az = size * srand1(&s);
			double bx = size * srand1(&s), by = size * srand1(&s), bz = size * srand1(&s);
			if ((splitmix(&s) & 7) == 0) { /* sliver: e2 almost parallel to e1 */
				double e = logu(&s, 1e-7, 1e-2);
				bx = ax * (1 + e * srand1(&s)) + e * size * srand1(&s), by = ay * (1 + e * srand1(&s)), bz = az + e * size * srand1(&s);
			}
			v3 v0 = f3d(cx, cy, cz);
			v3 v1 = f3d(cx + ax, cy + ay, cz + az), v2 = f3d(cx + bx, cy + by, cz + bz);
			v3 e1 = v3_sub(v1, v0), e2 = v3_sub(v2, v0); /* the stored operands, render.cl:247-248 */
			/* ray: origin at distance 0.1 .. 100 sizes, aimed at a chosen point of the triangle's plane */
			double dist = size * logu(&s, 0.1, 100.0);
			double ox = cx + dist * srand1(&s) + off * 0.1 * srand1(&s), oy = cy + dist * srand1(&s), oz = cz + dist * srand1(&s);
			int mode = (int)(splitmix(&s) % (uv ? 7 : 5));
			double u, v;
			if (mode == 0) { u = 3 * srand1(&s), v = 3 * srand1(&s); }                       /* anywhere */
			else if (mode == 1) { u = logu(&s, 1e-9, 1e-3) * srand1(&s), v = 2 * srand1(&s); } /* u ~ 0 */
			else if (mode == 2) { u = 1 + logu(&s, 1e-9, 1e-3) * srand1(&s), v = 2 * srand1(&s); } /* u ~ 1 */
			else if (mode == 5) { v = logu(&s, 1e-9, 1e-3) * srand1(&s), u = 2 * srand1(&s); }     /* v ~ 0 */
			else if (mode == 6) { v = 1 + logu(&s, 1e-9, 1e-3) * srand1(&s), u = 2 * srand1(&s); } /* v ~ 1 */
			else { u = 1.5 * srand1(&s), v = 1.5 * srand1(&s); }
			double px = cx + u * ax + v * bx, py = cy + u * ay + v * by, pz = cz + u * az + v * bz;
			double dx = px - ox, dy = py - oy, dz = pz - oz;
			if (mode == 3 || mode == 4) { /* grazing: direction almost in the triangle's plane */
				double e = logu(&s, 1e-7, 1e-2), a = srand1(&s), b = srand1(&s);
				double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
				double nn = __builtin_sqrt(nx * nx + ny * ny + nz * nz) + 1e-300;
				dx = a * ax + b * bx + e * size * nx / nn, dy = a * ay + b * by + e * size * ny / nn, dz = a * az + b * bz + e * size * nz / nn;
			}
			v3 o = f3d(ox, oy, oz);
			v3 d = v3_normalize(f3d(dx, dy, dz)); /* directions reach the intersection code normalised */
			if (!(d.x == d.x)) continue;
			pairs++;
			int fr = uv ? filter_rejects_uv(v0, e1, e2, o, d, margin_scale) : filter_rejects(v0, e1, e2, o, d, margin_scale);
			int rr = uv ? reference_rejects_uv(v0, e1, e2, o, d) : reference_rejects_u(v0, e1, e2, o, d);
			rejects += (uint64_t)fr;
			ref_rejects += (uint64_t)rr;
			if (fr && !rr) violations++;
		}
	}
	out[0] = pairs, out[1] = rejects, out[2] = violations, out[3] = ref_rejects;
}

void filter_check(uint64_t n_pairs, uint64_t seed, float margin_scale, uint64_t out[4]) {
	filter_check_impl(n_pairs, seed, margin_scale, out, 0);
}
/* the two-strip (u and v) filter against the reference's u, v, u+v and t tests together */
void filter_check_uv(uint64_t n_pairs, uint64_t seed, float margin_scale, uint64_t out[4]) {
	filter_check_impl(n_pairs, seed, margin_scale, out, 1);
}
