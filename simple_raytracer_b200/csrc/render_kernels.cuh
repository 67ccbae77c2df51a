// render_kernels.cuh -- the path-tracing hot path for sm_100a.
//
// Replaces the OpenCL kernels `render` and `average` of reference src/render.cl:483-535 (and every
// helper they call, :114-481).  This is not a translation of that file: the device data is SoA with
// model triangles pre-transformed to world space (v0, e1, e2) at upload, path state lives in
// registers, and the bounce loop is FLATTENED -- one persistent thread pulls (pixel, sample) work
// items and runs exactly one bounce per loop trip, starting the next camera path in place the moment
// its current path ends.  A warp therefore never waits on its longest path: terminated lanes are
// refilled through a ballot-aggregated atomic on a global item cursor, and warps stay full until the
// frame runs dry.  Every path writes its radiance to a per-launch scratch buffer; accumulate_kernel
// then adds the samples of each pixel in sample order (render.cl:495-522), so results are bit-identical
// to the sequential formulation while the unit of scheduling is one path, not one pixel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_math.cuh"
#include "fastdiv.hpp"

// Developer build with device-side invariant checks (compute-sanitizer is not available on the GPU pool):
// scripts/build_variant.sh checks "-DSRT_DEBUG_CHECKS=1", then run the GPU tests with SRT_LIB pointing at it.
#ifndef SRT_DEBUG_CHECKS
#define SRT_DEBUG_CHECKS 0
#endif
#if SRT_DEBUG_CHECKS
#include <cstdio>
#define SRT_ASSERT(cond)                                                       \
	do {                                                                       \
		if (!(cond)) {                                                         \
			printf("SRT_ASSERT failed: %s (line %d)\n", #cond, __LINE__);      \
			__trap();                                                          \
		}                                                                      \
	} while (0)
#else
#define SRT_ASSERT(cond) ((void)0)
#endif

namespace srt {

enum { SHAPE_SPHERE = 0, SHAPE_PLANE = 1, SHAPE_MODEL = 2 };

// Device scene: SoA primitive buffers (see DESIGN.md "Data layout in HBM").
struct DevScene {
	int num_shapes;
	int has_models;
	const int4 *shape_hdr;    // {type, material, soa_tri_begin, tri_count} in reference array order
	const float4 *shape_a;    // sphere {c.xyz, r} | plane {p0.xyz, -} | model {bmin.xyz, -}
	const float4 *shape_b;    // sphere {-}        | plane {n.xyz, -}  | model {bmax.xyz, -}
	const float4 *tri_hot;    // 3 per triangle, 48 B stride: world-space v0, e1 = v1-v0, e2 = v2-v0 (+1 pad triangle)
	const float2 *tri_flt;    // 5 per triangle, 40 B stride: the sweep filter's record n', g, e2, m (tri_filter_sweep)
	const float4 *tri_uv;     // 5 per triangle, 80 B stride: the two-strip filter's record (tri_filter_sweep_uv)
	const float *model_k;     // per shape slot: max over the model's triangles of |v0|_1 + 3 max(|e1|_1, |e2|_1) (filter margin)
	const float4 *tri_n;      // 3 per triangle: object-space vertex normals (cold: winner only)
	const float4 *model_xf;   // 4 per shape slot: model matrix columns      (cold)
	const float4 *materials;  // 4 per material: the reference's 64-byte record as 4 x float4
	const float4 *sky;        // RGBA f32 texels, row 0 = v 0
	// optional acceleration structure (srt_set_accel(SRT_ACCEL_BVH); null otherwise): bvh_build.hpp
	const float4 *bvh_nodes;  // 4 per node: {lo0, c0} {hi0, n0} {lo1, c1} {hi1, n1}
	const int *bvh_order;     // leaf slots -> SoA triangle index
	const int *bvh_root;      // per shape slot: entry node of the model's hierarchy, -1 = none (brute force)
	int sky_w, sky_h;
	float sun_focus, sun_intensity;
	float sun_color[3];
	float sun_dir[3];
};

// The first CONST_SHAPES shape records once more, as a KERNEL PARAMETER (constant bank): the closest-hit scan of a small
// analytic scene walks the shape list with a warp-uniform index, so from here every record is a uniform constant-bank
// load (LDCU) feeding uniform-register operands -- no per-lane address arithmetic, no LDG latency in the scan loop
// (7 shapes x ~20 instructions of loop overhead per bounce on BASELINE config 2).  finish_hit keeps reading the global
// arrays: its index differs from lane to lane.
constexpr int CONST_SHAPES = 16;
// SRT_PAIR_SCAN: the wavefront build of such a scene scans the shapes TWO AT A TIME (scan_pairs below).  Consecutive
// shapes of the same type form one "op" whose operands sit interleaved {A, B} in the table, so that every floating-point
// operation before the per-shape decision is one packed FP32x2 instruction whose halves belong to shape A and shape B
// (the ray operand is the broadcast scalar); a shape without a partner is an op whose B half repeats A and is ignored.
#ifndef SRT_PAIR_SCAN
#define SRT_PAIR_SCAN 1
#endif
#ifndef SRT_RING_V4   // render_wavefront: hit and sky records through 16-byte shared-memory accesses
#define SRT_RING_V4 1
#endif
#ifndef SRT_WF_CARRY  // render_wavefront: path-state registers carried across trips instead of re-initialised
#define SRT_WF_CARRY 1
#endif
#ifndef SRT_PAIR_ANY  // scan_pairs: one combined miss test per sphere pair
#define SRT_PAIR_ANY 1
#endif
#ifndef SRT_FASTDIV   // start_path: the item -> (launch, pixel, sample, row, column) divisions by precomputed multipliers
#define SRT_FASTDIV 1
#endif
struct ShapeTable {
	int4 hdr[CONST_SHAPES];
	float4 a[CONST_SHAPES];
	float4 b[CONST_SHAPES];
#if SRT_PAIR_SCAN
	int n_ops, pad_[3];
	int4 op_hdr[CONST_SHAPES];   // {type, index of A, index of B or -1, -}
	// spheres: {c.x A, c.x B, c.y A, c.y B} {c.z A, c.z B, -r*r A, -r*r B} {-}
	// planes:  {p0.x A, p0.x B, p0.y A, p0.y B} {p0.z A, p0.z B, n.x A, n.x B} {n.y A, n.y B, n.z A, n.z B}
	float4 op[CONST_SHAPES][3];
#endif
};

// A kernel launch covers `num_launches` consecutive launches of the reference's kernel `render` that differ in
// nothing but `time` (srt_render_batch): one persistent grid pulls items of launch 0, then launch 1, ..., so the
// ragged end of one launch -- the last long paths, and sweeps with only a few parked rays per warp -- is filled with
// the fresh rays of the next instead of idling lanes (a 1080p x 4-sample launch is 4.5 % slower per sample than a
// 64-sample one; an eighth of it, one GPU's share under tile sharding, 25 %).
constexpr int MAX_BATCH = 16;
// Work items (launch, pixel, sample) of one kernel.  Item indices are 32 bits; the global cursor they are dealt from
// is 64 bits, so overshooting it is harmless, and base + lane (base < MAX_ITEMS, lane < 32) cannot wrap either.
constexpr unsigned long long MAX_ITEMS = 0xffffff00ull;
struct RenderParams {
	int width, height, num_samples, num_bounces;
	float aspect_ratio, fov_scale;
	int show_normals;
	float c2w[16];  // column-major
	int num_launches;
	unsigned int items_per_launch;  // total_pixels * num_samples
	uint32_t times[MAX_BATCH];      // RenderData::time of each launch
	// row-band tiling (srt_set_row_bands)
	int band_h, band_i, band_n;
	int my_rows;               // number of rows this launch renders
	unsigned int total_pixels; // my_rows * width
	unsigned int total_items;  // num_launches * items_per_launch: item = (launch * total_pixels + local_pixel) * num_samples + sample
	float inv_ns;              // 1/num_samples when that is exact (num_samples a power of two), else 0
	int uv_max_tris;           // models of at most this many triangles are swept with the two-strip filter
#if SRT_FASTDIV
	// n / d for ANY 32-bit n without a division (fastdiv.hpp): d = items_per_launch, num_samples, width
	uint32_t dv_m[3], dv_s[3];
#endif
};

struct Counters {
	unsigned long long samples, bounces, tri_tests, aabb_pass, hits, sky;
};

struct Hit {
	float t;
	int shape;  // index into the shape list, -1 = miss
	int tri;    // SoA triangle index of the winning triangle (models)
};

#ifndef SRT_PHASE_PATIENCE
#define SRT_PHASE_PATIENCE 1
#endif
constexpr int PHASE_PATIENCE = SRT_PHASE_PATIENCE;
#ifndef SRT_SHADE_SYNC  // lanes ready to shade wait for parked ones unless this many are ready (0 = never wait: the default)
#define SRT_SHADE_SYNC 0
#endif
constexpr int SHADE_SYNC = SRT_SHADE_SYNC;
// Models with at most this many triangles are intersected inline during the shape scan; larger
// ones park the lane until the warp runs a dense triangle phase (see render_kernel).
constexpr int INLINE_MODEL_TRIS = 32;

// One ray x triangle test: Moller-Trumbore (render.cl:243-275) on pre-transformed (v0, e1, e2), split in two.
// tri_filter is a division-free conservative reject: with x = su * sign(det),
//   x >  |det| * (1 + 1e-6)  =>  the exact u = (1/det) * su is certainly > 1
//   x < -1e-6                =>  the exact u is certainly < 0 (and cannot underflow to -0)
// so a triangle is dropped only when the exact test is certain to fail its u range check (proof in
// DESIGN.md "Triangle filter").  Everything that survives runs tri_exact, the reference arithmetic.
__device__ __forceinline__ bool tri_filter(const float4 v0, const float4 e1, const float4 e2, vec3 o, vec3 d) {
	vec3 h = cross(d, xyz(e2));
	vec3 s = o - xyz(v0);
	const float2 ds = dot_x2(h, xyz(e1), s);  // {dot(e1, h), dot(s, h)} as one packed chain (products commute exactly)
	float det = ds.x, su = ds.y;
	float x = __int_as_float(__float_as_int(su) ^ (__float_as_int(det) & 0x80000000));
	float lim = fabsf(det) * 1.000001f;
	return !(x > lim || x < -1e-6f);
}
__device__ __forceinline__ void tri_exact(const float4 v0, const float4 e1, const float4 e2, vec3 o, vec3 d, int shape,
                                          int tri, Hit &hit) {
	vec3 h = cross(d, xyz(e2));
	vec3 s = o - xyz(v0);
	// the four dot products as two packed FP32x2 chains: {e1.h, s.h} and {d.q, e2.q} (products commute exactly, the
	// accumulation order x, y, z is dot()'s)
	const float2 ds = dot_x2(h, xyz(e1), s);
	float det = ds.x;
	// det == 0 (render.cl:253) needs no test of its own: then f = +-inf and t below is +-inf or NaN,
	// which can never satisfy t < hit.t
	float f = rcp_(det);
	float u = f * ds.y;
	if (u < 0.0f || u > 1.0f) return;
	vec3 q = cross(s, xyz(e1));
	const float2 dq = dot_x2(q, d, xyz(e2));
	float v = f * dq.x;
	if (v < 0.0f || u + v > 1.0f) return;
	float t = f * dq.y;
	if (t > 0.0f && t < hit.t) {
		hit.t = t;
		hit.shape = shape;
		hit.tri = tri;
	}
}
// The filter of the dense sweep decides the same u-range question from PRE-MULTIPLIED operands.  By the cyclic
// identities of the triple product,
//   det = e1 . (d x e2) = d . n',  n' = e2 x e1           su = (o - v0) . (d x e2) = e2 . (o x d) - d . m,  m = e2 x v0
// so with n', m stored per triangle and c = o x d computed once per ray, det and su cost 9 multiply-adds instead
// of the 22 operations of tri_filter.  These values differ from the reference's det / su by rounding only, by at
// most   |d det| <= 22u |e1||e2|   and   |d su| <= 25u |e2| (|o| + |v0|)   (u = 2^-24; DESIGN.md "Triangle
// filter" derives the bounds), so the reject thresholds are widened by
//   M = g * R + 2e-6,   g = 48u |e2|_1 (per triangle),   R = |o|_1 + K (per ray),   K = max_model(|v0|_1 + 3 |e1|_1)
// which covers both error terms (the det term three times over, so that a wrong sign of a near-zero det is
// harmless too).  A triangle is dropped only when the reference's u test is certain to fail; NaNs compare false
// and survive; every survivor runs tri_exact on the reference operands.
//   record (40 B): a = {n'.x, m.x}  b = {n'.y, m.y}  c = {n'.z, m.z}  e = {e2.x, e2.y}  f = {e2.z, g}
// (n' and m interleaved: the layout of an earlier form of the sweep that evaluated {det, t} as one packed chain per ray;
// the present form -- two RAYS per packed operation, tri_filter_sweep below -- reads every field as a scalar, so the
// order is immaterial)
struct TriFlt {
	float2 a, b, c, e, f;
};
#ifndef SRT_MARGIN_SCALE  // developer mutation test only: 0 removes the error margins of the sweep filter
#define SRT_MARGIN_SCALE 1.0f
#endif
// The margin is formed once per triangle and tile from the LARGEST R among the parked rays (M = g * Rmax + 2e-6):
// a larger margin only lets more triangles through.
// m: the margin M of this triangle for the phase (g * Rmax + 2e-6, formed once per tile); cz = r2.x
// The sweep evaluates the filter for TWO RAYS at once: every floating-point operation below is one packed FP32x2
// instruction whose halves belong to ray A and ray B, with the ray operands as register pairs {A, B} (loop-invariant over
// a lane's triangles) and the triangle operand as ONE scalar register broadcast to both halves (SASS operand form
// `R.F32`).  That is the cheapest form an FMA can take on this machine: scripts/microbench/fma_operands.cu measures 2.16
// cycles per FFMA2 (1.08 per FMA) for {shared pair, broadcast scalar, accumulator pair} against 1.31 for a scalar FFMA
// with one shared operand, 3.05 for an FFMA2 of three distinct pairs -- and 2.8 per instruction when packed and scalar
// FMAs alternate, which is what the previous form of this loop ({det, t} packed per ray, the rest scalar) did.
// Each half performs exactly the operations the scalar filter performed, in the same order, so its decisions -- and the
// bound derived in DESIGN 4.1, restated operation for operation in oracle/filter_check.c -- are unchanged.
__device__ __forceinline__ float2 bc(float x) { return make_float2(x, x); }
struct RayPair {  // the operands of two parked rays, interleaved: {A, B} per component; c = o x d
	float2 dx, dy, dz, cx, cy, cz;
};
// m: the margin M of this triangle for the phase (g * Rmax + 2e-6, formed once per tile).
// Returns bit 0: ray A survives, bit 1: ray B survives.
__device__ __forceinline__ uint32_t tri_filter_sweep(const TriFlt &r, const float m, const RayPair &p) {
	// det = d . n',  -t = d . (-m)   (negating every product negates the correctly rounded sum exactly)
	const float2 det = __ffma2_rn(p.dz, bc(r.c.x), __ffma2_rn(p.dy, bc(r.b.x), __fmul2_rn(p.dx, bc(r.a.x))));
	const float2 nt = __ffma2_rn(p.dz, bc(-r.c.y), __ffma2_rn(p.dy, bc(-r.b.y), __fmul2_rn(p.dx, bc(-r.a.y))));
	const float2 su = __ffma2_rn(bc(r.f.x), p.cz, __ffma2_rn(bc(r.e.y), p.cy, __ffma2_rn(bc(r.e.x), p.cx, nt)));
	// x = su sign(det) in [-M, |det| (1 + 2e-6) + M]  <=>  |su - det k| <= |det| k + M  with k = (1 + 2e-6) / 2: the
	// interval test as a distance from its centre.  Same decision up to a few u |det| of rounding in the centre and
	// half-width, which the slack in k and M absorbs; centre and half-width are one FMA each (the filter
	// is not render.cl arithmetic, so nothing forbids fusing: one rounding instead of two only tightens the bound).
	// A NaN fails `>` and survives.
	const float2 diff = __ffma2_rn(make_float2(-det.x, -det.y), bc(0.500001f), su);
	const float2 w = __ffma2_rn(make_float2(fabsf(det.x), fabsf(det.y)), bc(0.500001f), bc(m));
	return (fabsf(diff.x) > w.x ? 0u : 1u) | (fabsf(diff.y) > w.y ? 0u : 2u);
}
// Two-strip filter for models with few, large triangles.  The u strip alone lets through every ray that crosses the
// infinite band between the triangle's edge e2 and its parallel through v1 -- on a ~1k-triangle mesh about 18 pairs for
// each real hit, and pushing / exact-testing those survivors costs more than the sweep itself.  The v range is bounded
// the same way: with m1 = e1 x v0,
//   v det = d . ((o - v0) x e1) = d . m1 - e1 . (o x d)
// so -v det =: sv costs six more multiply-adds, evaluated TOGETHER with su as packed FP32x2 chains:
//   {-t, -t1} = d . {-m, -m1}      {su, sv} = c . {e2, e1} + {-t, -t1}      (c = o x d)
// and the pair is dropped when u or v is certainly outside [0, 1] (then render.cl:261 / :266 rejects it: v > 1 with
// u >= 0 gives u + v > 1).  The error bounds are those of the u test with e1 in the place of e2, so ONE margin
// M = g R + 2e-6 with g = 48u max(|e1|_1, |e2|_1) and K = max(|v0|_1 + 3 max(|e1|_1, |e2|_1)) covers both
// (oracle/filter_check.c checks this filter too: 1e8 adversarial pairs, no wrong reject).  Survivors: ~2 per real hit.
//   record (80 B): q0 = {n'.x, n'.y, n'.z, g}  q1 = {-m.x, -m1.x, -m.y, -m1.y}  q2 = {-m.z, -m1.z, e2.x, e1.x}
//                  q3 = {e2.y, e1.y, e2.z, e1.z}  (+16 B pad: the 80 B stride keeps the lanes' LDS.128 conflict-free)
// evaluated for two rays per packed operation like tri_filter_sweep (per ray, per half: det, -t, -t1 = d . (-m1), su, sv)
struct TriUV {
	float4 q0, q1, q2, q3;
};
__device__ __forceinline__ uint32_t tri_filter_sweep_uv(const TriUV &r, const float m, const RayPair &p) {
	// two rays per operation, like tri_filter_sweep; per half the operations and their order are the scalar filter's
	const float2 det = __ffma2_rn(p.dz, bc(r.q0.z), __ffma2_rn(p.dy, bc(r.q0.y), __fmul2_rn(p.dx, bc(r.q0.x))));
	const float2 nt = __ffma2_rn(p.dz, bc(r.q2.x), __ffma2_rn(p.dy, bc(r.q1.z), __fmul2_rn(p.dx, bc(r.q1.x))));   // d . (-m)
	const float2 nt1 = __ffma2_rn(p.dz, bc(r.q2.y), __ffma2_rn(p.dy, bc(r.q1.w), __fmul2_rn(p.dx, bc(r.q1.y))));  // d . (-m1)
	const float2 su = __ffma2_rn(p.cz, bc(r.q3.z), __ffma2_rn(p.cy, bc(r.q3.x), __ffma2_rn(p.cx, bc(r.q2.z), nt)));   // c . e2 - t
	const float2 sv = __ffma2_rn(p.cz, bc(r.q3.w), __ffma2_rn(p.cy, bc(r.q3.y), __ffma2_rn(p.cx, bc(r.q2.w), nt1)));  // c . e1 - t1
	const float2 du = __ffma2_rn(make_float2(-det.x, -det.y), bc(0.500001f), su);  // u det - det k
	const float2 dv = __ffma2_rn(det, bc(0.500001f), sv);                          // -(v det - det k)
	const float2 w = __ffma2_rn(make_float2(fabsf(det.x), fabsf(det.y)), bc(0.500001f), bc(m));
	// NaNs fail `>` and survive
	return ((fabsf(du.x) > w.x || fabsf(dv.x) > w.x) ? 0u : 1u) | ((fabsf(du.y) > w.y || fabsf(dv.y) > w.y) ? 0u : 2u);
}
__device__ __forceinline__ void test_triangle(const float4 v0, const float4 e1, const float4 e2, vec3 o, vec3 d,
                                              int shape, int tri, Hit &hit) {
	if (tri_filter(v0, e1, e2, o, d)) tri_exact(v0, e1, e2, o, d, shape, tri, hit);
}

// ---- optional BVH traversal (a labelled extension outside the parity path: srt_set_accel, bvh_build.hpp) ----------
// Same exact test, same operands (tri_hot) as the brute-force loop; only the SET of triangles tested differs: those
// in leaves whose (padded) boxes the ray enters no farther than the closest hit so far.  Closest hit wins; on equal t
// the lowest triangle index of THIS model wins and a hit of an earlier shape is kept (render.cl:332 with its strict
// `<`, evaluated in triangle order) -- so traversal order does not matter.
constexpr int BVH_STACK = 48;
__device__ __forceinline__ void tri_exact_tie(const float4 v0, const float4 e1, const float4 e2, vec3 o, vec3 d, int shape,
                                              int tri, Hit &hit) {
	vec3 h = cross(d, xyz(e2));
	float det = dot(xyz(e1), h);
	float f = rcp_(det);
	vec3 s = o - xyz(v0);
	float u = f * dot(s, h);
	if (u < 0.0f || u > 1.0f) return;
	vec3 q = cross(s, xyz(e1));
	float v = f * dot(d, q);
	if (v < 0.0f || u + v > 1.0f) return;
	float t = f * dot(xyz(e2), q);
	if (t > 0.0f && (t < hit.t || (t == hit.t && hit.shape == shape && tri < hit.tri))) {
		hit.t = t;
		hit.shape = shape;
		hit.tri = tri;
	}
}
// entry distance of the ray into a box, or +inf when it misses it or enters beyond tmax.  fminf / fmaxf drop NaNs
// (0 * inf on a slab boundary), and the exit distance is widened by 4 ulp: a box is never culled by rounding.
__device__ __forceinline__ float bvh_slab(const float4 lo, const float4 hi, vec3 o, vec3 inv, float tmax) {
	const float x0 = (lo.x - o.x) * inv.x, x1 = (hi.x - o.x) * inv.x;
	const float y0 = (lo.y - o.y) * inv.y, y1 = (hi.y - o.y) * inv.y;
	const float z0 = (lo.z - o.z) * inv.z, z1 = (hi.z - o.z) * inv.z;
	const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
	const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax)) * 1.0000005f;
	return tn <= tf ? tn : __int_as_float(0x7f800000);
}
template <bool COUNT>
__device__ __forceinline__ void bvh_leaf(const DevScene &sc, int first, int n, vec3 o, vec3 d, int shape, Hit &hit,
                                         Counters &cnt) {
	for (int k = 0; k < n; ++k) {
		const int tri = __ldg(&sc.bvh_order[first + k]);
		const float4 *tp = sc.tri_hot + 3 * (size_t)tri;
		if (COUNT) cnt.tri_tests += 1;
		tri_exact_tie(__ldg(tp), __ldg(tp + 1), __ldg(tp + 2), o, d, shape, tri, hit);
	}
}
template <bool COUNT>
__device__ __noinline__ void bvh_traverse(const DevScene &sc, int root, vec3 o, vec3 d, vec3 inv, int shape, Hit &hit,
                                          Counters &cnt) {
	const float INF = __int_as_float(0x7f800000);
	int stack[BVH_STACK];
	int sp = 0;
	int node = root;
	for (;;) {
		const float4 a = __ldg(&sc.bvh_nodes[4 * (size_t)node + 0]), b = __ldg(&sc.bvh_nodes[4 * (size_t)node + 1]);
		const float4 c = __ldg(&sc.bvh_nodes[4 * (size_t)node + 2]), e = __ldg(&sc.bvh_nodes[4 * (size_t)node + 3]);
		const int c0 = __float_as_int(a.w), n0 = __float_as_int(b.w), c1 = __float_as_int(c.w), n1 = __float_as_int(e.w);
		float t0 = n0 < 0 ? INF : bvh_slab(a, b, o, inv, hit.t);
		float t1 = n1 < 0 ? INF : bvh_slab(c, e, o, inv, hit.t);
		if (t0 < INF && n0 > 0) {  // leaves are tested on the spot
			bvh_leaf<COUNT>(sc, c0, n0, o, d, shape, hit, cnt);
			t0 = INF;
		}
		if (t1 < INF && n1 > 0) {
			if (t1 <= hit.t) bvh_leaf<COUNT>(sc, c1, n1, o, d, shape, hit, cnt);
			t1 = INF;
		}
		if (t0 < INF && t1 < INF) {  // two inner children: nearer first, the other waits on the stack
			const bool first0 = t0 <= t1;
			if (sp < BVH_STACK) stack[sp++] = first0 ? c1 : c0;
			node = first0 ? c0 : c1;
		} else if (t0 < INF) {
			node = c0;
		} else if (t1 < INF) {
			node = c1;
		} else {
			if (sp == 0) return;
			node = stack[--sp];
		}
	}
}

// ---- shape scan ------------------------------------------------------------------------------
// reference closest_intersection, render.cl:293-378, as a resumable scan.  Shapes are visited in
// array order starting at `cursor`, and a hit replaces the current one only if strictly closer
// (:306,:332,:356), so the lowest index wins ties.  A model whose AABB test passes (:319, with
// tmax = the closest t so far) and that is too large to intersect inline stops the scan: the function
// returns that shape index (the lane "parks" there) and the caller runs the triangles later, then
// resumes at index + 1.  Returns -1 when the scan reached the end of the list.
// Normal / position of the winner are reconstructed afterwards (finish_hit) instead of at every
// improvement; only the last improvement is observable.
template <bool COUNT, bool PARK, bool MODELS = true, bool BVH = false, bool CONST = false>
__device__ __forceinline__ int scan_shapes(const DevScene &sc, vec3 o, vec3 d, vec3 inv, int cursor, Hit &hit,
                                           Counters &cnt, const ShapeTable *tab = nullptr) {
	for (int i = cursor; i < sc.num_shapes; ++i) {
		const int4 hdr = CONST ? tab->hdr[i] : __ldg(&sc.shape_hdr[i]);
		const float4 a = CONST ? tab->a[i] : __ldg(&sc.shape_a[i]);
		if (hdr.x == SHAPE_SPHERE) {
			// intersect_sphere, :180-204
			vec3 L = xyz(a) - o;
			// dot(L, d) and dot(L, L) as ONE packed FP32x2 chain (each half is the scalar dot's FMUL, FFMA, FFMA)
			const float2 bl = dot_x2(L, d, L);
			float b = bl.x;
			float c = cfma_(-a.w, a.w, bl.y);
			float disc = cfma_(b, b, -c);
			if (disc >= 0.0f) {
				float sq = sqrt_(disc);
				float t = b - sq;
				if (t < 0.0f) t = b + sq;
				if (t >= 0.0f && t < hit.t) {
					hit.t = t;
					hit.shape = i;
				}
			}
		} else if (hdr.x == SHAPE_PLANE) {
			// intersect_plane, :206-221
			const float4 nb = CONST ? tab->b[i] : __ldg(&sc.shape_b[i]);
			vec3 n = xyz(nb);
			// dot(n, d) and dot(n, p0 - o) as one packed chain (the second is only used when the first is not zero)
			const float2 dn = dot_x2(n, d, xyz(a) - o);
			float denom = dn.x;
			if (fabsf(denom) != 0.0f) {
				float t = div_(dn.y, denom);
				if (t >= 0.0f && t < hit.t) {
					hit.t = t;
					hit.shape = i;
				}
			}
		} else if (MODELS && hdr.x == SHAPE_MODEL) {
			// intersection_aabb, :279-290, with tmax = current closest t (:319)
			const float4 bb = __ldg(&sc.shape_b[i]);
			float tmin = 0.0f, tmax = hit.t;
			{
				float t1 = (a.x - o.x) * inv.x, t2 = (bb.x - o.x) * inv.x;
				tmin = max_(tmin, min_(t1, t2));
				tmax = min_(tmax, max_(t1, t2));
				t1 = (a.y - o.y) * inv.y, t2 = (bb.y - o.y) * inv.y;
				tmin = max_(tmin, min_(t1, t2));
				tmax = min_(tmax, max_(t1, t2));
				t1 = (a.z - o.z) * inv.z, t2 = (bb.z - o.z) * inv.z;
				tmin = max_(tmin, min_(t1, t2));
				tmax = min_(tmax, max_(t1, t2));
			}
			if (tmin < tmax) {
				if (BVH) {  // labelled extension: only the triangles of the leaves the ray enters are tested
					const int root = __ldg(&sc.bvh_root[i]);
					if (root >= 0) {
						if (COUNT) cnt.aabb_pass += 1;
						bvh_traverse<COUNT>(sc, root, o, d, inv, i, hit, cnt);
						continue;
					}
				}
				if (COUNT) {
					cnt.aabb_pass += 1;
					cnt.tri_tests += (unsigned)hdr.w;
				}
				if (PARK && hdr.w > INLINE_MODEL_TRIS) return i;
				// brute-force triangle loop, :324-350
				const float4 *tp = sc.tri_hot + 3 * (size_t)hdr.z;
				for (int k = 0; k < hdr.w; ++k, tp += 3)
					test_triangle(__ldg(tp), __ldg(tp + 1), __ldg(tp + 2), o, d, i, hdr.z + k, hit);
			}
		}
	}
	return -1;
}

#if SRT_PAIR_SCAN
// closest_intersection (render.cl:293-378) for a scene of spheres and planes only, two shapes per trip of the loop.
// Per half, every operation is the one scan_shapes performs for that shape, in the same order (L = c - o; the dot
// products' FMUL, FFMA, FFMA; c = -r*r + L.L with the product rounded at upload by the same single multiplication;
// b*b as a packed product whose halves feed two SCALAR adds -- ptxas contracts a packed multiply feeding a packed add
// whatever --fmad says); the decisions are taken shape A first, then shape B, so hits replace each other in
// array order exactly as in the sequential loop (strict `<`: the lower index wins a tie).  The loop runs for the whole
// warp (a lane without a ray scans a dummy ray whose result is dropped), which makes the loop control, the table loads
// and the type dispatch warp-uniform.
__device__ __forceinline__ void sphere_decide(float disc, float b, int index, Hit &hit) {
	if (disc >= 0.0f) {  // intersect_sphere, :190-203
		float sq = sqrt_(disc);
		float t = b - sq;
		if (t < 0.0f) t = b + sq;
		if (t >= 0.0f && t < hit.t) {
			hit.t = t;
			hit.shape = index;
		}
	}
}
__device__ __forceinline__ void plane_decide(float denom, float num, int index, Hit &hit) {
	if (fabsf(denom) != 0.0f) {  // intersect_plane, :211-220
		float t = div_(num, denom);
		if (t >= 0.0f && t < hit.t) {
			hit.t = t;
			hit.shape = index;
		}
	}
}
__device__ __forceinline__ void scan_pairs(const ShapeTable &tab, vec3 o, vec3 d, Hit &hit) {
	const float2 nox = make_float2(-o.x, -o.x), noy = make_float2(-o.y, -o.y), noz = make_float2(-o.z, -o.z);
	const float2 dx = make_float2(d.x, d.x), dy = make_float2(d.y, d.y), dz = make_float2(d.z, d.z);
	const int n_ops = tab.n_ops;
#pragma unroll 1  // (unrolled by two: 1 % slower)
	for (int k = 0; k < n_ops; ++k) {
		const int4 h = tab.op_hdr[k];
		const float4 q0 = tab.op[k][0], q1 = tab.op[k][1];
		// position - origin for both shapes (x - o is x + (-o) exactly)
		const float2 Lx = __fadd2_rn(make_float2(q0.x, q0.y), nox), Ly = __fadd2_rn(make_float2(q0.z, q0.w), noy);
		const float2 Lz = __fadd2_rn(make_float2(q1.x, q1.y), noz);
		if (h.x == SHAPE_SPHERE) {
			const float2 b = __ffma2_rn(Lz, dz, __ffma2_rn(Ly, dy, __fmul2_rn(Lx, dx)));   // dot(L, d)
			const float2 ll = __ffma2_rn(Lz, Lz, __ffma2_rn(Ly, Ly, __fmul2_rn(Lx, Lx)));  // dot(L, L)
			const float2 c = __fadd2_rn(make_float2(q1.z, q1.w), ll);                       // -r*r + dot(L, L), :187
			const float2 bb = __fmul2_rn(b, b);
			const float disc_a = __fadd_rn(bb.x, -c.x), disc_b = __fadd_rn(bb.y, -c.y);     // b*b - c, :188 (two roundings)
			// (both square roots through one packed copy of sqrt's fast path, evaluated before the `disc >= 0` tests,
			// measured 5 % SLOWER than this: most ray x sphere pairs miss, and a miss costs a compare and a branch here)
#if SRT_PAIR_ANY
			const bool ca = disc_a >= 0.0f, cb = h.z >= 0 && disc_b >= 0.0f;
			if (ca || cb) {  // one test for the common case that the ray misses both
				if (ca) sphere_decide(disc_a, b.x, h.y, hit);
				if (cb) sphere_decide(disc_b, b.y, h.z, hit);
			}
#else
			sphere_decide(disc_a, b.x, h.y, hit);
			if (h.z >= 0) sphere_decide(disc_b, b.y, h.z, hit);
#endif
		} else {
			const float4 q2 = tab.op[k][2];
			const float2 nx = make_float2(q1.z, q1.w), ny = make_float2(q2.x, q2.y), nz = make_float2(q2.z, q2.w);
			const float2 denom = __ffma2_rn(nz, dz, __ffma2_rn(ny, dy, __fmul2_rn(nx, dx)));  // dot(n, d)
			const float2 num = __ffma2_rn(nz, Lz, __ffma2_rn(ny, Ly, __fmul2_rn(nx, Lx)));    // dot(n, p0 - o)
			plane_decide(denom.x, num.x, h.y, hit);
			if (h.z >= 0) plane_decide(denom.y, num.y, h.z, hit);
		}
	}
}
#endif

// Position and shading normal of the winning hit (render.cl:311-312, :337-343, :361-362) followed by
// the front-face flip (:372-375).
// (finish_hit_at: the same from a position already computed as origin + direction * t, :311 / :337 / :361)
__device__ __forceinline__ void finish_hit_at(const DevScene &sc, const Hit &hit, vec3 pos, vec3 d, vec3 &n, bool &front,
                                              int &material);
__device__ __forceinline__ void finish_hit(const DevScene &sc, const Hit &hit, vec3 o, vec3 d, vec3 &pos,
                                           vec3 &n, bool &front, int &material) {
	pos = cfma3(d, hit.t, o);
	finish_hit_at(sc, hit, pos, d, n, front, material);
}
__device__ __forceinline__ void finish_hit_at(const DevScene &sc, const Hit &hit, vec3 pos, vec3 d, vec3 &n, bool &front,
                                              int &material) {
	const int4 hdr = __ldg(&sc.shape_hdr[hit.shape]);
	const float4 a = __ldg(&sc.shape_a[hit.shape]);
	material = hdr.y;
	if (hdr.x == SHAPE_SPHERE) {
		vec3 r = pos - xyz(a);
		n = mk(div_(r.x, a.w), div_(r.y, a.w), div_(r.z, a.w));
	} else if (hdr.x == SHAPE_PLANE) {
		n = xyz(__ldg(&sc.shape_b[hit.shape]));
	} else {
		// barycentric_weights, :223-241 (weights come back rotated: (w2, w0, w1))
		vec3 v0 = xyz(__ldg(&sc.tri_hot[3 * (size_t)hit.tri + 0]));
		vec3 e1 = xyz(__ldg(&sc.tri_hot[3 * (size_t)hit.tri + 1]));
		vec3 e2 = xyz(__ldg(&sc.tri_hot[3 * (size_t)hit.tri + 2]));
		vec3 v2 = pos - v0;
		float d00 = dot(e1, e1), d01 = dot(e1, e2), d11 = dot(e2, e2);
		float d20 = dot(v2, e1), d21 = dot(v2, e2);
		float denom = cfma_(d00, d11, -(d01 * d01));
		float w0 = div_(cfma_(d11, d20, -(d01 * d21)), denom);
		float w1 = div_(cfma_(d00, d21, -(d01 * d20)), denom);
		float w2 = (1.0f - w0) - w1;
		vec3 n0 = xyz(__ldg(&sc.tri_n[3 * hit.tri + 0]));
		vec3 n1 = xyz(__ldg(&sc.tri_n[3 * hit.tri + 1]));
		vec3 n2 = xyz(__ldg(&sc.tri_n[3 * hit.tri + 2]));
		// n0*w2 + n1*w0 + n2*w1  (:341 with the rotated weights)
		vec3 ns = mk(cfma_(n2.x, w1, cfma_(n1.x, w0, n0.x * w2)), cfma_(n2.y, w1, cfma_(n1.y, w0, n0.y * w2)),
		             cfma_(n2.z, w1, cfma_(n1.z, w0, n0.z * w2)));
		// transform_mat(model->transform, n, false), :342 (model matrix, w = 0)
		const float4 m0 = __ldg(&sc.model_xf[4 * hit.shape + 0]);
		const float4 m1 = __ldg(&sc.model_xf[4 * hit.shape + 1]);
		const float4 m2 = __ldg(&sc.model_xf[4 * hit.shape + 2]);
		const float4 m3 = __ldg(&sc.model_xf[4 * hit.shape + 3]);
		vec3 nt = mk(cfma_(m3.x, 0.0f, cfma_(m2.x, ns.z, cfma_(m1.x, ns.y, m0.x * ns.x))),
		             cfma_(m3.y, 0.0f, cfma_(m2.y, ns.z, cfma_(m1.y, ns.y, m0.y * ns.x))),
		             cfma_(m3.z, 0.0f, cfma_(m2.z, ns.z, cfma_(m1.z, ns.y, m0.z * ns.x))));
		n = normalize(nt);
	}
	front = dot(n, d) < 0.0f;
	if (!front) n = -n;
}

// sky_box, render.cl:380-394, with read_imagef(linear, clamp-to-edge, normalised) done as a manual
// FP32 bilinear fetch (the texture unit's 9-bit weights would break parity with a CPU OpenCL device).
__device__ __forceinline__ vec3 sky_box(const DevScene &sc, vec3 d) {
	vec3 sun_dir = mk(sc.sun_dir[0], sc.sun_dir[1], sc.sun_dir[2]);
	float sd = max_(dot(d, -sun_dir), 0.0f);
	float pw = pow_(sd, sc.sun_focus);
	vec3 sun = (mk(sc.sun_color[0], sc.sun_color[1], sc.sun_color[2]) * pw) * sc.sun_intensity;
	float u = cfma_(atan2pi_(d.z, d.x), 0.5f, 0.5f);
	float v = cfma_(d.y, 0.5f, 0.5f);
	// read_imagef: the arithmetic inside the builtin stays fused
	const int w = sc.sky_w, h = sc.sky_h;
	float fu = fma_(u, (float)w, -0.5f), fv = fma_(v, (float)h, -0.5f);
	float flu = floorf(fu), flv = floorf(fv);
	float a = fu - flu, b = fv - flv;
	int i0 = (int)flu, j0 = (int)flv;
	int i1 = i0 + 1, j1 = j0 + 1;
	i0 = min(max(i0, 0), w - 1);
	i1 = min(max(i1, 0), w - 1);
	j0 = min(max(j0, 0), h - 1);
	j1 = min(max(j1, 0), h - 1);
	const float4 t00 = __ldg(&sc.sky[(size_t)j0 * w + i0]);
	const float4 t10 = __ldg(&sc.sky[(size_t)j0 * w + i1]);
	const float4 t01 = __ldg(&sc.sky[(size_t)j1 * w + i0]);
	const float4 t11 = __ldg(&sc.sky[(size_t)j1 * w + i1]);
	float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
	// red and green as one packed FP32x2 chain (the texels' .xy arrive as register pairs from the 16-byte loads)
	const float2 rg = __ffma2_rn(make_float2(w11, w11), make_float2(t11.x, t11.y),
	                             __ffma2_rn(make_float2(w01, w01), make_float2(t01.x, t01.y),
	                                        __ffma2_rn(make_float2(w10, w10), make_float2(t10.x, t10.y),
	                                                   __fmul2_rn(make_float2(w00, w00), make_float2(t00.x, t00.y)))));
	vec3 tex = mk(rg.x, rg.y, fma_(w11, t11.z, fma_(w01, t01.z, fma_(w10, t10.z, w00 * t00.z))));
	return tex + sun;
}

// camera ray for pixel (gx, gy), render.cl:496-516.  Returns the advanced seed in `seed`.
__device__ __forceinline__ void camera_ray(const RenderParams &p, int gx, int gy, uint32_t &seed, vec3 &o,
                                           vec3 &d) {
	float u0 = random_float(seed);
	float u1 = random_float(seed);
	float ndc_x = div_((float)gx + u0, (float)p.width);
	float ndc_y = div_((float)gy + u1, (float)p.height);
	float sx = (cfma_(2.0f, ndc_x, -1.0f) * p.aspect_ratio) * p.fov_scale;
	float sy = cfma_(-2.0f, ndc_y, 1.0f) * p.fov_scale;
	const float *m = p.c2w;
	o = mk(m[12], m[13], m[14]);
	// matrix_by_vector(camera_to_world, (sx, sy, -1, 0)), :114-120
	vec3 t = mk(cfma_(m[12], 0.0f, cfma_(m[8], -1.0f, cfma_(m[4], sy, m[0] * sx))),
	            cfma_(m[13], 0.0f, cfma_(m[9], -1.0f, cfma_(m[5], sy, m[1] * sx))),
	            cfma_(m[14], 0.0f, cfma_(m[10], -1.0f, cfma_(m[6], sy, m[2] * sx))));
	d = normalize(t);
}

// Scatter at a hit, render.cl:418-462.  Updates o, d, mask; consumes 9 or 10 random numbers.
__device__ __forceinline__ void scatter(const DevScene &sc, int material, vec3 pos, vec3 n, bool front,
                                        uint32_t &seed, vec3 &o, vec3 &d, vec3 &mask, float4 m0, float4 m1) {
	o = pos;
	// random_direction_hemisphere, :156-163
	const float2 g = random_float_normal_x2(seed);  // two draws as packed FP32x2 chains: fewer issue slots
	const float gx = g.x, gy = g.y;
	const float gz = random_float_normal(seed);
	vec3 rd = normalize(mk(gx, gy, gz));
	const float2 nn = dot_x2(n, rd, d);                         // {dot(n, rd), dot(n, d)}: one packed chain
	rd = rd * sign_(nn.x);
	vec3 random_dir = normalize(n + rd);                       // :421
	float k2 = 2.0f * nn.y;                                     // reflect, :139-141: 2 dot(d, n) (products commute exactly)
	vec3 reflected = mk(cfma_(-k2, n.x, d.x), cfma_(-k2, n.y, d.y), cfma_(-k2, n.z, d.z));
	const bool is_metallic = m0.y > random_float(seed);         // :424
	const bool is_specular = m0.z > random_float(seed);         // :425
	vec3 rough = mix3(random_dir, reflected, m0.x);             // :427
	const bool is_transparent = m1.x > random_float(seed);      // :429
	vec3 nd;
	if (!is_transparent) {
		nd = mix3(random_dir, rough, (is_metallic || is_specular) ? 1.0f : 0.0f);  // :432
		vec3 col = xyz(__ldg(&sc.materials[4 * material + 2]));
		float sp = is_specular ? 1.0f : 0.0f;
		mask = mask * mk(mix_(col.x, 1.0f, sp), mix_(col.y, 1.0f, sp), mix_(col.z, 1.0f, sp));  // :436
	} else {
		float ki = 2.0f * dot(rough, n);                        // in_dir = reflect(rough, n), :440
		vec3 in_dir = mk(cfma_(-ki, n.x, rough.x), cfma_(-ki, n.y, rough.y), cfma_(-ki, n.z, rough.z));
		float mu = front ? rcp_(m1.y) : m1.y;             // :442
		const float r0 = front ? m1.z : m1.w;             // shlick's r0 for this mu, precomputed per material (:174-175)
		float cos_theta = min_(1.0f, dot(in_dir, -n));          // :443
		float sin_theta = sqrt_(cfma_(-cos_theta, cos_theta, 1.0f));
		bool reflected_t = mu * sin_theta > 1.0f;               // :446
		if (!reflected_t) reflected_t = schlick_from_r0(r0, cos_theta) > random_float(seed);  // :447 (short-circuit)
		if (reflected_t) {
			nd = rough;                                         // :450
		} else {
			vec3 out_perp = cfma3(n, cos_theta, in_dir) * mu;   // :452
			float kp = -sqrt_(fabsf(1.0f - length_squared(out_perp)));  // :453
			nd = cfma3(n, kp, out_perp);                        // :454
			mask = mask * xyz(__ldg(&sc.materials[4 * material + 2]));  // :457
		}
	}
	d = normalize(nd);                                          // :461
	float sg = sign_(dot(n, d)) * 0.001f;                       // :462
	o = cfma3(n, sg, o);
}

// ---- kernel `render` ---------------------------------------------------------------------------
#ifndef SRT_RENDER_THREADS
#define SRT_RENDER_THREADS 128
#endif
#ifndef SRT_MIN_BLOCKS
#define SRT_MIN_BLOCKS 4
#endif
#ifndef SRT_MIN_BLOCKS_ANALYTIC
#define SRT_MIN_BLOCKS_ANALYTIC 7
#endif
#ifndef SRT_MIN_BLOCKS_WAVEFRONT  // wavefront schedule, scenes without models: 8 x 128 threads x 64 registers = the register file
#define SRT_MIN_BLOCKS_WAVEFRONT 8
#endif
#ifndef SRT_MIN_BLOCKS_BVH
#define SRT_MIN_BLOCKS_BVH 6
#endif
constexpr int RENDER_THREADS = SRT_RENDER_THREADS;

// Dense triangle phase: a register-tiled outer product of (parked rays) x (one model's triangles), in two stages --
// a cheap conservative FILTER over all pairs, then the reference's exact Moller-Trumbore on the few it lets through.
//
//   * the model's 40-byte filter records (TriFlt: n' = e2 x e1, m = e2 x v0, e2, margin scale g) arrive in a per-warp
//     ring of shared-memory tiles filled by 1-D bulk async copies (cp.async.bulk -> SASS UBLKCP) completing on an
//     mbarrier, TILE_STAGES tiles ahead of the consumer: no lane issues a global load for the sweep;
//   * every lane lifts TRIS_PER_LANE triangles of the current tile into REGISTERS (40 B stride: conflict-free
//     LDS.64), so all 32 lanes are busy however few rays are parked;
//   * the parked rays publish their operands once per phase, two rays to a record (12 floats {A, B} x (d, c = o x d));
//     a trip takes one record and every lane runs tri_filter_sweep for BOTH rays on its own triangles, in packed
//     FP32x2 instructions whose halves are the two rays; a lane keeps, per triangle slot, the bit mask of the rays
//     its triangle survived;
//   * after a tile the lanes expand their masks into a per-warp ring of (ray << 27 | triangle) PAIRS, and the ring
//     is drained 32 pairs at a time by ALL lanes: tri_exact on the reference operands (tri_hot, LDG.128 x 3 through
//     L1/L2), folded into best[ray] by a shared-memory atomicMin on (t bits << 32 | triangle + 1), which keeps the
//     closest hit and, on equal t, the lowest triangle index -- the order of the sequential loop, render.cl:324-350.
// Shared-memory traffic is 24 B per ray per tile instead of 40 B per ray per triangle; what bounds the loop is the FMA
// pipe's register-operand bandwidth (scripts/microbench/fma_operands.cu).
#ifndef SRT_TRIS_PER_LANE
#define SRT_TRIS_PER_LANE 4
#endif
#ifndef SRT_TILE_STAGES  // mbarriers per warp = the most stages either record size gets out of the ring (see RING_BYTES)
#define SRT_TILE_STAGES 2
#endif
constexpr int TRIS_PER_LANE = SRT_TRIS_PER_LANE;
constexpr int TILE_TRIS = 32 * TRIS_PER_LANE;
constexpr int TILE_STAGES = SRT_TILE_STAGES;
constexpr int FLT_BYTES = 40;                 // filter record of one triangle (TriFlt)
// Two-strip (u AND v) variant of the filter for models of at most UV_MAX_TRIS triangles (tri_filter_sweep_uv):
// 80-byte records (TriUV), fewer triangles per lane because a record is 16 registers instead of 10.
#ifndef SRT_TRIS_PER_LANE_UV
#define SRT_TRIS_PER_LANE_UV 4
#endif
#ifndef SRT_UV_MAX_TRIS
#define SRT_UV_MAX_TRIS 20000
#endif
constexpr int TRIS_PER_LANE_UV = SRT_TRIS_PER_LANE_UV;
constexpr int TILE_TRIS_UV = 32 * TRIS_PER_LANE_UV;
constexpr int UV_BYTES = 80;
constexpr int UV_MAX_TRIS = SRT_UV_MAX_TRIS;
static_assert(TRIS_PER_LANE_UV <= TRIS_PER_LANE, "the survivor masks are sized for TRIS_PER_LANE slots");
constexpr int TILE_BYTES = TILE_TRIS * FLT_BYTES > TILE_TRIS_UV * UV_BYTES ? TILE_TRIS * FLT_BYTES : TILE_TRIS_UV * UV_BYTES;
// The ring is ONE tile of the larger (two-strip, 80-byte) records = TWO tiles of the one-strip (40-byte) records: the
// one-strip sweep runs two half-size stages, the two-strip sweep a single one (triangle_phase).
constexpr int RING_BYTES = TILE_BYTES;
static_assert(TILE_STAGES == 2 && 2 * TILE_TRIS * FLT_BYTES <= RING_BYTES && TILE_TRIS_UV * UV_BYTES <= RING_BYTES, "tile ring geometry");
constexpr int RAYS_BYTES = 32 * 64;           // 32 rays x {origin, direction} (exact test) + 16 ray pairs x 12 floats (the filter)
constexpr int PAIR_SLOTS = 256;               // ring of filter survivors (ray << 27 | triangle) awaiting the exact test
constexpr int PAIRS_BYTES = PAIR_SLOTS * 4;
constexpr int BEST_BYTES = 32 * 8;            // per parked ray: (t bits << 32 | triangle + 1), minimised atomically
constexpr int RENDER_WARPS = RENDER_THREADS / 32;
constexpr int WARP_SMEM_BYTES = RING_BYTES + RAYS_BYTES + PAIRS_BYTES + BEST_BYTES;
constexpr int MAX_SWEEP_TRIS = 1 << 27;       // a pair packs the triangle's index within its model in 27 bits
constexpr int RENDER_SMEM_BYTES = RENDER_WARPS * (WARP_SMEM_BYTES + TILE_STAGES * 8);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void bulk_load_tile(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
	             "l"(src), "r"(bytes), "r"(bar)
	             : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "WAIT_%=:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@!p bra WAIT_%=;\n"
	    "}\n" ::"r"(bar),
	    "r"(parity)
	    : "memory");
}

// Before shared memory that was just READ by ordinary loads is handed to an asynchronous bulk copy, the loads must have
// completed -- issued is not enough: under shared-memory contention a load can still be queued when the copy's first
// bytes land.  Two measures, both needed in principle and cheap in practice:
//   * a real instruction chain that reads one destination register of EVERY load (an XOR folded three at a time by LOP3)
//     and a store of its result: instructions issue in order and only when their operands are there, so everything after
//     the store is issued after the data of every load has arrived;
//   * fence.proxy.async, which orders this thread's generic-proxy accesses before later async-proxy ones.
__device__ __forceinline__ uint32_t fbits(float x) { return __float_as_uint(x); }
__device__ __forceinline__ void loads_done_before_async_copy(uint32_t acc, uint32_t *sink) {
	*sink = acc;
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Exact test of one filter survivor by whichever lane picked it up: the reference arithmetic (tri_exact) on the
// reference operands, fetched through L1/L2.  A hit is folded into best[ray] with an atomic min on
// (t bits << 32 | triangle + 1): for positive floats the bit pattern orders like the value, so the minimum is the
// closest hit and, among equal t, the lowest triangle index -- what the sequential loop of render.cl:324-350
// with its strict `t_i < tmin` keeps.  best[ray] starts at (hit.t << 32 | 0), so an equal t never replaces the hit
// of an earlier shape either.
__device__ __forceinline__ void exact_pair(const float4 v0, const float4 e1, const float4 e2, const float4 o4,
                                           const float4 d4, unsigned long long *best_slot, int tri_local) {
	Hit h = {__int_as_float(0x7f800000), -1, -1};
	tri_exact(v0, e1, e2, xyz(o4), xyz(d4), 0, tri_local, h);
	if (h.shape == 0) {
		const unsigned long long key = ((unsigned long long)__float_as_uint(h.t) << 32) | (unsigned)(tri_local + 1);
		if (key < *best_slot) atomicMin(best_slot, key);
	}
}

// n, tri_begin, shape: the model being swept (warp-uniform); active: this lane's ray is parked at it.
// UV: which filter the model is swept with (and with it the record size and the tile geometry); two instantiations,
// of which a scene normally exercises one, so the hot instruction footprint stays that of a single phase.
#ifndef SRT_PHASE_ATTR
#define SRT_PHASE_ATTR __forceinline__
#endif
template <bool UV>
__device__ SRT_PHASE_ATTR void triangle_phase(const DevScene &sc, int n, int tri_begin, int shape, bool active, vec3 o,
                                               vec3 d, Hit &hit, unsigned char *wsmem, uint32_t wsmem_s, uint32_t bars_s,
                                               uint32_t &parity, int lane) {
	const unsigned FULL = 0xffffffffu;
	const unsigned char *tiles = wsmem;
	float4 *rays = reinterpret_cast<float4 *>(wsmem + RING_BYTES);
	uint32_t *pairs = reinterpret_cast<uint32_t *>(wsmem + RING_BYTES + RAYS_BYTES);
	unsigned long long *best = reinterpret_cast<unsigned long long *>(wsmem + RING_BYTES + RAYS_BYTES + PAIRS_BYTES);
	constexpr bool uv = UV;
	constexpr int tile_tris = uv ? TILE_TRIS_UV : TILE_TRIS;
	constexpr int rec_bytes = uv ? UV_BYTES : FLT_BYTES;
	constexpr int tile_bytes = tile_tris * rec_bytes;
	// tri_begin is even: both record arrays start 16-byte aligned
	const char *src = uv ? reinterpret_cast<const char *>(sc.tri_uv + 5 * (size_t)tri_begin)
	                     : reinterpret_cast<const char *>(sc.tri_flt + 5 * (size_t)tri_begin);
	const float4 *exact = sc.tri_hot + 3 * (size_t)tri_begin;  // survivors fetch the reference operands from L1/L2
	const int ntiles = (n + tile_tris - 1) / tile_tris;
	// One-strip records: two stages -- tile t + 1 is copied into the stage that held tile t - 1, whose registers the
	// previous sweep has consumed to the last instruction, while tile t is being swept.  Two-strip records fill the
	// ring with one tile, so the stage that is refilled is the one that was just lifted into registers, and the loads
	// have to be known complete first (loads_done_before_async_copy).
	constexpr int stages = uv ? 1 : 2;
	auto issue = [&](int t) {
		const int st = t % stages;
		const uint32_t bytes = ((uint32_t)min(tile_tris, n - t * tile_tris) * rec_bytes + 15u) & ~15u;  // the arrays are padded
		SRT_ASSERT(bytes > 0 && bytes <= (uint32_t)TILE_BYTES && ((size_t)(src + (size_t)t * tile_bytes) & 15) == 0);
		bulk_load_tile(wsmem_s + st * tile_bytes, src + (size_t)t * tile_bytes, bytes, bars_s + st * 8);
	};
	if (lane == 0) issue(0);  // always ONE tile ahead of the sweep
	// the parked rays take DENSE slots 0 .. nrays-1 (slot = rank of the lane among the parked ones)
	const unsigned ray_mask = __ballot_sync(FULL, active);
	const int nrays = __popc(ray_mask);
	const int slot = __popc(ray_mask & ((1u << lane) - 1u));
	const uint32_t ray_bits = nrays >= 32 ? 0xffffffffu : (1u << nrays) - 1u;
	// the filter's operands, two rays to a record: 12 floats {A, B} x (d.x, d.y, d.z, c.x, c.y, c.z), c = o x d, for the
	// rays in slots 2k and 2k + 1; and per ray its origin and direction for the exact test
	float *pairbuf = reinterpret_cast<float *>(rays + 64);  // behind the 32 x 2 float4 of origins and directions
	uint32_t *sink = reinterpret_cast<uint32_t *>(pairbuf + 16 * 12);  // 32 words behind the 16 pair records (loads_done_...)
	if (active) {
		const vec3 c = cross(o, d);
		float *pr = pairbuf + 12 * (slot >> 1) + (slot & 1);
		pr[0] = d.x, pr[2] = d.y, pr[4] = d.z, pr[6] = c.x, pr[8] = c.y, pr[10] = c.z;
		rays[2 * slot] = make_float4(o.x, o.y, o.z, 0.f);
		rays[2 * slot + 1] = make_float4(d.x, d.y, d.z, 0.f);
		best[slot] = (unsigned long long)__float_as_uint(hit.t) << 32;
	}
	if ((nrays & 1) && lane == 0) {  // pad to an even count: a null ray (its survivor bits are masked off)
		float *pr = pairbuf + 12 * (nrays >> 1) + 1;
		pr[0] = pr[2] = pr[4] = pr[6] = pr[8] = pr[10] = 0.f;
	}
	int pair_head = 0, pair_count = 0;  // warp-uniform
	// R = |o|_1 + K of my ray; its maximum over the parked rays (NaN -- a ray the filter cannot decide -- must win)
	float rmax = active ? fabsf(o.x) + fabsf(o.y) + fabsf(o.z) + __ldg(&sc.model_k[shape]) : 0.0f;
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) {
		const float other = __shfl_xor_sync(FULL, rmax, off);
		rmax = (other > rmax || other != other) ? other : rmax;
	}
	__syncwarp();

	// Code size matters here (the kernel's instruction footprint must stay cache resident while warps sit in
	// different phases), so there is ONE instance of the exact test: the loop below drains the ring 32 pairs at a
	// time, and every path that produces survivors -- including a ray the filter cannot decide at all -- feeds
	// the ring.
	for (int t = 0; t <= ntiles; ++t) {
		uint32_t cand[TRIS_PER_LANE];
#pragma unroll
		for (int q = 0; q < TRIS_PER_LANE; ++q) cand[q] = 0;
		if (t < ntiles) {
			const int st = t % stages;
			mbar_wait(bars_s + st * 8, (parity >> st) & 1u);
			parity ^= 1u << st;
			const unsigned char *tile = tiles + st * tile_bytes;
			if (!uv) {
				// my triangles of this tile: slot q holds triangle q*32 + lane
				TriFlt tf[TRIS_PER_LANE];
#pragma unroll
				for (int q = 0; q < TRIS_PER_LANE; ++q) {
					const float2 *r = reinterpret_cast<const float2 *>(tile) + 5 * (q * 32 + lane);  // 40 B stride: conflict-free LDS.64
					tf[q].a = r[0], tf[q].b = r[1], tf[q].c = r[2], tf[q].e = r[3], tf[q].f = r[4];
					tf[q].f.y = fma_(tf[q].f.y, rmax, 2e-6f * SRT_MARGIN_SCALE);  // g -> M
				}
				// the tile now lives in registers; the NEXT tile goes into the other stage -- the one the previous sweep
				// has finished with -- and has this whole sweep to land
				__syncwarp();
				if (lane == 0 && t + 1 < ntiles) issue(t + 1);
				// two parked rays per trip, evaluated TOGETHER (tri_filter_sweep: one packed operation serves both); the
				// slots are dense, an odd count is padded with a null ray whose bit is masked off below.  A lane keeps,
				// per triangle slot, the bit mask of the RAYS its triangle survived -- no vote, no hand-off to an owner lane.
				const float4 *rp = reinterpret_cast<const float4 *>(pairbuf);
				int sh = 0;
				for (int i = 0; i < nrays; i += 2, rp += 3, sh += 2) {  // warp-uniform
					const float4 p0 = rp[0], p1 = rp[1], p2 = rp[2];
					const RayPair pr = {make_float2(p0.x, p0.y), make_float2(p0.z, p0.w), make_float2(p1.x, p1.y),
					                    make_float2(p1.z, p1.w), make_float2(p2.x, p2.y), make_float2(p2.z, p2.w)};
#pragma unroll
					for (int q = 0; q < TRIS_PER_LANE; ++q) cand[q] |= tri_filter_sweep(tf[q], tf[q].f.y, pr) << sh;
				}
			} else {
				// the same sweep with the two-strip filter: 80-byte records, TRIS_PER_LANE_UV triangles per lane
				TriUV tv[TRIS_PER_LANE_UV];
#pragma unroll
				for (int q = 0; q < TRIS_PER_LANE_UV; ++q) {
					const float4 *r = reinterpret_cast<const float4 *>(tile) + 5 * (q * 32 + lane);  // 80 B stride: conflict-free LDS.128
					tv[q].q0 = r[0], tv[q].q1 = r[1], tv[q].q2 = r[2], tv[q].q3 = r[3];
					tv[q].q0.w = fma_(tv[q].q0.w, rmax, 2e-6f * SRT_MARGIN_SCALE);  // g -> M
				}
				// single stage: the next tile overwrites the one just lifted into registers, which has to be literally
				// true first -- the loads above are only ISSUED at this point, and under shared-memory contention (16
				// warps per SM lifting 10 KB each) a load can still be queued when the copy's first bytes land; the lane
				// would then sweep a mixture of two tiles (seen as rare missed hits once the sweep got faster)
				uint32_t acc = 0;
#pragma unroll
				for (int q = 0; q < TRIS_PER_LANE_UV; ++q)  // one register of each of the four loads of the record
					acc ^= fbits(tv[q].q0.x) ^ fbits(tv[q].q1.x) ^ fbits(tv[q].q2.x) ^ fbits(tv[q].q3.x);
				loads_done_before_async_copy(acc, sink + lane);
				__syncwarp();
				if (lane == 0 && t + 1 < ntiles) issue(t + 1);
				const float4 *rp = reinterpret_cast<const float4 *>(pairbuf);
				int sh = 0;
				for (int i = 0; i < nrays; i += 2, rp += 3, sh += 2) {  // warp-uniform
					const float4 p0 = rp[0], p1 = rp[1], p2 = rp[2];
					const RayPair pr = {make_float2(p0.x, p0.y), make_float2(p0.z, p0.w), make_float2(p1.x, p1.y),
					                    make_float2(p1.z, p1.w), make_float2(p2.x, p2.y), make_float2(p2.z, p2.w)};
#pragma unroll
					for (int q = 0; q < TRIS_PER_LANE_UV; ++q) cand[q] |= tri_filter_sweep_uv(tv[q], tv[q].q0.w, pr) << sh;
				}
			}
			// the padding ray's bit, and triangles beyond the end of the list (the tile holds stale shared memory there)
			const int cnt = min(n - t * tile_tris, tile_tris);
#pragma unroll
			for (int q = 0; q < TRIS_PER_LANE; ++q) cand[q] = q * 32 + lane < cnt ? cand[q] & ray_bits : 0u;
		}
		// survivors -> pair ring -> exact tests, 32 at a time; after the last tile (t == ntiles) the ring is emptied
		for (;;) {
			int mine = 0;
#pragma unroll
			for (int q = 0; q < TRIS_PER_LANE; ++q) mine += __popc(cand[q]);
			int total = 0;
			const int room = PAIR_SLOTS - pair_count;
			if (__any_sync(FULL, mine != 0)) {  // (a tile without survivors skips the scan altogether)
				int incl = mine;  // inclusive prefix sum over the lanes
#pragma unroll
				for (int off = 1; off < 32; off <<= 1) {
					const int up = __shfl_up_sync(FULL, incl, off);
					if (lane >= off) incl += up;
				}
				total = __shfl_sync(FULL, incl, 31);
				int rank = incl - mine;  // my first survivor's rank among the warp's
#pragma unroll
				for (int q = 0; q < TRIS_PER_LANE; ++q) {
					uint32_t c = cand[q];  // bit r: my triangle of slot q survived the filter for the ray in slot r
					const uint32_t base = (uint32_t)(t * tile_tris + q * 32 + lane);
					// (one loop over all four masks of a lane was tried: fewer trips, but the slot bookkeeping made each
					// trip dearer -- 9 % slower on the ~1k-triangle meshes; so was an L1 prefetch of the survivors' operands)
					while (c && rank < room) {
						SRT_ASSERT(rank >= 0 && pair_count + rank < PAIR_SLOTS && (int)base < n && __ffs(c) - 1 < nrays);
						pairs[(pair_head + pair_count + rank) & (PAIR_SLOTS - 1)] = base | ((uint32_t)(__ffs(c) - 1) << 27);
						c &= c - 1;
						++rank;
					}
					cand[q] = c;  // what did not fit waits for the next round
				}
				pair_count += min(total, room);
				__syncwarp();
			}
			const bool flush = t == ntiles || total > room;
			while (pair_count >= 32 || (flush && pair_count > 0)) {
				const int m = min(pair_count, 32);
				if (lane < m) {
					const uint32_t pr = pairs[(pair_head + lane) & (PAIR_SLOTS - 1)];
					const int r = pr >> 27, j = pr & (MAX_SWEEP_TRIS - 1);
					SRT_ASSERT(j >= 0 && j < n && r < nrays);
					exact_pair(__ldg(exact + 3 * (size_t)j), __ldg(exact + 3 * (size_t)j + 1), __ldg(exact + 3 * (size_t)j + 2),
					           rays[2 * r], rays[2 * r + 1], best + r, j);
				}
				pair_head = (pair_head + m) & (PAIR_SLOTS - 1);
				pair_count -= m;
				__syncwarp();
			}
			if (total <= room) break;
		}
	}
	SRT_ASSERT(pair_count == 0);
	if (active) {
		const unsigned long long k = best[slot];
		const unsigned tri1 = (unsigned)k;
		SRT_ASSERT(tri1 <= (unsigned)n && (k >> 32) <= (unsigned long long)__float_as_uint(hit.t));
		if (tri1) {
			hit.t = __uint_as_float((unsigned)(k >> 32));
			hit.shape = shape;
			hit.tri = tri_begin + (int)tri1 - 1;
		}
	}
	__syncwarp();
}

// ---- the per-pixel part of kernel `average`, render.cl:525-535 (aces :473-481) -----------------
__device__ __forceinline__ float aces_sqrt(float x) {
	const float a = 2.51f, b = 0.03f, c = 2.43f, dd = 0.59f, e = 0.14f;
	float num = x * cfma_(x, a, b);
	float den = cfma_(x, cfma_(x, c, dd), e);
	float r = div_(num, den);
	r = r > 0.0f ? r : 0.0f;  // clamp; NaN -> 0
	r = r < 1.0f ? r : 1.0f;
	return sqrt_(r);
}
__device__ __forceinline__ uchar4 argb_pixel(float4 c, float steps) {
	float r = aces_sqrt(div_(c.x, steps)) * 255.0f;
	float g = aces_sqrt(div_(c.y, steps)) * 255.0f;
	float b = aces_sqrt(div_(c.z, steps)) * 255.0f;
	// uchar4(255, r, g, b): A,R,G,B byte order, float -> uchar by truncation
	return make_uchar4(255, (unsigned char)(int)r, (unsigned char)(int)g, (unsigned char)(int)b);
}
// Start the camera path of work item `item` = local_pixel * num_samples + sample (render.cl:488-516).
__device__ __forceinline__ void start_path(const RenderParams &p, unsigned int item, uint32_t &seed, vec3 &o, vec3 &d) {
#if SRT_FASTDIV
	const unsigned int launch = p.num_launches > 1 ? fast_div(item, p.dv_m[0], p.dv_s[0]) : 0u;
#else
	const unsigned int launch = p.num_launches > 1 ? item / p.items_per_launch : 0u;
#endif
	const unsigned int in_launch = item - launch * p.items_per_launch;
	SRT_ASSERT(launch < (unsigned)p.num_launches && launch < (unsigned)MAX_BATCH && item < p.total_items);
#if SRT_FASTDIV
	const unsigned int lp = fast_div(in_launch, p.dv_m[1], p.dv_s[1]);  // local pixel
	const unsigned int sample = in_launch - lp * (unsigned)p.num_samples;
	const int row = (int)fast_div(lp, p.dv_m[2], p.dv_s[2]);
#else
	const unsigned int lp = in_launch / (unsigned)p.num_samples;  // local pixel
	const unsigned int sample = in_launch - lp * (unsigned)p.num_samples;
	const int row = (int)(lp / (unsigned)p.width);
#endif
	const int gx = (int)(lp - (unsigned)row * (unsigned)p.width);
	const int gy = p.band_n > 1 ? ((row / p.band_h) * p.band_n + p.band_i) * p.band_h + (row % p.band_h) : row;
	const uint32_t pix = (uint32_t)gx + (uint32_t)gy * (uint32_t)p.width;
	seed = (sample + pix * (uint32_t)p.num_samples) * p.times[launch] * 5304u;
	camera_ray(p, gx, gy, seed, o, d);
}

// Per-warp shared-memory queues of the analytic builds (MODE_ANALYTIC, MODE_SMALL_MODELS).
// The two things only a minority of lanes needs in any one trip -- a fresh camera path (about one lane in five)
// and a sky-box evaluation (paths that just escaped) -- are not executed by that minority under divergence:
// camera paths are generated 32 at a time into a ring the finishing lanes pop from, and escaped paths push
// {item, radiance, throughput, direction} into a ring that is evaluated 32 at a time.  Both run with all lanes
// active.  Results go to scratch[item], so the order in which records are flushed does not matter.
#ifndef SRT_BIG_SKYQ
#define SRT_BIG_SKYQ 0
#endif
constexpr int QUEUE_SLOTS = 64;  // < 32 left over + <= 32 pushed per trip
constexpr int RAYQ_WORDS = 5;    // item, seed, d.xyz          (origin = camera position)
constexpr int SKYQ_WORDS = 10;   // item, color.xyz, mask.xyz, d.xyz   (wavefront schedule: grouped into two float4 + two words, SRT_RING_V4)
// Wavefront schedule of the queue builds (SRT_WAVEFRONT, render_wavefront): hits go through a third ring, so that hit
// shading -- 60 % of the instructions of an analytic scene -- always runs with 32 lanes instead of the ~25 whose ray hit
// something in that trip.
#ifndef SRT_WAVEFRONT
#define SRT_WAVEFRONT 1
#endif
constexpr int HITQ_WORDS = 16;   // item, seed, position.xyz, d.xyz, mask.xyz, color.xyz, shape << 8 | bounce, triangle (grouped into three
                                 // float4 + three / four words under SRT_RING_V4: render_wavefront)
constexpr int QUEUE_WARP_BYTES = (RAYQ_WORDS + SKYQ_WORDS) * QUEUE_SLOTS * 4;  // plain schedule: ray and sky rings
constexpr int QUEUE_SMEM_BYTES = (SRT_RENDER_THREADS / 32) * QUEUE_WARP_BYTES;
// wavefront schedule: a ONE-batch ray ring (see its refill step), the sky ring, and the hit ring -- whose triangle word
// only the builds that know models carry.  Without it a warp needs 7040 bytes, and 8 CTAs x (4 x 7040 + 1024 reserved)
// are exactly the 233472 bytes of shared memory an SM has: scenes without models run 8 CTAs (32 warps) per SM.
constexpr int RAYQ_SLOTS = 32;
__host__ __device__ constexpr int wavefront_warp_bytes(bool models) {
	return (RAYQ_WORDS * RAYQ_SLOTS + SKYQ_WORDS * QUEUE_SLOTS + (models ? HITQ_WORDS : HITQ_WORDS - 1) * QUEUE_SLOTS) * 4;
}
__host__ __device__ constexpr int wavefront_smem_bytes(bool models) { return (SRT_RENDER_THREADS / 32) * wavefront_warp_bytes(models); }
constexpr int BIG_SKYQ_BYTES = SRT_BIG_SKYQ ? (SRT_RENDER_THREADS / 32) * SKYQ_WORDS * QUEUE_SLOTS * 4 : 0;

// MODE selects the build: MODE_ANALYTIC for scenes without any model shape (no triangle code at all),
// MODE_SMALL_MODELS when every model is small enough to be intersected inline during the scan (no
// parking, no shared memory), MODE_BIG_MODELS for the full machine.  The first two need fewer registers,
// i.e. more resident warps for the latency-bound analytic path.
// MODE_BVH (srt_set_accel(SRT_ACCEL_BVH), NOT the parity path) is the SMALL_MODELS machine with big models traversed
// through their hierarchy inside the scan.
// MODE_ANALYTIC_CONST is MODE_ANALYTIC for scenes of at most CONST_SHAPES shapes, scanning the constant-bank table.
enum { MODE_ANALYTIC = 0, MODE_SMALL_MODELS = 1, MODE_BIG_MODELS = 2, MODE_BVH = 3, MODE_ANALYTIC_CONST = 4 };
// ---- wavefront schedule of the queue builds (analytic scenes, small models, BVH) ----------------------------------
// One trip of a warp:
//   1. SHADE   if 32 hit records are queued (or the frame is running dry): every lane pops one {item, seed, position,
//              direction, throughput, radiance, shape, bounce} and runs finish_hit + emission + scatter (render.cl:406-462).
//              A lane whose path goes on now holds its next ray.
//   2. REFILL  the other lanes pop fresh camera rays (generated 32 at a time, all lanes, as before).
//   3. SCAN    closest_intersection for all 32 rays (render.cl:293-378).
//   4. PUSH    hits into the hit ring, escaped paths into the sky ring (evaluated 32 at a time, render.cl:463-466).
// Path state lives in registers only WITHIN a trip (ray -> scan -> record; record -> shade -> ray) and in shared memory
// between scan and shade, so the three expensive stages each run with all 32 lanes: the scan always did, the sky box and
// the camera rays did through their rings, and hit shading -- the bulk of the instructions -- no longer runs with only
// the ~25 lanes whose ray happened to hit something in that trip.  Every path executes exactly the operations it
// executed before, in the same order; only which lane executes them changes.
template <bool COUNT, int MODE>
__device__ __forceinline__ void render_wavefront(const RenderParams &p, const DevScene &sc, const ShapeTable &tab,
                                                 float4 *__restrict__ scratch, unsigned long long *__restrict__ cursor,
                                                 Counters &cnt, unsigned char *smem_raw) {
	const unsigned FULL = 0xffffffffu;
	constexpr bool MODELS = MODE != MODE_ANALYTIC && MODE != MODE_ANALYTIC_CONST;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const unsigned lt_mask = (1u << lane) - 1u;
	uint32_t *rayq = reinterpret_cast<uint32_t *>(smem_raw + warp * wavefront_warp_bytes(MODELS));
	float *skyq = reinterpret_cast<float *>(rayq + RAYQ_WORDS * RAYQ_SLOTS);
	uint32_t *hitq = reinterpret_cast<uint32_t *>(skyq + SKYQ_WORDS * QUEUE_SLOTS);
#if SRT_RING_V4
	// Same rings, same bytes, but twelve of a hit record's words and eight of a sky record's travel as 16-byte vectors
	// (slot-major float4 arrays: consecutive slots are 16 bytes apart, conflict-free LDS.128 / STS.128); the remaining
	// words stay word arrays behind them.
	//   hit4[0] = {pos.xyz, item}  hit4[1] = {d.xyz, seed}  hit4[2] = {mask.xyz, color.z}   words 12 .. 14 (15): color.xy, shape | bounce
	//   sky4[0] = {color.xyz, item}  sky4[1] = {mask.xyz, d.z}                            words 8, 9: d.xy
	float4 *hit4 = reinterpret_cast<float4 *>(hitq);
	float4 *sky4 = reinterpret_cast<float4 *>(skyq);
#endif
	int ray_head = 0, ray_count = 0, sky_head = 0, sky_count = 0, hit_head = 0, hit_count = 0;  // warp-uniform
	bool exhausted = false;
	const vec3 cam_origin = mk(p.c2w[12], p.c2w[13], p.c2w[14]);
	auto flush_sky = [&](int n) {  // mask *= sky, color += mask (:464-465) for n <= 32 queued records, one per lane
		unsigned int it = 0;
		if (lane < n) {
			const int sl = (sky_head + lane) & (QUEUE_SLOTS - 1);
#if SRT_RING_V4
			const float4 s0 = sky4[0 * QUEUE_SLOTS + sl], s1 = sky4[1 * QUEUE_SLOTS + sl];
			it = __float_as_uint(s0.w);
			const vec3 c = mk(s0.x, s0.y, s0.z);
			vec3 m = mk(s1.x, s1.y, s1.z);
			const vec3 dir = mk(skyq[8 * QUEUE_SLOTS + sl], skyq[9 * QUEUE_SLOTS + sl], s1.w);
#else
			it = __float_as_uint(skyq[0 * QUEUE_SLOTS + sl]);
			const vec3 c = mk(skyq[1 * QUEUE_SLOTS + sl], skyq[2 * QUEUE_SLOTS + sl], skyq[3 * QUEUE_SLOTS + sl]);
			vec3 m = mk(skyq[4 * QUEUE_SLOTS + sl], skyq[5 * QUEUE_SLOTS + sl], skyq[6 * QUEUE_SLOTS + sl]);
			const vec3 dir = mk(skyq[7 * QUEUE_SLOTS + sl], skyq[8 * QUEUE_SLOTS + sl], skyq[9 * QUEUE_SLOTS + sl]);
#endif
			m = m * sky_box(sc, dir);
			const vec3 r = c + m;
			scratch[it] = make_float4(r.x, r.y, r.z, 0.f);
		}
		sky_head = (sky_head + n) & (QUEUE_SLOTS - 1);
		sky_count -= n;
		__syncwarp();
	};

#if SRT_WF_CARRY
	// Path state of the trip.  Declared OUTSIDE the loop: a lane without a ray keeps whatever its last path left here
	// (every use is guarded by has_ray; the scan of a stale ray is computed and dropped), which saves re-creating
	// fifteen defaults at the top of every trip.
	unsigned int item = 0;
	int bounce = 0;
	uint32_t seed = 0;
	vec3 o = mk(0, 0, 0), d = mk(0, 0, 1), mask = mk(1, 1, 1), color = mk(0, 0, 0);
#endif
	for (;;) {
		bool has_ray = false;
#if !SRT_WF_CARRY
		unsigned int item = 0;
		int bounce = 0;
		uint32_t seed = 0;
		vec3 o = mk(0, 0, 0), d = mk(0, 0, 1), mask = mk(1, 1, 1), color = mk(0, 0, 0);
#endif

		// -- 1. shade 32 queued hits (fewer only when no fresh ray is left to wait for)
		const bool dry = exhausted && ray_count == 0;
		if (hit_count >= 32 || (dry && hit_count > 0)) {
			const int n = min(hit_count, 32);
			if (lane < n) {
				const int sl = (hit_head + lane) & (QUEUE_SLOTS - 1);
#if SRT_RING_V4
				const float4 h0 = hit4[0 * QUEUE_SLOTS + sl], h1 = hit4[1 * QUEUE_SLOTS + sl], h2 = hit4[2 * QUEUE_SLOTS + sl];
				const vec3 pos = mk(h0.x, h0.y, h0.z);
				item = __float_as_uint(h0.w);
				d = mk(h1.x, h1.y, h1.z);
				seed = __float_as_uint(h1.w);
				mask = mk(h2.x, h2.y, h2.z);
				color = mk(__uint_as_float(hitq[12 * QUEUE_SLOTS + sl]), __uint_as_float(hitq[13 * QUEUE_SLOTS + sl]), h2.w);
#else
				item = hitq[0 * QUEUE_SLOTS + sl];
				seed = hitq[1 * QUEUE_SLOTS + sl];
				const vec3 pos = mk(__uint_as_float(hitq[2 * QUEUE_SLOTS + sl]), __uint_as_float(hitq[3 * QUEUE_SLOTS + sl]),
				                    __uint_as_float(hitq[4 * QUEUE_SLOTS + sl]));
				d = mk(__uint_as_float(hitq[5 * QUEUE_SLOTS + sl]), __uint_as_float(hitq[6 * QUEUE_SLOTS + sl]),
				       __uint_as_float(hitq[7 * QUEUE_SLOTS + sl]));
				mask = mk(__uint_as_float(hitq[8 * QUEUE_SLOTS + sl]), __uint_as_float(hitq[9 * QUEUE_SLOTS + sl]),
				          __uint_as_float(hitq[10 * QUEUE_SLOTS + sl]));
				color = mk(__uint_as_float(hitq[11 * QUEUE_SLOTS + sl]), __uint_as_float(hitq[12 * QUEUE_SLOTS + sl]),
				           __uint_as_float(hitq[13 * QUEUE_SLOTS + sl]));
#endif
				const uint32_t sb = hitq[14 * QUEUE_SLOTS + sl];
				Hit hit = {0.f, (int)(sb >> 8), MODELS ? (int)hitq[15 * QUEUE_SLOTS + sl] : -1};
				bounce = (int)(sb & 255u);
				if (COUNT) cnt.hits += 1;
				vec3 n_;
				bool front;
				int material;
				finish_hit_at(sc, hit, pos, d, n_, front, material);
				bool done;
				if (p.show_normals) {  // :407-410
					color = mk(cfma_(n_.x, 0.5f, 0.5f), cfma_(n_.y, 0.5f, 0.5f), cfma_(n_.z, 0.5f, 0.5f));
					done = true;
				} else {
					const float4 m0 = __ldg(&sc.materials[4 * material + 0]);
					const float4 em = __ldg(&sc.materials[4 * material + 3]);
					color = color + (mask * xyz(em)) * m0.w;  // :413
					if (bounce == p.num_bounces - 1) {         // :415-416
						done = true;
					} else {
						const float4 m1 = __ldg(&sc.materials[4 * material + 1]);
						scatter(sc, material, pos, n_, front, seed, o, d, mask, m0, m1);
						bounce += 1;
						done = false;
					}
				}
				if (done) scratch[item] = make_float4(color.x, color.y, color.z, 0.f);  // summed per pixel in sample order later
				else has_ray = true;
			}
			hit_head = (hit_head + n) & (QUEUE_SLOTS - 1);
			hit_count -= n;
			__syncwarp();
		}

		// -- 2. the other lanes take fresh camera rays (start_path, 32 at a time with every lane active)
		// The ray ring holds ONE batch (32 slots): lanes first take what is left in it, and only when that runs out is a new
		// batch of 32 generated -- into the then empty ring -- for the remaining lanes.  (Generating ahead of the pops, as
		// the plain schedule does, needs 64 slots; the 640 bytes saved per warp are what lets 8 CTAs share an SM.)
		const unsigned need = __ballot_sync(FULL, !has_ray);
		if (need) {
			const int want = __popc(need);
			int served = 0;  // needing lanes already given a ray, in lane order
#pragma unroll 1
			for (int pass = 0; pass < 2; ++pass) {
				const int r = __popc(need & lt_mask) - served;
				if (!has_ray && r >= 0 && r < ray_count) {
					const int sl = (ray_head + r) & (RAYQ_SLOTS - 1);
					item = rayq[0 * RAYQ_SLOTS + sl];
					seed = rayq[1 * RAYQ_SLOTS + sl];
					d = mk(__uint_as_float(rayq[2 * RAYQ_SLOTS + sl]), __uint_as_float(rayq[3 * RAYQ_SLOTS + sl]),
					       __uint_as_float(rayq[4 * RAYQ_SLOTS + sl]));
					o = cam_origin;
					mask = mk(1, 1, 1);
					color = mk(0, 0, 0);
					bounce = 0;
					has_ray = true;
				}
				const int popped = min(want - served, ray_count);
				ray_head = (ray_head + popped) & (RAYQ_SLOTS - 1);
				ray_count -= popped;
				served += popped;
				__syncwarp();
				if (served == want || exhausted || pass == 1) break;
				// the ring is empty: the next 32 work items (start_path with every lane active)
				unsigned int base = 0;
				if (lane == 0) {  // 64-bit cursor: stepping past the end can never wrap into the item range
					const unsigned long long b = atomicAdd(cursor, 32ull);
					base = b < (unsigned long long)p.total_items ? (unsigned int)b : 0xffffffffu;
				}
				base = __shfl_sync(FULL, base, 0);
				const unsigned int it = base + lane;
				const bool valid = base != 0xffffffffu && it < p.total_items;
				const unsigned vm = __ballot_sync(FULL, valid);
				if (valid) {
					uint32_t sd;
					vec3 oo, dd;
					start_path(p, it, sd, oo, dd);
					const int sl = __popc(vm & lt_mask);
					rayq[0 * RAYQ_SLOTS + sl] = it;
					rayq[1 * RAYQ_SLOTS + sl] = sd;
					rayq[2 * RAYQ_SLOTS + sl] = __float_as_uint(dd.x);
					rayq[3 * RAYQ_SLOTS + sl] = __float_as_uint(dd.y);
					rayq[4 * RAYQ_SLOTS + sl] = __float_as_uint(dd.z);
					if (COUNT) cnt.samples += 1;
				}
				ray_head = 0;
				ray_count = __popc(vm);
				exhausted = vm != FULL;
				__syncwarp();
			}
		}
		const bool any_ray = __any_sync(FULL, has_ray);
		if (!any_ray && hit_count > 0) continue;  // only queued hits are left: the next trip shades them
		// !any_ray: nothing in flight, the frame is done for this warp once the sky ring is empty (ONE flush site below:
		// the sky box is the largest block of code in the kernel, and a second inlined copy costs instruction cache)

		// -- 3. closest_intersection for every ray of the trip, :293-378
		Hit hit = {__int_as_float(0x7f800000), -1, -1};
#if SRT_PAIR_SCAN
		if (MODE == MODE_ANALYTIC_CONST) {  // two shapes per packed operation, every lane (scan_pairs)
			if (COUNT && has_ray) cnt.bounces += 1;
			scan_pairs(tab, o, d, hit);
		} else
#endif
		if (has_ray) {
			if (COUNT) cnt.bounces += 1;
			const vec3 inv = MODELS ? mk(rcp_(d.x), rcp_(d.y), rcp_(d.z)) : mk(0, 0, 0);
			scan_shapes<COUNT, false, MODELS, MODE == MODE_BVH, MODE == MODE_ANALYTIC_CONST>(sc, o, d, inv, 0, hit, cnt, &tab);
		}

		// -- 4. hits -> hit ring (position = origin + direction * t, :311 / :337 / :361); escaped paths -> sky ring
		const bool is_hit = has_ray && hit.shape >= 0, is_miss = has_ray && hit.shape < 0;
		const unsigned hm = __ballot_sync(FULL, is_hit), sm = __ballot_sync(FULL, is_miss);
		if (is_hit) {
			const vec3 pos = cfma3(d, hit.t, o);
			const int sl = (hit_head + hit_count + __popc(hm & lt_mask)) & (QUEUE_SLOTS - 1);
#if SRT_RING_V4
			hit4[0 * QUEUE_SLOTS + sl] = make_float4(pos.x, pos.y, pos.z, __uint_as_float(item));
			hit4[1 * QUEUE_SLOTS + sl] = make_float4(d.x, d.y, d.z, __uint_as_float(seed));
			hit4[2 * QUEUE_SLOTS + sl] = make_float4(mask.x, mask.y, mask.z, color.z);
			hitq[12 * QUEUE_SLOTS + sl] = __float_as_uint(color.x);
			hitq[13 * QUEUE_SLOTS + sl] = __float_as_uint(color.y);
#else
			hitq[0 * QUEUE_SLOTS + sl] = item;
			hitq[1 * QUEUE_SLOTS + sl] = seed;
			hitq[2 * QUEUE_SLOTS + sl] = __float_as_uint(pos.x), hitq[3 * QUEUE_SLOTS + sl] = __float_as_uint(pos.y);
			hitq[4 * QUEUE_SLOTS + sl] = __float_as_uint(pos.z);
			hitq[5 * QUEUE_SLOTS + sl] = __float_as_uint(d.x), hitq[6 * QUEUE_SLOTS + sl] = __float_as_uint(d.y);
			hitq[7 * QUEUE_SLOTS + sl] = __float_as_uint(d.z);
			hitq[8 * QUEUE_SLOTS + sl] = __float_as_uint(mask.x), hitq[9 * QUEUE_SLOTS + sl] = __float_as_uint(mask.y);
			hitq[10 * QUEUE_SLOTS + sl] = __float_as_uint(mask.z);
			hitq[11 * QUEUE_SLOTS + sl] = __float_as_uint(color.x), hitq[12 * QUEUE_SLOTS + sl] = __float_as_uint(color.y);
			hitq[13 * QUEUE_SLOTS + sl] = __float_as_uint(color.z);
#endif
			hitq[14 * QUEUE_SLOTS + sl] = ((uint32_t)hit.shape << 8) | (uint32_t)bounce;
			if (MODELS) hitq[15 * QUEUE_SLOTS + sl] = (uint32_t)hit.tri;
		}
		if (is_miss) {
			if (COUNT) cnt.sky += 1;
			const int sl = (sky_head + sky_count + __popc(sm & lt_mask)) & (QUEUE_SLOTS - 1);
#if SRT_RING_V4
			sky4[0 * QUEUE_SLOTS + sl] = make_float4(color.x, color.y, color.z, __uint_as_float(item));
			sky4[1 * QUEUE_SLOTS + sl] = make_float4(mask.x, mask.y, mask.z, d.z);
			skyq[8 * QUEUE_SLOTS + sl] = d.x;
			skyq[9 * QUEUE_SLOTS + sl] = d.y;
#else
			skyq[0 * QUEUE_SLOTS + sl] = __uint_as_float(item);
			skyq[1 * QUEUE_SLOTS + sl] = color.x, skyq[2 * QUEUE_SLOTS + sl] = color.y, skyq[3 * QUEUE_SLOTS + sl] = color.z;
			skyq[4 * QUEUE_SLOTS + sl] = mask.x, skyq[5 * QUEUE_SLOTS + sl] = mask.y, skyq[6 * QUEUE_SLOTS + sl] = mask.z;
			skyq[7 * QUEUE_SLOTS + sl] = d.x, skyq[8 * QUEUE_SLOTS + sl] = d.y, skyq[9 * QUEUE_SLOTS + sl] = d.z;
#endif
		}
		hit_count += __popc(hm);
		sky_count += __popc(sm);
		__syncwarp();
		if (sky_count >= 32 || (!any_ray && sky_count > 0)) flush_sky(min(sky_count, 32));  // (< 32 are left after a trip)
		if (!any_ray) break;
	}
}

// the three by-value kernel parameters + three pointers stay inside the 4 KB every CUDA kernel may take
static_assert(sizeof(RenderParams) + sizeof(DevScene) + sizeof(ShapeTable) + 3 * sizeof(void *) <= 4096, "kernel parameter space");
// WF: the queue builds exist with both schedules -- wavefront (hits shaded 32 at a time through the hit ring) for
// launches that keep every thread busy for many items, plain (hits shaded in place) for short ones, where queueing a
// hit until 32 are there only lengthens the ragged end (BASELINE config 1, 3 items per thread: +6 % plain) and for
// launches whose bounce count does not fit the hit record's 8 bits.
template <bool COUNT, int MODE, bool WF>
__global__ void __launch_bounds__(RENDER_THREADS, MODE == MODE_BIG_MODELS ? SRT_MIN_BLOCKS
                                                  : MODE == MODE_BVH      ? SRT_MIN_BLOCKS_BVH
                                                  : (WF && (MODE == MODE_ANALYTIC || MODE == MODE_ANALYTIC_CONST)) ? SRT_MIN_BLOCKS_WAVEFRONT
                                                                                                                    : SRT_MIN_BLOCKS_ANALYTIC)
render_kernel(const __grid_constant__ RenderParams p, const __grid_constant__ DevScene sc,
              const __grid_constant__ ShapeTable tab, float4 *__restrict__ scratch,
              unsigned long long *__restrict__ cursor, Counters *__restrict__ counters) {
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	Counters cnt = {0, 0, 0, 0, 0, 0};
	constexpr bool MODELS = MODE != MODE_ANALYTIC && MODE != MODE_ANALYTIC_CONST;  // the scan knows about model shapes
	constexpr bool PHASES = MODE == MODE_BIG_MODELS;  // lanes park and the warp runs dense triangle phases
	constexpr bool QUEUES = !PHASES;                  // dense camera-path batches through a warp queue
	constexpr bool SKYQ = QUEUES || SRT_BIG_SKYQ;     // dense sky-box batches through a warp queue

	// per-warp ring of triangle tiles + one mbarrier per stage
	extern __shared__ __align__(128) unsigned char smem_raw[];
	if (QUEUES && WF) {  // the wavefront schedule (the loop below is the plain schedule and the dense-sweep build)
		render_wavefront<COUNT, MODE>(p, sc, tab, scratch, cursor, cnt, smem_raw);
		if (COUNT) {
			unsigned long long *c = reinterpret_cast<unsigned long long *>(&cnt);
			for (int k = 0; k < 6; ++k) {
				unsigned long long v = c[k];
				for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
				if (lane == 0 && v) atomicAdd(reinterpret_cast<unsigned long long *>(counters) + k, v);
			}
		}
		return;
	}
	const int warp = threadIdx.x >> 5;
	unsigned char *wsmem = smem_raw + warp * WARP_SMEM_BYTES;
	const uint32_t wsmem_s = smem_u32(wsmem);
	const uint32_t bars_s = smem_u32(smem_raw + RENDER_WARPS * WARP_SMEM_BYTES + warp * (TILE_STAGES * 8));
	uint32_t parity = 0;
	if (PHASES) {
		if (lane == 0) {
			for (int st = 0; st < TILE_STAGES; ++st) mbar_init(bars_s + st * 8, 1);
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		}
		__syncwarp();
	}

	bool alive = true;
	bool fresh = true;  // the lane needs a new work item (and a camera ray)
	unsigned int item = 0;
	int bounce = 0;
	uint32_t seed = 0;
	vec3 o = mk(0, 0, 0), d = mk(0, 0, 1), mask = mk(1, 1, 1), color = mk(0, 0, 0);
	// closest-hit scan state of the current bounce
	vec3 inv = mk(0, 0, 0);
	Hit hit = {0.f, -1, -1};
	int scan_at = -1;  // next shape to visit; -1 = a new bounce has to be started
	int park = -1;     // shape index of the model this lane waits to run triangles for
	bool ready = false;  // the scan of the current bounce is complete: the lane waits to shade it
	int waited = 0;    // warp-uniform: trips spent with parked lanes waiting for company
	// warp-uniform queue state (QUEUES builds)
	uint32_t *rayq = reinterpret_cast<uint32_t *>(smem_raw + warp * QUEUE_WARP_BYTES);
	float *skyq = PHASES ? reinterpret_cast<float *>(smem_raw + RENDER_SMEM_BYTES + warp * (SKYQ_WORDS * QUEUE_SLOTS * 4))
	                     : reinterpret_cast<float *>(rayq + RAYQ_WORDS * QUEUE_SLOTS);
	int ray_head = 0, ray_count = 0, sky_head = 0, sky_count = 0;
	bool exhausted = false;
	const unsigned lt_mask = (1u << lane) - 1u;
	const vec3 cam_origin = mk(p.c2w[12], p.c2w[13], p.c2w[14]);
	// evaluate `n` queued sky records (n <= 32), one per lane: mask *= sky, color += mask (:464-465)
	auto flush_sky = [&](int n) {
		unsigned int it = 0;
		if (lane < n) {
			const int sl = (sky_head + lane) & (QUEUE_SLOTS - 1);
			it = __float_as_uint(skyq[0 * QUEUE_SLOTS + sl]);
			const vec3 c = mk(skyq[1 * QUEUE_SLOTS + sl], skyq[2 * QUEUE_SLOTS + sl], skyq[3 * QUEUE_SLOTS + sl]);
			vec3 m = mk(skyq[4 * QUEUE_SLOTS + sl], skyq[5 * QUEUE_SLOTS + sl], skyq[6 * QUEUE_SLOTS + sl]);
			const vec3 dir = mk(skyq[7 * QUEUE_SLOTS + sl], skyq[8 * QUEUE_SLOTS + sl], skyq[9 * QUEUE_SLOTS + sl]);
			m = m * sky_box(sc, dir);
			const vec3 r = c + m;
			scratch[it] = make_float4(r.x, r.y, r.z, 0.f);
		}
		sky_head = (sky_head + n) & (QUEUE_SLOTS - 1);
		sky_count -= n;
		__syncwarp();
	};

	for (;;) {
		// -- refill: lanes whose path ended pull the next (pixel, sample) item and start its camera path
		const bool need_item = alive && fresh;
		const unsigned need = __ballot_sync(FULL, need_item);
		if (QUEUES && need) {
			const int want = __popc(need);
			if (ray_count < want && !exhausted) {  // generate 32 camera paths with every lane active
				// the cursor is 64 bits wide, so the one step every warp takes past the end can never wrap it back
				// into the item range; past the end the base collapses to a sentinel (total_items <= MAX_ITEMS)
				unsigned int base = 0;
				if (lane == 0) {
					const unsigned long long b = atomicAdd(cursor, 32ull);
					base = b < (unsigned long long)p.total_items ? (unsigned int)b : 0xffffffffu;
				}
				base = __shfl_sync(FULL, base, 0);
				const unsigned int it = base + lane;
				const bool valid = base != 0xffffffffu && it < p.total_items;
				const unsigned vm = __ballot_sync(FULL, valid);
				if (valid) {
					uint32_t sd;
					vec3 oo, dd;
					start_path(p, it, sd, oo, dd);
					const int sl = (ray_head + ray_count + __popc(vm & lt_mask)) & (QUEUE_SLOTS - 1);
					rayq[0 * QUEUE_SLOTS + sl] = it;
					rayq[1 * QUEUE_SLOTS + sl] = sd;
					rayq[2 * QUEUE_SLOTS + sl] = __float_as_uint(dd.x);
					rayq[3 * QUEUE_SLOTS + sl] = __float_as_uint(dd.y);
					rayq[4 * QUEUE_SLOTS + sl] = __float_as_uint(dd.z);
					if (COUNT) cnt.samples += 1;
				}
				ray_count += __popc(vm);
				exhausted = vm != FULL;
				__syncwarp();
			}
			if (need_item) {
				const int r = __popc(need & lt_mask);
				if (r < ray_count) {
					const int sl = (ray_head + r) & (QUEUE_SLOTS - 1);
					item = rayq[0 * QUEUE_SLOTS + sl];
					seed = rayq[1 * QUEUE_SLOTS + sl];
					d = mk(__uint_as_float(rayq[2 * QUEUE_SLOTS + sl]), __uint_as_float(rayq[3 * QUEUE_SLOTS + sl]),
					       __uint_as_float(rayq[4 * QUEUE_SLOTS + sl]));
					o = cam_origin;
					mask = mk(1, 1, 1);
					color = mk(0, 0, 0);
					bounce = 0;
					fresh = false;
					scan_at = -1;
				} else {
					alive = false;  // the frame has run dry
				}
			}
			const int popped = min(want, ray_count);
			ray_head = (ray_head + popped) & (QUEUE_SLOTS - 1);
			ray_count -= popped;
			__syncwarp();
		}
		if (!QUEUES && need) {
			unsigned int base = 0;
			if (lane == 0) {  // 64-bit cursor, see above
				const unsigned long long b = atomicAdd(cursor, (unsigned long long)__popc(need));
				base = b < (unsigned long long)p.total_items ? (unsigned int)b : 0xffffffffu;
			}
			base = __shfl_sync(FULL, base, 0);
			if (need_item) {
				item = base + __popc(need & ((1u << lane) - 1u));
				if (base != 0xffffffffu && item < p.total_items) {  // start path `sample` of pixel `pix`, :496-516
					start_path(p, item, seed, o, d);
					mask = mk(1, 1, 1);
					color = mk(0, 0, 0);
					bounce = 0;
					fresh = false;
					scan_at = -1;
					if (COUNT) cnt.samples += 1;
				} else {
					alive = false;
				}
			}
		}
		// nobody alive: the frame is done for this warp once the sky ring is empty (ONE flush site below: the sky box is
		// the largest block of code in the kernel, and a second inlined copy costs instruction cache)
		const bool any_alive = __any_sync(FULL, alive);

		bool push_sky = false;
		if (alive && park < 0 && !ready) {
			if (scan_at < 0) {  // new bounce: closest_intersection prologue, :294-297
				if (COUNT) cnt.bounces += 1;
				hit.t = __int_as_float(0x7f800000);
				hit.shape = -1;
				hit.tri = -1;
				if (MODELS) inv = mk(rcp_(d.x), rcp_(d.y), rcp_(d.z));
				scan_at = 0;
			}
			park = scan_shapes<COUNT, PHASES, MODELS, MODE == MODE_BVH, MODE == MODE_ANALYTIC_CONST>(sc, o, d, inv, scan_at, hit, cnt, &tab);
			if (park < 0) {  // scan complete
				scan_at = -1;
				ready = true;
			}
		}
		// Shading.  A lane shades in the trip its scan completes.  With dense triangle phases lanes finish their scans at
		// different trips (they park at different models), so hit shading runs at ~11 of 32 lanes on the two-mesh config.
		// SHADE_SYNC > 0 makes ready lanes WAIT while any lane is still parked unless that many are ready (phases first,
		// then a nearly full warp shades).  Measured: 20 / 28 / 33 -> +0 / -1 / -5 % on config 3, -2 / -3 / -7 % on
		// config 5 -- waiting lanes delay the NEXT bounce's parked rays, and thinner phases cost more than fuller
		// shading saves -- so it is off; the flattened "everybody advances when it can" schedule stays.
		bool shade_now = true;
		if (PHASES && SHADE_SYNC > 0) {
			const unsigned rd = __ballot_sync(FULL, ready);
			const unsigned pk = __ballot_sync(FULL, park >= 0);
			shade_now = pk == 0u || __popc(rd) >= SHADE_SYNC;
		}
		if (ready && shade_now) {  // shade this bounce, :404-468
			ready = false;
			{
				bool done;
				if (hit.shape >= 0) {
					if (COUNT) cnt.hits += 1;
					vec3 pos, n;
					bool front;
					int material;
					finish_hit(sc, hit, o, d, pos, n, front, material);
					if (p.show_normals) {  // :407-410
						color = mk(cfma_(n.x, 0.5f, 0.5f), cfma_(n.y, 0.5f, 0.5f), cfma_(n.z, 0.5f, 0.5f));
						done = true;
					} else {
						const float4 m0 = __ldg(&sc.materials[4 * material + 0]);
						const float4 em = __ldg(&sc.materials[4 * material + 3]);
						color = color + (mask * xyz(em)) * m0.w;  // :413
						if (bounce == p.num_bounces - 1) {         // :415-416
							done = true;
						} else {
							const float4 m1 = __ldg(&sc.materials[4 * material + 1]);
							scatter(sc, material, pos, n, front, seed, o, d, mask, m0, m1);
							bounce += 1;
							done = false;
						}
					}
				} else {  // :463-467
					if (COUNT) cnt.sky += 1;
					if (SKYQ) {
						push_sky = true;  // evaluated 32 at a time below
						fresh = true;
						done = false;
					} else {
						mask = mask * sky_box(sc, d);
						color = color + mask;
						done = true;
					}
				}
				if (done) {  // the sample's radiance; accumulate_kernel sums a pixel's samples in order (:518-522)
					scratch[item] = make_float4(color.x, color.y, color.z, 0.f);
					fresh = true;
				}
			}
		}

		if (SKYQ) {  // escaped paths: queue {item, color, mask, direction}; evaluate the sky box 32 at a time
			const unsigned sm = __ballot_sync(FULL, push_sky);
			if (sm) {
				if (push_sky) {
					const int sl = (sky_head + sky_count + __popc(sm & lt_mask)) & (QUEUE_SLOTS - 1);
					skyq[0 * QUEUE_SLOTS + sl] = __uint_as_float(item);
					skyq[1 * QUEUE_SLOTS + sl] = color.x, skyq[2 * QUEUE_SLOTS + sl] = color.y, skyq[3 * QUEUE_SLOTS + sl] = color.z;
					skyq[4 * QUEUE_SLOTS + sl] = mask.x, skyq[5 * QUEUE_SLOTS + sl] = mask.y, skyq[6 * QUEUE_SLOTS + sl] = mask.z;
					skyq[7 * QUEUE_SLOTS + sl] = d.x, skyq[8 * QUEUE_SLOTS + sl] = d.y, skyq[9 * QUEUE_SLOTS + sl] = d.z;
				}
				sky_count += __popc(sm);
				__syncwarp();
			}
			if (sky_count >= 32 || (!any_alive && sky_count > 0)) flush_sky(min(sky_count, 32));  // (< 32 are left after a trip)
		}
		if (!any_alive) break;

		// -- dense triangle phase.  Lane utilisation inside the phase does not depend on how many rays are
		// parked (the lanes hold triangles there), so it runs after at most PHASE_PATIENCE trips of waiting
		// for company; parked lanes get back to tracing as soon as possible.
		const unsigned parked = PHASES ? __ballot_sync(FULL, park >= 0) : 0u;
		if (PHASES && parked) {
			const unsigned movable = __ballot_sync(FULL, alive && park < 0 && !ready);
			if (movable == 0 || waited >= PHASE_PATIENCE) {
				// the model most lanes wait for (ties: the lower shape index)
				const unsigned same = __match_any_sync(FULL, park);
				const int votes = park >= 0 ? (__popc(same) << 20) | (0xfffff - min(park, 0xfffff)) : 0;
				const int best = __reduce_max_sync(FULL, votes);
				const int model = 0xfffff - (best & 0xfffff);
				const int4 hdr = __ldg(&sc.shape_hdr[model]);
				const bool active = park == model;
				if (hdr.w <= p.uv_max_tris)  // few, large triangles: the two-strip filter pays for itself (warp-uniform)
					triangle_phase<true>(sc, hdr.w, hdr.z, model, active, o, d, hit, wsmem, wsmem_s, bars_s, parity, lane);
				else
					triangle_phase<false>(sc, hdr.w, hdr.z, model, active, o, d, hit, wsmem, wsmem_s, bars_s, parity, lane);
				if (active) {
					scan_at = park + 1;
					park = -1;
				}
				waited = 0;
			} else {
				waited += 1;
			}
		}
	}

	if (COUNT) {
		unsigned long long *c = reinterpret_cast<unsigned long long *>(&cnt);
		for (int k = 0; k < 6; ++k) {
			unsigned long long v = c[k];
			for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
			if (lane == 0 && v) atomicAdd(reinterpret_cast<unsigned long long *>(counters) + k, v);
		}
	}
}

// ---- per-launch epilogue of `render`: color = sum of the pixel's samples in sample order (:494-519),
// color /= num_samples (:520), canvas[id] += color (:522) -- launch after launch for a batch, exactly the sequence of
// additions separate launches would perform.  One thread per local pixel.
__global__ void __launch_bounds__(256)
accumulate_kernel(const __grid_constant__ RenderParams p, const float4 *__restrict__ scratch, float4 *__restrict__ canvas) {
	const unsigned int lp = blockIdx.x * blockDim.x + threadIdx.x;
	if (lp >= p.total_pixels) return;
	const int row = (int)(lp / (unsigned)p.width);
	const int gx = (int)(lp - (unsigned)row * (unsigned)p.width);
	const int gy = p.band_n > 1 ? ((row / p.band_h) * p.band_n + p.band_i) * p.band_h + (row % p.band_h) : row;
	const size_t pix = (size_t)gx + (size_t)gy * p.width;
	float4 c = canvas[pix];
	const float ns_f = (float)p.num_samples;
	for (int l = 0; l < p.num_launches; ++l) {
		const float4 *s = scratch + (size_t)l * p.items_per_launch + (size_t)lp * p.num_samples;
		vec3 color = mk(0, 0, 0);
		for (int k = 0; k < p.num_samples; ++k) color = color + xyz(s[k]);
		if (p.inv_ns != 0.0f) {  // x / 2^k == x * 2^-k exactly
			c.x += color.x * p.inv_ns, c.y += color.y * p.inv_ns, c.z += color.z * p.inv_ns;
		} else {
			c.x += div_(color.x, ns_f), c.y += div_(color.y, ns_f), c.z += div_(color.z, ns_f);
		}
	}
	canvas[pix] = c;
}

// ---- frame epilogue: accumulate_kernel + kernel `average` + the read-back in ONE kernel ---------------------------
// For srt_render_frame into a page-locked caller vector (srt_pin_output): one thread per FOUR consecutive local pixels
// performs accumulate_kernel's additions for them, applies `average` (render.cl:525-535) and stores the four ARGB8
// pixels as one 16-byte word both to the device image and straight into the caller's vector over PCIe (a warp writes
// 512 contiguous bytes), so no copy engine transfer follows the kernels.  Same arithmetic, same order: bit-identical.
__global__ void __launch_bounds__(256)
frame_epilogue_kernel(const __grid_constant__ RenderParams p, const float4 *__restrict__ scratch, float4 *__restrict__ canvas,
                      uchar4 *__restrict__ output, uchar4 *__restrict__ host_out, uint32_t num_steps) {
	const unsigned int q = blockIdx.x * blockDim.x + threadIdx.x;  // quad of pixels (full frame: local pixel == pixel id)
	const unsigned int first = 4u * q;
	if (first >= p.total_pixels) return;
	const float steps = (float)num_steps, ns_f = (float)p.num_samples;
	uchar4 px[4];
	const int n = min(4u, p.total_pixels - first);
	for (int j = 0; j < n; ++j) {
		const size_t pix = (size_t)first + j;
		float4 c = canvas[pix];
		const float4 *s = scratch + pix * p.num_samples;
		vec3 color = mk(0, 0, 0);
		for (int k = 0; k < p.num_samples; ++k) color = color + xyz(s[k]);
		if (p.inv_ns != 0.0f) {
			c.x += color.x * p.inv_ns, c.y += color.y * p.inv_ns, c.z += color.z * p.inv_ns;
		} else {
			c.x += div_(color.x, ns_f), c.y += div_(color.y, ns_f), c.z += div_(color.z, ns_f);
		}
		canvas[pix] = c;
		px[j] = argb_pixel(c, steps);
	}
	if (n == 4) {  // total_pixels * 4 bytes: both images are 16-byte aligned at every quad
		const uint4 w = make_uint4(*reinterpret_cast<uint32_t *>(&px[0]), *reinterpret_cast<uint32_t *>(&px[1]),
		                           *reinterpret_cast<uint32_t *>(&px[2]), *reinterpret_cast<uint32_t *>(&px[3]));
		reinterpret_cast<uint4 *>(output)[q] = w;
		reinterpret_cast<uint4 *>(host_out)[q] = w;
	} else {
		for (int j = 0; j < n; ++j) output[first + j] = px[j], host_out[first + j] = px[j];
	}
}

// ---- kernel `average`, render.cl:525-535 (argb_pixel above) ----------------------------------------
__global__ void __launch_bounds__(256)
average_kernel(uint32_t num_steps, const float4 *__restrict__ canvas, uchar4 *__restrict__ output, int n) {
	int id = blockIdx.x * blockDim.x + threadIdx.x;
	if (id >= n) return;
	const float steps = (float)num_steps;
	output[id] = argb_pixel(canvas[id], steps);
}

// ---- debug: primary hit of every pixel's sample-0 camera ray -----------------------------------
template <bool BVH>
__global__ void __launch_bounds__(256)
primary_kernel(const __grid_constant__ RenderParams p, const __grid_constant__ DevScene sc,
               int *__restrict__ shape_idx, float *__restrict__ t_out) {
	int id = blockIdx.x * blockDim.x + threadIdx.x;
	if (id >= p.width * p.height) return;
	int gx = id % p.width, gy = id / p.width;
	uint32_t seed = (0u + (uint32_t)id * (uint32_t)p.num_samples) * p.times[0] * 5304u;
	vec3 o, d;
	camera_ray(p, gx, gy, seed, o, d);
	Counters cnt = {0, 0, 0, 0, 0, 0};
	Hit hit = {__int_as_float(0x7f800000), -1, -1};
	vec3 inv = mk(rcp_(d.x), rcp_(d.y), rcp_(d.z));
	scan_shapes<false, false, true, BVH>(sc, o, d, inv, 0, hit, cnt);
	shape_idx[id] = hit.shape;
	t_out[id] = hit.t;
}

// ---- scene upload: AoS Triangle[] -> pre-transformed SoA ---------------------------------------
// One thread per (model instance, triangle).  Positions are transformed exactly as render.cl:326-328
// does per ray (transform_mat(model->transform, pos, true), :114-120 operation order) and the edges
// are the subtractions of :247-248, so every later intersection sees the same bits as the reference.
struct ModelSpan {
	int shape;      // shape slot (index into model_xf / 4)
	int src_begin;  // first triangle in the AoS array (Model::triangle_index)
	int dst_begin;  // first triangle in the SoA arrays
	int count;
};
__global__ void __launch_bounds__(256)
prepare_triangles_kernel(const float4 *__restrict__ aos /* 6 float4 per triangle */, const ModelSpan *__restrict__ spans,
                         int n_spans, int total, const float4 *__restrict__ model_xf, float4 *__restrict__ hot_out,
                         float2 *__restrict__ flt_out, float4 *__restrict__ uv_out, float *__restrict__ model_k,
                         float4 *__restrict__ n_out) {
	int g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= total) return;
	int lo = 0, hi = n_spans - 1;  // last span with dst_begin <= g
	while (lo < hi) {
		int mid = (lo + hi + 1) >> 1;
		if (spans[mid].dst_begin <= g) lo = mid;
		else hi = mid - 1;
	}
	const ModelSpan sp = spans[lo];
	if (g - sp.dst_begin >= sp.count) return;  // alignment gap between two models' ranges
	const float4 *tri = aos + 6 * (size_t)(sp.src_begin + (g - sp.dst_begin));
	const float4 m0 = model_xf[4 * sp.shape + 0], m1 = model_xf[4 * sp.shape + 1];
	const float4 m2 = model_xf[4 * sp.shape + 2], m3 = model_xf[4 * sp.shape + 3];
	vec3 w[3];
	for (int j = 0; j < 3; ++j) {
		const float4 pj = tri[2 * j + 1];
		w[j] = mk(cfma_(m3.x, 1.0f, cfma_(m2.x, pj.z, cfma_(m1.x, pj.y, m0.x * pj.x))),
		          cfma_(m3.y, 1.0f, cfma_(m2.y, pj.z, cfma_(m1.y, pj.y, m0.y * pj.x))),
		          cfma_(m3.z, 1.0f, cfma_(m2.z, pj.z, cfma_(m1.z, pj.y, m0.z * pj.x))));
		n_out[3 * (size_t)g + j] = tri[2 * j];
	}
	vec3 e1 = w[1] - w[0], e2 = w[2] - w[0];
	hot_out[3 * (size_t)g + 0] = make_float4(w[0].x, w[0].y, w[0].z, 0.f);
	hot_out[3 * (size_t)g + 1] = make_float4(e1.x, e1.y, e1.z, 0.f);
	hot_out[3 * (size_t)g + 2] = make_float4(e2.x, e2.y, e2.z, 0.f);
	// filter record of the dense sweep (tri_filter_sweep): n' = e2 x e1, m = e2 x v0, margin scale g; the model's K
	const vec3 np = cross(e2, e1), m = cross(e2, w[0]);
	const float U = 5.9604644775390625e-8f;  // 2^-24
	const float n1e1 = fabsf(e1.x) + fabsf(e1.y) + fabsf(e1.z), n1e2 = fabsf(e2.x) + fabsf(e2.y) + fabsf(e2.z);
	const float n1v0 = fabsf(w[0].x) + fabsf(w[0].y) + fabsf(w[0].z);
	float2 *r = flt_out + 5 * (size_t)g;
	r[0] = make_float2(np.x, m.x);
	r[1] = make_float2(np.y, m.y);
	r[2] = make_float2(np.z, m.z);
	r[3] = make_float2(e2.x, e2.y);
	r[4] = make_float2(e2.z, 48.0f * SRT_MARGIN_SCALE * U * n1e2);
	// record of the two-strip filter (tri_filter_sweep_uv): n', -m, -m1 = -(e1 x v0), e2, e1, one margin scale for both strips
	const vec3 mv = cross(e1, w[0]);
	const float emax = n1e1 > n1e2 || n1e1 != n1e1 ? n1e1 : n1e2;  // a NaN wins: the filter then passes everything
	float4 *q = uv_out + 5 * (size_t)g;
	q[0] = make_float4(np.x, np.y, np.z, 48.0f * SRT_MARGIN_SCALE * U * emax);
	q[1] = make_float4(-m.x, -mv.x, -m.y, -mv.y);
	q[2] = make_float4(-m.z, -mv.z, e2.x, e1.x);
	q[3] = make_float4(e2.y, e1.y, e2.z, e1.z);
	q[4] = make_float4(0.f, 0.f, 0.f, 0.f);
	const float k = n1v0 + 3.0f * emax;  // non-negative (or NaN, which the filter then passes): orders like its bits
	atomicMax(reinterpret_cast<unsigned int *>(model_k + sp.shape), __float_as_uint(k == k ? k : __int_as_float(0x7f800000)));
}

// ---- scene upload: per-material constants --------------------------------------------------------------------------
// The device copy of a Material record keeps the reference's 64 bytes; its two padding floats (bytes 24..31, unused by
// render.cl:17-27) receive shlick_reflectance's r0 = ((1 - mu) / (1 + mu))^2 (:174-175, FP64 division) for the two
// values mu can take at that material -- 1 / refraction_index entering (:442), refraction_index leaving -- computed by
// the same device function the per-hit code used to call, so that no hit pays for a double-precision division.
__global__ void prepare_materials_kernel(float4 *__restrict__ materials, int n) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float4 m1 = materials[4 * i + 1];
	m1.z = schlick_r0(rcp_(m1.y));
	m1.w = schlick_r0(m1.y);
	materials[4 * i + 1] = m1;
}

// ---- device math self-test -------------------------------------------------------------------
__global__ void math_kernel(int op, const float *__restrict__ x, const float *__restrict__ y, float *__restrict__ out,
                            size_t n) {
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float r = 0.f;
	switch (op) {
	case 0: r = log_(x[i]); break;
	case 1: r = cos_(x[i]); break;
	case 2: r = atan2pi_(x[i], y[i]); break;
	case 3: r = pow_(x[i], y[i]); break;
	case 4: r = sqrt_(x[i]); break;
	case 5: r = schlick_(x[i], y[i]); break;
	// the packed FP32x2 forms of random_float_normal_x2: op 6/7 = halves of log_x2({x, y}), 8/9 = halves of cos_x2({x, y})
	case 6: r = log_x2(make_float2(x[i], y[i])).x; break;
	case 7: r = log_x2(make_float2(x[i], y[i])).y; break;
	case 8: r = cos_x2(make_float2(x[i], y[i])).x; break;
	case 9: r = cos_x2(make_float2(x[i], y[i])).y; break;
	case 10: r = rcp_sqrt_(x[i]); break;
	case 11: r = rcp_(sqrt_(x[i])); break;
	case 12: r = sqrt_x2(make_float2(x[i], y[i])).x; break;
	case 13: r = sqrt_x2(make_float2(x[i], y[i])).y; break;
	// EXHAUSTIVE checks: thread i compares the two forms on the bit patterns i * 2^32 / n ... (n = 2^20 threads cover all
	// 2^32); the result is the number of patterns on which they differ (NaN payloads aside)
	case 14:
	case 15: {
		const uint32_t span = (uint32_t)(0x100000000ull / n);
		uint32_t bad = 0;
		for (uint32_t k = 0; k < span; ++k) {
			const float v = __uint_as_float((uint32_t)i * span + k);
			float a, b;
			if (op == 14) {
				a = rcp_sqrt_(v), b = rcp_(sqrt_(v));
			} else {
				const float w = __uint_as_float(((uint32_t)i * span + k) * 2654435761u);  // some other pattern in the other half
				const float2 p2 = sqrt_x2(make_float2(v, w));
				a = p2.x, b = sqrt_(v);
				if (!(__float_as_uint(p2.y) == __float_as_uint(sqrt_(w)) || (p2.y != p2.y && w != w) || (p2.y != p2.y && sqrt_(w) != sqrt_(w)))) ++bad;
			}
			if (!(__float_as_uint(a) == __float_as_uint(b) || (a != a && b != b))) ++bad;
		}
		r = (float)bad;
		break;
	}
	}
	out[i] = r;
}

// ---- FP32 FMA-chain micro-benchmark (the roofline denominator) ----------
__global__ void __launch_bounds__(256) fma_peak_kernel(float *out, int iters, float a, float b) {
	float r[16];
#pragma unroll
	for (int k = 0; k < 16; ++k) r[k] = (float)(threadIdx.x + k);
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int k = 0; k < 16; ++k) r[k] = __fmaf_rn(r[k], a, b);
	}
	float s = 0.f;
#pragma unroll
	for (int k = 0; k < 16; ++k) s += r[k];
	if (s == 123.456f) out[0] = s;
}

}  // namespace srt
