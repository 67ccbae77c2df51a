// cl_shim.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Just enough of the OpenCL C language for g++ to compile the reference kernel source
// /root/reference/src/render.cl AS IT LIES (see ref_driver.cpp and Makefile `_ref`): the address-space
// qualifiers, the vector types float2/float3/float4/uchar4 with OpenCL's size/alignment rules
// (float3 = 16 bytes, 16-aligned), component-wise operators, the `.xyz` swizzle, work-item ids and the
// builtins render.cl calls.  Nothing in here restates render.cl: every struct, every function and both
// kernels come from the reference file itself.
//
// The builtins are the part of an OpenCL implementation that is third-party to the reference (whichever
// driver JIT-compiles the kernel supplies them, src/tracer.cpp:13); they forward to the ONE set of
// definitions this repository documents, oracle/oracle_math.h, shared with
// oracle.c, so that any difference between this build and the hand-written oracle is a difference in the
// restatement of render.cl and nothing else.
//
// Everything lives in namespace refcl so that `log`, `cos`, `sqrt`, `min`, ... resolve to the functions
// below (unqualified lookup stops at the enclosing namespace) and never to libm.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../oracle_math.h"

namespace refcl {

// ---- qualifiers / scalar typedefs ------------------------------------------------------------
#define __kernel
#define __global
#define __generic
#define __constant const
#define _Bool bool
typedef unsigned int uint;
typedef unsigned char uchar;
#ifdef UINT_MAX
#undef UINT_MAX
#endif
#define UINT_MAX 4294967295u
#ifdef FLT_MAX
#undef FLT_MAX
#endif
#define FLT_MAX 3.402823466e+38f
#ifdef INFINITY
#undef INFINITY
#endif
#define INFINITY (__builtin_inff())
#define M_PI_F 3.14159274101257f

// ---- vector types ----------------------------------------------------------------------------
// Plain aggregates (render.cl puts them inside anonymous structs/unions, :44-53, where g++ accepts no type
// with a constructor).  A vector literal `(float3)(a, b, c)` reaches C++ as `float3(float3_lit{a, b, c})`
// (rewrite_cl.py): the braced list is evaluated left to right, like OpenCL C / clang does, and the _lit
// helper implements OpenCL's literal forms (scalar splat, vector + scalar concatenation).
struct alignas(8) float2 {
	float x, y;
};

struct alignas(16) float3 {
	float x, y, z, pad_;
	float3 xyz() const { return *this; }
	float &operator[](int i) { return (&x)[i]; }
	const float &operator[](int i) const { return (&x)[i]; }
	// compound assignments touch x, y, z only: the 4th lane of a float3 in memory is left alone
	float3 &operator+=(float3 b) { x += b.x, y += b.y, z += b.z; return *this; }
	float3 &operator*=(float3 b) { x *= b.x, y *= b.y, z *= b.z; return *this; }
	float3 &operator*=(float s) { x *= s, y *= s, z *= s; return *this; }
	float3 &operator/=(float s) { x /= s, y /= s, z /= s; return *this; }
};
inline float3 mk3(float a, float b, float c) { return float3{a, b, c, 0.0f}; }

struct alignas(16) float4 {
	float x, y, z, w;
	float3 xyz() const { return mk3(x, y, z); }
};

struct alignas(4) uchar4 {
	uchar x, y, z, w;
};

static_assert(sizeof(float2) == 8 && sizeof(float3) == 16 && sizeof(float4) == 16 && sizeof(uchar4) == 4, "vector sizes");

struct float2_lit {
	float x, y;
	float2_lit(float a, float b) : x(a), y(b) {}
	operator float2() const { return float2{x, y}; }
};
struct float3_lit {
	float x, y, z;
	float3_lit(float s) : x(s), y(s), z(s) {}
	float3_lit(float a, float b, float c) : x(a), y(b), z(c) {}
	float3_lit(float2 a, float c) : x(a.x), y(a.y), z(c) {}
	operator float3() const { return mk3(x, y, z); }
};
struct float4_lit {
	float x, y, z, w;
	float4_lit(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
	float4_lit(float3 a, float d) : x(a.x), y(a.y), z(a.z), w(d) {}
	operator float4() const { return float4{x, y, z, w}; }
};
struct uchar4_lit {
	uchar x, y, z, w;
	// float -> uchar: C conversion (truncation toward zero), as for an OpenCL vector literal
	uchar4_lit(uchar a, uchar b, uchar c, uchar d) : x(a), y(b), z(c), w(d) {}
	operator uchar4() const { return uchar4{x, y, z, w}; }
};

// the swizzle used by render.cl (`q.xyz`, `m(...).xyz`, `v.xyz`) becomes a member call
#define xyz xyz()

// component-wise operators (OpenCL 6.3); scalar operands widen to the vector
inline float3 operator+(float3 a, float3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator-(float3 a, float3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator*(float3 a, float3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float3 operator/(float3 a, float3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline float3 operator+(float3 a, float s) { return mk3(a.x + s, a.y + s, a.z + s); }
inline float3 operator-(float3 a, float s) { return mk3(a.x - s, a.y - s, a.z - s); }
inline float3 operator*(float3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
inline float3 operator/(float3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
inline float3 operator+(float s, float3 a) { return mk3(s + a.x, s + a.y, s + a.z); }
inline float3 operator-(float s, float3 a) { return mk3(s - a.x, s - a.y, s - a.z); }
inline float3 operator*(float s, float3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
inline float3 operator/(float s, float3 a) { return mk3(s / a.x, s / a.y, s / a.z); }
inline float3 operator-(float3 a) { return mk3(-a.x, -a.y, -a.z); }

// ---- builtins: forwarded to oracle_math.h ------------------------------------------------------
inline float sqrt(float a) { return om_sqrt(a); }
inline float3 sqrt(float3 a) { return mk3(om_sqrt(a.x), om_sqrt(a.y), om_sqrt(a.z)); }
inline float fabs(float a) { return __builtin_fabsf(a); }
inline float log(float a) { return om_log(a); }
inline float cos(float a) { return om_cos(a); }
inline float pow(float a, float b) { return om_pow(a, b); }
inline float atan2pi(float y, float x) { return om_atan2pi(y, x); }
inline double pown(double x, int n) {  // only called with n = 5 (render.cl:177)
	double r = x;
	for (int i = 1; i < n; ++i) r = r * x;
	return r;
}
inline float min(float a, float b) { return om_min(a, b); }
inline float max(float a, float b) { return om_max(a, b); }
inline float sign(float a) { return om_sign(a); }
inline float dot(float3 a, float3 b) { return v3_dot(v3_make(a.x, a.y, a.z), v3_make(b.x, b.y, b.z)); }
inline float3 cross(float3 a, float3 b) {
	v3 r = v3_cross(v3_make(a.x, a.y, a.z), v3_make(b.x, b.y, b.z));
	return mk3(r.x, r.y, r.z);
}
inline float3 normalize(float3 a) {
	v3 r = v3_normalize(v3_make(a.x, a.y, a.z));
	return mk3(r.x, r.y, r.z);
}
inline float3 mix(float3 a, float3 b, float t) { return mk3(om_mix(a.x, b.x, t), om_mix(a.y, b.y, t), om_mix(a.z, b.z, t)); }
inline float3 mix(float3 a, float3 b, float3 t) {
	return mk3(om_mix(a.x, b.x, t.x), om_mix(a.y, b.y, t.y), om_mix(a.z, b.z, t.z));
}
inline float3 clamp(float3 v, float3 lo, float3 hi) {
	return mk3(om_clamp(v.x, lo.x, hi.x), om_clamp(v.y, lo.y, hi.y), om_clamp(v.z, lo.z, hi.z));
}

// ---- images ------------------------------------------------------------------------------------
// image2d_t = CL_RGBA / CL_FLOAT texels, row 0 first (src/tracer.cpp:42-52); the only sampler the host ever
// creates is normalised coordinates + clamp-to-edge + linear (src/tracer.cpp:47-48).
struct image2d_desc {
	const float *texels;
	int width, height;
};
typedef const image2d_desc *image2d_t;
typedef int sampler_t;
inline float4 read_imagef(image2d_t img, sampler_t, float2 uv) {
	float r[4];
	om_read_imagef_linear_clamp(img->texels, img->width, img->height, uv.x, uv.y, r);
	return float4{r[0], r[1], r[2], r[3]};
}

// ---- work-item ids -----------------------------------------------------------------------------
extern thread_local size_t g_global_id[2];
inline size_t get_global_id(uint dim) { return g_global_id[dim]; }

}  // namespace refcl
