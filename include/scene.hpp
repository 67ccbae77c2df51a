// scene.hpp -- header-only C++ scene helpers with the reference's names, over the plain-C records of srt.h.
//
// What main.cpp / interface.cpp use to BUILD a scene before handing it to Tracer::update_scene:
//   Sphere, Plane, Triangle, Model, Box, Shape        reference include/shape.hpp:15-111, src/shape.cpp
//   Material, color::from_hex                         reference include/material.hpp:10-37, include/color.hpp:8-14
//   load_stl_model, load_obj_model, save_ppm          reference include/parser.hpp:14-28, src/parser.cpp
//   Camera::camera_matrix                             reference include/helper.hpp:14-27
// Every struct DERIVES from its srt.h record and adds constructors only, so sizeof and layout are the record's
// (static_asserted below) and a std::vector<Shape> is passed to the C ABI as is.  glm is not available in this image:
// `vec3` is a 3-float stand-in with the few operations the helpers need; a tree that has glm converts at the call site.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "srt.h"

namespace srt_facade {

struct vec3 {
	float x = 0.f, y = 0.f, z = 0.f;
	vec3() = default;
	vec3(float s) : x(s), y(s), z(s) {}
	vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline vec3 normalize(vec3 v) { return v * (1.0f / std::sqrt(dot(v, v))); }  // glm::normalize = v * inversesqrt(dot(v, v))
inline srt_float3 to_record(vec3 v) { return srt_float3{v.x, v.y, v.z, 0.f}; }

using Color = vec3;
namespace color {
const Color white(1.f), black(0.f), gray(.5f);
inline Color from_hex(uint32_t v) {  // include/color.hpp:12-14
	return Color(((v & 0xFF0000) >> 16) / 255.f, ((v & 0xFF00) >> 8) / 255.f, (v & 0xFF) / 255.f);
}
}  // namespace color

struct Material : srt_material {  // include/material.hpp:23-37 (same argument order and defaults)
	Material(const Color &color = color::white, float smoothness = 0.0f, float metallic = 0.0f, float specular = 0.0f,
	         float transmittance = 0.0f, float refraction_index = 1.0f, const Color &emission = color::black,
	         float emission_strength = 0.0f) {
		std::memset(static_cast<srt_material *>(this), 0, sizeof(srt_material));
		this->color = to_record(color), this->emission = to_record(emission);
		this->smoothness = smoothness, this->metallic = metallic, this->specular = specular;
		this->transmittance = transmittance, this->refraction_index = refraction_index;
		this->emission_strength = emission_strength;
	}
};

struct Sphere : srt_sphere {  // shape.hpp:15-20
	Sphere(const vec3 &position, float radius) {
		std::memset(static_cast<srt_sphere *>(this), 0, sizeof(srt_sphere));
		this->position = to_record(position), this->radius = radius;
	}
};
struct Plane : srt_plane {  // shape.hpp:22-27
	Plane(const vec3 &position, const vec3 &normal) {
		this->position = to_record(position), this->normal = to_record(normal);
	}
};
struct Triangle : srt_triangle {  // shape.hpp:29-44
	Triangle() { std::memset(static_cast<srt_triangle *>(this), 0, sizeof(srt_triangle)); }
	Triangle(vec3 normal, vec3 pos0, vec3 pos1, vec3 pos2) {  // flat shaded
		const vec3 p[3] = {pos0, pos1, pos2};
		for (int i = 0; i < 3; ++i) vertices[i].normal = to_record(normal), vertices[i].pos = to_record(p[i]);
	}
	explicit Triangle(const srt_triangle &t) : srt_triangle(t) {}
};

struct Model : srt_model {  // shape.hpp:47-69, src/shape.cpp:35-58
	Model() { std::memset(static_cast<srt_model *>(this), 0, sizeof(srt_model)); }
	// identity transform + bounding box of its triangles
	Model(const std::vector<Triangle> &triangles, uint32_t triangle_index_, uint32_t num_triangles_) : Model() {
		triangle_index = triangle_index_, num_triangles = num_triangles_;
		transform[0].x = transform[1].y = transform[2].z = transform[3].w = 1.0f;
		compute_bounding_box(triangles);
	}
	// src/shape.cpp:45-58: min / max over transform * (pos, 1) of every vertex.  srt_model_bounds evaluates the product
	// with the operations of the kernel's own vertex transform (render.cl:114-120), so the box encloses exactly what
	// the kernel intersects.
	void compute_bounding_box(const std::vector<Triangle> &triangles) {
		if (srt_model_bounds(triangles.data(), triangles.size(), this) != SRT_OK)
			throw std::out_of_range("Model: triangle range outside the triangle list");
	}
};

struct Box {  // shape.hpp:71-77, src/shape.cpp:74-119
	static int &triangle_index() {
		static int index = -1;
		return index;
	}
	// The 12 triangles of the cube [-1, 1]^3 the reference shares between all boxes: corner i sits at
	// (i & 4 ? +1 : -1, i & 1 ? +1 : -1, i & 2 ? -1 : +1), the faces are listed in the reference's order (their order
	// decides which of two coplanar triangles wins a tie on a shared edge), normals flat and outward.
	static void create_triangle(std::vector<Triangle> &triangles) {
		static const char faces[] = "120362746504602357132376754510640315";  // src/shape.cpp:100-101, one digit per corner
		triangle_index() = static_cast<int>(triangles.size());
		auto corner = [](int i) { return vec3(i & 4 ? 1.f : -1.f, i & 1 ? 1.f : -1.f, i & 2 ? -1.f : 1.f); };
		for (int f = 0; f < 12; ++f) {
			const vec3 v1 = corner(faces[3 * f] - '0'), v2 = corner(faces[3 * f + 1] - '0'), v3 = corner(faces[3 * f + 2] - '0');
			vec3 n = cross(v2 - v1, v3 - v1);
			n = n * (dot(v1, n) > 0.0f ? 1.0f : -1.0f);  // outward
			triangles.push_back(Triangle(normalize(n), v1, v2, v3));
		}
	}
	// src/shape.cpp:76-89 as written there: the transform is translate(position) only -- `size` enters the box but not
	// the matrix, so this is self-consistent for size = 2 (what the UI passes, src/interface.cpp:162)
	static Model model(const vec3 &position, const vec3 &size) {
		if (triangle_index() == -1) throw std::runtime_error("uninitialized box model, you forgot to call Box::create_triangle");
		Model m;
		m.triangle_index = static_cast<uint32_t>(triangle_index()), m.num_triangles = 12;
		m.bounding_min = to_record(position - size * 0.5f), m.bounding_max = to_record(position + size * 0.5f);
		m.transform[0].x = m.transform[1].y = m.transform[2].z = 1.0f;
		m.transform[3] = srt_float4{position.x, position.y, position.z, 1.0f};
		return m;
	}
};

enum ShapeType { SHAPE_SPHERE = SRT_SHAPE_SPHERE, SHAPE_PLANE = SRT_SHAPE_PLANE, SHAPE_MODEL = SRT_SHAPE_MODEL };

struct Shape : srt_shape {  // shape.hpp:85-111
	Shape() { std::memset(static_cast<srt_shape *>(this), 0, sizeof(srt_shape)); }
	Shape(int32_t material_index, const Sphere &s) : Shape() { type = SHAPE_SPHERE, material = material_index, shape.sphere = s; }
	Shape(int32_t material_index, const Plane &p) : Shape() { type = SHAPE_PLANE, material = material_index, shape.plane = p; }
	Shape(int32_t material_index, const Model &m) : Shape() { type = SHAPE_MODEL, material = material_index, shape.model = m; }
};

static_assert(sizeof(Shape) == sizeof(srt_shape) && sizeof(Triangle) == sizeof(srt_triangle) &&
              sizeof(Material) == sizeof(srt_material) && sizeof(Model) == sizeof(srt_model), "helpers add no bytes");

// ---- include/parser.hpp:14-28 -------------------------------------------------------------------------------------
using ModelPair = std::pair<uint32_t, uint32_t>;  // (first triangle, count)

namespace detail {
inline std::optional<ModelPair> append_loaded(int rc, srt_triangle *loaded, size_t n, std::vector<Triangle> &triangles) {
	if (rc != SRT_OK) return std::nullopt;  // the reference returns nullopt for a file it cannot open
	const uint32_t first = static_cast<uint32_t>(triangles.size());
	for (size_t i = 0; i < n; ++i) triangles.push_back(Triangle(loaded[i]));
	srt_free(loaded);
	return ModelPair{first, static_cast<uint32_t>(n)};
}
}  // namespace detail

// Appends the file's triangles; returns where they start and how many there are, nullopt if the file cannot be read.
inline std::optional<ModelPair> load_stl_model(const std::string &filename, std::vector<Triangle> &triangles) {
	srt_triangle *loaded = nullptr;
	size_t n = 0;
	const int rc = srt_load_stl(filename.c_str(), &loaded, &n);
	return detail::append_loaded(rc, loaded, n, triangles);
}
inline std::optional<ModelPair> load_obj_model(const std::string &filename, std::vector<Triangle> &triangles) {
	srt_triangle *loaded = nullptr;
	size_t n = 0;
	const int rc = srt_load_obj(filename.c_str(), &loaded, &n);
	return detail::append_loaded(rc, loaded, n, triangles);
}
inline void save_ppm(const std::string &filename, const std::vector<uint8_t> &pixels, int width, int height) {
	if (pixels.size() < static_cast<size_t>(width) * height * 4 || srt_save_ppm(filename.c_str(), pixels.data(), width, height) != SRT_OK)
		throw std::runtime_error("save_ppm: cannot write " + filename);
}

// ---- include/helper.hpp:14-27: camera_to_world = translate(position) * eulerAngleYXZ(yaw, pitch, 0), column-major ----
struct Camera {
	vec3 position;
	float yaw = 0.f, pitch = 0.f;
	void camera_matrix(srt_float4 out[4]) const {
		const float cy = std::cos(yaw), sy = std::sin(yaw), cp = std::cos(pitch), sp = std::sin(pitch);
		out[0] = srt_float4{cy, 0.f, -sy, 0.f};            // R = Ry(yaw) * Rx(pitch), columns
		out[1] = srt_float4{sy * sp, cp, cy * sp, 0.f};
		out[2] = srt_float4{sy * cp, -sp, cy * cp, 0.f};
		out[3] = srt_float4{position.x, position.y, position.z, 1.f};
	}
};

}  // namespace srt_facade
