// mesh_io.cpp -- data formats either side of the hot path.
//
// Re-implements what reference src/parser.cpp provides -- load_stl_model (:17-53), load_obj_model
// (:55-135), save_ppm (:4-15) -- and Model::compute_bounding_box (src/shape.cpp:45-58), producing the
// same 96-byte Triangle records (include/srt.h) that srt_upload_scene turns into device SoA buffers.
// Differences from the reference, all on inputs the reference mishandles:
//   * files are validated (short reads, a triangle count the file is too short for, index ranges) and errors are
//     returned instead of crashing; no C++ exception leaves an extern "C" function;
//   * OBJ negative indices count back from the end of the list (the reference computes len-idx+1);
//   * OBJ faces without normals get the flat geometric normal (the reference reads an
//     uninitialised index); faces with more than 3 corners are fan-triangulated (the reference
//     keeps the first three corners only).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#include "../../include/srt.h"

namespace {

struct V3 {
	float x, y, z;
};

srt_triangle make_triangle(const V3 p[3], const V3 n[3]) {
	srt_triangle t;
	memset(&t, 0, sizeof t);
	for (int i = 0; i < 3; ++i) {
		t.vertices[i].pos.x = p[i].x, t.vertices[i].pos.y = p[i].y, t.vertices[i].pos.z = p[i].z;
		t.vertices[i].normal.x = n[i].x, t.vertices[i].normal.y = n[i].y, t.vertices[i].normal.z = n[i].z;
	}
	return t;
}

V3 normalized(V3 v) {  // glm::normalize = v * inversesqrt(dot(v, v)), parser.cpp:83
	float inv = 1.0f / std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
	return V3{v.x * inv, v.y * inv, v.z * inv};
}

int hand_over(std::vector<srt_triangle> &tris, srt_triangle **out, size_t *count) {
	*count = tris.size();
	*out = nullptr;
	if (tris.empty()) return SRT_OK;
	*out = static_cast<srt_triangle *>(malloc(tris.size() * sizeof(srt_triangle)));
	if (!*out) return SRT_ERR_INVALID;
	memcpy(*out, tris.data(), tris.size() * sizeof(srt_triangle));
	return SRT_OK;
}

// one "v", "v/t", "v//n" or "v/t/n" corner; returns false on a malformed token
bool parse_corner(const char *&s, long &v, long &n, bool &has_n) {
	char *end;
	v = strtol(s, &end, 10);
	if (end == s) return false;
	s = end;
	has_n = false;
	if (*s == '/') {
		++s;
		if (*s != '/') {  // texture index, ignored (parser.cpp:91-94)
			strtol(s, &end, 10);
			s = end;
		}
		if (*s == '/') {
			++s;
			n = strtol(s, &end, 10);
			if (end == s) return false;
			s = end;
			has_n = true;
		}
	}
	return true;
}

bool resolve(long idx, size_t len, size_t &out) {  // 1-based; negative = from the end
	long r = idx > 0 ? idx - 1 : (long)len + idx;
	if (idx == 0 || r < 0 || (size_t)r >= len) return false;
	out = (size_t)r;
	return true;
}

}  // namespace

extern "C" {

void srt_free(void *p) { free(p); }

static int load_stl(const char *path, srt_triangle **triangles, size_t *count) {
	FILE *f = fopen(path, "rb");
	if (!f) return SRT_ERR_INVALID;  // reference returns nullopt, parser.cpp:20-22
	uint8_t header[84];
	if (fread(header, 1, 84, f) != 84) {
		fclose(f);
		return SRT_ERR_INVALID;
	}
	uint32_t n;
	memcpy(&n, header + 80, 4);
	// the count comes from the file: it must fit the bytes that are actually there before anything is reserved
	long size = -1;
	if (fseek(f, 0, SEEK_END) == 0) size = ftell(f);
	if (size < 84 || (uint64_t)n > (uint64_t)(size - 84) / 50 || fseek(f, 84, SEEK_SET) != 0) {
		fclose(f);
		return SRT_ERR_INVALID;
	}
	std::vector<srt_triangle> tris;
	tris.reserve(n);
	for (uint32_t i = 0; i < n; ++i) {
		uint8_t rec[50];  // float normal[3], v1[3], v2[3], v3[3]; uint16 attribute (packed)
		if (fread(rec, 1, 50, f) != 50) {
			fclose(f);
			return SRT_ERR_INVALID;
		}
		float v[12];
		memcpy(v, rec, 48);
		V3 nrm = {v[0], v[1], v[2]};  // facet normal copied as-is to all three vertices (parser.cpp:46-50)
		V3 p[3] = {{v[3], v[4], v[5]}, {v[6], v[7], v[8]}, {v[9], v[10], v[11]}};
		V3 nn[3] = {nrm, nrm, nrm};
		tris.push_back(make_triangle(p, nn));
	}
	fclose(f);
	return hand_over(tris, triangles, count);
}

int srt_load_stl(const char *path, srt_triangle **triangles, size_t *count) {
	if (!path || !triangles || !count) return SRT_ERR_INVALID;
	*triangles = nullptr;
	*count = 0;
	try {
		return load_stl(path, triangles, count);
	} catch (const std::exception &) {  // bad_alloc and friends must not cross the C boundary
		return SRT_ERR_INVALID;
	}
}

static int load_obj(const char *path, srt_triangle **triangles, size_t *count) {
	FILE *f = fopen(path, "r");
	if (!f) return SRT_ERR_INVALID;
	std::vector<V3> verts, normals;
	struct Corner {
		long v, n;
		bool has_n;
	};
	std::vector<std::vector<Corner>> faces;
	std::string line;
	char buf[4096];
	while (fgets(buf, sizeof buf, f)) {
		line = buf;
		while (!line.empty() && line.back() != '\n' && fgets(buf, sizeof buf, f)) line += buf;
		const char *s = line.c_str();
		while (*s == ' ' || *s == '\t') ++s;
		if (s[0] == 'v' && (s[1] == ' ' || s[1] == '\t')) {
			V3 v;
			if (sscanf(s + 2, "%f %f %f", &v.x, &v.y, &v.z) != 3) {
				fclose(f);
				return SRT_ERR_INVALID;
			}
			verts.push_back(v);
		} else if (s[0] == 'v' && s[1] == 'n' && (s[2] == ' ' || s[2] == '\t')) {
			V3 v;
			if (sscanf(s + 3, "%f %f %f", &v.x, &v.y, &v.z) != 3) {
				fclose(f);
				return SRT_ERR_INVALID;
			}
			normals.push_back(normalized(v));
		} else if (s[0] == 'f' && (s[1] == ' ' || s[1] == '\t')) {
			s += 2;
			std::vector<Corner> face;
			for (;;) {
				while (*s == ' ' || *s == '\t') ++s;
				if (*s == '\0' || *s == '\n' || *s == '\r') break;
				Corner c{0, 0, false};
				if (!parse_corner(s, c.v, c.n, c.has_n)) {
					fclose(f);
					return SRT_ERR_INVALID;
				}
				face.push_back(c);
			}
			if (face.size() < 3) {
				fclose(f);
				return SRT_ERR_INVALID;
			}
			faces.push_back(face);
		}  // '#', 's', 'vt', 'o', 'g', 'usemtl', ... ignored (parser.cpp:74-76,107-109)
	}
	fclose(f);

	std::vector<srt_triangle> tris;
	tris.reserve(faces.size());
	for (const auto &face : faces) {
		for (size_t k = 1; k + 1 < face.size(); ++k) {
			const Corner *c[3] = {&face[0], &face[k], &face[k + 1]};
			V3 p[3], n[3];
			bool all_n = true;
			for (int i = 0; i < 3; ++i) {
				size_t vi;
				if (!resolve(c[i]->v, verts.size(), vi)) return SRT_ERR_INVALID;
				p[i] = verts[vi];
				all_n = all_n && c[i]->has_n;
			}
			if (all_n) {
				for (int i = 0; i < 3; ++i) {
					size_t ni;
					if (!resolve(c[i]->n, normals.size(), ni)) return SRT_ERR_INVALID;
					n[i] = normals[ni];
				}
			} else {
				V3 a = {p[1].x - p[0].x, p[1].y - p[0].y, p[1].z - p[0].z};
				V3 b = {p[2].x - p[0].x, p[2].y - p[0].y, p[2].z - p[0].z};
				V3 g = normalized(V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x});
				n[0] = n[1] = n[2] = g;
			}
			tris.push_back(make_triangle(p, n));
		}
	}
	return hand_over(tris, triangles, count);
}

int srt_load_obj(const char *path, srt_triangle **triangles, size_t *count) {
	if (!path || !triangles || !count) return SRT_ERR_INVALID;
	*triangles = nullptr;
	*count = 0;
	try {
		return load_obj(path, triangles, count);
	} catch (const std::exception &) {
		return SRT_ERR_INVALID;
	}
}

static int save_ppm(const char *path, const uint8_t *argb, int width, int height) {
	FILE *f = fopen(path, "wb");
	if (!f) return SRT_ERR_INVALID;
	fprintf(f, "P6 %d %d 255\n", width, height);  // parser.cpp:7-8
	std::vector<uint8_t> row((size_t)width * 3);
	for (int y = 0; y < height; ++y) {
		const uint8_t *p = argb + (size_t)y * width * 4;
		for (int x = 0; x < width; ++x) {  // bytes 1..3 of every A,R,G,B pixel, parser.cpp:10-14
			row[3 * x + 0] = p[4 * x + 1];
			row[3 * x + 1] = p[4 * x + 2];
			row[3 * x + 2] = p[4 * x + 3];
		}
		if (fwrite(row.data(), 1, row.size(), f) != row.size()) {
			fclose(f);
			return SRT_ERR_INVALID;
		}
	}
	fclose(f);
	return SRT_OK;
}

int srt_save_ppm(const char *path, const uint8_t *argb, int width, int height) {
	if (!path || !argb || width <= 0 || height <= 0) return SRT_ERR_INVALID;
	try {
		return save_ppm(path, argb, width, height);
	} catch (const std::exception &) {
		return SRT_ERR_INVALID;
	}
}

int srt_model_bounds(const srt_triangle *triangles, size_t n_triangles, srt_model *model) {
	if (!model || (!triangles && n_triangles)) return SRT_ERR_INVALID;
	if ((size_t)model->triangle_index + model->num_triangles > n_triangles) return SRT_ERR_INVALID;
	float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
	const srt_float4 *m = model->transform;
	for (uint32_t i = 0; i < model->num_triangles; ++i) {
		const srt_triangle &t = triangles[model->triangle_index + i];
		for (int j = 0; j < 3; ++j) {
			const srt_float3 &p = t.vertices[j].pos;
			// the operations of the device pre-transform (prepare_triangles_kernel; render.cl:114-120 with w = 1):
			// every product and sum rounded on its own (this file is built with -ffp-contract=off), so the box
			// encloses exactly the world-space vertices the kernel intersects
			float w[3] = {m[0].x * p.x + m[1].x * p.y + m[2].x * p.z + m[3].x * 1.0f,
			              m[0].y * p.x + m[1].y * p.y + m[2].y * p.z + m[3].y * 1.0f,
			              m[0].z * p.x + m[1].z * p.y + m[2].z * p.z + m[3].z * 1.0f};
			for (int c = 0; c < 3; ++c) {
				mn[c] = w[c] < mn[c] ? w[c] : mn[c];
				mx[c] = w[c] > mx[c] ? w[c] : mx[c];
			}
		}
	}
	model->bounding_min.x = mn[0], model->bounding_min.y = mn[1], model->bounding_min.z = mn[2];
	model->bounding_max.x = mx[0], model->bounding_max.y = mx[1], model->bounding_max.z = mx[2];
	return SRT_OK;
}

}  // extern "C"
