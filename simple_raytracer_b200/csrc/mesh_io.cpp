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

#include <zlib.h>

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

// ---- sky-box image: PNG -> RGBA float32 the way the reference prepares it (src/tracer.cpp:42-52) --------------------
// stbi_set_flip_vertically_on_load(1) + stbi_loadf_from_file(..., 4): the 8-bit image is expanded to 4 channels, rows
// bottom-up, colour = (float)pow(x / 255.0f, 2.2f) evaluated in double, alpha = x / 255.0f (lib/stb_image.h:1868-1874).
// stb_image is the reference's vendored third-party code and is not copied: this is a small PNG reader of its own
// (8-bit grey / grey+alpha / RGB / RGBA / palette, non-interlaced -- what image editors write), zlib does the inflate.
static uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

static int load_skybox_png(const char *path, float **rgba, int *width, int *height) {
	FILE *f = fopen(path, "rb");
	if (!f) return SRT_ERR_INVALID;
	std::vector<uint8_t> file;
	uint8_t buf[65536];
	for (size_t n; (n = fread(buf, 1, sizeof buf, f)) > 0;) file.insert(file.end(), buf, buf + n);
	fclose(f);
	static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
	if (file.size() < 8 + 25 || memcmp(file.data(), sig, 8) != 0) return SRT_ERR_INVALID;
	uint32_t w = 0, h = 0;
	int depth = 0, ctype = -1, interlace = 0;
	std::vector<uint8_t> idat, palette, trns;
	for (size_t at = 8; at + 12 <= file.size();) {
		const uint32_t len = be32(&file[at]);
		const uint8_t *type = &file[at + 4], *data = &file[at + 8];
		if (len > file.size() - at - 12) return SRT_ERR_INVALID;
		if (!memcmp(type, "IHDR", 4) && len == 13) {
			w = be32(data), h = be32(data + 4), depth = data[8], ctype = data[9], interlace = data[12];
		} else if (!memcmp(type, "PLTE", 4)) {
			palette.assign(data, data + len);
		} else if (!memcmp(type, "tRNS", 4)) {
			trns.assign(data, data + len);
		} else if (!memcmp(type, "IDAT", 4)) {
			idat.insert(idat.end(), data, data + len);
		} else if (!memcmp(type, "IEND", 4)) {
			break;
		}
		at += 12 + (size_t)len;
	}
	const int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
	if (!w || !h || w > 32768 || h > 32768 || depth != 8 || !channels || interlace || idat.empty()) return SRT_ERR_INVALID;
	if (ctype == 3 && palette.size() < 3) return SRT_ERR_INVALID;
	const size_t stride = (size_t)w * channels;
	std::vector<uint8_t> raw((stride + 1) * h);
	uLongf raw_len = (uLongf)raw.size();
	if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size()) return SRT_ERR_INVALID;
	// undo the per-row filters (PNG specification, section 9), in place over a zero row above the first
	std::vector<uint8_t> img(stride * h), zero(stride, 0);
	for (uint32_t y = 0; y < h; ++y) {
		const uint8_t *in = &raw[(stride + 1) * y];
		uint8_t *out = &img[stride * y];
		const uint8_t *up = y ? out - stride : zero.data();
		const int ft = in[0];
		if (ft > 4) return SRT_ERR_INVALID;
		for (size_t i = 0; i < stride; ++i) {
			const int a = i >= (size_t)channels ? out[i - channels] : 0, b = up[i], c = i >= (size_t)channels ? up[i - channels] : 0;
			int pred = 0;
			if (ft == 1) pred = a;
			else if (ft == 2) pred = b;
			else if (ft == 3) pred = (a + b) >> 1;
			else if (ft == 4) {
				const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
				pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
			}
			out[i] = (uint8_t)(in[1 + i] + pred);
		}
	}
	float lut[256], lin[256];
	for (int i = 0; i < 256; ++i) {
		lut[i] = (float)pow((double)(i / 255.0f), (double)2.2f);  // stbi__ldr_to_hdr, lib/stb_image.h:1868
		lin[i] = i / 255.0f;
	}
	float *px = static_cast<float *>(malloc((size_t)w * h * 4 * sizeof(float)));
	if (!px) return SRT_ERR_INVALID;
	for (uint32_t y = 0; y < h; ++y) {
		const uint8_t *row = &img[stride * y];
		float *dst = px + (size_t)(h - 1 - y) * w * 4;  // memory row 0 = image bottom (flip on load)
		for (uint32_t x = 0; x < w; ++x) {
			uint8_t r, g, b, a = 255;
			if (ctype == 0) r = g = b = row[x];
			else if (ctype == 4) r = g = b = row[2 * x], a = row[2 * x + 1];
			else if (ctype == 2) r = row[3 * x], g = row[3 * x + 1], b = row[3 * x + 2];
			else if (ctype == 6) r = row[4 * x], g = row[4 * x + 1], b = row[4 * x + 2], a = row[4 * x + 3];
			else {
				const size_t k = row[x];
				if (3 * k + 2 >= palette.size()) {
					free(px);
					return SRT_ERR_INVALID;
				}
				r = palette[3 * k], g = palette[3 * k + 1], b = palette[3 * k + 2];
				if (k < trns.size()) a = trns[k];
			}
			dst[4 * x] = lut[r], dst[4 * x + 1] = lut[g], dst[4 * x + 2] = lut[b], dst[4 * x + 3] = lin[a];
		}
	}
	*rgba = px, *width = (int)w, *height = (int)h;
	return SRT_OK;
}

int srt_load_skybox_png(const char *path, float **rgba, int *width, int *height) {
	if (!path || !rgba || !width || !height) return SRT_ERR_INVALID;
	*rgba = nullptr;
	*width = *height = 0;
	try {
		return load_skybox_png(path, rgba, width, height);
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
