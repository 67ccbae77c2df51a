// bvh_build.hpp -- host-side builder of the OPTIONAL bounding-volume hierarchy (srt_set_accel(SRT_ACCEL_BVH)).
//
// The reference has no acceleration structure: every ray tests every triangle of every model whose box it enters
// (reference src/render.cl:299-367, loop :324); a BVH is the first item of its own "Future plans"
// (reference README.md:41).  This is that extension, kept OUT of the parity-graded path: it is off by default, and
// when it is on the kernel visits only the triangles whose boxes the ray enters, so the rounding-noise "hits" the
// brute-force loop can produce on far-away, nearly edge-on triangles are not reproduced.
// What IS preserved: the exact test runs the reference arithmetic on the reference operands (tri_hot), the closest hit
// wins, equal t goes to the lowest triangle index, and an earlier shape keeps an equal t.
//
// Input: the world-space triangles exactly as the kernel intersects them (tri_hot: v0, e1, e2 as 3 x float4).
// Output: 64-byte nodes holding BOTH children's boxes, and the triangle order of the leaves.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace srt_bvh {

struct Node {  // 4 x float4 on the device
	float lo0[3];
	int32_t c0;  // child 0: node index (n0 == 0) or first slot in the leaf order (n0 > 0); absolute indices
	float hi0[3];
	int32_t n0;  // number of triangles if child 0 is a leaf, 0 for an inner node, -1 for "no child"
	float lo1[3];
	int32_t c1;
	float hi1[3];
	int32_t n1;
};
static_assert(sizeof(Node) == 64, "node");

constexpr int LEAF_TRIS = 4;
constexpr int SAH_BINS = 16;
constexpr int SAH_DEPTH = 16;  // below this depth: median splits, which bound the depth by SAH_DEPTH + log2(n)
constexpr int MAX_DEPTH = 48;  // traversal stack of the kernel

struct Box {
	float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
	void grow(const float p[3]) {
		for (int k = 0; k < 3; ++k) lo[k] = std::min(lo[k], p[k]), hi[k] = std::max(hi[k], p[k]);
	}
	void grow(const Box &b) {
		for (int k = 0; k < 3; ++k) lo[k] = std::min(lo[k], b.lo[k]), hi[k] = std::max(hi[k], b.hi[k]);
	}
	float area() const {
		const float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
		return x < 0 ? 0.0f : 2.0f * (x * y + y * z + z * x);
	}
};

struct Prim {
	Box box;
	float c[3];
	int32_t tri;
};

struct Builder {
	std::vector<Prim> prims;
	std::vector<Node> &nodes;
	std::vector<int32_t> &order;
	float pad;
	int max_depth = 0;

	Builder(std::vector<Node> &n, std::vector<int32_t> &o) : nodes(n), order(o), pad(0) {}

	Box bounds(int b, int e) const {
		Box x;
		for (int i = b; i < e; ++i) x.grow(prims[i].box);
		return x;
	}

	// Fills child slot (lo, hi, c, n) of a parent for the primitive range [b, e): a leaf or a new inner node.
	void make_child(int b, int e, int depth, float *lo, float *hi, int32_t &c, int32_t &n) {
		const Box bx = bounds(b, e);
		for (int k = 0; k < 3; ++k) lo[k] = bx.lo[k] - pad, hi[k] = bx.hi[k] + pad;
		max_depth = std::max(max_depth, depth);
		if (e - b <= LEAF_TRIS) {
			// leaf: triangles in ascending index order (ties on equal t resolve to the lowest index either way)
			std::sort(prims.begin() + b, prims.begin() + e, [](const Prim &x, const Prim &y) { return x.tri < y.tri; });
			c = (int32_t)order.size();
			n = e - b;
			for (int i = b; i < e; ++i) order.push_back(prims[i].tri);
			return;
		}
		const int me = (int)nodes.size();
		nodes.emplace_back();
		c = me;
		n = 0;
		const int mid = split(b, e, depth);
		Node nd{};
		make_child(b, mid, depth + 1, nd.lo0, nd.hi0, nd.c0, nd.n0);
		make_child(mid, e, depth + 1, nd.lo1, nd.hi1, nd.c1, nd.n1);
		nodes[me] = nd;
	}

	int split(int b, int e, int depth) {
		Box cb;
		for (int i = b; i < e; ++i) cb.grow(prims[i].c);
		int axis = 0;
		float ext[3] = {cb.hi[0] - cb.lo[0], cb.hi[1] - cb.lo[1], cb.hi[2] - cb.lo[2]};
		if (ext[1] > ext[axis]) axis = 1;
		if (ext[2] > ext[axis]) axis = 2;
		const int median = b + (e - b) / 2;
		auto by_centroid = [axis](const Prim &x, const Prim &y) { return x.c[axis] < y.c[axis] || (x.c[axis] == y.c[axis] && x.tri < y.tri); };
		if (depth >= SAH_DEPTH || !(ext[axis] > 0.0f) || !std::isfinite(ext[axis])) {
			std::nth_element(prims.begin() + b, prims.begin() + median, prims.begin() + e, by_centroid);
			return median;
		}
		// binned surface-area heuristic along the longest centroid axis
		Box bin_box[SAH_BINS];
		int bin_n[SAH_BINS] = {0};
		const float k1 = SAH_BINS * (1.0f - 1e-6f) / ext[axis];
		auto bin_of = [&](const Prim &p) { return std::min(SAH_BINS - 1, std::max(0, (int)(k1 * (p.c[axis] - cb.lo[axis])))); };
		for (int i = b; i < e; ++i) {
			const int k = bin_of(prims[i]);
			bin_box[k].grow(prims[i].box);
			bin_n[k]++;
		}
		float right_area[SAH_BINS];
		Box acc;
		for (int k = SAH_BINS - 1; k > 0; --k) {
			acc.grow(bin_box[k]);
			right_area[k] = acc.area();
		}
		Box left;
		int left_n = 0, best = -1;
		float best_cost = INFINITY;
		for (int k = 0; k + 1 < SAH_BINS; ++k) {
			left.grow(bin_box[k]);
			left_n += bin_n[k];
			const int right_n = (e - b) - left_n;
			if (left_n == 0 || right_n == 0) continue;
			const float cost = left.area() * left_n + right_area[k + 1] * right_n;
			if (cost < best_cost) best_cost = cost, best = k;
		}
		if (best < 0) {
			std::nth_element(prims.begin() + b, prims.begin() + median, prims.begin() + e, by_centroid);
			return median;
		}
		auto it = std::partition(prims.begin() + b, prims.begin() + e, [&](const Prim &p) { return bin_of(p) <= best; });
		const int mid = (int)(it - prims.begin());
		if (mid == b || mid == e) {
			std::nth_element(prims.begin() + b, prims.begin() + median, prims.begin() + e, by_centroid);
			return median;
		}
		return mid;
	}
};

// Builds the hierarchy of one model over triangles [first, first + count) of `hot` (12 floats per triangle: v0.xyzw,
// e1.xyzw, e2.xyzw).  Appends to nodes / order (all indices absolute; the model's entry node is the first one appended)
// and returns the depth.
// Boxes are padded by 2^-18 of the model's largest coordinate magnitude (plus its extent): the exact test's u, v, t
// carry relative errors of a few 2^-24 of those magnitudes, and the kernel's slab test adds its own few ulps.
inline int build(const float *hot, int first, int count, std::vector<Node> &nodes, std::vector<int32_t> &order) {
	Builder bl(nodes, order);
	bl.prims.resize(count);
	float mag = 0.0f;
	for (int i = 0; i < count; ++i) {
		const float *t = hot + 12 * (size_t)(first + i);
		Prim &p = bl.prims[i];
		p.tri = first + i;
		const float v1[3] = {t[0] + t[4], t[1] + t[5], t[2] + t[6]}, v2[3] = {t[0] + t[8], t[1] + t[9], t[2] + t[10]};
		p.box.grow(t);
		p.box.grow(v1);
		p.box.grow(v2);
		for (int k = 0; k < 3; ++k) {
			p.c[k] = 0.5f * (p.box.lo[k] + p.box.hi[k]);
			const float m = std::max(std::fabs(p.box.lo[k]), std::fabs(p.box.hi[k]));
			if (std::isfinite(m)) mag = std::max(mag, m);
			if (!std::isfinite(p.c[k])) p.c[k] = 0.0f;  // NaN / inf vertices: the triangle can never be hit; park it anywhere
		}
	}
	bl.pad = mag * (1.0f / 262144.0f) + 1e-30f;
	Node root{};  // node 0 of the model is a pseudo-parent whose child 0 is the real root and child 1 is empty
	const int me = (int)nodes.size();
	nodes.emplace_back();
	bl.make_child(0, count, 1, root.lo0, root.hi0, root.c0, root.n0);
	for (int k = 0; k < 3; ++k) root.lo1[k] = INFINITY, root.hi1[k] = -INFINITY;
	root.c1 = 0, root.n1 = -1;
	nodes[me] = root;
	return bl.max_depth;
}

}  // namespace srt_bvh
