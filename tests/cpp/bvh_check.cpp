// CPU check of the BVH builder (simple_raytracer_b200/csrc/bvh_build.hpp): structural invariants on seeded triangle
// soups.  Prints "ok <nodes> <leaf slots> <depth>" or a diagnostic and exits non-zero.
//   bvh_check <n_triangles> <seed> <mode>     mode 0 = random soup, 1 = all triangles identical, 2 = with NaN / inf vertices
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "bvh_build.hpp"

using srt_bvh::Node;

static bool inside(const float *lo, const float *hi, const float *p) {
	for (int k = 0; k < 3; ++k)
		if (std::isfinite(p[k]) && !(p[k] >= lo[k] && p[k] <= hi[k])) return false;
	return true;
}

struct Walk {
	const std::vector<Node> &nodes;
	const std::vector<int32_t> &order;
	const std::vector<float> &hot;
	std::vector<int> seen;
	int max_depth = 0;
	bool ok = true;
	void child(const float *lo, const float *hi, int c, int n, int depth, const float *plo, const float *phi) {
		if (n < 0) return;
		max_depth = std::max(max_depth, depth);
		if (plo)  // a child's box lies inside its parent's (both padded by the same amount)
			for (int k = 0; k < 3; ++k)
				if (lo[k] <= hi[k] && !(lo[k] >= plo[k] && hi[k] <= phi[k])) ok = false, std::printf("child box outside parent\n");
		if (n > 0) {
			if (n > srt_bvh::LEAF_TRIS) ok = false, std::printf("leaf of %d triangles\n", n);
			for (int i = 0; i < n; ++i) {
				const int tri = order[c + i];
				if (i && order[c + i - 1] >= tri) ok = false, std::printf("leaf not in ascending triangle order\n");
				seen[tri]++;
				const float *t = &hot[12 * (size_t)tri];
				const float v1[3] = {t[0] + t[4], t[1] + t[5], t[2] + t[6]}, v2[3] = {t[0] + t[8], t[1] + t[9], t[2] + t[10]};
				if (!inside(lo, hi, t) || !inside(lo, hi, v1) || !inside(lo, hi, v2)) ok = false, std::printf("vertex outside its leaf box\n");
			}
			return;
		}
		const Node &nd = nodes[c];
		child(nd.lo0, nd.hi0, nd.c0, nd.n0, depth + 1, lo, hi);
		child(nd.lo1, nd.hi1, nd.c1, nd.n1, depth + 1, lo, hi);
	}
};

int main(int argc, char **argv) {
	const int n = argc > 1 ? std::atoi(argv[1]) : 1000, seed = argc > 2 ? std::atoi(argv[2]) : 1, mode = argc > 3 ? std::atoi(argv[3]) : 0;
	std::mt19937 rng(seed);
	std::uniform_real_distribution<float> pos(-5.f, 5.f), edge(-0.3f, 0.3f);
	const int first = 7;  // the model does not start at triangle 0
	std::vector<float> hot(12 * (size_t)(first + n), 0.f);
	for (int i = 0; i < n; ++i) {
		float *t = &hot[12 * (size_t)(first + i)];
		for (int k = 0; k < 3; ++k) {
			t[k] = mode == 1 ? 1.0f : pos(rng);
			t[4 + k] = mode == 1 ? 0.5f : edge(rng);
			t[8 + k] = mode == 1 ? -0.25f : edge(rng);
		}
		if (mode == 2 && i % 97 == 0) t[i % 3] = NAN;
		if (mode == 2 && i % 89 == 0) t[4 + i % 3] = INFINITY;
	}
	std::vector<Node> nodes(3);  // another model's nodes come first: indices must be absolute
	std::vector<int32_t> order(5, -1);
	const int root = (int)nodes.size();
	const int depth = srt_bvh::build(hot.data(), first, n, nodes, order);
	Walk w{nodes, order, hot, std::vector<int>(first + n, 0)};
	const Node &r = nodes[root];
	if (r.n1 != -1) w.ok = false, std::printf("entry node's second child must be empty\n");
	w.child(r.lo0, r.hi0, r.c0, r.n0, 1, nullptr, nullptr);
	for (int i = 0; i < first; ++i)
		if (w.seen[i]) w.ok = false, std::printf("triangle %d outside the model appears\n", i);
	for (int i = first; i < first + n; ++i)
		if (w.seen[i] != 1) w.ok = false, std::printf("triangle %d appears %d times\n", i, w.seen[i]);
	if (w.max_depth != depth || depth > srt_bvh::MAX_DEPTH) w.ok = false, std::printf("depth %d (walk %d)\n", depth, w.max_depth);
	if (!w.ok) return 1;
	std::printf("ok %zu %zu %d\n", nodes.size() - root, order.size() - 5, depth);
	return 0;
}
