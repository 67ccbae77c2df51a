"""Exact division by launch constants (simple_raytracer_b200/csrc/fastdiv.hpp): the item -> (launch, pixel, sample, row)
mapping of start_path (render.cl:488-496 derives the same numbers from get_global_id) must equal `/` for every 32-bit
dividend.  The header is plain C++, so it is compiled here with g++ and driven against the machine's division."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r"""
#include <cstdio>
#include <cstdlib>
#include <random>
#include "fastdiv.hpp"
static unsigned long long bad = 0, checked = 0;
static void check(uint32_t n, uint32_t d, srt::FastDiv f) {
	++checked;
	if (srt::fast_div(n, f.m, f.s) != (d <= 1u ? n : n / d)) {
		if (bad++ < 5) std::printf("BAD n=%u d=%u got=%u\n", n, d, srt::fast_div(n, f.m, f.s));
	}
}
static void divisor(uint32_t d, std::mt19937 &rng, int exhaustive) {
	const srt::FastDiv f = srt::fast_div_make(d);
	if (exhaustive) {  // every 32-bit dividend
		unsigned long long wrong = 0;
#pragma omp parallel for reduction(+ : wrong) schedule(static)
		for (long long n = 0; n <= 0xffffffffll; ++n) wrong += srt::fast_div((uint32_t)n, f.m, f.s) != (d <= 1u ? (uint32_t)n : (uint32_t)n / d);
		checked += 1ull << 32;
		bad += wrong;
		if (wrong) std::printf("BAD d=%u: %llu dividends\n", d, wrong);
		return;
	}
	const uint32_t edge[] = {0u, 1u, d - 1u, d, d + 1u, 2u * d - 1u, 2u * d, 0x7fffffffu, 0x80000000u, 0xfffffffeu, 0xffffffffu};
	for (uint32_t n : edge) check(n, d, f);
	const uint32_t dd = d ? d : 1u, qmax = 0xffffffffu / dd;
	for (int k = 0; k < 64; ++k) {  // multiples of d and their neighbours: where a wrong multiplier shows
		const uint32_t q = qmax == 0xffffffffu ? rng() : rng() % (qmax + 1u);
		const unsigned long long base = (unsigned long long)q * dd;
		for (int e = -1; e <= 1; ++e) {
			const long long n = (long long)base + e;
			if (n >= 0 && n <= 0xffffffffll) check((uint32_t)n, d, f);
		}
		check(rng(), d, f);
	}
}
int main(int argc, char **argv) {
	std::mt19937 rng(12345);
	for (uint32_t d = 0; d < 70000u; ++d) divisor(d, rng, 0);
	for (int k = 1; k < 32; ++k)
		for (int e = -2; e <= 2; ++e) divisor((1u << k) + (uint32_t)e, rng, 0);
	for (int k = 0; k < 200000; ++k) divisor(rng(), rng, 0);
	const uint32_t named[] = {0xffffffffu, 0xfffffffeu, 0x80000001u, 1920u, 3840u, 1920u * 1080u * 4u, 3840u * 2160u * 16u, 800u * 600u};
	for (uint32_t d : named) divisor(d, rng, 0);
	// every 32-bit dividend for a few divisors of each kind (identity, powers of two, small odd, frame widths, an item count)
	const uint32_t full[] = {1u, 4u, 7u, 1920u, 1920u * 1080u * 4u};
	if (argc > 1)
		for (uint32_t d : full) divisor(d, rng, 1);
	std::printf("checked %llu bad %llu\n", checked, bad);
	return bad ? 1 : 0;
}
"""


def test_fast_div_equals_division(tmp_path):
    src = tmp_path / "fastdiv_check.cpp"
    src.write_text(HARNESS)
    exe = tmp_path / "fastdiv_check"
    subprocess.run(["g++", "-std=c++17", "-O2", "-fopenmp", "-I", os.path.join(ROOT, "simple_raytracer_b200", "csrc"), "-o", str(exe), str(src)],
                   check=True)
    out = subprocess.run([str(exe), "full"], capture_output=True, text=True, timeout=300)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert " bad 0" in out.stdout
