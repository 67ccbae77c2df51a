// fastdiv.hpp -- exact unsigned 32-bit division by a launch constant without a division instruction.
//
// Granlund-Montgomery round-up multiplier in the overflow-free "add and halve" form: with l = floor(log2 d) and
//   m = floor(2^(33+l) / d) - 2^32 + 1   (the low 32 bits of the 33-bit round-up multiplier),
//   n / d = (((n - t) >> 1) + t) >> l    where t = umulhi(n, m),
// for EVERY 32-bit n and every d >= 2 that is not a power of two; a power of two 2^k takes m = 0 and the shift k - 1
// (t = 0, (n >> 1) >> (k - 1)).  s = 32 marks d <= 1: the quotient is n itself.
// Plain C++ (no CUDA) so that tests/test_fastdiv.py can compile it with g++ and check it against `/`.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SRT_HD __host__ __device__ __forceinline__
#else
#define SRT_HD inline
#endif

namespace srt {

struct FastDiv {
	uint32_t m, s;
};

inline FastDiv fast_div_make(uint32_t d) {
	if (d <= 1u) return FastDiv{0u, 32u};
	uint32_t l = 31;
	while ((d >> l) == 0u) --l;  // floor(log2 d), d >= 2
	if ((d & (d - 1u)) == 0u) return FastDiv{0u, l - 1u};
	const unsigned long long two = 1ull << (32 + l);  // l <= 31: fits
	unsigned long long pm = two / d;
	const unsigned long long rem = two - pm * d;
	pm *= 2ull;
	if (rem * 2ull >= d) pm += 1ull;  // floor(2^(33+l) / d)
	return FastDiv{(uint32_t)(pm + 1ull), l};
}

SRT_HD uint32_t fast_div(uint32_t n, uint32_t m, uint32_t s) {
#if defined(__CUDA_ARCH__)
	const uint32_t t = __umulhi(n, m);
#else
	const uint32_t t = (uint32_t)(((unsigned long long)n * m) >> 32);
#endif
	const uint32_t q = (((n - t) >> 1) + t) >> (s & 31u);
	return s >= 32u ? n : q;
}

}  // namespace srt
