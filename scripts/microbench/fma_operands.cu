// Developer micro-benchmark (sm_100a): cost of FP32 FMA-pipe instructions as a function of how many DISTINCT vector
// registers they read, scalar (FFMA) vs packed (FFMA2), at W warps per SM sub-partition.  Prints SM cycles per
// warp-instruction per sub-partition (1.0 = the issue limit).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o build/fma_operands scripts/microbench/fma_operands.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 512
#define UNROLL 16
#define N 8  // independent chains per thread

template <int MODE>
__global__ void __launch_bounds__(128) k(float *out, const float *in) {
	float a[N], b[N], c[N];
	float2 a2[N], b2[N], c2[N];
	for (int i = 0; i < N; ++i) {  // opaque values: nothing folds into immediates
		a[i] = in[threadIdx.x + i], b[i] = in[64 + threadIdx.x + i], c[i] = in[128 + i];
		a2[i] = make_float2(a[i], in[200 + i]), b2[i] = make_float2(b[i], in[220 + i]), c2[i] = make_float2(c[i], in[240 + i]);
	}
	float u = in[300 + threadIdx.x], v = in[301];
	float2 u2 = make_float2(u, in[302]);
#pragma unroll 1
	for (int it = 0; it < ITERS; ++it) {
#pragma unroll
		for (int r = 0; r < UNROLL; ++r) {
#pragma unroll
			for (int i = 0; i < N; ++i) {
				if (MODE == 0) c[i] = __fmaf_rn(a[i], b[i], c[i]);        // 3 distinct registers
				if (MODE == 1) c[i] = __fmaf_rn(u, b[i], c[i]);           // one operand shared by consecutive instructions
				if (MODE == 2) c[i] = __fmaf_rn(u, v, c[i]);              // two shared
				if (MODE == 3) c[i] = __fmaf_rn(c[i], 1.0000001f, 1e-9f); // immediates
				if (MODE == 4) c2[i] = __ffma2_rn(a2[i], b2[i], c2[i]);   // packed, 3 distinct register pairs
				if (MODE == 5) c2[i] = __ffma2_rn(u2, b2[i], c2[i]);      // packed, one pair shared
				if (MODE == 6) { c2[i] = __ffma2_rn(u2, b2[i], c2[i]); c[i] = __fmaf_rn(u, b[i], c[i]); a[i] = __fmaf_rn(v, b[i], a[i]); }  // 1 packed : 2 scalar
				if (MODE == 7) { c2[i] = __ffma2_rn(u2, b2[i], c2[i]); c[i] = __fmaf_rn(u, b[i], c[i]); }  // 1 packed : 1 scalar
				// packed with ONE SCALAR operand broadcast to both halves (SASS operand form R.F32)
				if (MODE == 8) c2[i] = __ffma2_rn(a2[i], make_float2(b[i], b[i]), c2[i]);   // pair, scalar, pair: all distinct
				if (MODE == 9) c2[i] = __ffma2_rn(a2[i], make_float2(u, u), c2[i]);         // distinct pair, shared scalar
				if (MODE == 10) c2[i] = __ffma2_rn(u2, make_float2(b[i], b[i]), c2[i]);     // shared pair, distinct scalar
			}
		}
	}
	float r = 0;
	for (int i = 0; i < N; ++i) r += c[i] + c2[i].x + c2[i].y + a[i];
	if (r == 123.456f) out[0] = r;
}

template <int MODE>
void run(const char *name, int per_iter, int w, float *out, float *in, double mhz) {
	int blocks = 148 * w;  // 128 threads = one warp per sub-partition; w blocks per SM
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0), cudaEventCreate(&e1);
	k<MODE><<<blocks, 128>>>(out, in);
	cudaEventRecord(e0);
	k<MODE><<<blocks, 128>>>(out, in);
	cudaEventRecord(e1);
	cudaDeviceSynchronize();
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	double inst = (double)ITERS * UNROLL * per_iter * w;  // per sub-partition
	printf("%-30s W=%d  %.3f cycles / warp-inst / SMSP   (%.1f us)\n", name, w, ms * 1e-3 * mhz * 1e6 / inst, ms * 1e3);
}

int main() {
	float *out, *in;
	cudaMalloc(&out, 4);
	cudaMalloc(&in, 4096);
	float h[1024];
	for (int i = 0; i < 1024; ++i) h[i] = 1.0f + 1e-6f * i;
	cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
	int khz;
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	double mhz = khz / 1e3;
	printf("clock %.0f MHz (nominal max; cycles assume the GPU runs at it)\n", mhz);
	for (int w : {2, 4, 6, 8}) {
		run<0>("FFMA  3 distinct regs", N, w, out, in, mhz);
		run<1>("FFMA  1 shared operand", N, w, out, in, mhz);
		run<2>("FFMA  2 shared operands", N, w, out, in, mhz);
		run<3>("FFMA  immediates", N, w, out, in, mhz);
		run<4>("FFMA2 3 distinct pairs", N, w, out, in, mhz);
		run<5>("FFMA2 1 shared pair", N, w, out, in, mhz);
		run<6>("mix 1 FFMA2 : 2 FFMA", 3 * N, w, out, in, mhz);
		run<7>("mix 1 FFMA2 : 1 FFMA", 2 * N, w, out, in, mhz);
		run<8>("FFMA2 pair, scalar bcast, pair", N, w, out, in, mhz);
		run<9>("FFMA2 pair, SHARED scalar bcast", N, w, out, in, mhz);
		run<10>("FFMA2 SHARED pair, scalar bcast", N, w, out, in, mhz);
	}
	return 0;
}
