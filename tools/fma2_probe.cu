// tools/fma2_probe.cu -- packed FP32 (FFMA2) issue-rate micro-probe (measurement aid).
#include <cstdio>
#include <cuda_runtime.h>

template <int V> __global__ void __launch_bounds__(1024) k(float *out, int iters, float c0, float c1)
{
	float2 acc[8], a[8];
#pragma unroll
	for (int i = 0; i < 8; ++i) {
		acc[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
		a[i] = make_float2(1.0f + 1e-6f * (threadIdx.x + i), 1.0f - 1e-6f * (threadIdx.x + i));
	}
	float2 x = make_float2(c0, c0);
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int r = 0; r < 8; ++r) {
			if (V == 0) {           // FFMA2, shared operand pair
#pragma unroll
				for (int i = 0; i < 8; ++i)
					acc[i] = __ffma2_rn(a[i], x, acc[i]);
				x = make_float2(acc[r].x, acc[r].x);
			} else if (V == 1) {    // scalar FFMA, same math
#pragma unroll
				for (int i = 0; i < 8; ++i) {
					acc[i].x = fmaf(a[i].x, x.x, acc[i].x);
					acc[i].y = fmaf(a[i].y, x.x, acc[i].y);
				}
				x.x = acc[r].x;
			} else if (V == 2) {    // FFMA2 + interleaved FMNMX (ALU pipe) 1:1
#pragma unroll
				for (int i = 0; i < 8; ++i) {
					acc[i] = __ffma2_rn(a[i], x, acc[i]);
					a[i].x = fminf(a[i].x, acc[i].y);
				}
				x = make_float2(acc[r].x, acc[r].x);
			} else {                // FADD2
#pragma unroll
				for (int i = 0; i < 8; ++i)
					acc[i] = __fadd2_rn(a[i], acc[i]);
			}
		}
	}
	float s = 0;
#pragma unroll
	for (int i = 0; i < 8; ++i)
		s += acc[i].x + acc[i].y + a[i].x;
	if (s == 123.456f)
		out[0] = s;
}

template <int V> void run(const char *name, int sms, double flops_per_iter_thread)
{
	float *d;
	cudaMalloc(&d, 4);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	const int iters = 20000;
	for (int warps : { 4, 8, 16, 64 }) {
		dim3 block(warps >= 32 ? 1024 : warps * 32), grid(sms * (warps >= 32 ? warps / 32 : 1));
		k<V><<<grid, block>>>(d, 100, 0.999f, 0.001f);
		cudaEventRecord(e0);
		k<V><<<grid, block>>>(d, iters, 0.999f, 0.001f);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms;
		cudaEventElapsedTime(&ms, e0, e1);
		double tfl = flops_per_iter_thread * iters * (double)grid.x * block.x / (ms * 1e-3) / 1e12;
		printf("%-34s warps/SM %2d  %.1f TFLOP/s (%.3f ms)\n", name, warps, tfl, ms);
	}
}

int main()
{
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	run<0>("FFMA2 shared-pair", p.multiProcessorCount, 8 * 8 * 4.0);
	run<1>("FFMA scalar same math", p.multiProcessorCount, 8 * 8 * 4.0);
	run<2>("FFMA2 + FMNMX 1:1 (flops of FFMA2)", p.multiProcessorCount, 8 * 8 * 4.0);
	run<3>("FADD2", p.multiProcessorCount, 8 * 8 * 2.0);
	return 0;
}
