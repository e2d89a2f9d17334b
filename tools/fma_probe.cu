// tools/fma_probe.cu -- FP32 issue-rate micro-probe (measurement aid, not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fma_probe tools/fma_probe.cu
// Variants:
//   0  v = fma(v, c0, c1)           constant-bank operands (the bench's FP32 peak probe)
//   1  acc[i] = fma(a[i], x, acc[i]) x shared by a run of FMAs (the triangular-solve pattern)
//   2  acc[i] = fma(a[i], b[i], acc[i]) three distinct registers per FMA
// for several warps/SM, printing warp-instructions per cycle per SM (peak 4.0).
#include <cstdio>
#include <cuda_runtime.h>

template <int V> __global__ void __launch_bounds__(1024) k(float *out, int iters, float c0, float c1)
{
	float acc[16], a[16], b[16];
#pragma unroll
	for (int i = 0; i < 16; ++i) {
		acc[i] = threadIdx.x + i;
		a[i] = 1.0f + 1e-6f * (threadIdx.x + i);
		b[i] = 1.0f - 1e-6f * (threadIdx.x * 3 + i);
	}
	float x = c0;
	for (int it = 0; it < iters; ++it) {
		if (V == 0) {
#pragma unroll
			for (int r = 0; r < 4; ++r)
#pragma unroll
				for (int i = 0; i < 16; ++i)
					acc[i] = fmaf(acc[i], c0, c1);
		} else if (V == 1) {
#pragma unroll
			for (int r = 0; r < 4; ++r) {
#pragma unroll
				for (int i = 0; i < 16; ++i)
					acc[i] = fmaf(a[i], x, acc[i]);
				x = acc[r]; // new shared operand for the next run, like w[k]
			}
		} else {
#pragma unroll
			for (int r = 0; r < 4; ++r)
#pragma unroll
				for (int i = 0; i < 16; ++i)
					acc[i] = fmaf(a[i], b[(i + r) & 15], acc[i]);
		}
	}
	float s = 0;
#pragma unroll
	for (int i = 0; i < 16; ++i)
		s += acc[i];
	if (s == 123.456f)
		out[0] = s;
}

template <int V> void run(const char *name, int sms)
{
	float *d;
	cudaMalloc(&d, 4);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	int clk_khz = 0;
	cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
	const int iters = 20000;
	for (int warps : { 4, 8, 12, 16, 32, 64 }) {
		dim3 block(warps >= 32 ? 1024 : warps * 32), grid(sms * (warps >= 32 ? warps / 32 : 1));
		k<V><<<grid, block>>>(d, 100, 0.999f, 0.001f);
		cudaEventRecord(e0);
		k<V><<<grid, block>>>(d, iters, 0.999f, 0.001f);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms;
		cudaEventElapsedTime(&ms, e0, e1);
		double winst = 64.0 * iters * (double)grid.x * block.x / 32.0;
		double tfl = winst * 32 * 2 / (ms * 1e-3) / 1e12;
		double ipc = winst / sms / (ms * 1e-3 * clk_khz * 1e3);
		printf("%-28s warps/SM %2d  %.1f TFLOP/s  FMA warp-inst/cycle/SM %.2f (at %d MHz nominal)\n", name, warps,
		       tfl, ipc, clk_khz / 1000);
	}
}

int main()
{
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	run<0>("fma(v, const, const)", p.multiProcessorCount);
	run<1>("fma(a[i], x_shared, acc[i])", p.multiProcessorCount);
	run<2>("fma(a[i], b[j], acc[i])", p.multiProcessorCount);
	return 0;
}
