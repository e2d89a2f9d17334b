// tools/ffma2_lat_probe.cu -- latency / issue-interval micro-probe for packed FP32 (FFMA2)
// and scalar FFMA on one warp per SM sub-partition (measurement aid, not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma2_lat_probe tools/ffma2_lat_probe.cu
// MODE 0: FFMA2, pair operands        d = a*b+d   (a,b,d float2)
// MODE 1: FFMA2, scalar-broadcast b   d = a*(s,s)+d
// MODE 2: scalar FFMA                 d = a*s+d
// CH = number of independent chains interleaved; reports cycles per instruction.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int CH> __global__ void k(float *out, long long *cyc, int iters, float s0)
{
	float2 d[CH], a[CH];
#pragma unroll
	for (int i = 0; i < CH; ++i) {
		d[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
		a[i] = make_float2(1.0f + 1e-7f * (threadIdx.x + i), 1.0f - 1e-7f * i);
	}
	float s = s0;
	float2 b = make_float2(s0, s0 * 0.5f);
	long long t0 = clock64();
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int r = 0; r < 16; ++r) {
#pragma unroll
			for (int i = 0; i < CH; ++i) {
				if (MODE == 0)
					d[i] = __ffma2_rn(a[i], b, d[i]);
				else if (MODE == 1)
					d[i] = __ffma2_rn(a[i], make_float2(s, s), d[i]);
				else
					d[i].x = fmaf(a[i].x, s, d[i].x);
			}
		}
	}
	long long t1 = clock64();
	float acc = 0;
#pragma unroll
	for (int i = 0; i < CH; ++i)
		acc += d[i].x + d[i].y;
	if (acc == 123.456f)
		out[0] = acc;
	if (threadIdx.x == 0 && blockIdx.x == 0)
		cyc[0] = t1 - t0;
}

template <int MODE, int CH> void run(const char *name, int warps)
{
	float *d;
	long long *c, h;
	cudaMalloc(&d, 4);
	cudaMalloc(&c, 8);
	const int iters = 2000;
	k<MODE, CH><<<1, warps * 32>>>(d, c, iters, 0.999f);
	k<MODE, CH><<<1, warps * 32>>>(d, c, iters, 0.999f);
	cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
	const double per = double(h) / (double(iters) * 16 * CH);
	printf("%-28s chains %2d  warps/SM %2d (%d per sub-partition): %.2f cycles per warp-instruction per warp, %.2f per sub-partition\n",
	       name, CH, warps, warps / 4, per, per / (warps / 4));
	cudaFree(d);
	cudaFree(c);
}

int main()
{
	run<0, 1>("FFMA2 pair operands", 4);
	run<0, 2>("FFMA2 pair operands", 4);
	run<0, 4>("FFMA2 pair operands", 4);
	run<0, 8>("FFMA2 pair operands", 4);
	run<0, 4>("FFMA2 pair operands", 8);
	run<0, 8>("FFMA2 pair operands", 8);
	run<1, 1>("FFMA2 scalar broadcast", 4);
	run<1, 2>("FFMA2 scalar broadcast", 4);
	run<1, 4>("FFMA2 scalar broadcast", 4);
	run<1, 8>("FFMA2 scalar broadcast", 4);
	run<1, 4>("FFMA2 scalar broadcast", 8);
	run<1, 8>("FFMA2 scalar broadcast", 8);
	run<1, 8>("FFMA2 scalar broadcast", 16);
	run<2, 1>("FFMA scalar", 4);
	run<2, 2>("FFMA scalar", 4);
	run<2, 4>("FFMA scalar", 4);
	run<2, 8>("FFMA scalar", 4);
	run<2, 8>("FFMA scalar", 8);
	return 0;
}
