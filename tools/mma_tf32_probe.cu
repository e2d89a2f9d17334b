// tools/mma_tf32_probe.cu -- issue interval of the legacy warp-level TF32 MMA
// (mma.sync.m16n8k8 / m16n8k4, SASS HMMA.1688.F32.TF32 ...) on sm_100a (measurement aid).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_tf32_probe tools/mma_tf32_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int K8, int CH> __global__ void k(float *out, long long *cyc, int iters)
{
	unsigned a[4], b[2];
	float d[CH][4];
	for (int i = 0; i < 4; ++i)
		a[i] = __float_as_uint(1.0f + 1e-3f * (threadIdx.x + i));
	for (int i = 0; i < 2; ++i)
		b[i] = __float_as_uint(1.0f - 1e-3f * (threadIdx.x + i));
#pragma unroll
	for (int c = 0; c < CH; ++c)
		for (int i = 0; i < 4; ++i)
			d[c][i] = float(c + i);
	long long t0 = clock64();
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int r = 0; r < 8; ++r) {
#pragma unroll
			for (int c = 0; c < CH; ++c) {
				if (K8)
					asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 "
						     "{%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
						     : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
						     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
				else
					asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 "
						     "{%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
						     : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
						     : "r"(a[0]), "r"(a[1]), "r"(b[0]));
			}
		}
	}
	long long t1 = clock64();
	float acc = 0;
#pragma unroll
	for (int c = 0; c < CH; ++c)
		acc += d[c][0] + d[c][1] + d[c][2] + d[c][3];
	if (acc == 123.456f)
		out[0] = acc;
	if (threadIdx.x == 0 && blockIdx.x == 0)
		cyc[0] = t1 - t0;
}

template <int K8, int CH> void run(const char *name, int warps)
{
	float *d;
	long long *c, h;
	cudaMalloc(&d, 4);
	cudaMalloc(&c, 8);
	const int iters = 2000;
	k<K8, CH><<<1, warps * 32>>>(d, c, iters);
	k<K8, CH><<<1, warps * 32>>>(d, c, iters);
	cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
	const double per = double(h) / (double(iters) * 8 * CH);
	const double flops = K8 ? 2.0 * 16 * 8 * 8 : 2.0 * 16 * 8 * 4;
	printf("%-10s chains %d warps/SM %2d: %.2f cycles per MMA per warp, %.2f per sub-partition, %.0f flop/cycle/sub-partition (FFMA: 64)\n",
	       name, CH, warps, per, per / (warps / 4.0), flops / (per / (warps / 4.0)));
	cudaFree(d);
	cudaFree(c);
}

int main()
{
	run<1, 1>("m16n8k8", 4);
	run<1, 4>("m16n8k8", 4);
	run<1, 8>("m16n8k8", 4);
	run<1, 8>("m16n8k8", 8);
	run<1, 8>("m16n8k8", 16);
	run<0, 1>("m16n8k4", 4);
	run<0, 8>("m16n8k4", 4);
	run<0, 8>("m16n8k4", 8);
	return 0;
}
