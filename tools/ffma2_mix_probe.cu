// tools/ffma2_mix_probe.cu -- does FFMA2 leave issue slots free for other pipes?
// (measurement aid, not product code)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma2_mix_probe tools/ffma2_mix_probe.cu
// Each variant runs 8 FMA-pipe instructions (FFMA2 or FFMA) per group, optionally
// interleaved 1:1 with 8 independent instructions of another pipe, 16 warps/SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE> __global__ void __launch_bounds__(512) k(float *out, int iters, float s0, int i0)
{
	float2 d[8], a[8];
	int v[8];
	float f[8];
	extern __shared__ float sm[];
#pragma unroll
	for (int i = 0; i < 8; ++i) {
		d[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
		a[i] = make_float2(1.0f + 1e-7f * (threadIdx.x + i), 1.0f - 1e-7f * i);
		v[i] = threadIdx.x * 3 + i;
		f[i] = threadIdx.x * 0.5f + i;
	}
	sm[threadIdx.x] = s0;
	__syncthreads();
	const float s = s0;
	float lsum = 0.f;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int r = 0; r < 8; ++r) {
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				// FMA-pipe part
				if (MODE <= 4)
					d[i] = __ffma2_rn(a[i], make_float2(s, s), d[i]);
				else {
					d[i].x = fmaf(a[i].x, s, d[i].x);
					if (MODE == 6 || MODE == 7)
						d[i].y = fmaf(a[i].y, s, d[i].y);
				}
				// other-pipe part (independent of the FMA results)
				if (MODE == 1 || MODE == 7)
					v[i] = (v[i] ^ i0) + r;          // ALU: LOP3 / IADD3
				if (MODE == 2 || MODE == 5)
					f[i] = fminf(f[i], s + r);         // ALU: FMNMX (+ maybe FADD hoisted)
				if (MODE == 3)
					lsum += sm[(threadIdx.x + i * 32 + r) & 511]; // LSU + FADD
			}
		}
	}
	float acc = lsum;
#pragma unroll
	for (int i = 0; i < 8; ++i)
		acc += d[i].x + d[i].y + float(v[i]) + f[i];
	if (acc == 123.456f)
		out[0] = acc;
}

template <int MODE> void run(const char *name, int sms, double fma_inst_per_group)
{
	float *d;
	cudaMalloc(&d, 4);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	const int iters = 4000;
	int clk_khz = 0;
	cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
	k<MODE><<<sms, 512, 2048>>>(d, 10, 0.999f, 5);
	cudaEventRecord(e0);
	k<MODE><<<sms, 512, 2048>>>(d, iters, 0.999f, 5);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	// cycles per group of (8 FMA-pipe instr [+ 8 others]) per sub-partition (4 warps each)
	const double cycles = ms * 1e-3 * clk_khz * 1e3;
	const double groups_per_smsp = double(iters) * 8 * 4; // 4 warps per sub-partition
	printf("%-44s %.2f cycles per 8-instruction FMA group per sub-partition (%.2f per FMA-pipe instr)\n",
	       name, cycles / groups_per_smsp, cycles / groups_per_smsp / fma_inst_per_group);
	cudaFree(d);
}

int main()
{
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	const int sms = p.multiProcessorCount;
	run<0>("8 FFMA2", sms, 8);
	run<1>("8 FFMA2 + 8 x (LOP3, IADD3)", sms, 8);
	run<2>("8 FFMA2 + 8 FMNMX", sms, 8);
	run<3>("8 FFMA2 + 8 x (LDS, FADD)", sms, 8);
	run<5>("8 FFMA  + 8 FMNMX", sms, 8);
	run<6>("16 FFMA", sms, 16);
	run<7>("16 FFMA + 8 x (LOP3, IADD3)", sms, 16);
	return 0;
}
