// k_aux.cu -- the small kernels of libqpb200.so and their launchers: objective evaluation
// (quadratic_form_eval, /root/reference/qp_solvers/qp.c:9-27), the Newton / gradient-descent
// kernel (qp_newton.cuh), the general-rows kernel (qp_gen.cuh) and the roofline probes.
#include <algorithm>

#include "qp_internal.h"
#include "qp_gen.cuh"
#include "qp_newton.cuh"

namespace qpb {

__global__ void qp_eval_kernel(const float *__restrict__ P, const float *__restrict__ q,
			       const float *__restrict__ x, float r, float *__restrict__ f, int n,
			       unsigned long long count)
{
	const int lane = threadIdx.x & 31;
	const unsigned long long w = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
	const unsigned long long nw = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
	for (unsigned long long c = w; c < count; c += nw) {
		const float *Pc = P + c * (unsigned long long)n * n;
		const float *xc = x + c * (unsigned long long)n;
		const float *qc = q + c * (unsigned long long)n;
		float quad = 0.0f, lin = 0.0f;
		// lane j accumulates column j: sum_i x_i P_ij, then times x_j (P symmetric or not)
		for (int j = lane; j < n; j += 32) {
			float t = 0.0f;
			for (int i = 0; i < n; ++i)
				t = fmaf(xc[i], Pc[(unsigned long long)i * n + j], t);
			quad = fmaf(t, xc[j], quad);
			lin = fmaf(qc[j], xc[j], lin);
		}
		quad = warp_sum(quad);
		lin = warp_sum(lin);
		if (lane == 0)
			f[c] = fmaf(0.5f, quad, lin) + r;
	}
}

int launch_eval(const float *P, const float *q, const float *x, float r, float *f, int n,
		unsigned long long count, int dev, cudaStream_t st)
{
	const unsigned grid = (unsigned)std::min<size_t>((count + 7) / 8, size_t(dev_sms(dev)) * 8);
	qp_eval_kernel<<<grid, 256, 0, st>>>(P, q, x, r, f, n, count);
	QP_CUDA(cudaGetLastError());
	return QP_OK;
}

int launch_newton(const NewtonArgs &a, int dev, cudaStream_t st)
{
	const size_t smem = NEWTON_WARPS * newton_smem_floats_per_warp(a.n) * sizeof(float);
	static std::atomic<size_t> attr_done[MAX_DEVICES];
	if (attr_done[dev] < smem) {
		QP_CUDA(cudaFuncSetAttribute(qp_newton_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
					     int(NEWTON_WARPS * newton_smem_floats_per_warp(QP_MAX_DIM) *
						 sizeof(float))));
		attr_done[dev] = NEWTON_WARPS * newton_smem_floats_per_warp(QP_MAX_DIM) * sizeof(float);
	}
	int per_sm = 1;
	QP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qp_newton_kernel,
							      NEWTON_WARPS * 32, smem));
	per_sm = std::max(1, per_sm);
	const unsigned long long need = (a.count + NEWTON_WARPS - 1) / NEWTON_WARPS;
	const unsigned long long grid =
		std::max(1ull, std::min((unsigned long long)dev_sms(dev) * per_sm, need));
	qp_newton_kernel<<<dim3((unsigned)grid), dim3(NEWTON_WARPS * 32), smem, st>>>(a);
	QP_CUDA(cudaGetLastError());
	return QP_OK;
}

template <int T> static int launch_gen_t(const GenArgs &a, int dev, cudaStream_t st)
{
	constexpr int G = 32 / T;
	const size_t smem = GEN_WARPS * G * gen_smem_floats_per_group(a.n, a.m, T) * sizeof(float);
	// the largest problem this group width takes: n = 2T, m = 4T
	const size_t smem_max = GEN_WARPS * G * gen_smem_floats_per_group(2 * T, 4 * T, T) * sizeof(float);
	static PerDeviceOnce attr_done;
	if (!attr_done.get(dev)) {
		QP_CUDA(cudaFuncSetAttribute(qp_gen_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
					     int(smem_max)));
		attr_done.set(dev, 1);
	}
	int per_sm = 1;
	QP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qp_gen_kernel<T>, GEN_WARPS * 32, smem));
	per_sm = std::max(1, per_sm);
	const unsigned long long per_cta = GEN_WARPS * G;
	const unsigned long long need = (a.count + per_cta - 1) / per_cta;
	const unsigned long long grid =
		std::max(1ull, std::min((unsigned long long)dev_sms(dev) * per_sm, need));
	unsigned long long *ticket = nullptr;
	if (need > grid) {
		const int rc = gen_ticket(dev, st, &ticket);
		if (rc != QP_OK)
			return rc;
	}
	qp_gen_kernel<T><<<dim3((unsigned)grid), dim3(GEN_WARPS * 32), smem, st>>>(a, ticket);
	QP_CUDA(cudaGetLastError());
	return QP_OK;
}

int launch_gen(const GenArgs &a, int dev, cudaStream_t st)
{
	if (a.force != 1) {
		const int rc = launch_genr(a, dev, st);
		if (rc != QP_ERR_UNSUPPORTED || a.force == 2)
			return rc;
	}
	switch (gen_lanes(a.n, a.m)) {
	case 4:  return launch_gen_t<4>(a, dev, st);
	case 8:  return launch_gen_t<8>(a, dev, st);
	case 16: return launch_gen_t<16>(a, dev, st);
	}
	return launch_gen_t<32>(a, dev, st);
}

} // namespace qpb

using namespace qpb;

// ----------------------------------------------------------------------------------------
// roofline probes
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) qp_probe_fma_kernel(float *out, int iters, float a, float b)
{
	float v[8];
#pragma unroll
	for (int k = 0; k < 8; ++k)
		v[k] = float(threadIdx.x + k);
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int r = 0; r < 8; ++r) {
#pragma unroll
			for (int k = 0; k < 8; ++k)
				v[k] = fmaf(v[k], a, b);
		}
	}
	float s = 0.0f;
#pragma unroll
	for (int k = 0; k < 8; ++k)
		s += v[k];
	if (s == 12345.678f)
		out[0] = s;
}

extern "C" double qp_probe_fp32_tflops(int device, int repeats)
{
	const int dev = resolve_device(device);
	if (dev < 0)
		return 0.0;
	DeviceGuard g;
	if (g.enter(dev) != QP_OK)
		return 0.0;
	float *d = nullptr;
	if (cudaMalloc(&d, 256) != cudaSuccess)
		return 0.0;
	const int sms = dev_sms(dev);
	const int iters = 4096;
	const dim3 grid(sms * 8), block(256);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	double best = 0.0;
	for (int r = 0; r < std::max(1, repeats) + 1; ++r) {
		cudaEventRecord(e0);
		qp_probe_fma_kernel<<<grid, block>>>(d, iters, 0.999f, 0.001f);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e1);
		const double flops = 2.0 * 64.0 * iters * double(grid.x) * block.x;
		if (r > 0 && ms > 0)
			best = std::max(best, flops / (ms * 1e-3) / 1e12);
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(d);
	return best;
}

__global__ void qp_probe_copy_kernel(const float4 *__restrict__ src, float4 *__restrict__ dst,
				     size_t n4)
{
	const size_t stride = size_t(gridDim.x) * blockDim.x;
	for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n4; i += stride)
		dst[i] = src[i];
}

extern "C" double qp_probe_hbm_gbs(int device, size_t bytes, int repeats)
{
	const int dev = resolve_device(device);
	if (dev < 0)
		return 0.0;
	DeviceGuard g;
	if (g.enter(dev) != QP_OK)
		return 0.0;
	bytes = std::max<size_t>(bytes, 1 << 20) / 16 * 16;
	float4 *a = nullptr, *b = nullptr;
	if (cudaMalloc(&a, bytes) != cudaSuccess)
		return 0.0;
	if (cudaMalloc(&b, bytes) != cudaSuccess) {
		cudaFree(a);
		return 0.0;
	}
	cudaMemset(a, 1, bytes);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	double best = 0.0;
	const int sms = dev_sms(dev);
	for (int r = 0; r < std::max(1, repeats) + 1; ++r) {
		cudaEventRecord(e0);
		qp_probe_copy_kernel<<<sms * 16, 512>>>(a, b, bytes / 16);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e1);
		if (r > 0 && ms > 0)
			best = std::max(best, 2.0 * double(bytes) / (ms * 1e-3) / 1e9);
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(a);
	cudaFree(b);
	return best;
}