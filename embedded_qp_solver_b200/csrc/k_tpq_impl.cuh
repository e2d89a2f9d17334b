// k_tpq_impl.cuh -- launcher of the thread-per-QP kernels for ONE dimension N, included by
// k_tpq4.cu, k_tpq8.cu, k_tpq12.cu, k_tpq16.cu (one translation unit per N: the 24 to 48
// instantiations of a dimension compile in parallel with the others).
#pragma once

#include <algorithm>

#include "qp_internal.h"
#include "qp_tpq.cuh"

namespace qpb {
namespace {

template <int N, bool B, bool W, bool R, bool PK, int MODE>
int launch_tpq_m(const KernelArgs &a, int dev, cudaStream_t st)
{
	using L = TpqLayout<N, B, W>;
	using Cfg = TpqCfg<N, B, W>;
	auto kern = qp_tpq_kernel<N, B, W, R, PK, MODE>;
	// CTAs per SM as the hardware will really place them (0 = not asked yet)
	static PerDeviceOnce per_sm;
	int b = per_sm.get(dev);
	if (!b) {
		QP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
					     int(L::BYTES)));
		QP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, Cfg::WARPS * 32, L::BYTES));
		b = std::max(1, std::min(b, int(Cfg::MINB)));
		per_sm.set(dev, b);
	}
	// (a list pass is sized for the whole batch: how long its list is, only the device knows)
	const unsigned long long nbatch = (a.count + 31ull) / 32ull;
	unsigned long long grid = (unsigned long long)dev_sms(dev) * b;
	// a small batch spreads over the SMs (one warp batch per CTA, the other warps leave at once)
	// rather than filling a few CTAs whose warps would share issue slots: it is latency that
	// counts there
	grid = std::max(1ull, std::min(grid, nbatch));
	kern<<<dim3((unsigned)grid), dim3(Cfg::WARPS * 32), L::BYTES, st>>>(a);
	QP_CUDA(cudaGetLastError());
	return QP_OK;
}

// a pass mode is a separate instantiation: the plain kernel is untouched by the others.  A
// list pass takes its (z, u) from scratch, so it has no warm-start variant of its own.
template <int N, bool B, bool W, bool R, bool PK>
int launch_tpq(const KernelArgs &a, int dev, cudaStream_t st, int mode)
{
	switch (mode) {
	case PASS_FIRST: return launch_tpq_m<N, B, W, R, PK, PASS_FIRST>(a, dev, st);
	case PASS_LIST:  return launch_tpq_m<N, B, false, R, PK, PASS_LIST>(a, dev, st);
	}
	return launch_tpq_m<N, B, W, R, PK, PASS_PLAIN>(a, dev, st);
}

template <int N, bool PK> int launch_tpq_v(const KernelArgs &a, int dev, cudaStream_t st, int mode)
{
	const bool b = a.lb != nullptr, w = a.z0 != nullptr;
	if (a.alpha != 1.0f) { // over-relaxation: separate instantiation, not a branch in the loop
		if (b && w)
			return launch_tpq<N, true, true, true, PK>(a, dev, st, mode);
		if (b)
			return launch_tpq<N, true, false, true, PK>(a, dev, st, mode);
		if (w)
			return launch_tpq<N, false, true, true, PK>(a, dev, st, mode);
		return launch_tpq<N, false, false, true, PK>(a, dev, st, mode);
	}
	if (b && w)
		return launch_tpq<N, true, true, false, PK>(a, dev, st, mode);
	if (b)
		return launch_tpq<N, true, false, false, PK>(a, dev, st, mode);
	if (w)
		return launch_tpq<N, false, true, false, PK>(a, dev, st, mode);
	return launch_tpq<N, false, false, false, PK>(a, dev, st, mode);
}

// packed upper triangles (QP_P_UPPER, or the host path's own packing): n = 8, 12, 16
template <int N> int launch_tpq_n(const KernelArgs &a, int dev, cudaStream_t st, int mode)
{
	if constexpr (N >= 8) {
		if (a.p_packed)
			return launch_tpq_v<N, true>(a, dev, st, mode);
	}
	if (a.p_packed)
		return QP_ERR_UNSUPPORTED;
	return launch_tpq_v<N, false>(a, dev, st, mode);
}

} // namespace
} // namespace qpb
