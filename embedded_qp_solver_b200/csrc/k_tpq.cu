// k_tpq.cu -- instantiations and launcher of the thread-per-QP kernels (qp_tpq.cuh),
// n = 4, 8, 12, 16.  One translation unit per kernel family so that the library builds in
// parallel (embedded_qp_solver_b200/build.py).
#include <algorithm>

#include "qp_internal.h"
#include "qp_tpq.cuh"

namespace qpb {

bool tpq_eligible(const KernelArgs &a)
{
	const int n = a.n;
	if (!(n == 4 || n == 8 || n == 12 || n == 16))
		return false;
	if ((a.lb == nullptr) != (a.ub == nullptr))
		return false;
	if ((a.z0 == nullptr) != (a.u0 == nullptr))
		return false;
	const void *ptrs[] = { a.P, a.q, a.lb, a.ub, a.z0, a.u0, a.x, a.z, a.lam };
	for (const void *p : ptrs)
		if (p && !aligned16(p))
			return false;
	return true;
}

namespace {

template <int N, bool B, bool W, bool R, bool PK>
int launch_tpq(const KernelArgs &a, int dev, cudaStream_t st)
{
	using L = TpqLayout<N, B, W>;
	using Cfg = TpqCfg<N, B, W>;
	auto kern = qp_tpq_kernel<N, B, W, R, PK>;
	// CTAs per SM as the hardware will really place them (0 = not asked yet): the level
	// scheduler's grid barrier needs every CTA resident
	static PerDeviceOnce per_sm;
	int b = per_sm.get(dev);
	if (!b) {
		QP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
					     int(L::BYTES)));
		QP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, Cfg::WARPS * 32, L::BYTES));
		b = std::max(1, std::min(b, int(Cfg::MINB)));
		per_sm.set(dev, b);
	}
	const unsigned long long nbatch = (a.count + 31ull) / 32ull;
	unsigned long long grid = (unsigned long long)dev_sms(dev) * b;
	// every warp of a CTA should have a batch to work on
	const unsigned long long need = (nbatch + Cfg::WARPS - 1) / Cfg::WARPS;
	grid = std::max(1ull, std::min(grid, need));
	QP_CUDA(launch_kernel(kern, dim3((unsigned)grid), dim3(Cfg::WARPS * 32), L::BYTES, st, a));
	return QP_OK;
}

template <int N, bool PK> int launch_tpq_v(const KernelArgs &a, int dev, cudaStream_t st)
{
	const bool b = a.lb != nullptr, w = a.z0 != nullptr;
	if (a.alpha != 1.0f) { // over-relaxation: separate instantiation, not a branch in the loop
		if (b && w)
			return launch_tpq<N, true, true, true, PK>(a, dev, st);
		if (b)
			return launch_tpq<N, true, false, true, PK>(a, dev, st);
		if (w)
			return launch_tpq<N, false, true, true, PK>(a, dev, st);
		return launch_tpq<N, false, false, true, PK>(a, dev, st);
	}
	if (b && w)
		return launch_tpq<N, true, true, false, PK>(a, dev, st);
	if (b)
		return launch_tpq<N, true, false, false, PK>(a, dev, st);
	if (w)
		return launch_tpq<N, false, true, false, PK>(a, dev, st);
	return launch_tpq<N, false, false, false, PK>(a, dev, st);
}

// packed upper triangles (QP_P_UPPER, or the host path's own packing): n = 8, 12, 16
template <int N> int launch_tpq_n(const KernelArgs &a, int dev, cudaStream_t st)
{
	if constexpr (N >= 8) {
		if (a.p_packed)
			return launch_tpq_v<N, true>(a, dev, st);
	}
	if (a.p_packed)
		return QP_ERR_UNSUPPORTED;
	return launch_tpq_v<N, false>(a, dev, st);
}

} // namespace

int launch_tpq_family(const KernelArgs &a, int dev, cudaStream_t st)
{
	switch (a.n) {
	case 4:  return launch_tpq_n<4>(a, dev, st);
	case 8:  return launch_tpq_n<8>(a, dev, st);
	case 12: return launch_tpq_n<12>(a, dev, st);
	case 16: return launch_tpq_n<16>(a, dev, st);
	}
	return QP_ERR_UNSUPPORTED;
}

} // namespace qpb
