// k_bpq.cu -- instantiations and launcher of the four-rows-per-lane kernels (qp_bpq.cuh),
// n = 8, 16, 24, 32, 40.
#include <algorithm>

#include "qp_internal.h"
#include "qp_bpq.cuh"

#ifndef QPB_BPQ_PFI_FROM
#define QPB_BPQ_PFI_FROM 24
#endif

namespace qpb {

bool bpq_eligible(const KernelArgs &a)
{
	const int n = a.n;
	if (n % 8 != 0 || n > 40)
		return false;
	const void *ptrs[] = { a.P, a.q, a.lb, a.ub, a.z0, a.u0, a.x, a.z, a.lam };
	for (const void *p : ptrs)
		if (p && !aligned16(p))
			return false;
	return true;
}

namespace {

template <int NP, bool PFI> int launch_bpq_form(const KernelArgs &a, int dev, cudaStream_t st)
{
	auto kern = qp_bpq_kernel<NP, PFI>;
	static PerDeviceOnce per_sm;
	int b = per_sm.get(dev);
	if (!b) {
		QP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, BPQ_WARPS * 32, 0));
		b = std::max(1, b);
		per_sm.set(dev, b);
	}
	constexpr unsigned long long per_cta = BPQ_WARPS * BpqCfg<NP>::G;
	const unsigned long long need = (a.count + per_cta - 1) / per_cta;
	unsigned long long grid = (unsigned long long)dev_sms(dev) * b;
	grid = std::max(1ull, std::min(grid, need));
	QP_CUDA(launch_kernel(kern, dim3((unsigned)grid), dim3(BPQ_WARPS * 32), 0, st, a));
	return QP_OK;
}

template <int NP> int launch_bpq(const KernelArgs &a, int dev, cudaStream_t st)
{
	if constexpr (NP >= QPB_BPQ_PFI_FROM) {
		if (a.z0 != nullptr)
			return launch_bpq_form<NP, true>(a, dev, st);
	}
	return launch_bpq_form<NP, false>(a, dev, st);
}

} // namespace

int launch_bpq_family(const KernelArgs &a, int dev, cudaStream_t st)
{
	if (a.p_packed)
		return QP_ERR_UNSUPPORTED;
	switch (a.n) {
	case 8:  return launch_bpq<8>(a, dev, st);
	case 16: return launch_bpq<16>(a, dev, st);
	case 24: return launch_bpq<24>(a, dev, st);
	case 32: return launch_bpq<32>(a, dev, st);
	case 40: return launch_bpq<40>(a, dev, st);
	}
	return QP_ERR_UNSUPPORTED;
}

} // namespace qpb
