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

// the streaming second pass keeps the full inverse in registers: not for the batches that
// launch_bpq() below solves in product form
bool bpq_stream_eligible(const KernelArgs &a)
{
	// (and it takes lb and ub as a pair of arrays or not at all)
	return bpq_eligible(a) && !(a.n >= QPB_BPQ_PFI_FROM && a.z0 != nullptr) &&
	       ((a.lb == nullptr) == (a.ub == nullptr));
}

namespace {

template <int NP, bool PFI, int MODE> int launch_bpq_m(const KernelArgs &a, int dev, cudaStream_t st)
{
	auto kern = qp_bpq_kernel<NP, PFI, MODE>;
	static PerDeviceOnce per_sm;
	int b = per_sm.get(dev);
	if (!b) {
		QP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, BPQ_WARPS * 32, 0));
		b = std::max(1, b);
		per_sm.set(dev, b);
	}
	// (a list pass is sized for the whole batch: how long its list is, only the device knows)
	// (a small batch spreads over the SMs, one warp's worth per CTA: latency counts there)
	constexpr unsigned long long per_warp = BpqCfg<NP>::G;
	const unsigned long long need = (a.count + per_warp - 1) / per_warp;
	unsigned long long grid = (unsigned long long)dev_sms(dev) * b;
	grid = std::max(1ull, std::min(grid, need));
	kern<<<dim3((unsigned)grid), dim3(BPQ_WARPS * 32), 0, st>>>(a);
	QP_CUDA(cudaGetLastError());
	return QP_OK;
}

// a pass mode is a separate instantiation: the plain kernel is untouched by the others
template <int NP, bool PFI> int launch_bpq_form(const KernelArgs &a, int dev, cudaStream_t st, int mode)
{
	switch (mode) {
	case PASS_FIRST: return launch_bpq_m<NP, PFI, PASS_FIRST>(a, dev, st);
	case PASS_LIST:  return launch_bpq_m<NP, PFI, PASS_LIST>(a, dev, st);
	case PASS_SETUP:
		if constexpr (!PFI)
			return launch_bpq_m<NP, false, PASS_SETUP>(a, dev, st);
		return QP_ERR_UNSUPPORTED;
	}
	return launch_bpq_m<NP, PFI, PASS_PLAIN>(a, dev, st);
}

// the product form is chosen by the caller's warm start; every pass of a solve uses the same
// form (the saved "inverse" of an evicted problem is whatever the first pass built)
template <int NP> int launch_bpq(const KernelArgs &a, int dev, cudaStream_t st, int mode)
{
	if constexpr (NP >= QPB_BPQ_PFI_FROM) {
		if (a.z0 != nullptr)
			return launch_bpq_form<NP, true>(a, dev, st, mode);
	}
	return launch_bpq_form<NP, false>(a, dev, st, mode);
}

} // namespace

int launch_bpq_family(const KernelArgs &a, int dev, cudaStream_t st, int mode)
{
	if (a.p_packed)
		return QP_ERR_UNSUPPORTED;
	switch (a.n) {
	case 8:  return launch_bpq<8>(a, dev, st, mode);
	case 16: return launch_bpq<16>(a, dev, st, mode);
	case 24: return launch_bpq<24>(a, dev, st, mode);
	case 32: return launch_bpq<32>(a, dev, st, mode);
	case 40: return launch_bpq<40>(a, dev, st, mode);
	}
	return QP_ERR_UNSUPPORTED;
}

} // namespace qpb
