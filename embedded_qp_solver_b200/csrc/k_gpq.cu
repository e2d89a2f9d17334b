// k_gpq.cu -- instantiations and launcher of the lane-group kernels with cyclic rows
// (qp_gpq.cuh), any n <= 64.
#include <algorithm>

#include "qp_internal.h"
#include "qp_gpq.cuh"

#ifndef QPB_GPQ_PFI_FROM
#define QPB_GPQ_PFI_FROM 48
#endif

namespace qpb {
namespace {

template <int NP, bool PFI> int launch_gpq_form(const KernelArgs &a, int dev, cudaStream_t st)
{
	constexpr int T = GpqLanes<NP>::T; // lanes per QP; 32 / T problems per warp
	auto kern = qp_gpq_kernel<NP, T, PFI>;
	static PerDeviceOnce per_sm;
	int b = per_sm.get(dev);
	if (!b) {
		QP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, GPQ_WARPS * 32, 0));
		b = std::max(1, b);
		per_sm.set(dev, b);
	}
	constexpr unsigned long long per_cta = GPQ_WARPS * (32 / T);
	const unsigned long long need = (a.count + per_cta - 1) / per_cta;
	unsigned long long grid = (unsigned long long)dev_sms(dev) * b;
	grid = std::max(1ull, std::min(grid, need));
	kern<<<dim3((unsigned)grid), dim3(GPQ_WARPS * 32), 0, st>>>(a);
	QP_CUDA(cudaGetLastError());
	return QP_OK;
}

template <int NP> int launch_gpq(const KernelArgs &a, int dev, cudaStream_t st)
{
	// n >= 48: the elimination dominates up to ~20-30 iterations: always the product form
	if constexpr (NP >= QPB_GPQ_PFI_FROM)
		return launch_gpq_form<NP, true>(a, dev, st);
	else
		return launch_gpq_form<NP, false>(a, dev, st);
}

} // namespace

int launch_gpq_family(const KernelArgs &a, int dev, cudaStream_t st)
{
	if (a.p_packed)
		return QP_ERR_UNSUPPORTED;
	switch ((a.n + 7) / 8 * 8) {
	case 8:  return launch_gpq<8>(a, dev, st);
	case 16: return launch_gpq<16>(a, dev, st);
	case 24: return launch_gpq<24>(a, dev, st);
	case 32: return launch_gpq<32>(a, dev, st);
	case 40: return launch_gpq<40>(a, dev, st);
	case 48: return launch_gpq<48>(a, dev, st);
	case 56: return launch_gpq<56>(a, dev, st);
	case 64: return launch_gpq<64>(a, dev, st);
	}
	return QP_ERR_UNSUPPORTED;
}

} // namespace qpb
