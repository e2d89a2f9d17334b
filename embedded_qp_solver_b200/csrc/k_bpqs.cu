// k_bpqs.cu -- instantiations and launcher of the streaming second pass of the
// four-rows-per-lane kernels (qp_bpq_stream.cuh), n = 8, 16, 24, 32, 40.
#include <algorithm>

#include "qp_internal.h"
#include "qp_bpq_stream.cuh"

namespace qpb {

namespace {

template <int NP, bool BOUNDS, bool FRESH> int launch_bpq_stream_b(const KernelArgs &a, int dev, cudaStream_t st)
{
	using SC = BpqStreamCfg<NP>;
	auto kern = qp_bpq_stream_kernel<NP, BOUNDS, FRESH>;
	static PerDeviceOnce per_sm;
	int b = per_sm.get(dev);
	if (!b) {
		QP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SC::BYTES)));
		QP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, BPQ_WARPS * 32, SC::BYTES));
		b = std::max(1, b);
		per_sm.set(dev, b);
	}
	// sized for the whole batch: how long the list is, only the device knows (a group that finds
	// no entry for itself leaves at once)
	constexpr unsigned long long per_warp = BpqCfg<NP>::G;
	const unsigned long long need = (a.count + per_warp - 1) / per_warp;
	unsigned long long grid = (unsigned long long)dev_sms(dev) * b;
	grid = std::max(1ull, std::min(grid, need));
	kern<<<dim3((unsigned)grid), dim3(BPQ_WARPS * 32), SC::BYTES, st>>>(a);
	QP_CUDA(cudaGetLastError());
	return QP_OK;
}

// (bpq_stream_eligible: lb and ub are both arrays or both absent)
template <int NP> int launch_bpq_stream(const KernelArgs &a, int dev, cudaStream_t st, bool fresh)
{
	const bool b = a.lb && a.ub;
	if (fresh)
		return b ? launch_bpq_stream_b<NP, true, true>(a, dev, st) : launch_bpq_stream_b<NP, false, true>(a, dev, st);
	return b ? launch_bpq_stream_b<NP, true, false>(a, dev, st) : launch_bpq_stream_b<NP, false, false>(a, dev, st);
}

} // namespace

int launch_bpq_stream_family(const KernelArgs &a, int dev, cudaStream_t st, bool fresh)
{
	switch (a.n) {
	case 8:  return launch_bpq_stream<8>(a, dev, st, fresh);
	case 16: return launch_bpq_stream<16>(a, dev, st, fresh);
	case 24: return launch_bpq_stream<24>(a, dev, st, fresh);
	case 32: return launch_bpq_stream<32>(a, dev, st, fresh);
	case 40: return launch_bpq_stream<40>(a, dev, st, fresh);
	}
	return QP_ERR_UNSUPPORTED;
}

} // namespace qpb
