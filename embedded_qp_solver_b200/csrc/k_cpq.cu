// k_cpq.cu -- instantiations and launcher of the CTA-per-QP kernels (qp_cpq.cuh), n <= 128.
#include <algorithm>

#include "qp_internal.h"
#include "qp_cpq.cuh"

namespace qpb {
namespace {

template <int TT> int launch_cpq(const KernelArgs &a, int dev, cudaStream_t st)
{
	const unsigned long long grid = std::min<unsigned long long>(a.count, 2ull * dev_sms(dev));
	qp_cpq_kernel<TT><<<dim3((unsigned)grid), dim3(CPQ_THREADS), 0, st>>>(a);
	QP_CUDA(cudaGetLastError());
	return QP_OK;
}

} // namespace

int launch_cpq_family(const KernelArgs &a, int dev, cudaStream_t st)
{
	if (a.p_packed)
		return QP_ERR_UNSUPPORTED;
	switch ((a.n + 15) / 16) {
	case 1: return launch_cpq<1>(a, dev, st);
	case 2: return launch_cpq<2>(a, dev, st);
	case 3: return launch_cpq<3>(a, dev, st);
	case 4: return launch_cpq<4>(a, dev, st);
	case 5: return launch_cpq<5>(a, dev, st);
	case 6: return launch_cpq<6>(a, dev, st);
	case 7: return launch_cpq<7>(a, dev, st);
	case 8: return launch_cpq<8>(a, dev, st);
	}
	return QP_ERR_UNSUPPORTED;
}

} // namespace qpb
