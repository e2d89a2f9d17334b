// k_genr.cu -- the register-resident general-rows kernel (qp_genr.cuh) and its launcher.
#include <algorithm>
#include <cstdlib>

#include "qp_internal.h"
#include "qp_genr.cuh"

namespace qpb {

template <int NP, int MP, int T> static int launch_genr_t(const GenArgs &a, int dev, cudaStream_t st)
{
	using C = GenrCfg<NP, MP, T>;
	static PerDeviceOnce attr_done, ctas_per_sm;
	if (!attr_done.get(dev)) {
		QP_CUDA(cudaFuncSetAttribute(qp_genr_kernel<NP, MP, T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
					     int(C::SMEM)));
		int per_sm = 1;
		QP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qp_genr_kernel<NP, MP, T>,
								      GENR_WARPS * 32, C::SMEM));
		ctas_per_sm.set(dev, std::max(1, per_sm));
		attr_done.set(dev, 1);
	}
	const unsigned long long per_cta = GENR_WARPS * C::G;
	const unsigned long long need = (a.count + per_cta - 1) / per_cta;
	const unsigned long long grid =
		std::max(1ull, std::min((unsigned long long)dev_sms(dev) * ctas_per_sm.get(dev), need));
	qp_genr_kernel<NP, MP, T><<<dim3((unsigned)grid), dim3(GENR_WARPS * 32), C::SMEM, st>>>(a);
	QP_CUDA(cudaGetLastError());
	return QP_OK;
}

int launch_genr(const GenArgs &a, int dev, cudaStream_t st)
{
	// measurements: QPB200_GENR_WIDE=1 doubles the lanes per problem of the two smaller shapes
	static const bool wide = [] { const char *e = std::getenv("QPB200_GENR_WIDE"); return e && e[0] == '1'; }();
	if (a.n <= 8 && a.m <= 16)
		return wide ? launch_genr_t<8, 16, 8>(a, dev, st) : launch_genr_t<8, 16, 4>(a, dev, st);
	if (a.n <= 16 && a.m <= 32)
		return wide ? launch_genr_t<16, 32, 16>(a, dev, st) : launch_genr_t<16, 32, 8>(a, dev, st);
	if (a.n <= 32 && a.m <= 64)
		return launch_genr_t<32, 64, 32>(a, dev, st);
	return QP_ERR_UNSUPPORTED;
}

} // namespace qpb
