// k_f64.cu -- launcher of the binary64 box-ADMM kernel (qp_f64.cuh).
#include <algorithm>

#include "qp_internal.h"
#include "qp_f64.cuh"

namespace qpb {

int launch_admm64(const Admm64Args &a, int dev, cudaStream_t st)
{
	if (a.n < 1 || a.n > QP_MAX_DIM)
		return QP_ERR_UNSUPPORTED;
	const size_t smem = F64_WARPS * f64_smem_doubles_per_warp(a.n) * sizeof(double);
	static PerDeviceOnce attr_done;
	if (!attr_done.get(dev)) {
		QP_CUDA(cudaFuncSetAttribute(qp_admm64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
					     int(F64_WARPS * f64_smem_doubles_per_warp(QP_MAX_DIM) * sizeof(double))));
		attr_done.set(dev, 1);
	}
	int per_sm = 1;
	QP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qp_admm64_kernel, F64_WARPS * 32, smem));
	per_sm = std::max(1, per_sm);
	const unsigned long long need = (a.count + F64_WARPS - 1) / F64_WARPS;
	const unsigned long long grid = std::max(1ull, std::min((unsigned long long)dev_sms(dev) * per_sm, need));
	qp_admm64_kernel<<<dim3((unsigned)grid), dim3(F64_WARPS * 32), smem, st>>>(a);
	QP_CUDA(cudaGetLastError());
	return QP_OK;
}

} // namespace qpb
