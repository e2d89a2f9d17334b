// k_tpq.cu -- dispatcher of the thread-per-QP kernels (qp_tpq.cuh), n = 4, 8, 12, 16; the
// kernels themselves are instantiated in k_tpq4.cu ... k_tpq16.cu.
#include "qp_internal.h"

namespace qpb {

int launch_tpq4(const KernelArgs &a, int dev, cudaStream_t st, int mode);
int launch_tpq8(const KernelArgs &a, int dev, cudaStream_t st, int mode);
int launch_tpq12(const KernelArgs &a, int dev, cudaStream_t st, int mode);
int launch_tpq16(const KernelArgs &a, int dev, cudaStream_t st, int mode);

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

int launch_tpq_family(const KernelArgs &a, int dev, cudaStream_t st, int mode)
{
	switch (a.n) {
	case 4:  return launch_tpq4(a, dev, st, mode);
	case 8:  return launch_tpq8(a, dev, st, mode);
	case 12: return launch_tpq12(a, dev, st, mode);
	case 16: return launch_tpq16(a, dev, st, mode);
	}
	return QP_ERR_UNSUPPORTED;
}

} // namespace qpb
