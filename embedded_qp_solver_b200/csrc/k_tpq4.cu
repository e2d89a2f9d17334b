// k_tpq4.cu -- the thread-per-QP kernels (qp_tpq.cuh) for n = 4.
#include "k_tpq_impl.cuh"

namespace qpb {
int launch_tpq4(const KernelArgs &a, int dev, cudaStream_t st, int mode)
{
	return launch_tpq_n<4>(a, dev, st, mode);
}
} // namespace qpb
