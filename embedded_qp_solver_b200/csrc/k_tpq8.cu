// k_tpq8.cu -- the thread-per-QP kernels (qp_tpq.cuh) for n = 8.
#include "k_tpq_impl.cuh"

namespace qpb {
int launch_tpq8(const KernelArgs &a, int dev, cudaStream_t st, int mode)
{
	return launch_tpq_n<8>(a, dev, st, mode);
}
} // namespace qpb
