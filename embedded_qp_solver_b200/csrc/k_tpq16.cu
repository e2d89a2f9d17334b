// k_tpq16.cu -- the thread-per-QP kernels (qp_tpq.cuh) for n = 16.
#include "k_tpq_impl.cuh"

namespace qpb {
int launch_tpq16(const KernelArgs &a, int dev, cudaStream_t st, int mode)
{
	return launch_tpq_n<16>(a, dev, st, mode);
}
} // namespace qpb
