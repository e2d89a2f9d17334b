// k_tpq12.cu -- the thread-per-QP kernels (qp_tpq.cuh) for n = 12.
#include "k_tpq_impl.cuh"

namespace qpb {
int launch_tpq12(const KernelArgs &a, int dev, cudaStream_t st, int mode)
{
	return launch_tpq_n<12>(a, dev, st, mode);
}
} // namespace qpb
