// k_genr.cu -- the register-resident general-rows kernel (qp_genr.cuh) and its launcher.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <mutex>

#include "qp_internal.h"
#include "qp_genr.cuh"

namespace qpb {

constexpr unsigned GENR_TICKETS = 256;
static std::atomic<unsigned long long *> g_ticket_pool[MAX_DEVICES];
static std::atomic<unsigned> g_ticket_next[MAX_DEVICES];
static std::mutex g_ticket_mu;

// the device's ticket counters, allocated on first use (never freed: 2 KB per device); null
// when the allocation fails -- the launch then falls back to the static stride
static unsigned long long *ticket_pool(int dev)
{
	unsigned long long *p = g_ticket_pool[dev].load(std::memory_order_acquire);
	if (p)
		return p;
	std::lock_guard<std::mutex> lock(g_ticket_mu);
	p = g_ticket_pool[dev].load(std::memory_order_acquire);
	if (!p) {
		if (cudaMalloc(reinterpret_cast<void **>(&p), GENR_TICKETS * sizeof(unsigned long long)) != cudaSuccess) {
			cudaGetLastError();
			return nullptr;
		}
		g_ticket_pool[dev].store(p, std::memory_order_release);
	}
	return p;
}

// A zeroed ticket counter for a launch on `st` that is longer than one wave (qp_genr.cuh,
// qp_gen.cuh: the first round is dealt by position, the kernel adds its size to the tickets).
// The counter is one of GENR_TICKETS slots of a per-device array, taken in rotation and zeroed
// on the launch's stream: a slot is reused after GENR_TICKETS further launches on this device,
// by which time its launch has long ended unless that many general-rows solves are in flight
// at once on different streams.  *out stays null (static stride) with QPB200_GENR_TICKET=0 or
// when the array cannot be allocated.
int gen_ticket(int dev, cudaStream_t st, unsigned long long **out)
{
	static const bool no_ticket = [] { const char *e = std::getenv("QPB200_GENR_TICKET"); return e && e[0] == '0'; }();
	*out = nullptr;
	if (no_ticket)
		return QP_OK;
	unsigned long long *pool = ticket_pool(dev);
	if (!pool)
		return QP_OK;
	unsigned long long *t = pool + g_ticket_next[dev].fetch_add(1u, std::memory_order_relaxed) % GENR_TICKETS;
	QP_CUDA(cudaMemsetAsync(t, 0, sizeof(unsigned long long), st));
	*out = t;
	return QP_OK;
}

template <int NP, int MP, int T, bool RELAX> static int launch_genr_r(const GenArgs &a, int dev, cudaStream_t st)
{
	using C = GenrCfg<NP, MP, T>;
	static PerDeviceOnce attr_done, ctas_per_sm;
	if (!attr_done.get(dev)) {
		QP_CUDA(cudaFuncSetAttribute(qp_genr_kernel<NP, MP, T, RELAX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
					     int(C::SMEM)));
		int per_sm = 1;
		QP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qp_genr_kernel<NP, MP, T, RELAX>,
								      GENR_WARPS * 32, C::SMEM));
		ctas_per_sm.set(dev, std::max(1, per_sm));
		attr_done.set(dev, 1);
	}
	const unsigned long long per_cta = GENR_WARPS * C::G;
	const unsigned long long need = (a.count + per_cta - 1) / per_cta;
	const unsigned long long grid =
		std::max(1ull, std::min((unsigned long long)dev_sms(dev) * ctas_per_sm.get(dev), need));
	unsigned long long *ticket = nullptr;
	if (need > grid) {
		const int rc = gen_ticket(dev, st, &ticket);
		if (rc != QP_OK)
			return rc;
	}
	qp_genr_kernel<NP, MP, T, RELAX><<<dim3((unsigned)grid), dim3(GENR_WARPS * 32), C::SMEM, st>>>(a, ticket);
	QP_CUDA(cudaGetLastError());
	return QP_OK;
}

template <int NP, int MP, int T> static int launch_genr_t(const GenArgs &a, int dev, cudaStream_t st)
{
	return a.alpha == 1.0f ? launch_genr_r<NP, MP, T, false>(a, dev, st) : launch_genr_r<NP, MP, T, true>(a, dev, st);
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
