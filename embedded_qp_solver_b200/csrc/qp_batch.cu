// qp_batch.cu -- host runtime and C ABI of libqpb200.so (include/qp_batch.h).
//
// Replaces, for a whole batch, the reference's per-problem call
//   struct _matrix *admm(x0, iterations, qf)     /root/reference/qp_solvers/qp_solvers.c:255-319
// and the static slab pools behind it (/root/reference/kmalloc/kmalloc.c:37-66) with
// preallocated device workspaces.  The kernels live in one translation unit per family
// (k_tpq.cu, k_bpq.cu, k_gpq.cu, k_cpq.cu, k_aux.cu; qp_internal.h declares their launchers).
//
// Two data paths:
//   QP_MEM_DEVICE  arrays already in HBM: pick the kernel for n, launch it on the
//                  caller's stream, return without synchronising.
//   QP_MEM_HOST    arrays in host memory: the batch is split into contiguous shards,
//                  one per GPU in device_mask (one host thread per GPU), and every
//                  shard into chunks that flow through a DEPTH-deep ring of
//                  {H2D copies, kernel, D2H copies} on separate streams, so the
//                  copy engines and the SMs overlap.  Pinned caller memory is used
//                  in place; pageable memory is staged through pinned buffers.
//                  No collective anywhere: problems are independent.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <functional>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#if defined(__SSE__)
#include <xmmintrin.h>
#endif

#include "qp_internal.h"

using namespace qpb;

namespace qpb { // qp_hostpack.cpp
bool hostpack_have_avx512();
void hostpack_upper16_avx512(const float *P, float *out, size_t count);
void hostpack_upper16_avx512_cached(const float *P, float *out, size_t count);
}

// ----------------------------------------------------------------------------------------
// errors
// ----------------------------------------------------------------------------------------
namespace qpb {
thread_local std::string g_err;
}

extern "C" const char *qp_last_error(void)
{
	return g_err.c_str();
}

extern "C" int qp_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess)
		return 0;
	return std::min(n, MAX_DEVICES);
}

extern "C" void qp_options_default(qp_options *o)
{
	if (!o)
		return;
	std::memset(o, 0, sizeof(*o));
	o->rho = 1.0f;        // qp_solvers.c:281
	o->alpha = 1.0f;      // qp_solvers.c:282
	o->abstol = 1e-4f;    // qp_solvers.c:17
	o->reltol = 1e-2f;    // qp_solvers.c:18
	o->max_iter = 10000;  // ADMM_ITERATIONS, config.h:38
	o->device_mask = 0;
	o->device = -1;
	o->stream = nullptr;
	o->chunk = 0;
	o->kernel = QP_KERNEL_AUTO;
	o->schedule = QP_SCHEDULE_AUTO;
}

extern "C" size_t qp_packed_stride(uint32_t n)
{
	return size_t(tpq_packed_stride(int(n)));
}

// ----------------------------------------------------------------------------------------
// kernel selection and launch
// ----------------------------------------------------------------------------------------
namespace qpb {

int dev_sms(int dev)
{
	static std::atomic<int> sms[MAX_DEVICES];
	int v = sms[dev].load(std::memory_order_acquire);
	if (!v) {
		cudaDeviceProp pr;
		if (cudaGetDeviceProperties(&pr, dev) != cudaSuccess)
			return 0;
		v = pr.multiProcessorCount;
		sms[dev].store(v, std::memory_order_release);
	}
	return v;
}

int resolve_device(int dev)
{
	if (dev >= 0)
		return dev < qp_device_count() ? dev : -1;
	int cur = 0;
	if (cudaGetDevice(&cur) != cudaSuccess)
		return -1;
	return cur < MAX_DEVICES ? cur : -1;
}

bool family_has_passes(int family, int n)
{
	return family == QP_KERNEL_TPQ || (family == QP_KERNEL_BPQ && n >= 8);
}

} // namespace qpb

namespace {

// which kernel family solves this batch (QP_KERNEL_*), 0 = none
int pick_kernel(const KernelArgs &a, uint32_t forced)
{
	const bool t = tpq_eligible(a), b = bpq_eligible(a), g = a.n >= 1 && a.n <= QP_MAX_DIM;
	const bool c = a.n >= 1 && a.n <= QP_MAX_DIM_BOX;
	if (a.p_packed) // packed upper triangles: the thread-per-QP kernels only
		return (t && a.n >= 8 && (forced == QP_KERNEL_AUTO || forced == QP_KERNEL_TPQ)) ? QP_KERNEL_TPQ : 0;
	if (forced > QP_KERNEL_CPQ) // the general-rows kernels do not take a box batch
		return 0;
	switch (forced) {
	case QP_KERNEL_TPQ: return t ? QP_KERNEL_TPQ : 0;
	case QP_KERNEL_BPQ: return b ? QP_KERNEL_BPQ : 0;
	case QP_KERNEL_GPQ: return g ? QP_KERNEL_GPQ : 0;
	case QP_KERNEL_CPQ: return c ? QP_KERNEL_CPQ : 0;
	}
	if (a.n > QP_MAX_DIM)
		return c ? QP_KERNEL_CPQ : 0;
	if (t)
		return QP_KERNEL_TPQ;
	if (b && a.n >= 24) // measured at n = 24, 32, 40: bpq ahead of gpq by 12, 11 and 27 % (profiles/README.md)
		return QP_KERNEL_BPQ;
	return g ? QP_KERNEL_GPQ : 0;
}

// ---- passes for heavy-tailed iteration counts (qp_common.cuh, PassArgs) -------------------
//
// Per device: a grow-only scratch arena (index lists, list lengths, the saved z / u /
// iteration count of evicted problems and, for qp_bpq.cuh, their inverses) and three
// counters the kernels keep (problems that outlived the first window, problems looked at,
// evictions by later passes).  The arena belongs to the library (the reference's pools are
// static too, kmalloc/kmalloc.c:37-42); solves that use it are chained by an event, so
// callers on different streams of one device never share it in time.
//
// QP_SCHEDULE_AUTO follows what the device has seen, PER DIMENSION n (a host that solves
// well-conditioned n = 16 problems and heavy-tailed n = 40 ones on one GPU gets a schedule for
// each): the kernels add to the counters of their n; behind every such solve the row is
// copied to pinned host memory on a side stream (nobody waits for it), alternately into two
// buffers, and the next solve of that n looks at the DIFFERENCE of the two -- what the most
// recent solves did, not what the device has done since the last decision (a burst of 20
// MPC solves would otherwise outvote the well-conditioned batch that follows it for as long
// again).  The schedule never changes a result (bit for bit, tests/test_passes_gpu.py), only
// where the time goes.
constexpr int MAX_PASSES = 16;
constexpr int STAT_WORDS = 8; // PassArgs::stats
constexpr int STAT_ROWS = QP_MAX_DIM_BOX + 1; // one row of counters per n
struct DevArena {
	std::mutex mu;
	char *base = nullptr;
	size_t bytes = 0;
	cudaEvent_t last = nullptr;  // recorded behind the latest solve that used the scratch
	unsigned long long *stat_dev = nullptr;  // [STAT_ROWS][STAT_WORDS] counters in device memory
	unsigned long long *stat_host = nullptr; // pinned [2][STAT_ROWS][STAT_WORDS]: the two latest copies
	cudaStream_t side = nullptr;             // carries the copies
	cudaEvent_t ev_side = nullptr;
	unsigned char flip[STAT_ROWS] = {};      // which buffer the next copy of a row goes to
	bool heavy[STAT_ROWS] = {};              // AUTO's current belief about the work of this n
	std::chrono::steady_clock::time_point last_copy; // when the counters were last sent for
	bool copied_once = false;
};
DevArena g_arena[MAX_DEVICES];

struct PassPolicy {
	unsigned w0, grow; // window of pass k = w0 << (grow k)
};

// thread-per-QP and lane groups: 24, 48, 96, ... iterations (the benchmark distribution needs
// at most 18, so a first pass evicts nothing of it).  QPB200_PASSES="w0,grow" overrides
// (measurements, tests).
// Lane groups followed by the streaming pass (qp_bpq_stream.cuh): a first window of 6 -- it only
// has to tell the expected stragglers from the rest (the classifier in qp_bpq.cuh needs about 6
// iterations), every iteration beyond that is cheaper in the streaming kernel, where no group
// waits for another (condensed-MPC set, 2^18 problems: window 24 64 M QP/s, 12 75 M, 8 78 M,
// 6 and 4 80-81 M).
PassPolicy pass_policy(int, bool stream)
{
	PassPolicy pol = { stream ? 6u : 24u, 1u };
	if (const char *e = std::getenv("QPB200_PASSES")) {
		unsigned a = 0, b = 0;
		if (std::sscanf(e, "%u,%u", &a, &b) == 2 && a > 0)
			pol.w0 = a, pol.grow = std::min(b, 4u);
	}
	return pol;
}

size_t up256(size_t v)
{
	return (v + 255) / 256 * 256;
}

int arena_stats_init(DevArena &ar)
{
	if (ar.stat_dev)
		return QP_OK;
	QP_CUDA(cudaMalloc(&ar.stat_dev, STAT_ROWS * STAT_WORDS * 8));
	QP_CUDA(cudaMemset(ar.stat_dev, 0, STAT_ROWS * STAT_WORDS * 8));
	QP_CUDA(cudaHostAlloc(&ar.stat_host, 2 * STAT_ROWS * STAT_WORDS * 8, cudaHostAllocDefault));
	std::memset(ar.stat_host, 0, 2 * STAT_ROWS * STAT_WORDS * 8);
	QP_CUDA(cudaStreamCreateWithFlags(&ar.side, cudaStreamNonBlocking));
	QP_CUDA(cudaEventCreateWithFlags(&ar.ev_side, cudaEventDisableTiming));
	QP_CUDA(cudaEventCreateWithFlags(&ar.last, cudaEventDisableTiming));
	return QP_OK;
}

int launch_family(const KernelArgs &a, int fam, int dev, cudaStream_t st, int mode)
{
	switch (fam) {
	case QP_KERNEL_TPQ: return launch_tpq_family(a, dev, st, mode);
	case QP_KERNEL_BPQ: return launch_bpq_family(a, dev, st, mode);
	case QP_KERNEL_GPQ: return launch_gpq_family(a, dev, st);
	case QP_KERNEL_CPQ: return launch_cpq_family(a, dev, st);
	}
	return QP_ERR_UNSUPPORTED;
}

// Solve one device-resident batch on `dev`, stream `st` (the device is current).
int launch_solver(KernelArgs a, int dev, cudaStream_t st, uint32_t forced, uint32_t schedule)
{
	if (a.count == 0)
		return QP_OK;
	const int fam = pick_kernel(a, forced);
	if (!fam)
		return QP_ERR_UNSUPPORTED;
	const char *se = std::getenv("QPB200_BPQ_STREAM"); // =0: list passes instead of the streaming pass
	// (bit 31 of a problem index is a tag in the streaming kernel)
	const bool stream = fam == QP_KERNEL_BPQ && !(se && se[0] == '0') && bpq_stream_eligible(a) &&
			    a.count < 0x7ffffff0ull;
	const PassPolicy pol = pass_policy(fam, stream);

	// Shared matrices (qp_batch.p_share) in the lane-group kernels, where the elimination is
	// most of a solve: eliminate each MATRIX once (qp_bpq.cuh, PASS_SETUP), then the streaming
	// kernel solves every problem from scratch with the inverse of its group.  Not with
	// per-problem rho (the inverse depends on it) and not when the caller forces single launches.
	// Measured on random problems (K = 10), 2^18 of them, matrix per problem -> one per 2 / 16 / 1024
	// problems: n = 40 170 -> 194 / 312 / 343 M QP/s, n = 32 313 -> 301 / 429 / 458 M, n = 24
	// 550 -> 418 / 518 / 539 M (the elimination is too small a part there: plain kernel).
	const bool shared_pays = (a.n == 40 && a.p_share >= 2u) || (a.n == 32 && a.p_share >= 4u);
	if (stream && shared_pays && schedule != QP_SCHEDULE_SINGLE && !a.rho_v && a.count < 0x7ffffff0ull) {
		cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
		const bool capturing = cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone;
		if (capturing)
			cudaGetLastError();
		DevArena &ar = g_arena[dev];
		std::lock_guard<std::mutex> lk(ar.mu);
		if (int rc = arena_stats_init(ar))
			return rc;
		const size_t nmat = size_t((a.count + a.p_share_off + a.p_share - 1) / a.p_share), nn = size_t(a.n) * a.n;
		const size_t o_its = 256, o_R = o_its + up256(nmat * 4), need = o_R + up256(nmat * nn * 4);
		if (ar.bytes < need) {
			if (capturing)
				return QP_ERR_UNSUPPORTED; // (no allocation inside a capture: warm the arena up first)
			QP_CUDA(cudaEventSynchronize(ar.last));
			if (ar.base)
				cudaFree(ar.base);
			ar.base = nullptr, ar.bytes = 0;
			QP_CUDA(cudaMalloc(&ar.base, need));
			ar.bytes = need;
		}
		QP_CUDA(cudaStreamWaitEvent(st, ar.last, 0));
		unsigned *cnt = reinterpret_cast<unsigned *>(ar.base);
		QP_CUDA(cudaMemsetAsync(cnt, 0, MAX_PASSES * 4, st));
		KernelArgs su = a; // the matrices as a batch of their own
		su.count = nmat, su.p_share = 1u, su.p_share_off = 0;
		su.z0 = su.u0 = nullptr, su.rho_v = nullptr;
		su.x = su.z = su.lam = nullptr, su.status = su.iters = nullptr;
		su.ps = PassArgs{};
		su.ps.its = reinterpret_cast<int *>(ar.base + o_its);
		su.ps.Rs = reinterpret_cast<float *>(ar.base + o_R);
		if (int rc = launch_bpq_family(su, dev, st, PASS_SETUP))
			return rc;
		a.ps = PassArgs{};
		a.ps.its = su.ps.its, a.ps.Rs = su.ps.Rs;
		a.ps.stream_ctr = cnt + (MAX_PASSES - 1);
		if (int rc = launch_bpq_stream_family(a, dev, st, true))
			return rc;
		QP_CUDA(cudaEventRecord(ar.last, st));
		return QP_OK;
	}
	const bool can = family_has_passes(fam, a.n) && a.count < 0xfffffff0ull && a.max_iter > pol.w0;
	bool passes = can && schedule == QP_SCHEDULE_PASSES;
	// (the probe adds iteration counts per lane in 32 bits: with 148 SMs a lane sees fewer than
	// count / 1024 problems)
	bool look = can && schedule == QP_SCHEDULE_AUTO && a.max_iter > 64 && a.count >= 4096 &&
		    (a.count / 1024 + 1) * (unsigned long long)a.max_iter < 0xffffffffull;
	if (look) { // a stream that is being captured into a graph cannot take the side-stream copy
		cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
		if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone)
			look = false, cudaGetLastError();
	}
	if (!passes && !look)
		return launch_family(a, fam, dev, st, PASS_PLAIN);

	DevArena &ar = g_arena[dev];
	std::lock_guard<std::mutex> lk(ar.mu);
	if (int rc = arena_stats_init(ar))
		return rc;
	const int row = a.n; // 1 ... QP_MAX_DIM_BOX
	if (look) {
		// the two latest copies of this n's counters; they only grow, so the larger one is newer
		const unsigned long long *h0 = ar.stat_host + size_t(row) * STAT_WORDS;
		const unsigned long long *h1 = h0 + size_t(STAT_ROWS) * STAT_WORDS;
		unsigned long long b0[STAT_WORDS], b1[STAT_WORDS], d[STAT_WORDS];
		for (int k = 0; k < STAT_WORDS; ++k)
			b0[k] = h0[k], b1[k] = h1[k];
		const bool newer1 = b1[1] + b1[5] >= b0[1] + b0[5];
		for (int k = 0; k < STAT_WORDS; ++k)
			d[k] = newer1 ? b1[k] - b0[k] : b0[k] - b1[k];
		bool &heavy = ar.heavy[row];
		if (!heavy && d[5] > 0) {
			// single launches lately: passes pay when the warps mostly waited for their slowest
			// problem (less than half of the lane-iterations did work) AND the iterations are
			// what the time goes to (a pass repeats the factorisation, or moves the inverse
			// through HBM): mean iteration count >= 16 (thread per QP) / 64 (lane groups).
			// Measured both ways on the reference's distribution and the condensed MPC set,
			// profiles/README.md "Passes".
			const double mean = double(d[3]) / double(d[5]);
			const double worked = d[4] ? double(d[3]) / double(d[4]) : 1.0;
			// (the streaming pass has no list passes to pay for and no group waits in it: it is
			// ahead as soon as a quarter of the lane-iterations wait and the iterations, not the
			// elimination, are where the time goes -- MPC set: 0.58 worked, mean 36)
			heavy = stream ? (worked < 0.75 && mean >= 24.0)
				       : (worked < 0.5 && mean >= (fam == QP_KERNEL_TPQ ? 16.0 : 64.0));
		} else if (heavy && d[1] > 0) {
			// passes lately: back to single launches when fewer than 1 problem in 128 outlived
			// the first window -- or, with the streaming pass (whose first window nearly
			// everything outlives), when the streamed problems needed fewer than 20 iterations
			// on average
			heavy = stream ? (d[7] == 0 || d[6] >= 20ull * d[7]) : d[0] * 128ull > d[1];
		}
		passes = heavy;
	}
	a.ps.stats = ar.stat_dev + size_t(row) * STAT_WORDS;
	auto copy_stats = [&]() -> int {
		// an event between back-to-back launches costs a microsecond or two of the stream's
		// time: behind SHORT solves that are enqueued closely after each other, copy every
		// 200 us (of the host's time) only; a solve of a millisecond always gets its copy
		const auto now = std::chrono::steady_clock::now();
		const bool small = a.count * (unsigned long long)(a.n * a.n) < (1ull << 24);
		if (small && ar.copied_once && now - ar.last_copy < std::chrono::microseconds(200))
			return QP_OK;
		ar.last_copy = now, ar.copied_once = true;
		QP_CUDA(cudaEventRecord(ar.ev_side, st));
		QP_CUDA(cudaStreamWaitEvent(ar.side, ar.ev_side, 0));
		unsigned long long *dst = ar.stat_host + (size_t(ar.flip[row]) * STAT_ROWS + row) * STAT_WORDS;
		ar.flip[row] ^= 1;
		QP_CUDA(cudaMemcpyAsync(dst, a.ps.stats, STAT_WORDS * 8, cudaMemcpyDeviceToHost, ar.side));
		return QP_OK;
	};
	if (!passes) {
		if (int rc = launch_family(a, fam, dev, st, PASS_PLAIN))
			return rc;
		return copy_stats();
	}

	// windows of the passes: w0, w0 << grow, ... ; the last one runs every problem out
	unsigned win[MAX_PASSES];
	int npass = 0;
	for (unsigned long long cum = 0; npass < MAX_PASSES; ++npass) {
		const unsigned long long w = (unsigned long long)pol.w0 << std::min(pol.grow * unsigned(npass), 24u);
		if (cum + w >= a.max_iter || npass == MAX_PASSES - 1) {
			win[npass++] = a.max_iter;
			break;
		}
		win[npass] = unsigned(w);
		cum += w;
	}
	const size_t c = size_t(a.count), n = size_t(a.n);
	const bool want_R = fam == QP_KERNEL_BPQ;
	const size_t o_list0 = 256, o_list1 = o_list0 + up256(c * 4), o_its = o_list1 + up256(c * 4),
		     o_zs = o_its + up256(c * 4), o_us = o_zs + up256(c * n * 4),
		     o_xs = o_us + up256(c * n * 4), o_R = o_xs + (want_R ? up256(c * n * 4) : 0),
		     need = o_R + (want_R ? up256(c * n * n * 4) : 0);
	if (ar.bytes < need) {
		QP_CUDA(cudaEventSynchronize(ar.last)); // nobody is using the old one any more
		if (ar.base)
			cudaFree(ar.base);
		ar.base = nullptr, ar.bytes = 0;
		QP_CUDA(cudaMalloc(&ar.base, need));
		ar.bytes = need;
	}
	QP_CUDA(cudaStreamWaitEvent(st, ar.last, 0));
	unsigned *cnt = reinterpret_cast<unsigned *>(ar.base); // [npass] list lengths
	unsigned *list[2] = { reinterpret_cast<unsigned *>(ar.base + o_list0),
			      reinterpret_cast<unsigned *>(ar.base + o_list1) };
	QP_CUDA(cudaMemsetAsync(cnt, 0, MAX_PASSES * 4, st));
	a.ps.its = reinterpret_cast<int *>(ar.base + o_its);
	a.ps.zs = reinterpret_cast<float *>(ar.base + o_zs);
	a.ps.us = reinterpret_cast<float *>(ar.base + o_us);
	a.ps.Rs = want_R ? reinterpret_cast<float *>(ar.base + o_R) : nullptr;
	a.ps.xs = (want_R && stream) ? reinterpret_cast<float *>(ar.base + o_xs) : nullptr;
	// lane groups with the full inverse: two launches -- a first window for everybody, then the
	// streaming kernel (qp_bpq_stream.cuh) runs the unfinished problems out, each lane group
	// taking its next one as soon as its current one stops.
	if (stream && npass >= 2) {
		a.ps.window = win[0];
		a.ps.list = nullptr, a.ps.list_count = nullptr;
		a.ps.next_list = list[0], a.ps.next_count = cnt;
		a.ps.next_list2 = list[1], a.ps.next_count2 = cnt + 1;
		if (int rc = launch_family(a, fam, dev, st, PASS_FIRST))
			return rc;
		a.ps.window = a.max_iter;
		a.ps.list = list[0], a.ps.list_count = cnt;
		a.ps.list2 = list[1], a.ps.list_count2 = cnt + 1;
		a.ps.next_list = nullptr, a.ps.next_count = nullptr;
		a.ps.next_list2 = nullptr, a.ps.next_count2 = nullptr;
		a.ps.stream_ctr = cnt + (MAX_PASSES - 1);
		if (int rc = launch_bpq_stream_family(a, dev, st, false))
			return rc;
		QP_CUDA(cudaEventRecord(ar.last, st));
		return copy_stats();
	}
	for (int k = 0; k < npass; ++k) {
		const bool last = k == npass - 1;
		a.ps.window = win[k];
		a.ps.list = k ? list[(k - 1) & 1] : nullptr;
		a.ps.list_count = k ? cnt + (k - 1) : nullptr;
		a.ps.next_list = last ? nullptr : list[k & 1];
		a.ps.next_count = last ? nullptr : cnt + k;
		if (int rc = launch_family(a, fam, dev, st, k ? PASS_LIST : PASS_FIRST))
			return rc;
	}
	QP_CUDA(cudaEventRecord(ar.last, st));
	return copy_stats();
}

} // namespace

extern "C" int qp_schedule_stats(int device, unsigned long long stats[6])
{
	const int dev = resolve_device(device);
	if (dev < 0 || !stats)
		return QP_ERR_INVALID;
	std::vector<unsigned long long> v(size_t(STAT_ROWS) * STAT_WORDS, 0ull);
	DevArena &ar = g_arena[dev];
	std::lock_guard<std::mutex> lk(ar.mu);
	if (ar.stat_dev) {
		DeviceGuard g;
		if (g.enter(dev) != QP_OK)
			return QP_ERR_NO_DEVICE;
		QP_CUDA(cudaDeviceSynchronize());
		QP_CUDA(cudaMemcpy(v.data(), ar.stat_dev, v.size() * 8, cudaMemcpyDeviceToHost));
	}
	for (int k = 0; k < 6; ++k) { // (the counters are kept per dimension: their sums)
		stats[k] = 0;
		for (int r = 0; r < STAT_ROWS; ++r)
			stats[k] += v[size_t(r) * STAT_WORDS + k];
	}
	return QP_OK;
}

namespace {

int check_problem(const qp_batch *pb, const qp_batch_solution *sol)
{
	if (!pb || !sol)
		return QP_ERR_INVALID;
	if (pb->n == 0)
		return QP_ERR_INVALID;
	if (pb->n > QP_MAX_DIM_BOX)
		return QP_ERR_UNSUPPORTED;
	if (pb->mem != QP_MEM_HOST && pb->mem != QP_MEM_DEVICE)
		return QP_ERR_INVALID;
	if (pb->p_layout != QP_P_DENSE && pb->p_layout != QP_P_UPPER)
		return QP_ERR_INVALID;
	if (pb->p_layout == QP_P_UPPER && !(pb->n == 8 || pb->n == 12 || pb->n == 16))
		return QP_ERR_UNSUPPORTED;
	if (pb->count && (!pb->P || !pb->q))
		return QP_ERR_INVALID;
	return QP_OK;
}

int check_options(const qp_options &o)
{
	if (!(o.rho > 0.0f) || !std::isfinite(o.rho))
		return QP_ERR_INVALID;
	if (!std::isfinite(o.alpha) || !(o.alpha > 0.0f) || !(o.alpha < 2.0f))
		return QP_ERR_INVALID;
	if (!(o.abstol >= 0.0f) || !(o.reltol >= 0.0f))
		return QP_ERR_INVALID;
	if (o.schedule > QP_SCHEDULE_PASSES)
		return QP_ERR_INVALID;
	if (o.precision > QP_PRECISION_F64)
		return QP_ERR_INVALID;
	return QP_OK;
}

KernelArgs make_args(const qp_batch &pb, const qp_batch_solution &sol, const qp_options &o)
{
	KernelArgs a = {};
	a.P = pb.P, a.q = pb.q, a.lb = pb.lb, a.ub = pb.ub, a.z0 = pb.z0, a.u0 = pb.u0;
	a.x = sol.x, a.z = sol.z, a.lam = sol.lambda, a.status = sol.status, a.iters = sol.iters;
	a.count = pb.count;
	a.n = int(pb.n);
	a.lb_s = pb.lb_scalar, a.ub_s = pb.ub_scalar;
	a.rho = o.rho, a.alpha = o.alpha;
	a.rho_v = pb.rho;
	// sqrt(nrows) * abstol, evaluated in double like qp_solvers.c:242 and rounded once
	a.eps_abs = float(std::sqrt(double(pb.n)) * double(o.abstol));
	a.reltol = o.reltol;
	a.max_iter = o.max_iter;
	a.p_packed = pb.p_layout == QP_P_UPPER ? 1 : 0;
	a.p_share = pb.p_share > 1u ? pb.p_share : 1u;
	a.p_share_off = 0;
	return a;
}

} // namespace

// ----------------------------------------------------------------------------------------
// host-side packing of P for the host-memory path
//
// The path is bound by the PCIe copy of P (1 024 of the 1 088 input bytes per problem at
// n = 16).  P is symmetric and the thread-per-QP kernels read only its upper triangle, so
// for n = 8, 12, 16 the host packs the upper triangles of a chunk into pinned staging
// (a small pool of worker threads, overlapped with the previous chunk's DMA) and only those
// travel: 544 instead of 1 024 bytes per problem at n = 16.  QPB200_HOST_PACK=0 disables it.
// ----------------------------------------------------------------------------------------
namespace {

class HostPool {
public:
	explicit HostPool(unsigned nthreads)
	{
		for (unsigned t = 0; t < nthreads; ++t)
			workers_.emplace_back([this, t] { loop(t); });
	}
	~HostPool()
	{
		{
			std::lock_guard<std::mutex> lk(mu_);
			stop_ = true;
			++gen_;
		}
		cv_.notify_all();
		for (std::thread &w : workers_)
			w.join();
	}
	// body(part, parts) on every worker, asynchronously; wait() returns when all are done
	void start(std::function<void(unsigned, unsigned)> body)
	{
		std::unique_lock<std::mutex> lk(mu_);
		done_.wait(lk, [this] { return pending_ == 0; }); // a previous call that returned early
		body_ = std::move(body);
		pending_ = unsigned(workers_.size());
		t0_ = std::chrono::steady_clock::now();
		++gen_;
		cv_.notify_all();
	}
	// seconds from start() to the last worker's return, of the call that finished last
	double last_seconds()
	{
		std::lock_guard<std::mutex> lk(mu_);
		return last_s_;
	}
	void wait()
	{
		std::unique_lock<std::mutex> lk(mu_);
		done_.wait(lk, [this] { return pending_ == 0; });
	}
	bool idle()
	{
		std::lock_guard<std::mutex> lk(mu_);
		return pending_ == 0;
	}

private:
	void loop(unsigned id)
	{
		unsigned long seen = 0;
		for (;;) {
			{
				std::unique_lock<std::mutex> lk(mu_);
				cv_.wait(lk, [&] { return gen_ != seen; });
				seen = gen_;
				if (stop_)
					return;
			}
			// body_ is not replaced before every worker has reported back
			if (body_)
				body_(id, unsigned(workers_.size()));
			{
				std::lock_guard<std::mutex> lk(mu_);
				if (--pending_ == 0) {
					last_s_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count();
					done_.notify_all();
				}
			}
		}
	}
	std::vector<std::thread> workers_;
	std::mutex mu_;
	std::condition_variable cv_, done_;
	std::function<void(unsigned, unsigned)> body_;
	unsigned pending_ = 0;
	unsigned long gen_ = 0;
	bool stop_ = false;
	std::chrono::steady_clock::time_point t0_;
	double last_s_ = 0.0;
};

// upper triangle (row-major, i <= j) of `count` n x n matrices, `pk` floats apart.
// Every row is moved as one fixed-size block of N floats starting at its diagonal element
// (a few vector moves; the floats behind the row's end are overwritten by the next row)
// into a cache-resident record, which leaves with non-temporal stores: the staging buffer
// is only ever read by the copy engine.  Measured on the 16-vCPU hosts of this pool, n = 16,
// 2^20 problems end to end: 59 M QP/s; the same rows written straight to the staging buffer
// (-DQPB_PACK_MEMCPY) 55 M; no packing (full matrices over PCIe) 49 M.
template <int N> void pack_upper_n(const float *P, float *out, size_t count, int pk)
{
#if !defined(QPB_PACK_MEMCPY) && defined(__SSE__)
	alignas(64) float rec[N * (N + 1) / 2 + N + 4];
	for (size_t c = 0; c < count; ++c) {
		const float *src = P + c * size_t(N) * N;
		int k = 0;
		for (int i = 0; i < N - 1; ++i) {
			__builtin_memcpy(rec + k, src + i * N + i, N * sizeof(float));
			k += N - i;
		}
		rec[k++] = src[N * N - 1];
		for (; k < pk; ++k)
			rec[k] = 0.0f;
		float *dst = out + c * size_t(pk);
		for (int v = 0; v < pk; v += 4)
			_mm_stream_ps(dst + v, _mm_load_ps(rec + v));
	}
	_mm_sfence();
#else
	for (size_t c = 0; c + 1 < count; ++c) {
		const float *src = P + c * size_t(N) * N;
		float *dst = out + c * size_t(pk);
		int k = 0;
		for (int i = 0; i < N; ++i) {
			__builtin_memcpy(dst + k, src + i * N + i, N * sizeof(float));
			k += N - i;
		}
	}
	if (count) {
		const float *src = P + (count - 1) * size_t(N) * N;
		float *dst = out + (count - 1) * size_t(pk);
		int k = 0;
		for (int i = 0; i < N; ++i)
			for (int j = i; j < N; ++j)
				dst[k++] = src[i * N + j];
		for (; k < pk; ++k)
			dst[k] = 0.0f;
	}
#endif
}

// wide: -1 = AVX-512 when the CPU has it (QPB200_HOST_PACK_SSE=1 keeps the 128-bit version,
// for measurements), 0 = 128-bit, 1 = AVX-512 with ordinary stores, 2 = AVX-512 with
// non-temporal stores (the caller has checked the CPU)
void pack_upper(const float *P, float *out, size_t count, int n, int wide = -1)
{
	const int pk = tpq_packed_stride(n);
	// n = 16 with AVX-512 (qp_hostpack.cpp): about twice the rate per core
	static const bool have = qpb::hostpack_have_avx512() && !std::getenv("QPB200_HOST_PACK_SSE");
	if (n == 16 && pk == 136 && (wide < 0 ? have : wide > 0)) {
		// ordinary stores straight into the staging buffer beat the non-temporal ones by
		// ~5 % end to end with the 512-bit moves (65 M against 61 M QP/s; with the 128-bit
		// version it is the other way round); QPB200_PACK_NT=1 keeps the non-temporal ones
		static const bool nt = std::getenv("QPB200_PACK_NT") != nullptr;
		if (wide < 0 ? !nt : wide == 1)
			qpb::hostpack_upper16_avx512_cached(P, out, count);
		else
			qpb::hostpack_upper16_avx512(P, out, count);
		return;
	}
	switch (n) {
	case 8:  pack_upper_n<8>(P, out, count, pk); break;
	case 12: pack_upper_n<12>(P, out, count, pk); break;
	case 16: pack_upper_n<16>(P, out, count, pk); break;
	}
}

bool host_pack_enabled()
{
	static const bool on = [] {
		const char *e = std::getenv("QPB200_HOST_PACK");
		return !(e && e[0] == '0');
	}();
	return on;
}

// QPB200_HOST_PACK=2: pack regardless of the timing (measurements);
// QPB200_HOST_PACK=3: give packing up after three chunks regardless of the timing (tests of
// the switch in the middle of a call)
bool host_pack_forced()
{
	static const bool on = [] {
		const char *e = std::getenv("QPB200_HOST_PACK");
		return e && e[0] == '2';
	}();
	return on;
}

bool host_pack_test_switch()
{
	static const bool on = [] {
		const char *e = std::getenv("QPB200_HOST_PACK");
		return e && e[0] == '3';
	}();
	return on;
}

} // namespace

// One core packs ~10 GB/s and the copy engine wants ~85 GB/s of it: with fewer than 12
// hardware threads the direct copy of the full matrices is the faster path.
//
// Packing trades PCIe bytes for host-memory traffic (about 2.1 KB instead of 1.1 KB per
// problem at n = 16), so it only pays while the host's DRAM is not the shared bottleneck:
// it is used when ONE GPU is fed from this host -- counting the ranks of a torchrun job
// (LOCAL_WORLD_SIZE) and the devices of this process -- and needs 12 hardware threads.
// Measured with two ranks on a 24-thread host: 68.9 M QP/s packed against 92.2 M plain, and
// with the greedy mix of packed and plain chunks (QPB200_HOST_PACK=2 lifts this rule)
// 85-88 M against 93 M plain.
static unsigned host_threads_per_gpu(unsigned gpus_in_this_process)
{
	unsigned ranks = 1;
	if (const char *e = std::getenv("LOCAL_WORLD_SIZE"))
		ranks = unsigned(std::max(1, std::atoi(e)));
	const unsigned sharers = std::max(1u, ranks * std::max(1u, gpus_in_this_process));
	if (sharers > 1 && !host_pack_forced())
		return 0;
	return std::thread::hardware_concurrency() / sharers;
}

// set when a workspace of this process gave packing up (see run_shard)
static std::atomic<bool> g_pack_given_up{false};

// bytes the QP_MEM_HOST path has handed to the copy engine (host to device) in this process
static std::atomic<unsigned long long> g_h2d_bytes{0};

extern "C" unsigned long long qp_host_path_h2d_bytes(void)
{
	return g_h2d_bytes.load();
}

extern "C" int qp_host_pack_upper(const float *P, float *out, size_t count, uint32_t n, int wide)
{
	if (!P || !out || !(n == 8 || n == 12 || n == 16) || (reinterpret_cast<uintptr_t>(out) & 15u))
		return QP_ERR_INVALID;
	if (wide > 0 && !qpb::hostpack_have_avx512())
		return QP_ERR_UNSUPPORTED; // no AVX-512 on this CPU
	pack_upper(P, out, count, int(n), wide);
	return QP_OK;
}

extern "C" int qp_host_path_packs(uint32_t n)
{
	return host_pack_enabled() && !g_pack_given_up.load() && (n == 8 || n == 12 || n == 16) &&
	       host_threads_per_gpu(1) >= 12;
}

// ----------------------------------------------------------------------------------------
// workspace for the host-memory path
// ----------------------------------------------------------------------------------------
static constexpr int DEPTH = 3;
#ifndef QPB_ZERO_COPY_DEFAULT
#define QPB_ZERO_COPY_DEFAULT (256u << 10) // the whole range of the one-copy path (measured: faster up to there)
#endif

// One thread per GPU that finishes chunks whose results go to PAGEABLE caller memory: it waits
// for the chunk's event and copies the results from the slot's pinned buffer into the caller's
// arrays, so that the enqueueing thread (which also feeds the packing pool) does neither.  Tasks
// run in the order they were pushed -- the order the chunks were enqueued on their streams.
class OutCopier {
public:
	OutCopier() : th_([this] { loop(); }) {}
	~OutCopier()
	{
		{
			std::lock_guard<std::mutex> lk(mu_);
			stop_ = true;
		}
		cv_.notify_all();
		th_.join();
	}
	void push(std::function<void()> f)
	{
		{
			std::lock_guard<std::mutex> lk(mu_);
			q_.push_back(std::move(f));
			++open_;
		}
		cv_.notify_all();
	}
	// every task pushed so far has run
	void drain()
	{
		std::unique_lock<std::mutex> lk(mu_);
		done_.wait(lk, [this] { return open_ == 0; });
	}

private:
	void loop()
	{
		for (;;) {
			std::function<void()> f;
			{
				std::unique_lock<std::mutex> lk(mu_);
				cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
				if (q_.empty())
					return;
				f = std::move(q_.front());
				q_.pop_front();
			}
			f();
			{
				std::lock_guard<std::mutex> lk(mu_);
				if (--open_ == 0)
					done_.notify_all();
			}
		}
	}
	std::mutex mu_;
	std::condition_variable cv_, done_;
	std::deque<std::function<void()>> q_;
	size_t open_ = 0;
	bool stop_ = false;
	std::thread th_; // last: the thread starts when everything above exists
};

struct Slot {
	// device buffers for one chunk
	float *dP = nullptr, *dq = nullptr, *dlb = nullptr, *dub = nullptr, *dz0 = nullptr,
	      *du0 = nullptr, *drho = nullptr;
	float *dx = nullptr, *dz = nullptr, *dlam = nullptr;
	int32_t *dstatus = nullptr, *diters = nullptr;
	// pinned staging (allocated on first use with pageable caller memory)
	float *hin = nullptr;  // P | q | lb | ub | z0 | u0 | rho
	float *hout = nullptr; // x | z | lam | status | iters
	cudaStream_t stream = nullptr;
	cudaEvent_t done = nullptr;
	cudaEvent_t t0 = nullptr, t1 = nullptr; // around the copy of the packed triangles
	// bookkeeping of the chunk in flight
	size_t first = 0, cnt = 0;
	bool busy = false;
	bool timed = false; // t0, t1 were recorded for the chunk in flight
	// the chunk in flight is finished by the OutCopier: 1 while its task is open, then 0;
	// copy_rc = what cudaEventSynchronize returned there
	bool copier = false;
	int copy_open = 0; // (accessed with __atomic builtins: Slot stays movable)
	cudaError_t copy_rc = cudaSuccess;
};

struct DeviceCtx {
	int dev = 0;
	Slot slot[DEPTH];
	HostPool *pool = nullptr; // packing workers of this device's shard
	OutCopier *out = nullptr; // results to pageable caller memory (created on first use)
	// pinned buffers of packed upper triangles (n = 8, 12, 16), filled in turn by the pool
	// (one being filled, one filled and waiting, one being read by the copy engine);
	// hpk_free[b] fires when the copy engine has read buffer b
	static constexpr int NPK = 3;
	float *hpk[NPK] = { nullptr, nullptr, nullptr };
	cudaEvent_t hpk_free[NPK] = { nullptr, nullptr, nullptr };
	bool hpk_busy[NPK] = { false, false, false };
	// host-side packing is given up for this workspace when the workers turn out slower
	// than the copy engine would be on the full matrices (a slow or busy host)
	bool pack_off = false;
	unsigned slow_chunks = 0;     // consecutive chunks whose packing took longer than that
	double h2d_bytes_per_s = 0.0; // measured on the packed copies
};

// `cap` is the capacity of a slot in problems: every buffer is laid out for it, and it only
// grows.  A call cuts its shards into chunks of any size up to cap, so batches of different
// sizes share one workspace without reallocation.
struct qp_workspace {
	uint32_t n = 0;
	size_t cap = 0;
	std::vector<DeviceCtx> devs;
	std::mutex mu; // held for the whole of a solve
	// default-workspace cache (qp_solve_batch): key and age
	uint32_t key_mask = 0;
	int key_dev = -1;
	unsigned long stamp = 0;
};

namespace {

size_t in_floats(uint32_t n)
{
	return size_t(n) * n + 5ull * n + 1;
}
size_t out_floats(uint32_t n)
{
	return 3ull * n + 2;
}

size_t choose_chunk(uint32_t n, size_t per_dev)
{
	// ~32 MB of input per chunk, at least 4 chunks per device when the shard allows
	const size_t bytes = in_floats(n) * 4;
	size_t c = std::max<size_t>(1024, (32ull << 20) / bytes);
	if (per_dev / 4 >= 4096)
		c = std::min(c, per_dev / 4);
	c = std::max<size_t>(32, std::min(c, std::max<size_t>(per_dev, 32)));
	return (c + 31) / 32 * 32;
}

std::vector<int> mask_devices(uint32_t mask, int dflt)
{
	std::vector<int> v;
	const int nd = qp_device_count();
	if (mask == 0) {
		if (dflt >= 0 && dflt < nd)
			v.push_back(dflt);
		return v;
	}
	for (int i = 0; i < 32 && i < nd; ++i)
		if (mask & (1u << i))
			v.push_back(i);
	return v;
}

void free_ctx(DeviceCtx &c)
{
	int prev = 0;
	cudaGetDevice(&prev);
	cudaSetDevice(c.dev);
	if (c.pool)
		c.pool->wait();
	for (Slot &s : c.slot) {
		if (s.stream)
			cudaStreamSynchronize(s.stream);
		cudaFree(s.dP);
		cudaFreeHost(s.hin);
		cudaFreeHost(s.hout);
		if (s.done)
			cudaEventDestroy(s.done);
		if (s.t0)
			cudaEventDestroy(s.t0);
		if (s.t1)
			cudaEventDestroy(s.t1);
		if (s.stream)
			cudaStreamDestroy(s.stream);
		s = Slot();
	}
	delete c.pool;
	c.pool = nullptr;
	delete c.out;
	c.out = nullptr;
	for (int b = 0; b < DeviceCtx::NPK; ++b) {
		cudaFreeHost(c.hpk[b]);
		c.hpk[b] = nullptr;
		if (c.hpk_free[b])
			cudaEventDestroy(c.hpk_free[b]);
		c.hpk_free[b] = nullptr;
		c.hpk_busy[b] = false;
	}
	cudaSetDevice(prev);
}

int alloc_ctx(DeviceCtx &c, uint32_t n, size_t cap)
{
	DeviceGuard g;
	if (g.enter(c.dev) != QP_OK)
		return QP_ERR_NO_DEVICE;
	for (Slot &s : c.slot) {
		// one allocation per slot, carved up (all offsets are multiples of 16 bytes
		// because cap is a multiple of 32)
		const size_t fl = cap * (in_floats(n) + out_floats(n));
		float *base = nullptr;
		QP_CUDA(cudaMalloc(&base, fl * 4));
		s.dP = base;
		s.dq = s.dP + cap * n * n;
		s.dlb = s.dq + cap * n;
		s.dub = s.dlb + cap * n;
		s.dz0 = s.dub + cap * n;
		s.du0 = s.dz0 + cap * n;
		s.drho = s.du0 + cap * n;
		s.dx = s.drho + cap;
		s.dz = s.dx + cap * n;
		s.dlam = s.dz + cap * n;
		s.dstatus = reinterpret_cast<int32_t *>(s.dlam + cap * n);
		s.diters = s.dstatus + cap;
		QP_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
		QP_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
		QP_CUDA(cudaEventCreate(&s.t0));
		QP_CUDA(cudaEventCreate(&s.t1));
	}
	return QP_OK;
}

// (re)allocate every device context of ws for a slot capacity of `cap` problems.  On failure
// the workspace is left empty (cap = 0), never half built.
int resize_workspace(qp_workspace *ws, size_t cap)
{
	cap = (cap + 31) / 32 * 32;
	for (DeviceCtx &c : ws->devs)
		free_ctx(c);
	ws->cap = 0;
	for (DeviceCtx &c : ws->devs) {
		const int rc = alloc_ctx(c, ws->n, cap);
		if (rc != QP_OK) {
			const std::string keep = g_err;
			for (DeviceCtx &d : ws->devs)
				free_ctx(d);
			g_err = keep;
			return rc;
		}
	}
	ws->cap = cap;
	return QP_OK;
}

bool is_pinned(const void *p)
{
	if (!p)
		return true;
	cudaPointerAttributes at;
	if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
		cudaGetLastError();
		return false;
	}
	return at.type == cudaMemoryTypeHost;
}

// After a failure in the middle of a shard: nothing of this call may stay in flight, because
// the streams, the packing workers and the pinned buffers would go on touching the caller's
// arrays after the call has returned (and the next call on the workspace would retire the
// stale chunks into ITS caller's arrays).
void quiesce(DeviceCtx &ctx)
{
	if (ctx.pool)
		ctx.pool->wait();
	if (ctx.out)
		ctx.out->drain();
	for (Slot &s : ctx.slot) {
		if (s.stream)
			cudaStreamSynchronize(s.stream);
		s.busy = false, s.timed = false, s.copier = false;
	}
	for (bool &b : ctx.hpk_busy)
		b = false;
	cudaGetLastError();
}

// bytes (inputs + results of a call) up to which the one-copy path of run_shard_body lets the
// kernel work on the pinned staging buffers directly; QPB200_ZERO_COPY_BYTES overrides (0 = never)
size_t zero_copy_bytes()
{
	static const size_t v = [] {
		const char *e = std::getenv("QPB200_ZERO_COPY_BYTES");
		return e ? size_t(std::strtoull(e, nullptr, 10)) : size_t(QPB_ZERO_COPY_DEFAULT);
	}();
	return v;
}

// One GPU's share of a host-memory solve: problems [lo, hi) of the batch in chunks of `chunk`.
int run_shard_body(qp_workspace *ws, DeviceCtx &ctx, const qp_batch &pb, const qp_batch_solution &sol,
		   const qp_options &opt, size_t lo, size_t hi, size_t chunk)
{
	const uint32_t n = pb.n;
	const size_t cap = ws->cap;
	const bool upper_in = pb.p_layout == QP_P_UPPER; // the caller's P is packed already
	const int pk = tpq_packed_stride(int(n));
	const size_t p_per = upper_in ? size_t(pk) : size_t(n) * n; // floats of P per problem
	const size_t share = pb.p_share > 1u ? pb.p_share : 1u;     // problems per matrix

	// ---- a handful of problems: latency is what counts (the drop-in replaces a CPU call of
	// ~20 us, test/test.c:89-106).  One staging copy in, one launch, one copy out, one wait:
	// every array of the call is laid out back to back (16-byte aligned pieces) in the slot's
	// pinned buffer and in its device buffer, whatever memory the caller's pointers are in (no
	// cudaPointerGetAttributes on twelve pointers, no events, no worker threads).
	const size_t small_c = hi - lo;
	// (the pieces are rounded up to 16 bytes: a count that is a multiple of four, as a full slot
	// is, needs no rounding, any other count has two problems' worth of room to spare)
	if (small_c > 0 && (small_c + 2 <= cap || (small_c == cap && small_c % 4 == 0)) && n >= 3 &&
	    small_c * (in_floats(n) + out_floats(n)) * 4 <= (256u << 10)) {
		Slot &s = ctx.slot[0];
		if (s.busy) { // (a failed earlier call leaves nothing in flight, see quiesce)
			QP_CUDA(cudaEventSynchronize(s.done));
			s.busy = false;
		}
		if (!s.hin)
			QP_CUDA(cudaMallocHost(&s.hin, cap * in_floats(n) * 4));
		if (!s.hout)
			QP_CUDA(cudaMallocHost(&s.hout, cap * out_floats(n) * 4));
		const size_t c = small_c;
		auto up4 = [](size_t v) { return (v + 3) / 4 * 4; };
		struct Piece {
			const float *src;
			size_t per;
			size_t off;
		} in[] = { { pb.P, p_per, 0 }, { pb.q, n, 0 },  { pb.lb, n, 0 }, { pb.ub, n, 0 },
			   { pb.z0, n, 0 },    { pb.u0, n, 0 }, { pb.rho, 1, 0 } };
		// (shared matrices, qp_batch.p_share: the groups these problems belong to)
		const size_t g_first = lo / share, g_count = (hi - 1) / share - g_first + 1;
		size_t tot = 0;
		for (Piece &pc : in)
			if (pc.src) {
				const bool isP = pc.src == pb.P;
				const size_t from = isP ? g_first : lo, cnt_ = isP ? g_count : c;
				pc.off = tot;
				std::memcpy(s.hin + tot, pc.src + from * pc.per, cnt_ * pc.per * 4);
				tot += up4(cnt_ * pc.per);
			}
		cudaStream_t st = s.stream;
		// A few KB: no copy at all.  The slot's pinned buffers are mapped into the device's
		// address space (unified addressing), so the kernel reads its inputs from them over
		// PCIe and writes its results into them: one launch and one wait instead of copy,
		// launch, copy, wait -- two DMA round trips (~6 us each) less per call.
		const size_t out_fl = 3 * up4(c * n) + 2 * up4(c);
		const bool zero_copy = (tot + out_fl) * 4 <= zero_copy_bytes();
		float *din = zero_copy ? s.hin : s.dP; // else: the slot's one device allocation, inputs first
		if (!zero_copy)
			QP_CUDA(cudaMemcpyAsync(din, s.hin, tot * 4, cudaMemcpyHostToDevice, st));
		g_h2d_bytes.fetch_add(tot * 4);
		qp_batch dpb = pb;
		dpb.mem = QP_MEM_DEVICE;
		dpb.count = c;
		dpb.P = din + in[0].off, dpb.q = din + in[1].off;
		dpb.lb = pb.lb ? din + in[2].off : nullptr, dpb.ub = pb.ub ? din + in[3].off : nullptr;
		dpb.z0 = pb.z0 ? din + in[4].off : nullptr, dpb.u0 = pb.u0 ? din + in[5].off : nullptr;
		dpb.rho = pb.rho ? din + in[6].off : nullptr;
		float *dout = zero_copy ? s.hout : din + up4(tot);
		const size_t ov = up4(c * n), oi = up4(c);
		qp_batch_solution dsol;
		dsol.x = sol.x ? dout : nullptr, dsol.z = sol.z ? dout + ov : nullptr;
		dsol.lambda = sol.lambda ? dout + 2 * ov : nullptr;
		dsol.status = sol.status ? reinterpret_cast<int32_t *>(dout + 3 * ov) : nullptr;
		dsol.iters = sol.iters ? reinterpret_cast<int32_t *>(dout + 3 * ov + oi) : nullptr;
		KernelArgs ka = make_args(dpb, dsol, opt);
		ka.p_share_off = lo - g_first * share;
		if (int rc = launch_solver(ka, ctx.dev, st, opt.kernel, opt.schedule))
			return rc;
		if (!zero_copy)
			QP_CUDA(cudaMemcpyAsync(s.hout, dout, (3 * ov + 2 * oi) * 4, cudaMemcpyDeviceToHost, st));
		QP_CUDA(cudaStreamSynchronize(st));
		if (sol.x)
			std::memcpy(sol.x + lo * n, s.hout, c * n * 4);
		if (sol.z)
			std::memcpy(sol.z + lo * n, s.hout + ov, c * n * 4);
		if (sol.lambda)
			std::memcpy(sol.lambda + lo * n, s.hout + 2 * ov, c * n * 4);
		if (sol.status)
			std::memcpy(sol.status + lo, s.hout + 3 * ov, c * 4);
		if (sol.iters)
			std::memcpy(sol.iters + lo, s.hout + 3 * ov + oi, c * 4);
		return QP_OK;
	}

	const bool pin_in = is_pinned(pb.P) && is_pinned(pb.q) && is_pinned(pb.lb) &&
			    is_pinned(pb.ub) && is_pinned(pb.z0) && is_pinned(pb.u0) &&
			    is_pinned(pb.rho);
	const bool pin_out = is_pinned(sol.x) && is_pinned(sol.z) && is_pinned(sol.lambda) &&
			     is_pinned(sol.status) && is_pinned(sol.iters);

	auto retire = [&](Slot &s) -> int {
		if (!s.busy)
			return QP_OK;
		if (s.copier) {
			// (its task is at the front of the copier's queue or running: a short wait)
			while (__atomic_load_n(&s.copy_open, __ATOMIC_ACQUIRE))
				std::this_thread::sleep_for(std::chrono::microseconds(10));
			QP_CUDA(s.copy_rc);
		} else {
			QP_CUDA(cudaEventSynchronize(s.done));
		}
		if (s.timed) {
			float ms = 0.0f;
			if (cudaEventElapsedTime(&ms, s.t0, s.t1) == cudaSuccess && ms > 0.0f)
				ctx.h2d_bytes_per_s = double(s.cnt) * pk * 4.0 / (ms * 1e-3);
			s.timed = false;
		}
		if (!pin_out && !s.copier) {
			const size_t c = s.cnt, f = s.first;
			const float *h = s.hout;
			if (sol.x)
				std::memcpy(sol.x + f * n, h, c * n * 4);
			if (sol.z)
				std::memcpy(sol.z + f * n, h + cap * n, c * n * 4);
			if (sol.lambda)
				std::memcpy(sol.lambda + f * n, h + 2 * cap * n, c * n * 4);
			const int32_t *hi32 = reinterpret_cast<const int32_t *>(h + 3 * cap * n);
			if (sol.status)
				std::memcpy(sol.status + f, hi32, c * 4);
			if (sol.iters)
				std::memcpy(sol.iters + f, hi32 + cap, c * 4);
		}
		s.busy = false, s.copier = false;
		return QP_OK;
	};
	// (QPB200_OUT_COPIER=0, measurements: results copied out by this thread, plain chunks staged
	// by this thread while the pool packs -- the path as it was)
	static const bool copier_on = [] { const char *e = std::getenv("QPB200_OUT_COPIER"); return !(e && e[0] == '0'); }();
	const bool use_copier = !pin_out && copier_on;
	if (use_copier && !ctx.out)
		ctx.out = new OutCopier;

	// pack the upper triangles of P on the host when the thread-per-QP kernel will run
	const unsigned tpg = host_threads_per_gpu(unsigned(ws->devs.size()));
	const bool pack0 = !upper_in && share == 1 && !ctx.pack_off && qp_host_path_packs(n) && tpg >= 12 &&
			  (opt.kernel == QP_KERNEL_AUTO || opt.kernel == QP_KERNEL_TPQ) &&
			  ((pb.lb == nullptr) == (pb.ub == nullptr)) && ((pb.z0 == nullptr) == (pb.u0 == nullptr));
	if (pack0 && !ctx.pool)
	{
		// three quarters of the hardware threads: the path is bound by the host's memory
		// system, and the enqueueing thread and the driver's own threads want cores too
		// (16-thread host, 2^20 problems at n = 16: 4 workers 63 M, 8 65 M, 12 70 M,
		// 15 68 M QP/s end to end)
		unsigned workers = std::min(32u, tpg - tpg / 4);
		if (const char *e = std::getenv("QPB200_PACK_THREADS")) // measurements
			workers = unsigned(std::max(1, std::min(64, std::atoi(e))));
		ctx.pool = new HostPool(workers);
	}

	// Chunk scheduling.  The pool packs one chunk at a time, taken from the FRONT of the shard,
	// and goes straight on to the next one (a finished chunk waits in `ready`).  Each time a
	// slot is free this thread enqueues the waiting packed chunk if there is one, and
	// otherwise a plain chunk (full matrices) taken from the BACK of the shard: the host
	// cores and the copy engine are both kept busy, and the mix of packed and plain chunks
	// follows their relative speed (a fast host packs almost everything, a slow one almost
	// nothing).  While it waits for a slot this thread polls the pool, so that the next
	// chunk is started as soon as the previous one is packed.  Without packing the chunks
	// go front to back.
	size_t front = lo, back = hi;
	struct Job {
		size_t first = 0, cnt = 0;
		int buf = 0;
		bool valid = false;
	} job, ready; // being packed; packed and waiting for a slot
	int pbuf = 0;
	// non-blocking: starts the next chunk if there is one and a buffer is free
	auto begin_pack = [&]() -> int {
		if (!pack0 || ctx.pack_off || job.valid || front >= back)
			return QP_OK;
		if (!ctx.hpk[pbuf]) {
			QP_CUDA(cudaMallocHost(&ctx.hpk[pbuf], cap * size_t(pk) * 4));
			QP_CUDA(cudaEventCreateWithFlags(&ctx.hpk_free[pbuf], cudaEventDisableTiming));
		} else if (ctx.hpk_busy[pbuf]) {
			const cudaError_t e = cudaEventQuery(ctx.hpk_free[pbuf]);
			if (e == cudaErrorNotReady)
				return QP_OK; // the copy engine still reads it: try again later
			QP_CUDA(e);
			ctx.hpk_busy[pbuf] = false;
		}
		if (ready.valid && ready.buf == pbuf)
			return QP_OK;
		const size_t c = std::min(chunk, back - front);
		job.first = front, job.cnt = c, job.buf = pbuf, job.valid = true;
		front += c;
		pbuf = (pbuf + 1) % DeviceCtx::NPK;
		const float *src = pb.P + job.first * size_t(n) * n;
		float *dst = ctx.hpk[job.buf];
		const int nn = int(n), pkk = pk;
		ctx.pool->start([=](unsigned part, unsigned parts) {
			const size_t a = c * part / parts, b = c * (part + 1) / parts;
			pack_upper(src + a * size_t(nn) * nn, dst + a * size_t(pkk), b - a, nn);
		});
		return QP_OK;
	};
	// moves a finished chunk to `ready` and keeps the pool going
	auto pump = [&]() -> int {
		if (job.valid && !ready.valid && ctx.pool->idle()) {
			ctx.pool->wait();
			ready = job;
			job.valid = false;
			// Packing pays while it takes less time than about twice what the copy engine
			// needs for the full matrices of the same chunk; on a hopelessly slow or busy
			// host it only takes memory bandwidth away from the copies.
			if (!host_pack_forced() && ctx.h2d_bytes_per_s > 0.0 && ready.cnt == chunk) {
				const double t_full = double(chunk) * n * n * 4.0 / ctx.h2d_bytes_per_s;
				if (std::getenv("QPB200_DEBUG"))
					std::fprintf(stderr, "qpb200: pack %.3f ms, full copy %.3f ms at %.1f GB/s\n",
						     ctx.pool->last_seconds() * 1e3, t_full * 1e3,
						     ctx.h2d_bytes_per_s * 1e-9);
				// eight chunks in a row, so that one preempted worker does not decide
				ctx.slow_chunks = ctx.pool->last_seconds() > 3.0 * t_full ? ctx.slow_chunks + 1 : 0;
				if (ctx.slow_chunks >= 8) {
					ctx.pack_off = true;
					g_pack_given_up.store(true);
				}
			}
		}
		return begin_pack();
	};
	{
		int rc = begin_pack();
		if (rc != QP_OK)
			return rc;
	}

	size_t k = 0;
	while (job.valid || ready.valid || front < back) {
		Slot &s = ctx.slot[k % DEPTH];
		int rc = QP_OK;
		// wait for the slot; meanwhile keep the pool fed
		while (rc == QP_OK && s.busy && (job.valid || front < back) && pack0 && !ctx.pack_off &&
		       cudaEventQuery(s.done) == cudaErrorNotReady) {
			rc = pump();
			std::this_thread::sleep_for(std::chrono::microseconds(20));
		}
		if (rc == QP_OK)
			rc = retire(s);
		if (rc == QP_OK)
			rc = pump();
		if (rc != QP_OK)
			return rc;
		if (host_pack_test_switch() && k >= 3 && !ctx.pack_off) {
			ctx.pack_off = true;
			g_pack_given_up.store(true);
		}
		bool pack = false;
		size_t first = 0, c = 0;
		int buf = 0;
		if (ready.valid) {
			pack = true, first = ready.first, c = ready.cnt, buf = ready.buf;
			ready.valid = false;
		} else if (job.valid && front < back && (pin_in || !copier_on)) {
			// (pageable inputs: a plain chunk would be staged by THIS thread, one core's memcpy
			// of the full matrices, while the pool waits to be fed: packed chunks only)
			c = std::min(chunk, back - front); // the pool is busy: a plain chunk from the back
			first = back - c;
			back -= c;
		} else if (job.valid) {
			ctx.pool->wait(); // nothing else left to send: wait for the last packed chunk
			continue;
		} else {
			c = std::min(chunk, back - front); // no packing: front to back
			first = front;
			front += c;
		}
		s.first = first, s.cnt = c;
		cudaStream_t st = s.stream;

		// ---- inputs ----
		// (with shared matrices, qp_batch.p_share, P is copied per group: from / cnt)
		const size_t g_first = first / share, g_count = (first + c - 1) / share - g_first + 1;
		struct In {
			const float *src;
			float *dst;
			size_t per, from, cnt;
		} ins[] = { { pb.P, s.dP, p_per, g_first, g_count }, { pb.q, s.dq, n, first, c },
			    { pb.lb, s.dlb, n, first, c },           { pb.ub, s.dub, n, first, c },
			    { pb.z0, s.dz0, n, first, c },           { pb.u0, s.du0, n, first, c },
			    { pb.rho, s.drho, 1, first, c } };
		{
			size_t wire = pack ? c * size_t(pk) * 4 : 0;
			for (const In &i : ins)
				if (i.src && !(pack && i.src == pb.P))
					wire += i.cnt * i.per * 4;
			g_h2d_bytes.fetch_add(wire);
		}
		if (pack) {
			QP_CUDA(cudaEventRecord(s.t0, st));
			QP_CUDA(cudaMemcpyAsync(s.dP, ctx.hpk[buf], c * size_t(pk) * 4, cudaMemcpyHostToDevice, st));
			QP_CUDA(cudaEventRecord(s.t1, st));
			QP_CUDA(cudaEventRecord(ctx.hpk_free[buf], st));
			ctx.hpk_busy[buf] = true;
			s.timed = true;
		}
		if (pin_in) {
			for (const In &i : ins)
				if (i.src && !(pack && i.src == pb.P))
					QP_CUDA(cudaMemcpyAsync(i.dst, i.src + i.from * i.per, i.cnt * i.per * 4,
								cudaMemcpyHostToDevice, st));
		} else {
			if (!s.hin)
				QP_CUDA(cudaMallocHost(&s.hin, cap * in_floats(n) * 4));
			size_t off = 0;
			for (const In &i : ins) {
				if (i.src && !(pack && i.src == pb.P)) {
					std::memcpy(s.hin + off, i.src + i.from * i.per, i.cnt * i.per * 4);
					QP_CUDA(cudaMemcpyAsync(i.dst, s.hin + off, i.cnt * i.per * 4,
								cudaMemcpyHostToDevice, st));
				}
				off += cap * (i.src == pb.P ? size_t(n) * n : i.per);
			}
		}

		// ---- solve ----
		qp_batch dpb = pb;
		dpb.mem = QP_MEM_DEVICE;
		dpb.count = c;
		dpb.P = s.dP, dpb.q = s.dq;
		dpb.lb = pb.lb ? s.dlb : nullptr, dpb.ub = pb.ub ? s.dub : nullptr;
		dpb.z0 = pb.z0 ? s.dz0 : nullptr, dpb.u0 = pb.u0 ? s.du0 : nullptr;
		dpb.rho = pb.rho ? s.drho : nullptr;
		dpb.p_layout = (pack || upper_in) ? QP_P_UPPER : QP_P_DENSE;
		qp_batch_solution dsol;
		dsol.x = sol.x ? s.dx : nullptr, dsol.z = sol.z ? s.dz : nullptr;
		dsol.lambda = sol.lambda ? s.dlam : nullptr;
		dsol.status = sol.status ? s.dstatus : nullptr;
		dsol.iters = sol.iters ? s.diters : nullptr;
		KernelArgs ka = make_args(dpb, dsol, opt);
		ka.p_share_off = first - g_first * share;
		rc = launch_solver(ka, ctx.dev, st, pack ? uint32_t(QP_KERNEL_TPQ) : opt.kernel, opt.schedule);
		if (rc != QP_OK)
			return rc;

		// ---- outputs ----
		if (pin_out) {
			if (sol.x)
				QP_CUDA(cudaMemcpyAsync(sol.x + first * n, s.dx, c * n * 4,
							cudaMemcpyDeviceToHost, st));
			if (sol.z)
				QP_CUDA(cudaMemcpyAsync(sol.z + first * n, s.dz, c * n * 4,
							cudaMemcpyDeviceToHost, st));
			if (sol.lambda)
				QP_CUDA(cudaMemcpyAsync(sol.lambda + first * n, s.dlam, c * n * 4,
							cudaMemcpyDeviceToHost, st));
			if (sol.status)
				QP_CUDA(cudaMemcpyAsync(sol.status + first, s.dstatus, c * 4,
							cudaMemcpyDeviceToHost, st));
			if (sol.iters)
				QP_CUDA(cudaMemcpyAsync(sol.iters + first, s.diters, c * 4,
							cudaMemcpyDeviceToHost, st));
		} else {
			if (!s.hout)
				QP_CUDA(cudaMallocHost(&s.hout, cap * out_floats(n) * 4));
			float *h = s.hout;
			if (sol.x)
				QP_CUDA(cudaMemcpyAsync(h, s.dx, c * n * 4, cudaMemcpyDeviceToHost, st));
			if (sol.z)
				QP_CUDA(cudaMemcpyAsync(h + cap * n, s.dz, c * n * 4,
							cudaMemcpyDeviceToHost, st));
			if (sol.lambda)
				QP_CUDA(cudaMemcpyAsync(h + 2 * cap * n, s.dlam, c * n * 4,
							cudaMemcpyDeviceToHost, st));
			int32_t *hi32 = reinterpret_cast<int32_t *>(h + 3 * cap * n);
			if (sol.status)
				QP_CUDA(cudaMemcpyAsync(hi32, s.dstatus, c * 4, cudaMemcpyDeviceToHost,
							st));
			if (sol.iters)
				QP_CUDA(cudaMemcpyAsync(hi32 + cap, s.diters, c * 4,
							cudaMemcpyDeviceToHost, st));
		}
		QP_CUDA(cudaEventRecord(s.done, st));
		s.busy = true;
		if (use_copier) {
			s.copier = true;
			__atomic_store_n(&s.copy_open, 1, __ATOMIC_RELAXED);
			Slot *sp = &s;
			const qp_batch_solution so = sol;
			const size_t f = first, cc = c, nn = n, cp = cap;
			const int dev = ctx.dev;
			ctx.out->push([sp, so, f, cc, nn, cp, dev] {
				cudaSetDevice(dev);
				sp->copy_rc = cudaEventSynchronize(sp->done);
				if (sp->copy_rc == cudaSuccess) {
					const float *h = sp->hout;
					if (so.x)
						std::memcpy(so.x + f * nn, h, cc * nn * 4);
					if (so.z)
						std::memcpy(so.z + f * nn, h + cp * nn, cc * nn * 4);
					if (so.lambda)
						std::memcpy(so.lambda + f * nn, h + 2 * cp * nn, cc * nn * 4);
					const int32_t *hi32 = reinterpret_cast<const int32_t *>(h + 3 * cp * nn);
					if (so.status)
						std::memcpy(so.status + f, hi32, cc * 4);
					if (so.iters)
						std::memcpy(so.iters + f, hi32 + cp, cc * 4);
				}
				__atomic_store_n(&sp->copy_open, 0, __ATOMIC_RELEASE);
			});
		}
		++k;
	}
	// drain in issue order
	for (size_t d = 0; d < DEPTH; ++d) {
		int rc = retire(ctx.slot[(k + d) % DEPTH]);
		if (rc != QP_OK)
			return rc;
	}
	return QP_OK;
}

int run_shard(qp_workspace *ws, DeviceCtx &ctx, const qp_batch &pb, const qp_batch_solution &sol,
	      const qp_options &opt, size_t lo, size_t hi, size_t chunk)
{
	DeviceGuard g;
	if (g.enter(ctx.dev) != QP_OK)
		return QP_ERR_NO_DEVICE;
	const int rc = run_shard_body(ws, ctx, pb, sol, opt, lo, hi, chunk);
	if (rc != QP_OK) {
		const std::string keep = g_err;
		quiesce(ctx);
		g_err = keep;
	}
	return rc;
}

// chunk size of a call: the caller's, else ~32 MB of input, never more than the slots hold
size_t call_chunk(const qp_workspace *ws, const qp_batch &pb, const qp_options &opt)
{
	const size_t nd = std::max<size_t>(1, ws->devs.size());
	size_t c = opt.chunk ? (size_t(opt.chunk) + 31) / 32 * 32
			     : choose_chunk(pb.n, (pb.count + nd - 1) / nd);
	return std::min(c, ws->cap);
}

// the solve itself; ws->mu is held by the caller
int solve_host_locked(qp_workspace *ws, const qp_batch &pb, const qp_batch_solution &sol,
		      const qp_options &opt)
{
	if (ws->cap == 0)
		return QP_ERR_INVALID;
	const size_t chunk = call_chunk(ws, pb, opt);
	const size_t nd = ws->devs.size();
	if (nd == 1)
		return run_shard(ws, ws->devs[0], pb, sol, opt, 0, pb.count, chunk);

	// contiguous shards, remainder to the first devices: one host thread per
	// GPU, results land in the caller's arrays at the shard's offset (the "host-side gather")
	std::vector<int> rcs(nd, QP_OK);
	std::vector<std::string> errs(nd);
	std::vector<std::thread> th;
	const size_t base = pb.count / nd, rem = pb.count % nd;
	size_t lo = 0;
	for (size_t i = 0; i < nd; ++i) {
		const size_t hi = lo + base + (i < rem ? 1 : 0);
		if (hi > lo)
			th.emplace_back([&, i, lo, hi] {
				rcs[i] = run_shard(ws, ws->devs[i], pb, sol, opt, lo, hi, chunk);
				errs[i] = g_err;
			});
		lo = hi;
	}
	for (std::thread &t : th)
		t.join();
	for (size_t i = 0; i < nd; ++i)
		if (rcs[i] != QP_OK) {
			g_err = errs[i];
			return rcs[i];
		}
	return QP_OK;
}

int resolve_options(const qp_options *opt_in, qp_options &opt)
{
	if (opt_in)
		opt = *opt_in;
	else
		qp_options_default(&opt);
	return check_options(opt);
}

} // namespace

extern "C" int qp_workspace_create(uint32_t n, size_t max_count, uint32_t device_mask,
				   qp_workspace **out)
{
	if (!out || n == 0)
		return QP_ERR_INVALID;
	if (n > QP_MAX_DIM_BOX)
		return QP_ERR_UNSUPPORTED;
	std::vector<int> devs = mask_devices(device_mask, resolve_device(-1));
	if (devs.empty())
		return QP_ERR_NO_DEVICE;
	qp_workspace *ws = new qp_workspace;
	ws->n = n;
	ws->devs.resize(devs.size());
	for (size_t i = 0; i < devs.size(); ++i)
		ws->devs[i].dev = devs[i];
	const size_t per_dev = (std::max<size_t>(max_count, 1) + devs.size() - 1) / devs.size();
	const int rc = resize_workspace(ws, choose_chunk(n, per_dev));
	if (rc != QP_OK) {
		delete ws;
		return rc;
	}
	*out = ws;
	return QP_OK;
}

extern "C" void qp_workspace_destroy(qp_workspace *ws)
{
	if (!ws)
		return;
	{
		std::lock_guard<std::mutex> lk(ws->mu); // a solve still running on it finishes first
		for (DeviceCtx &c : ws->devs)
			free_ctx(c);
	}
	delete ws;
}

extern "C" int qp_solve_batch_ws(qp_workspace *ws, const qp_batch *pb,
				 const qp_batch_solution *sol, const qp_options *opt_in)
{
	int rc = check_problem(pb, sol);
	if (rc != QP_OK)
		return rc;
	qp_options opt;
	rc = resolve_options(opt_in, opt);
	if (rc != QP_OK)
		return rc;
	if (opt.precision != QP_PRECISION_F32) // binary64: the qf entry points only
		return QP_ERR_UNSUPPORTED;
	if (pb->count == 0)
		return QP_OK;

	if (pb->mem == QP_MEM_DEVICE) {
		const int dev = resolve_device(opt.device);
		if (dev < 0)
			return QP_ERR_NO_DEVICE;
		DeviceGuard g;
		if (g.enter(dev) != QP_OK)
			return QP_ERR_NO_DEVICE;
		return launch_solver(make_args(*pb, *sol, opt), dev, static_cast<cudaStream_t>(opt.stream),
				     opt.kernel, opt.schedule);
	}

	if (!ws || ws->n != pb->n)
		return QP_ERR_INVALID;
	std::lock_guard<std::mutex> lk(ws->mu);
	// an explicit chunk larger than the slots: grow (the capacity never shrinks)
	const size_t want = opt.chunk ? (size_t(opt.chunk) + 31) / 32 * 32 : 0;
	if (want > ws->cap) {
		rc = resize_workspace(ws, want);
		if (rc != QP_OK)
			return rc;
	}
	return solve_host_locked(ws, *pb, *sol, opt);
}

// Default workspaces of qp_solve_batch(): a small cache keyed by (n, device set, device).
// An entry is handed out as a shared_ptr and its mutex is held for the whole solve, so a
// second host thread with another shape gets another entry and a thread with the same shape
// waits its turn -- nobody's streams and pinned buffers are freed under a running solve.  The
// slot capacity of an entry only grows: alternating batch sizes reuse the same buffers.
namespace {

struct WsDeleter {
	void operator()(qp_workspace *ws) const { qp_workspace_destroy(ws); }
};
std::vector<std::shared_ptr<qp_workspace>> g_cache;
std::mutex g_cache_mu;
unsigned long g_cache_clock = 0;
constexpr size_t CACHE_ENTRIES = 6;

int default_workspace(uint32_t n, uint32_t mask, int dev, std::shared_ptr<qp_workspace> &out)
{
	std::lock_guard<std::mutex> lk(g_cache_mu);
	for (auto &e : g_cache)
		if (e->n == n && e->key_mask == mask && e->key_dev == dev) {
			e->stamp = ++g_cache_clock;
			out = e;
			return QP_OK;
		}
	std::vector<int> devs = mask_devices(mask, dev);
	if (devs.empty())
		return QP_ERR_NO_DEVICE;
	if (g_cache.size() >= CACHE_ENTRIES) {
		// drop the least recently used entry nobody holds (its owner frees it otherwise)
		size_t victim = g_cache.size();
		for (size_t i = 0; i < g_cache.size(); ++i)
			if (g_cache[i].use_count() == 1 &&
			    (victim == g_cache.size() || g_cache[i]->stamp < g_cache[victim]->stamp))
				victim = i;
		if (victim < g_cache.size())
			g_cache.erase(g_cache.begin() + victim);
	}
	std::shared_ptr<qp_workspace> ws(new qp_workspace, WsDeleter());
	ws->n = n, ws->key_mask = mask, ws->key_dev = dev, ws->stamp = ++g_cache_clock;
	ws->devs.resize(devs.size());
	for (size_t i = 0; i < devs.size(); ++i)
		ws->devs[i].dev = devs[i];
	g_cache.push_back(ws); // empty (cap = 0): sized under its own mutex by the first solve
	out = ws;
	return QP_OK;
}

} // namespace

extern "C" int qp_solve_batch(const qp_batch *pb, const qp_batch_solution *sol,
			      const qp_options *opt_in)
{
	int rc = check_problem(pb, sol);
	if (rc != QP_OK)
		return rc;
	if (pb->mem == QP_MEM_DEVICE || pb->count == 0)
		return qp_solve_batch_ws(nullptr, pb, sol, opt_in);

	qp_options opt;
	rc = resolve_options(opt_in, opt);
	if (rc != QP_OK)
		return rc;
	if (opt.precision != QP_PRECISION_F32)
		return QP_ERR_UNSUPPORTED;
	const int dev = resolve_device(opt.device);
	if (dev < 0)
		return QP_ERR_NO_DEVICE;
	std::shared_ptr<qp_workspace> ws;
	rc = default_workspace(pb->n, opt.device_mask, dev, ws);
	if (rc != QP_OK)
		return rc;
	std::lock_guard<std::mutex> lk(ws->mu);
	const size_t nd = ws->devs.size();
	const size_t want = opt.chunk ? (size_t(opt.chunk) + 31) / 32 * 32
				      : choose_chunk(pb->n, (pb->count + nd - 1) / nd);
	if (want > ws->cap) {
		rc = resize_workspace(ws.get(), want);
		if (rc != QP_OK)
			return rc;
	}
	return solve_host_locked(ws.get(), *pb, *sol, opt);
}

// ----------------------------------------------------------------------------------------
// objective evaluation (quadratic_form_eval, /root/reference/qp_solvers/qp.c:9-27)
// ----------------------------------------------------------------------------------------
extern "C" int qp_eval_batch(const qp_batch *pb, const float *x, float r, float *f,
			     const qp_options *opt_in)
{
	if (!pb || !x || !f || !pb->P || !pb->q || pb->n == 0)
		return QP_ERR_INVALID;
	if (pb->p_share > 1u) // shared matrices: qp_solve_batch() only
		return QP_ERR_UNSUPPORTED;
	if (pb->count == 0)
		return QP_OK;
	qp_options opt;
	if (opt_in)
		opt = *opt_in;
	else
		qp_options_default(&opt);
	const int dev = resolve_device(opt.device);
	if (dev < 0)
		return QP_ERR_NO_DEVICE;
	DeviceGuard g;
	if (g.enter(dev) != QP_OK)
		return QP_ERR_NO_DEVICE;
	const size_t n = pb->n, c = pb->count;
	if (pb->mem == QP_MEM_DEVICE) {
		return launch_eval(pb->P, pb->q, x, r, f, int(n), c, dev, static_cast<cudaStream_t>(opt.stream));
	}
	float *d = nullptr;
	QP_CUDA(cudaMalloc(&d, (c * n * n + 2 * c * n + c) * 4));
	float *dP = d, *dq = dP + c * n * n, *dx = dq + c * n, *df = dx + c * n;
	cudaError_t e = cudaMemcpy(dP, pb->P, c * n * n * 4, cudaMemcpyHostToDevice);
	if (e == cudaSuccess)
		e = cudaMemcpy(dq, pb->q, c * n * 4, cudaMemcpyHostToDevice);
	if (e == cudaSuccess)
		e = cudaMemcpy(dx, x, c * n * 4, cudaMemcpyHostToDevice);
	int lrc = QP_OK;
	if (e == cudaSuccess)
		lrc = launch_eval(dP, dq, dx, r, df, int(n), c, dev, nullptr);
	if (e == cudaSuccess && lrc == QP_OK)
		e = cudaMemcpy(f, df, c * 4, cudaMemcpyDeviceToHost);
	cudaFree(d);
	if (lrc != QP_OK)
		return lrc;
	if (e != cudaSuccess) {
		g_err = std::string("qp_eval_batch: ") + cudaGetErrorString(e);
		return QP_ERR_CUDA;
	}
	return QP_OK;
}

// ----------------------------------------------------------------------------------------
// batched Newton (newton_method_with_line_search, /root/reference/qp_solvers/qp_solvers.c:103-144)
// ----------------------------------------------------------------------------------------
static int descent_batch(int gradient, const qp_batch *pb, const float *x0, unsigned iterations,
			 float *x, int32_t *status, int32_t *steps, float *f, float r,
			 const qp_options *opt_in)
{
	if (!pb || !x || pb->n == 0)
		return QP_ERR_INVALID;
	if (pb->n > QP_MAX_DIM)
		return QP_ERR_UNSUPPORTED;
	if (pb->mem != QP_MEM_HOST && pb->mem != QP_MEM_DEVICE)
		return QP_ERR_INVALID;
	if (pb->p_share > 1u) // shared matrices: qp_solve_batch() only
		return QP_ERR_UNSUPPORTED;
	if (pb->count == 0)
		return QP_OK;
	if (!pb->P || !pb->q)
		return QP_ERR_INVALID;
	qp_options opt;
	if (opt_in)
		opt = *opt_in;
	else
		qp_options_default(&opt);
	const int dev = resolve_device(opt.device);
	if (dev < 0)
		return QP_ERR_NO_DEVICE;
	DeviceGuard g;
	if (g.enter(dev) != QP_OK)
		return QP_ERR_NO_DEVICE;
	const size_t n = pb->n, c = pb->count;
	NewtonArgs a = {};
	a.count = c, a.n = int(n), a.iterations = iterations, a.r = r, a.gradient = gradient;
	if (pb->mem == QP_MEM_DEVICE) {
		a.P = pb->P, a.q = pb->q, a.x0 = x0, a.x = x, a.fval = f, a.status = status, a.steps = steps;
		return launch_newton(a, dev, static_cast<cudaStream_t>(opt.stream));
	}
	// host pointers: one staging allocation, synchronous (this is the cross-check path)
	float *d = nullptr;
	const size_t fl = c * n * n + 3 * c * n + 3 * c;
	QP_CUDA(cudaMalloc(&d, fl * 4));
	float *dP = d, *dq = dP + c * n * n, *dx0 = dq + c * n, *dx = dx0 + c * n, *df = dx + c * n;
	int32_t *dst = reinterpret_cast<int32_t *>(df + c), *dsp = dst + c;
	cudaError_t e = cudaMemcpy(dP, pb->P, c * n * n * 4, cudaMemcpyHostToDevice);
	if (e == cudaSuccess)
		e = cudaMemcpy(dq, pb->q, c * n * 4, cudaMemcpyHostToDevice);
	if (e == cudaSuccess && x0)
		e = cudaMemcpy(dx0, x0, c * n * 4, cudaMemcpyHostToDevice);
	int rc = QP_OK;
	if (e == cudaSuccess) {
		a.P = dP, a.q = dq, a.x0 = x0 ? dx0 : nullptr, a.x = dx, a.fval = f ? df : nullptr;
		a.status = dst, a.steps = dsp;
		rc = launch_newton(a, dev, nullptr);
		if (rc == QP_OK)
			e = cudaMemcpy(x, dx, c * n * 4, cudaMemcpyDeviceToHost);
		if (rc == QP_OK && e == cudaSuccess && f)
			e = cudaMemcpy(f, df, c * 4, cudaMemcpyDeviceToHost);
		if (rc == QP_OK && e == cudaSuccess && status)
			e = cudaMemcpy(status, dst, c * 4, cudaMemcpyDeviceToHost);
		if (rc == QP_OK && e == cudaSuccess && steps)
			e = cudaMemcpy(steps, dsp, c * 4, cudaMemcpyDeviceToHost);
	}
	cudaFree(d);
	if (rc != QP_OK)
		return rc;
	if (e != cudaSuccess) {
		g_err = std::string(gradient ? "qp_gradient_batch: " : "qp_newton_batch: ") + cudaGetErrorString(e);
		return QP_ERR_CUDA;
	}
	return QP_OK;
}

extern "C" int qp_newton_batch(const qp_batch *pb, const float *x0, unsigned iterations, float *x,
			       int32_t *status, int32_t *steps, float *f, float r,
			       const qp_options *opt)
{
	return descent_batch(0, pb, x0, iterations, x, status, steps, f, r, opt);
}

extern "C" int qp_gradient_batch(const qp_batch *pb, const float *x0, unsigned iterations, float *x,
				 int32_t *status, int32_t *steps, float *f, float r,
				 const qp_options *opt)
{
	return descent_batch(1, pb, x0, iterations, x, status, steps, f, r, opt);
}

// ----------------------------------------------------------------------------------------
// general rows l <= A x <= u (qp_gen.cuh)
// ----------------------------------------------------------------------------------------
// device staging of qp_solve_general_batch() with host pointers: two buffers and two streams per
// device, kept for the life of the process (they only grow; 2 x ~32 MB unless one problem is larger)
namespace {
struct GenStage {
	std::mutex mu;
	float *d[2] = { nullptr, nullptr };
	size_t floats = 0; // capacity of each buffer
	cudaStream_t st[2] = { nullptr, nullptr };
};
GenStage g_gen_stage[qpb::MAX_DEVICES];
} // namespace

extern "C" int qp_solve_general_batch(const qp_general_batch *pb, const qp_general_solution *sol,
				      const qp_options *opt_in)
{
	if (!pb || !sol || pb->n == 0 || pb->m == 0)
		return QP_ERR_INVALID;
	if (pb->n > QP_MAX_DIM || pb->m > QP_MAX_ROWS)
		return QP_ERR_UNSUPPORTED;
	if (pb->mem != QP_MEM_HOST && pb->mem != QP_MEM_DEVICE)
		return QP_ERR_INVALID;
	if ((pb->z0 == nullptr) != (pb->u0 == nullptr))
		return QP_ERR_INVALID;
	qp_options opt;
	if (opt_in)
		opt = *opt_in;
	else
		qp_options_default(&opt);
	int rc = check_options(opt);
	if (rc != QP_OK)
		return rc;
	if (pb->count == 0)
		return QP_OK;
	if (!pb->P || !pb->q || !pb->A || !pb->l || !pb->u || !sol->x)
		return QP_ERR_INVALID;
	const int dev = resolve_device(opt.device);
	if (dev < 0)
		return QP_ERR_NO_DEVICE;
	DeviceGuard g;
	if (g.enter(dev) != QP_OK)
		return QP_ERR_NO_DEVICE;
	const size_t n = pb->n, m = pb->m, c = pb->count;
	GenArgs a = {};
	a.count = c, a.n = int(n), a.m = int(m);
	a.rho = opt.rho, a.alpha = opt.alpha, a.reltol = opt.reltol, a.max_iter = opt.max_iter;
	a.eps_abs_pri = float(std::sqrt(double(m)) * double(opt.abstol));
	a.eps_abs_dual = float(std::sqrt(double(n)) * double(opt.abstol));
	switch (opt.kernel) {
	case QP_KERNEL_AUTO: a.force = 0; break;
	case QP_KERNEL_GEN_SMEM: a.force = 1; break;
	case QP_KERNEL_GEN_REG: a.force = 2; break;
	default: return QP_ERR_UNSUPPORTED; // a box kernel family
	}
	if (pb->mem == QP_MEM_DEVICE) {
		a.P = pb->P, a.q = pb->q, a.A = pb->A, a.l = pb->l, a.u = pb->u, a.z0 = pb->z0, a.u0 = pb->u0;
		a.x = sol->x, a.z = sol->z, a.y = sol->y, a.status = sol->status, a.iters = sol->iters;
		return launch_gen(a, dev, static_cast<cudaStream_t>(opt.stream));
	}
	// host pointers: chunks of ~32 MB alternate between two device buffers and two streams, so
	// the copy-in of a chunk overlaps the solve and the copy-out of the one before (copies on one
	// stream are ordered: a buffer is not overwritten before its previous results have left).  The
	// buffers are kept per device and only grow; calls on one device take turns.
	GenStage &gs = g_gen_stage[dev];
	std::lock_guard<std::mutex> lock(gs.mu);
	const size_t per = n * n + n + m * n + 4 * m + n + 2 * m + 2; // floats per problem, in and out
	size_t capc = std::min<size_t>(c, std::max<size_t>(4, (size_t(32) << 20) / 4 / per));
	capc = (capc + 3) / 4 * 4; // every piece of a buffer stays 16-byte aligned
	if (gs.floats < capc * per) {
		for (float *&p : gs.d) {
			cudaFree(p);
			p = nullptr;
		}
		gs.floats = 0;
		for (float *&p : gs.d)
			QP_CUDA(cudaMalloc(&p, capc * per * 4));
		gs.floats = capc * per;
	}
	for (cudaStream_t &st : gs.st)
		if (!st)
			QP_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
	cudaError_t e = cudaSuccess;
	rc = QP_OK;
	size_t k = 0;
	for (size_t f = 0; f < c && e == cudaSuccess && rc == QP_OK; f += capc, ++k) {
		const size_t cc = std::min(capc, c - f);
		cudaStream_t st = gs.st[k & 1];
		float *d = gs.d[k & 1];
		float *dP = d, *dq = dP + capc * n * n, *dA = dq + capc * n, *dl = dA + capc * m * n,
		      *du = dl + capc * m, *dz0 = du + capc * m, *du0 = dz0 + capc * m, *dx = du0 + capc * m,
		      *dz = dx + capc * n, *dy = dz + capc * m;
		int32_t *dst = reinterpret_cast<int32_t *>(dy + capc * m), *dit = dst + capc;
		struct Cp {
			void *dst;
			const void *src;
			size_t bytes;
		} ins[] = { { dP, pb->P + f * n * n, cc * n * n * 4 },
			    { dq, pb->q + f * n, cc * n * 4 },
			    { dA, pb->A + f * m * n, cc * m * n * 4 },
			    { dl, pb->l + f * m, cc * m * 4 },
			    { du, pb->u + f * m, cc * m * 4 },
			    { dz0, pb->z0 ? pb->z0 + f * m : nullptr, cc * m * 4 },
			    { du0, pb->u0 ? pb->u0 + f * m : nullptr, cc * m * 4 } };
		for (const Cp &cp : ins)
			if (cp.src && e == cudaSuccess)
				e = cudaMemcpyAsync(cp.dst, cp.src, cp.bytes, cudaMemcpyHostToDevice, st);
		if (e != cudaSuccess)
			break;
		GenArgs ac = a;
		ac.count = cc;
		ac.P = dP, ac.q = dq, ac.A = dA, ac.l = dl, ac.u = du;
		ac.z0 = pb->z0 ? dz0 : nullptr, ac.u0 = pb->u0 ? du0 : nullptr;
		ac.x = dx, ac.z = dz, ac.y = dy, ac.status = dst, ac.iters = dit;
		rc = launch_gen(ac, dev, st);
		if (rc != QP_OK)
			break;
		Cp outs[] = { { sol->x + f * n, dx, cc * n * 4 },
			      { sol->z ? sol->z + f * m : nullptr, dz, cc * m * 4 },
			      { sol->y ? sol->y + f * m : nullptr, dy, cc * m * 4 },
			      { sol->status ? sol->status + f : nullptr, dst, cc * 4 },
			      { sol->iters ? sol->iters + f : nullptr, dit, cc * 4 } };
		for (const Cp &cp : outs)
			if (cp.dst && e == cudaSuccess)
				e = cudaMemcpyAsync(cp.dst, cp.src, cp.bytes, cudaMemcpyDeviceToHost, st);
	}
	for (cudaStream_t st : gs.st) {
		const cudaError_t es = cudaStreamSynchronize(st);
		if (e == cudaSuccess)
			e = es;
	}
	if (rc != QP_OK)
		return rc;
	if (e != cudaSuccess) {
		g_err = std::string("qp_solve_general_batch: ") + cudaGetErrorString(e);
		cudaGetLastError();
		return QP_ERR_CUDA;
	}
	return QP_OK;
}

// ----------------------------------------------------------------------------------------
// entry points on the reference's structs
// ----------------------------------------------------------------------------------------
static inline unsigned ref_rows(const qp_ref_matrix *m)
{
	return (m->dimensions >> 16) & 0xffffu; // MATRIX_GET_ROW, matrix_type.h:14
}
static inline unsigned ref_cols(const qp_ref_matrix *m)
{
	return m->dimensions & 0xffffu; // MATRIX_GET_COL, matrix_type.h:15
}

extern "C" int qp_admm_qf_batch(const struct qp_ref_quadratic_form *const *qfs, size_t count,
				unsigned iterations, double lb, double ub, double *const *x_out,
				int32_t *status, int32_t *iters, const qp_options *opt_in)
{
	if (!qfs || !x_out)
		return QP_ERR_INVALID;
	if (count == 0)
		return QP_OK;
	for (size_t i = 0; i < count; ++i)
		if (!qfs[i] || !qfs[i]->p || !qfs[i]->q || !qfs[i]->p->elements ||
		    !qfs[i]->q->elements || !x_out[i])
			return QP_ERR_INVALID;
	const unsigned n = ref_rows(qfs[0]->p);
	if (n == 0)
		return QP_ERR_INVALID;
	if (n > QP_MAX_DIM_BOX)
		return QP_ERR_UNSUPPORTED;
	for (size_t i = 0; i < count; ++i)
		if (ref_rows(qfs[i]->p) != n || ref_cols(qfs[i]->p) != n ||
		    ref_rows(qfs[i]->q) * ref_cols(qfs[i]->q) != n)
			return QP_ERR_INVALID;

	if (opt_in && opt_in->precision == QP_PRECISION_F64) {
		// ---- the reference's own arithmetic (qp_f64.cuh): doubles in, doubles out ------------
		if (n > QP_MAX_DIM)
			return QP_ERR_UNSUPPORTED;
		qp_options opt = *opt_in;
		if (int rc = check_options(opt))
			return rc;
		const int dev = resolve_device(opt.device);
		if (dev < 0)
			return QP_ERR_NO_DEVICE;
		DeviceGuard g;
		if (g.enter(dev) != QP_OK)
			return QP_ERR_NO_DEVICE;
		const size_t nn = size_t(n) * n;
		std::vector<double> hP(count * nn), hq(count * n), hx(count * n);
		std::vector<int32_t> hs(count), hi(count);
		for (size_t i = 0; i < count; ++i) {
			std::memcpy(hP.data() + i * nn, qfs[i]->p->elements, nn * 8);
			std::memcpy(hq.data() + i * n, qfs[i]->q->elements, size_t(n) * 8);
		}
		char *d = nullptr;
		const size_t oP = 0, oq = oP + count * nn * 8, ox = oq + count * n * 8, os = ox + count * n * 8,
			     oi = os + up256(count * 4), tot = oi + up256(count * 4);
		QP_CUDA(cudaMalloc(&d, tot));
		auto fail = [&](cudaError_t e, const char *what) {
			g_err = std::string(what) + ": " + cudaGetErrorString(e);
			cudaFree(d);
			return QP_ERR_CUDA;
		};
		cudaStream_t st = static_cast<cudaStream_t>(opt.stream);
		cudaError_t e;
		if ((e = cudaMemcpyAsync(d + oP, hP.data(), count * nn * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess)
			return fail(e, "cudaMemcpyAsync");
		if ((e = cudaMemcpyAsync(d + oq, hq.data(), count * n * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess)
			return fail(e, "cudaMemcpyAsync");
		Admm64Args a = {};
		a.P = reinterpret_cast<const double *>(d + oP), a.q = reinterpret_cast<const double *>(d + oq);
		a.lb = lb, a.ub = ub;
		a.x = reinterpret_cast<double *>(d + ox);
		a.status = reinterpret_cast<int32_t *>(d + os), a.iters = reinterpret_cast<int32_t *>(d + oi);
		a.count = count, a.n = int(n);
		a.rho = double(opt.rho), a.alpha = double(opt.alpha);
		a.eps_abs = std::sqrt(double(n)) * double(opt.abstol), a.reltol = double(opt.reltol);
		a.max_iter = iterations;
		if (int rc = launch_admm64(a, dev, st)) {
			cudaFree(d);
			return rc;
		}
		if ((e = cudaMemcpyAsync(hx.data(), d + ox, count * n * 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
		    (e = cudaMemcpyAsync(hs.data(), d + os, count * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
		    (e = cudaMemcpyAsync(hi.data(), d + oi, count * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
		    (e = cudaStreamSynchronize(st)) != cudaSuccess)
			return fail(e, "copy back");
		cudaFree(d);
		for (size_t i = 0; i < count; ++i) {
			std::memcpy(x_out[i], hx.data() + i * n, size_t(n) * 8);
			if (status)
				status[i] = hs[i];
			if (iters)
				iters[i] = hi[i];
		}
		return QP_OK;
	}

	// A host that solves for many right-hand sides with one model hands over the SAME qf->p
	// in runs of consecutive structs (struct _quadratic_form keeps p and q apart,
	// qp_solvers/qp.h:8-13): when all runs have the same length k > 1 -- the last one may be
	// shorter -- each matrix is converted and sent once (qp_batch.p_share = k)
	size_t share = 1;
	{
		size_t run = 1;
		while (run < count && qfs[run]->p == qfs[0]->p)
			++run;
		bool uniform = run > 1 && run <= 0xffffffffull;
		for (size_t i = 0; uniform && i < count; ++i)
			uniform = qfs[i]->p == qfs[i / run * run]->p;
		if (uniform)
			share = run;
	}
	const size_t nP = (count + share - 1) / share;
	std::vector<float> P(nP * n * n), q(count * n), x(count * n);
	std::vector<int32_t> st(count), it(count);
	// gather + binary64 -> binary32, split over the host cores for large batches (the
	// structs are scattered in the caller's pools, so this cannot be a device-side copy)
	auto for_ranges = [&](auto &&body) {
		const size_t hw = std::max(1u, std::thread::hardware_concurrency());
		const size_t nt = std::min<size_t>(std::min<size_t>(hw, 32), std::max<size_t>(1, count / 4096));
		if (nt <= 1) {
			body(size_t(0), count);
			return;
		}
		std::vector<std::thread> th;
		for (size_t t = 0; t < nt; ++t)
			th.emplace_back(body, count * t / nt, count * (t + 1) / nt);
		for (std::thread &t : th)
			t.join();
	};
	for_ranges([&](size_t lo, size_t hi) {
		for (size_t i = lo; i < hi; ++i) {
			const double *pe = qfs[i]->p->elements, *qe = qfs[i]->q->elements;
			if (i % share == 0)
				for (size_t k = 0; k < size_t(n) * n; ++k)
					P[i / share * n * n + k] = float(pe[k]);
			for (size_t k = 0; k < n; ++k)
				q[i * n + k] = float(qe[k]);
		}
	});
	qp_batch pb;
	std::memset(&pb, 0, sizeof(pb));
	pb.n = n, pb.mem = QP_MEM_HOST, pb.count = count;
	pb.P = P.data(), pb.q = q.data();
	pb.p_share = uint32_t(share);
	pb.lb_scalar = float(lb), pb.ub_scalar = float(ub);
	qp_batch_solution sol;
	std::memset(&sol, 0, sizeof(sol));
	sol.x = x.data(), sol.status = st.data(), sol.iters = it.data();
	qp_options opt;
	if (opt_in)
		opt = *opt_in;
	else
		qp_options_default(&opt);
	opt.max_iter = iterations;
	const int rc = qp_solve_batch(&pb, &sol, &opt);
	if (rc != QP_OK)
		return rc;
	for_ranges([&](size_t lo, size_t hi) {
		for (size_t i = lo; i < hi; ++i) {
			for (size_t k = 0; k < n; ++k)
				x_out[i][k] = double(x[i * n + k]);
			if (status)
				status[i] = st[i];
			if (iters)
				iters[i] = it[i];
		}
	});
	return QP_OK;
}

extern "C" int qp_admm_qf(const struct qp_ref_quadratic_form *qf, unsigned iterations, double lb,
			  double ub, double *x_out, int32_t *status, int32_t *iters)
{
	const struct qp_ref_quadratic_form *one[1] = { qf };
	double *xo[1] = { x_out };
	return qp_admm_qf_batch(one, 1, iterations, lb, ub, xo, status, iters, nullptr);
}

static int descent_qf(int gradient, const struct qp_ref_quadratic_form *qf, const double *x0,
		      unsigned iterations, double *x_out, int32_t *status, int32_t *steps)
{
	if (!qf || !qf->p || !qf->q || !qf->p->elements || !qf->q->elements || !x_out)
		return QP_ERR_INVALID;
	const unsigned n = ref_rows(qf->p);
	if (n == 0 || ref_cols(qf->p) != n || ref_rows(qf->q) * ref_cols(qf->q) != n)
		return QP_ERR_INVALID;
	if (n > QP_MAX_DIM)
		return QP_ERR_UNSUPPORTED;
	std::vector<float> P(size_t(n) * n), q(n), x0f(n, 0.0f), x(n);
	for (size_t k = 0; k < size_t(n) * n; ++k)
		P[k] = float(qf->p->elements[k]);
	for (size_t k = 0; k < n; ++k) {
		q[k] = float(qf->q->elements[k]);
		if (x0)
			x0f[k] = float(x0[k]);
	}
	qp_batch pb;
	std::memset(&pb, 0, sizeof(pb));
	pb.n = n, pb.mem = QP_MEM_HOST, pb.count = 1, pb.P = P.data(), pb.q = q.data();
	const int rc = descent_batch(gradient, &pb, x0f.data(), iterations, x.data(), status, steps,
				     nullptr, 0.0f, nullptr);
	if (rc != QP_OK)
		return rc;
	for (size_t k = 0; k < n; ++k)
		x_out[k] = double(x[k]);
	return QP_OK;
}

extern "C" int qp_newton_qf(const struct qp_ref_quadratic_form *qf, const double *x0,
			    unsigned iterations, double *x_out, int32_t *status, int32_t *steps)
{
	return descent_qf(0, qf, x0, iterations, x_out, status, steps);
}

extern "C" int qp_gradient_qf(const struct qp_ref_quadratic_form *qf, const double *x0,
			      unsigned iterations, double *x_out, int32_t *status, int32_t *steps)
{
	return descent_qf(1, qf, x0, iterations, x_out, status, steps);
}

extern "C" const char *qp_kernel_for(uint32_t n, int al16)
{
	static thread_local char buf[32];
	if (n == 0 || n > QP_MAX_DIM_BOX)
		return "unsupported";
	KernelArgs a = {};
	a.n = int(n);
	if (!al16)
		a.P = reinterpret_cast<const float *>(uintptr_t(4)); // a pointer that is not 16-byte aligned
	switch (pick_kernel(a, QP_KERNEL_AUTO)) {
	case QP_KERNEL_TPQ: std::snprintf(buf, sizeof buf, "tpq<%u>", n); break;
	case QP_KERNEL_BPQ: std::snprintf(buf, sizeof buf, "bpq<%u>", n); break;
	case QP_KERNEL_CPQ: std::snprintf(buf, sizeof buf, "cpq<%u>", (n + 15) / 16 * 16); break;
	default: std::snprintf(buf, sizeof buf, "gpq<%u>", (n + 7) / 8 * 8); break;
	}
	return buf;
}

// ----------------------------------------------------------------------------------------
// pure-copy ceiling of the host-memory path
// ----------------------------------------------------------------------------------------
extern "C" double qp_probe_copy_seconds(int device, size_t h2d_bytes, size_t d2h_bytes, int pinned,
					size_t piece_bytes, int repeats)
{
	const int dev = resolve_device(device);
	if (dev < 0)
		return double(QP_ERR_NO_DEVICE);
	DeviceGuard g;
	if (g.enter(dev) != QP_OK)
		return double(QP_ERR_NO_DEVICE);
	const size_t piece = piece_bytes ? piece_bytes : (32ull << 20);
	char *din = nullptr, *dout = nullptr, *hin = nullptr, *hout = nullptr;
	cudaStream_t s0 = nullptr, s1 = nullptr;
	double best = -1.0;
	bool ok = cudaMalloc(&din, std::max<size_t>(h2d_bytes, 16)) == cudaSuccess &&
		  cudaMalloc(&dout, std::max<size_t>(d2h_bytes, 16)) == cudaSuccess &&
		  cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking) == cudaSuccess &&
		  cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking) == cudaSuccess;
	if (ok && pinned)
		ok = cudaMallocHost(&hin, std::max<size_t>(h2d_bytes, 16)) == cudaSuccess &&
		     cudaMallocHost(&hout, std::max<size_t>(d2h_bytes, 16)) == cudaSuccess;
	else if (ok) {
		hin = static_cast<char *>(std::malloc(std::max<size_t>(h2d_bytes, 16)));
		hout = static_cast<char *>(std::malloc(std::max<size_t>(d2h_bytes, 16)));
		ok = hin && hout;
	}
	if (ok) {
		std::memset(hin, 1, h2d_bytes);
		std::memset(hout, 0, d2h_bytes);
		cudaMemset(dout, 2, d2h_bytes);
		for (int r = 0; r < std::max(1, repeats) + 1 && ok; ++r) {
			cudaDeviceSynchronize();
			const auto t0 = std::chrono::steady_clock::now();
			size_t a = 0, b = 0;
			while ((a < h2d_bytes || b < d2h_bytes) && ok) {
				if (a < h2d_bytes) {
					const size_t c = std::min(piece, h2d_bytes - a);
					ok = cudaMemcpyAsync(din + a, hin + a, c, cudaMemcpyHostToDevice, s0) == cudaSuccess;
					a += c;
				}
				if (b < d2h_bytes && ok) {
					// the results are about a sixth of the inputs: keep the two directions in step
					const size_t c = std::min(std::max<size_t>(piece * d2h_bytes / std::max<size_t>(h2d_bytes, 1), 1 << 16),
								  d2h_bytes - b);
					ok = cudaMemcpyAsync(hout + b, dout + b, c, cudaMemcpyDeviceToHost, s1) == cudaSuccess;
					b += c;
				}
			}
			ok = ok && cudaStreamSynchronize(s0) == cudaSuccess && cudaStreamSynchronize(s1) == cudaSuccess;
			const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
			if (ok && r > 0)
				best = best < 0.0 ? dt : std::min(best, dt);
		}
	}
	if (!ok) {
		g_err = std::string("qp_probe_copy_seconds: ") + cudaGetErrorString(cudaGetLastError());
		best = double(QP_ERR_CUDA);
	}
	if (pinned) {
		cudaFreeHost(hin);
		cudaFreeHost(hout);
	} else {
		std::free(hin);
		std::free(hout);
	}
	cudaFree(din);
	cudaFree(dout);
	if (s0)
		cudaStreamDestroy(s0);
	if (s1)
		cudaStreamDestroy(s1);
	return best;
}
