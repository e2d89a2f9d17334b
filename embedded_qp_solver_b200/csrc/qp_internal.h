// qp_internal.h -- declarations shared by the translation units of libqpb200.so.
//
// The library is built from one translation unit per kernel family (k_tpq.cu, k_bpq.cu,
// k_gpq.cu, k_cpq.cu, k_aux.cu), compiled in parallel, plus the host runtime and C ABI
// (qp_batch.cu).  Nothing here is visible to callers; the ABI is include/qp_batch.h.
#pragma once

#include <atomic>
#include <string>

#include "qp_aux_args.h"

namespace qpb {

// text of the last QP_ERR_CUDA, per host thread (qp_last_error())
extern thread_local std::string g_err;

#define QP_CUDA(call)                                                                          \
	do {                                                                                   \
		cudaError_t e_ = (call);                                                       \
		if (e_ != cudaSuccess) {                                                       \
			qpb::g_err = std::string(#call) + ": " + cudaGetErrorString(e_);       \
			return QP_ERR_CUDA;                                                    \
		}                                                                              \
	} while (0)

constexpr int MAX_DEVICES = 64;

// number of SMs of a device (cached)
int dev_sms(int dev);

// "has this function attribute / occupancy been set up on this device": written from the
// per-GPU host threads of a sharded solve, so it is an atomic, one per template instance
struct PerDeviceOnce {
	std::atomic<int> v[MAX_DEVICES];
	PerDeviceOnce()
	{
		for (auto &a : v)
			a.store(0, std::memory_order_relaxed);
	}
	int get(int dev) const { return v[dev].load(std::memory_order_acquire); }
	void set(int dev, int val) { v[dev].store(val, std::memory_order_release); }
};

// switch the calling host thread to a device for the lifetime of the guard
struct DeviceGuard {
	int prev = -1;
	bool changed = false;
	int enter(int dev)
	{
		if (cudaGetDevice(&prev) != cudaSuccess)
			return QP_ERR_NO_DEVICE;
		if (dev != prev) {
			if (cudaSetDevice(dev) != cudaSuccess)
				return QP_ERR_NO_DEVICE;
			changed = true;
		}
		return QP_OK;
	}
	~DeviceGuard()
	{
		if (changed)
			cudaSetDevice(prev);
	}
};

// device ordinal of a qp_options.device value (-1 = the current device); -1 when unusable
int resolve_device(int dev);

inline bool aligned16(const void *p)
{
	return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

// family launchers: QP_OK, QP_ERR_UNSUPPORTED (the family does not take this batch) or
// QP_ERR_CUDA
bool tpq_eligible(const KernelArgs &a);
bool bpq_eligible(const KernelArgs &a);
// mode = PASS_PLAIN / PASS_FIRST / PASS_LIST (qp_common.cuh)
int launch_tpq_family(const KernelArgs &a, int dev, cudaStream_t st, int mode);
int launch_bpq_family(const KernelArgs &a, int dev, cudaStream_t st, int mode);
int launch_gpq_family(const KernelArgs &a, int dev, cudaStream_t st);
int launch_cpq_family(const KernelArgs &a, int dev, cudaStream_t st);
// PASS_STREAM (qp_bpq_stream.cuh): can the second launch of a pass schedule be the streaming
// kernel for this batch (full inverse, i.e. no product form), and its launcher
bool bpq_stream_eligible(const KernelArgs &a);
// fresh: shared matrices -- every problem of the batch, from scratch, with the inverse that a
// PASS_SETUP launch of qp_bpq.cuh left for its group
int launch_bpq_stream_family(const KernelArgs &a, int dev, cudaStream_t st, bool fresh);
// which families have the pass modes (KernelArgs::ps)
bool family_has_passes(int family, int n);

int launch_newton(const NewtonArgs &a, int dev, cudaStream_t st);
// qp_f64.cuh: the box ADMM in binary64 (qf entry points, qp_options.precision)
int launch_admm64(const Admm64Args &a, int dev, cudaStream_t st);
int launch_gen(const GenArgs &a, int dev, cudaStream_t st);
// qp_genr.cuh (k_genr.cu): QP_ERR_UNSUPPORTED when n > 32 or m > 64
int launch_genr(const GenArgs &a, int dev, cudaStream_t st);
int gen_ticket(int dev, cudaStream_t st, unsigned long long **out);
int launch_eval(const float *P, const float *q, const float *x, float r, float *f, int n,
		unsigned long long count, int dev, cudaStream_t st);

} // namespace qpb
