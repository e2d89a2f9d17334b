// qp_common.cuh -- shared device-side definitions for the batched box-QP kernels.
//
// Target: sm_100a only (B200).  FP32 CUDA cores; bulk asynchronous copies
// (cp.async.bulk + mbarrier) move problem data between HBM and shared memory.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/qp_batch.h"

namespace qpb {

// Passes for heavy-tailed iteration counts.  The reference's own test
// distribution (matrix_ops.c:699-734) needs 3 ... 10 000 iterations per problem; a warp that
// keeps 32 problems until the slowest one stops idles 90 % of its lanes there.  With passes,
// a problem gets a WINDOW of iterations per launch (24 or 16 at first, then growing); one
// that has not converged when its window closes is EVICTED: its ADMM state (z, u, iteration
// count) goes to scratch memory in HBM and its index to a list in HBM that the NEXT launch
// (stream order is the barrier) works through 32 at a time -- stragglers share warps with
// stragglers, and no warp idles for longer than a window.  The list lengths stay on the
// device; the host enqueues a fixed, small number of launches and never synchronises.
// A continued problem repeats the factorisation (qp_bpq.cuh: reloads its saved inverse) and
// carries on from (z, u): the arithmetic is the one the plain kernel does, bit for bit --
// WHERE a problem is evicted changes the schedule and never the result.
//
//   PASS_PLAIN  the whole batch, every problem to the end (the round-1 kernel); optionally
//               adds up the iteration counts and what the warps waited for (stats[3..5]),
//               which is how the host notices that the next call should use passes
//   PASS_FIRST  the whole batch, `window` iterations each, the unfinished ones to next_list
//   PASS_LIST   the problems of `list`, state from scratch, `window` more iterations each,
//               the unfinished ones to next_list (null: the last pass runs everything out)
//   PASS_STREAM qp_bpq_stream.cuh: the problems of `list`, every one to the end, each lane group
//               taking its next problem by ticket (stream_ctr) as soon as its current one stops
//   PASS_SETUP  qp_bpq.cuh, shared matrices (qp_batch.p_share): the "problems" of the launch are the
//               MATRICES; each is eliminated once, its inverse goes to Rs, "not positive definite"
//               to its[]; no iteration.  The streaming kernel then solves every problem from
//               scratch with the inverse of its group.
constexpr int PASS_PLAIN = 0, PASS_FIRST = 1, PASS_LIST = 2, PASS_STREAM = 3, PASS_SETUP = 4;
struct PassArgs {
	unsigned long long *stats;  // null, or counters the kernels add to (qp_batch.cu reads them):
	                            // [0] problems that outlived the first window (PASS_FIRST: evicted)
	                            // [1] problems a PASS_FIRST launch looked at
	                            // [2] evictions by PASS_LIST launches
	                            // PASS_PLAIN (the probe of QP_SCHEDULE_AUTO), one add per CTA:
	                            // [3] sum of iteration counts, [4] sum over warps of (largest
	                            // iteration count in the warp x problems in the warp) -- [3] / [4]
	                            // is the fraction of lane-iterations that did work --, [5] problems
	                            // PASS_STREAM, one add per lane group: [6] sum of the final iteration
	                            // counts of the problems it finished, [7] how many those were
	unsigned window;            // PASS_FIRST / PASS_LIST: iterations per problem in this launch
	const unsigned *list;       // PASS_LIST: problem indices ...
	const unsigned *list_count; // ... and how many (written by the previous launch)
	unsigned *next_list;        // evicted problems are appended here (null: nobody is evicted)
	unsigned *next_count;
	float *zs, *us;             // [count][n] z and u of an evicted problem
	int *its;                   // [count]    iterations an evicted problem has done
	float *Rs;                  // bpq: [count][n][n] its inverse (or eta file)
	float *xs;                  // bpq, streaming pass next: [count][n] x-update of its next iteration
	unsigned *stream_ctr;       // PASS_STREAM: tickets handed out so far (zero at launch)
	// qp_bpq.cuh PASS_FIRST -> PASS_STREAM: a second list.  The first pass appends to next_list the
	// problems whose residuals had almost stopped shrinking when the window closed (they will run
	// long) and to next_list2 all the others; the streaming pass hands out `list` before `list2`.
	unsigned *next_list2, *next_count2; // null: everything goes to next_list
	const unsigned *list2, *list_count2;
};

// Kernel arguments: one packed batch resident in device memory.
// Mirrors qp_batch / qp_batch_solution / qp_options (include/qp_batch.h).
struct KernelArgs {
	const float *P, *q, *lb, *ub, *z0, *u0;
	float *x, *z, *lam;
	int32_t *status, *iters;
	unsigned long long count;
	int n;
	float lb_s, ub_s;      // scalar box when lb/ub are null
	float rho, alpha;
	const float *rho_v;    // [count] per-problem rho, or null: rho for all
	float eps_abs;         // sqrt(n) * abstol   (qp_solvers.c:242, :252)
	float reltol;
	unsigned max_iter;
	int p_packed;          // tpq: P holds packed upper triangles, tpq_packed_stride(n) floats each
	unsigned p_share;      // >= 1: problems (i + p_share_off) / p_share share one P (qp_batch.p_share)
	unsigned long long p_share_off; // position of problem 0 of this launch inside its group (host chunks)
	PassArgs ps;           // passes for heavy-tailed iteration counts (see above)
};

// which matrix of KernelArgs::P problem `prob` of the launch uses
__device__ __forceinline__ unsigned long long p_index(const KernelArgs &p, unsigned long long prob)
{
	return p.p_share > 1u ? (prob + p.p_share_off) / p.p_share : prob;
}

// floats per problem of the packed upper triangle (row-major, i <= j), padded to 16 bytes
__host__ __device__ constexpr int tpq_packed_stride(int n)
{
	return (n * (n + 1) / 2 + 3) / 4 * 4;
}

// ---- PTX wrappers ---------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
	return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals)
		     : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
		     "r"(bytes)
		     : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"QPB_WAIT:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra QPB_DONE;\n"
		"bra QPB_WAIT;\n"
		"QPB_DONE:\n"
		"}\n" ::"r"(smem_u32(bar)),
		"r"(parity)
		: "memory");
}

// global -> shared bulk copy (TMA engine, 1-D), completion on an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes,
					 uint64_t *bar)
{
	asm volatile(
		"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
			"r"(smem_u32(dst)),
		"l"(src), "r"(bytes), "r"(smem_u32(bar))
		: "memory");
}

// shared -> global bulk store, tracked by the per-thread bulk async-group.
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes)
{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
		     "r"(smem_u32(src)), "r"(bytes)
		     : "memory");
}
__device__ __forceinline__ void bulk_commit()
{
	asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// wait until the sources of all committed bulk stores have been read
__device__ __forceinline__ void bulk_wait_read()
{
	asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// wait until all committed bulk stores have completed (their writes performed)
__device__ __forceinline__ void bulk_wait_all()
{
	asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// make generic-proxy shared-memory accesses visible to / ordered with the
// async proxy (bulk copies)
__device__ __forceinline__ void fence_async_smem()
{
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// MUFU.SQRT: one instruction, no special-case branches (relative error ~2^-22,
// far inside the 1e-3 stop-test margin that defines a non-degenerate problem)
__device__ __forceinline__ float sqrt_approx(float v)
{
	float r;
	asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
	return r;
}

// MUFU.RCP: one instruction (1 ulp), for the reciprocals of the block pivots
__device__ __forceinline__ float rcp_approx(float v)
{
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
	return r;
}

// reciprocal to ~1 ulp without the range checks and slow path of __frcp_rn: MUFU.RCP and
// one Newton step (pivots of P + rho I are far from the denormal range)
__device__ __forceinline__ float rcp_newton(float v)
{
	const float r = rcp_approx(v);
	return fmaf(r, fmaf(-v, r, 1.0f), r);
}

// L D L^T of a symmetric positive definite 4x4 pivot block, for the block Gauss-Jordan of
// the lane-group kernels.  The multipliers of a block step, -c B^-1, are formed by two
// triangular solves with these factors (minus_solve below), NOT by multiplying with an
// explicit inverse of B: the solves are componentwise backward stable, the explicit inverse
// is not, and on ill-conditioned problems (condensed MPC Hessians, cond(P + rho I) ~ 1e4)
// the explicit form lost two digits in R (profiles/README.md, "block pivot accuracy").
// Same operation count: 6 + 4 + 6 against 16.
struct Ldl4 {
	float l10, l20, l21, l30, l31, l32; // unit lower triangle
	float n0, n1, n2, n3;               // -1 / d_k
};

// Reads the upper triangle of B.  Returns true when a pivot is not strictly positive.
__device__ __forceinline__ bool ldl4_factor(const float (&B)[4][4], Ldl4 &f)
{
	const float d0 = B[0][0];
	const float i0 = rcp_approx(d0);
	f.l10 = B[0][1] * i0, f.l20 = B[0][2] * i0, f.l30 = B[0][3] * i0;
	const float d1 = fmaf(-f.l10, B[0][1], B[1][1]);
	const float t21 = fmaf(-f.l10, B[0][2], B[1][2]);
	const float t31 = fmaf(-f.l10, B[0][3], B[1][3]);
	const float i1 = rcp_approx(d1);
	f.l21 = t21 * i1, f.l31 = t31 * i1;
	const float d2 = fmaf(-f.l21, t21, fmaf(-f.l20, B[0][2], B[2][2]));
	const float t32 = fmaf(-f.l21, t31, fmaf(-f.l20, B[0][3], B[2][3]));
	const float i2 = rcp_approx(d2);
	f.l32 = t32 * i2;
	const float d3 = fmaf(-f.l32, t32, fmaf(-f.l31, t31, fmaf(-f.l30, B[0][3], B[3][3])));
	const float i3 = rcp_approx(d3);
	f.n0 = -i0, f.n1 = -i1, f.n2 = -i2, f.n3 = -i3;
	return !(d0 > 0.0f) || !(d1 > 0.0f) || !(d2 > 0.0f) || !(d3 > 0.0f);
}

// h = -c B^-1 for the row vector c (B = L D L^T symmetric):  w = c L^-T, m = -w D^-1,
// h = m L^-1
__device__ __forceinline__ void ldl4_minus_solve(const Ldl4 &f, float c0, float c1, float c2, float c3,
						 float &h0, float &h1, float &h2, float &h3)
{
	const float w1 = fmaf(-f.l10, c0, c1);
	const float w2 = fmaf(-f.l21, w1, fmaf(-f.l20, c0, c2));
	const float w3 = fmaf(-f.l32, w2, fmaf(-f.l31, w1, fmaf(-f.l30, c0, c3)));
	const float m0 = c0 * f.n0, m1 = w1 * f.n1, m2 = w2 * f.n2;
	h3 = w3 * f.n3;
	h2 = fmaf(-f.l32, h3, m2);
	h1 = fmaf(-f.l31, h3, fmaf(-f.l21, h2, m1));
	h0 = fmaf(-f.l30, h3, fmaf(-f.l20, h2, fmaf(-f.l10, h1, m0)));
}

// the same for two rows at once, packed as (row a, row b)
__device__ __forceinline__ float2 fma2_s(float s, float2 a, float2 c)
{
	return __ffma2_rn(make_float2(s, s), a, c);
}
__device__ __forceinline__ void ldl4_minus_solve2(const Ldl4 &f, float2 c0, float2 c1, float2 c2, float2 c3,
						  float2 &h0, float2 &h1, float2 &h2, float2 &h3)
{
	const float2 w1 = fma2_s(-f.l10, c0, c1);
	const float2 w2 = fma2_s(-f.l21, w1, fma2_s(-f.l20, c0, c2));
	const float2 w3 = fma2_s(-f.l32, w2, fma2_s(-f.l31, w1, fma2_s(-f.l30, c0, c3)));
	const float2 m0 = __fmul2_rn(c0, make_float2(f.n0, f.n0));
	const float2 m1 = __fmul2_rn(w1, make_float2(f.n1, f.n1));
	const float2 m2 = __fmul2_rn(w2, make_float2(f.n2, f.n2));
	h3 = __fmul2_rn(w3, make_float2(f.n3, f.n3));
	h2 = fma2_s(-f.l32, h3, m2);
	h1 = fma2_s(-f.l31, h3, fma2_s(-f.l21, h2, m1));
	h0 = fma2_s(-f.l30, h3, fma2_s(-f.l20, h2, fma2_s(-f.l10, h1, m0)));
}

// shared-memory 128-bit load the compiler may not hoist or merge
__device__ __forceinline__ float4 lds128(const float4 *p)
{
	float4 v;
	asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
		     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
		     : "r"(smem_u32(p)));
	return v;
}

// the same with ld.volatile: ptxas will neither merge two of them nor replace one
// by register moves (used to fill registers with constants four at a time)
__device__ __forceinline__ float4 lds128_volatile(const float4 *p)
{
	float4 v;
	asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
		     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
		     : "r"(smem_u32(p)));
	return v;
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// In-place inverse of a symmetric positive definite n x n matrix in SHARED memory (row
// stride ld), by one warp: Gauss-Jordan without pivoting, lanes own columns lane and
// lane + 32 (n <= 64).  `col` is an n-float scratch vector in shared memory.  Returns
// true when a pivot is not strictly positive.  Used by the correctness-first kernels
// (qp_newton.cuh, qp_gen.cuh); the tuned kernels invert in registers.
__device__ __forceinline__ bool smem_gauss_jordan(float *W, int n, int ld, float *col, int lane)
{
	const int i0 = lane, i1 = lane + 32;
	const bool h0 = i0 < n, h1 = i1 < n;
	bool bad = false;
	for (int k = 0; k < n; ++k) {
		const float piv = W[k * ld + k];
		bad |= !(piv > 0.0f);
		const float dinv = 1.0f / piv;
		__syncwarp();
		// scaled pivot row, kept in registers for the elimination below
		float p0 = 0.f, p1 = 0.f;
		if (h0) {
			p0 = (i0 == k) ? dinv : W[k * ld + i0] * dinv;
			W[k * ld + i0] = p0;
		}
		if (h1) {
			p1 = (i1 == k) ? dinv : W[k * ld + i1] * dinv;
			W[k * ld + i1] = p1;
		}
		// column k (the multipliers) is copied out first, so that its owner can rewrite
		// it while the others still read it
		if (h0) col[i0] = W[i0 * ld + k];
		if (h1) col[i1] = W[i1 * ld + k];
		__syncwarp();
		for (int i = 0; i < n; ++i) {
			if (i == k)
				continue;
			const float f = col[i];
			if (h0)
				W[i * ld + i0] = (i0 == k) ? -f * p0 : fmaf(-f, p0, W[i * ld + i0]);
			if (h1)
				W[i * ld + i1] = (i1 == k) ? -f * p1 : fmaf(-f, p1, W[i * ld + i1]);
		}
		__syncwarp();
	}
	return bad;
}

__device__ __forceinline__ bool finite_f(float v)
{
	return (v - v) == 0.0f;
}

} // namespace qpb
