// qp_aux_args.h -- argument blocks and launch geometry of the Newton / gradient kernel
// (qp_newton.cuh) and of the general-rows kernel (qp_gen.cuh), shared with the host runtime.
#pragma once

#include "qp_common.cuh"

namespace qpb {

constexpr int NEWTON_WARPS = 4;      // warps (= problems in flight) per CTA
constexpr int NEWTON_MAX_SHRINKS = 512; // 0.9^512 ~ 4e-24: the step is gone long before

struct NewtonArgs {
	const float *P, *q, *x0; // x0 may be null (= 0)
	float *x, *fval;         // fval may be null
	int32_t *status, *steps; // may be null
	unsigned long long count;
	int n;
	unsigned iterations;
	float r; // constant term of the objective, for fval only
	int gradient; // 1: gradient_descent_with_line_search (qp_solvers.c:65-101): d = -g, no inverse
};

__host__ __device__ inline int newton_ld(int n)
{
	return n | 1; // odd row stride: a lane per row reads conflict-free
}

__host__ __device__ inline size_t newton_smem_floats_per_warp(int n)
{
	return 2 * size_t(n) * newton_ld(n) + 4 * 64;
}

constexpr int GEN_WARPS = 4;      // warps per CTA
constexpr int GEN_MAX_ROWS = 128; // m
constexpr int GEN_RPL = 4;        // rows per lane

struct GenArgs {
	const float *P, *q, *A, *l, *u; // A [count][m][n] row-major; l, u [count][m]
	const float *z0, *u0;           // [count][m] warm start or null
	float *x, *z, *y;               // x [count][n]; z, y = rho u [count][m]; z, y may be null
	int32_t *status, *iters;
	unsigned long long count;
	int n, m;
	float rho, alpha;
	float eps_abs_pri, eps_abs_dual; // sqrt(m) abstol, sqrt(n) abstol
	float reltol;
	unsigned max_iter;
	int force; // 0: the register kernel (qp_genr.cuh) where the shape fits; 1: qp_gen.cuh; 2: qp_genr.cuh or refuse
};

__host__ __device__ inline int gen_ld(int n)
{
	return n | 1; // odd row stride: one lane per row reads conflict-free
}

// lanes per problem: the smallest group with n <= 2T and m <= 4T
__host__ __device__ inline int gen_lanes(int n, int m)
{
	return (n <= 8 && m <= 16) ? 4 : ((n <= 16 && m <= 32) ? 8 : ((n <= 32 && m <= 64) ? 16 : 32));
}

// floats of shared memory per group, padded so that the groups of a warp start T banks
// apart: the same element of every group's matrices then lies in a different bank window
__host__ __device__ inline size_t gen_smem_floats_per_group(int n, int m, int T)
{
	const size_t need = size_t(n) * gen_ld(n) + size_t(m) * gen_ld(n) + 3 * (2 * T) + 2 * (GEN_RPL * T);
	const size_t r = need % 32, want = size_t(T) % 32;
	return need + (want + 32 - r) % 32;
}

// ---- qp_f64.cuh: the box ADMM in binary64 (qf entry points, qp_options.precision) -----------
constexpr int F64_WARPS = 4;

struct Admm64Args {
	const double *P, *q;       // [count][n][n], [count][n]
	double lb, ub;             // the scalar box (config.h:29-30)
	double *x;                 // [count][n]
	int32_t *status, *iters;   // may be null
	unsigned long long count;
	int n;
	double rho, alpha, eps_abs, reltol; // eps_abs = sqrt(n) * abstol (:242)
	unsigned max_iter;
};

__host__ __device__ inline int f64_ld(int n)
{
	return n | 1;
}
__host__ __device__ inline size_t f64_smem_doubles_per_warp(int n)
{
	return size_t(n) * f64_ld(n) + 2 * 64;
}

} // namespace qpb
