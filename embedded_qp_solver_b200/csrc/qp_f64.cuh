// qp_f64.cuh -- the box ADMM in binary64, one warp per problem, n <= 64.
//
// What it replaces: the reference's admm() (/root/reference/qp_solvers/qp_solvers.c:255-319) in
// the reference's own arithmetic type -- its struct _matrix holds doubles
// (matrix/matrix_type/matrix_type.h:20-23).  The batch kernels compute in binary32 (that is the
// product); on the reference's own ill-conditioned test distribution
// (matrix/matrix_ops/matrix_ops.c:699-734: 3 ... 10 000 iterations, reltol 1e-2 stop test) a
// binary32 run stops at the reference's iteration on about two problems in three, because the
// stop test compares quantities that agree to six digits.  This kernel is for callers of the
// qf entry points (doubles in, doubles out) who want the reference's iteration counts and its
// x to ~1e-12: qp_options.precision = QP_PRECISION_F64.  A correctness path, not a tuned one.
//
// Layout: R = (P + rho I)^-1 in shared memory (Gauss-Jordan without pivoting: the matrix is
// symmetric positive definite; the reference's LU with partial pivoting, matrix_ops.c:487-630,
// picks the diagonal on such a matrix up to rounding), odd row stride; lane l owns coordinates
// l and l + 32: x = R y row by row, the element-wise updates (:161-203) on the owned
// coordinates, the five sums of squares (:205-253) by a warp butterfly, the stop test (:308)
// with real square roots as the reference has them.
#pragma once

#include "qp_aux_args.h"

namespace qpb {

__device__ __forceinline__ double warp_sum_f64(double v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

__global__ void __launch_bounds__(F64_WARPS * 32) qp_admm64_kernel(const Admm64Args a)
{
	extern __shared__ __align__(16) double dsm[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int n = a.n, ld = f64_ld(n);
	double *R = dsm + warp * f64_smem_doubles_per_warp(n);
	double *y = R + size_t(n) * ld, *col = y + 64;
	const int i0 = lane, i1 = lane + 32;
	const bool h0 = i0 < n, h1 = i1 < n;
	const double rho = a.rho, alpha = a.alpha;

	const unsigned long long nw = (unsigned long long)gridDim.x * F64_WARPS;
	for (unsigned long long prob = (unsigned long long)blockIdx.x * F64_WARPS + warp; prob < a.count; prob += nw) {
		const double *Pp = a.P + prob * (unsigned long long)n * n;
		for (int e = lane; e < n * n; e += 32) {
			const int r = e / n, c = e - r * n;
			R[r * ld + c] = Pp[e] + (r == c ? rho : 0.0);
		}
		const double q0 = h0 ? a.q[prob * n + i0] : 0.0, q1 = h1 ? a.q[prob * n + i1] : 0.0;
		__syncwarp();
		// ---- R <- R^-1, lanes own columns i0, i1 --------------------------------------------
		bool bad = !(rho > 0.0);
		for (int k = 0; k < n; ++k) {
			const double piv = R[k * ld + k];
			bad |= !(piv > 0.0);
			const double dinv = 1.0 / piv;
			__syncwarp();
			double p0 = 0.0, p1 = 0.0;
			if (h0) {
				p0 = (i0 == k) ? dinv : R[k * ld + i0] * dinv;
				R[k * ld + i0] = p0;
			}
			if (h1) {
				p1 = (i1 == k) ? dinv : R[k * ld + i1] * dinv;
				R[k * ld + i1] = p1;
			}
			if (h0) col[i0] = R[i0 * ld + k];
			if (h1) col[i1] = R[i1 * ld + k];
			__syncwarp();
			for (int i = 0; i < n; ++i) {
				if (i == k)
					continue;
				const double f = col[i];
				if (h0)
					R[i * ld + i0] = (i0 == k) ? -f * p0 : fma(-f, p0, R[i * ld + i0]);
				if (h1)
					R[i * ld + i1] = (i1 == k) ? -f * p1 : fma(-f, p1, R[i * ld + i1]);
			}
			__syncwarp();
		}
		// ---- iterations (:284-310) ------------------------------------------------------------
		double z0 = 0.0, z1 = 0.0, u0 = 0.0, u1 = 0.0, x0 = 0.0, x1 = 0.0;
		int status = QP_STATUS_MAX_ITER;
		unsigned it = 0;
		if (!bad) {
			while (it < a.max_iter) {
				if (h0) y[i0] = rho * (z0 - u0) - q0;
				if (h1) y[i1] = rho * (z1 - u1) - q1;
				__syncwarp();
				double s0 = 0.0, s1 = 0.0;
				for (int j = 0; j < n; ++j) {
					const double yj = y[j];
					if (h0) s0 = fma(R[i0 * ld + j], yj, s0);
					if (h1) s1 = fma(R[i1 * ld + j], yj, s1);
				}
				__syncwarp();
				x0 = s0, x1 = s1;
				const double xh0 = alpha * x0 + (1.0 - alpha) * z0, xh1 = alpha * x1 + (1.0 - alpha) * z1;
				const double zn0 = h0 ? fmin(fmax(xh0 + u0, a.lb), a.ub) : 0.0;
				const double zn1 = h1 ? fmin(fmax(xh1 + u1, a.lb), a.ub) : 0.0;
				const double dz0 = zn0 - z0, dz1 = zn1 - z1;
				u0 += h0 ? xh0 - zn0 : 0.0, u1 += h1 ? xh1 - zn1 : 0.0;
				const double dr0 = h0 ? x0 - zn0 : 0.0, dr1 = h1 ? x1 - zn1 : 0.0;
				z0 = zn0, z1 = zn1;
				const double rr = warp_sum_f64(dr0 * dr0 + dr1 * dr1);
				const double ss = warp_sum_f64(dz0 * dz0 + dz1 * dz1);
				const double nx = warp_sum_f64((h0 ? x0 * x0 : 0.0) + (h1 ? x1 * x1 : 0.0));
				const double nz = warp_sum_f64(z0 * z0 + z1 * z1);
				const double nu = warp_sum_f64(u0 * u0 + u1 * u1);
				++it;
				const double eps_pri = a.eps_abs + a.reltol * fmax(sqrt(nx), sqrt(nz));
				const double eps_dual = a.eps_abs + a.reltol * rho * sqrt(nu);
				if (sqrt(rr) < eps_pri && rho * sqrt(ss) < eps_dual) {
					status = QP_STATUS_CONVERGED;
					break;
				}
			}
		}
		const double nf = warp_sum_f64(x0 * 0.0 + x1 * 0.0);
		int iters = int(it);
		if (bad)
			status = QP_STATUS_SINGULAR, iters = 0;
		else if (!(nf == 0.0))
			status = QP_STATUS_NONFINITE;
		if (h0) a.x[prob * n + i0] = x0;
		if (h1) a.x[prob * n + i1] = x1;
		if (lane == 0) {
			if (a.status) a.status[prob] = status;
			if (a.iters) a.iters[prob] = iters;
		}
		__syncwarp();
	}
}

} // namespace qpb
