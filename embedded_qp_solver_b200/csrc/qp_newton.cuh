// qp_newton.cuh -- batched unconstrained minimisers: the reference's
// newton_method_with_line_search() (/root/reference/qp_solvers/qp_solvers.c:103-144) and
// gradient_descent_with_line_search() (:65-101, the same loop with d = -grad and no
// inverse) with their line_search() (:37-63) and armijo() (:21-35), one warp per problem,
// any n <= 64.
//
// This is the cross-check path of SURVEY.md 8(f) rank 2 (with the box inactive the ADMM
// answer is the Newton answer), not the headline kernel: P and its inverse live in shared
// memory (two n x n tiles per warp), lanes own columns during the inversion and rows
// during the mat-vecs.
//
//   setup   hessian_inv = P^-1                      (:115-117, matrix_invert)
//           Gauss-Jordan without pivoting: P must be symmetric positive definite (a convex
//           QP); a pivot <= 0 is status 2, where the reference's LU would print
//           "singular matrix" (matrix_ops.c:511-515) or pivot around it.
//   loop    g = P x + q                             (:119, qp.c:29-44)
//           stop when 0.1 > ||g||                   (:125, MIN_GRAD_NEWTON)
//           d = -hessian_inv g                      (:129-132)
//           alpha = line_search(alpha0 = 1, shrink 0.9, c2 = 1e-4)   (:134)
//           x += alpha d                            (:136-137)
//
// The reference's Armijo test evaluates f at xk + d_alpha with xk = x + d_alpha already
// (:45-48 and :26-27), i.e. it accepts alpha when  f(x + 2 alpha d) <= f(x) + c2 alpha g'd.
// f is quadratic, so the kernel tests the same inequality in difference form,
//     2 alpha g'd + 2 alpha^2 d'Pd <= c2 alpha g'd,
// which needs one more mat-vec (P d) per Newton step instead of one objective evaluation
// per trial, and does not subtract two nearly equal objective values in binary32.
#pragma once

#include "qp_aux_args.h"

namespace qpb {

__global__ void __launch_bounds__(NEWTON_WARPS * 32) qp_newton_kernel(const NewtonArgs a)
{
	extern __shared__ __align__(16) float nsm[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int n = a.n, ld = newton_ld(n);
	float *A = nsm + warp * newton_smem_floats_per_warp(n); // P
	float *W = A + n * ld;                                   // becomes P^-1
	float *vx = W + n * ld, *vg = vx + 64, *vd = vg + 64;    // x, g, d (broadcast buffers)
	const int i0 = lane, i1 = lane + 32;                     // my rows / columns
	const bool h0 = i0 < n, h1 = i1 < n;

	const unsigned long long nw = (unsigned long long)gridDim.x * NEWTON_WARPS;
	for (unsigned long long prob = (unsigned long long)blockIdx.x * NEWTON_WARPS + warp;
	     prob < a.count; prob += nw) {
		// ---- load P (coalesced), x0, q -----------------------------------------------------
		const float *Pp = a.P + prob * (unsigned long long)n * n;
		for (int e = lane; e < n * n; e += 32) {
			const float v = __ldg(Pp + e);
			const int r = e / n, c = e - r * n;
			A[r * ld + c] = v, W[r * ld + c] = v;
		}
		const unsigned long long vo = prob * (unsigned long long)n;
		float q0 = h0 ? __ldg(a.q + vo + i0) : 0.f, q1 = h1 ? __ldg(a.q + vo + i1) : 0.f;
		float x0 = (h0 && a.x0) ? __ldg(a.x0 + vo + i0) : 0.f;
		float x1 = (h1 && a.x0) ? __ldg(a.x0 + vo + i1) : 0.f;
		__syncwarp();

		// ---- W <- P^-1 (Newton only) ---------------------------------------------------------
		const bool bad = a.gradient ? false : smem_gauss_jordan(W, n, ld, vg, lane);

		// ---- Newton steps ---------------------------------------------------------------------
		auto matvec = [&](const float *M, const float *v, float &o0, float &o1) {
			// rows i0, i1 of M times the vector v (shared memory, broadcast reads)
			float s0 = 0.f, s1 = 0.f;
			for (int j = 0; j < n; ++j) {
				const float vj = v[j];
				if (h0)
					s0 = fmaf(M[i0 * ld + j], vj, s0);
				if (h1)
					s1 = fmaf(M[i1 * ld + j], vj, s1);
			}
			o0 = s0, o1 = s1;
		};
		int status = QP_STATUS_MAX_ITER;
		unsigned steps = 0;
		if (!bad) {
			for (unsigned it = 0; it < a.iterations; ++it) {
				if (h0) vx[i0] = x0;
				if (h1) vx[i1] = x1;
				__syncwarp();
				float g0, g1;
				matvec(A, vx, g0, g1); // quadratic_form_eval_grad, qp.c:29-44
				g0 = h0 ? g0 + q0 : 0.f, g1 = h1 ? g1 + q1 : 0.f;
				const float gn2 = warp_sum(fmaf(g0, g0, g1 * g1));
				if (0.1f > sqrtf(gn2)) { // MIN_GRAD_NEWTON, :125
					status = QP_STATUS_CONVERGED;
					break;
				}
				if (h0) vg[i0] = g0;
				if (h1) vg[i1] = g1;
				__syncwarp();
				float d0 = g0, d1 = g1; // gradient descent: d = -grad (:87-90)
				if (!a.gradient)
					matvec(W, vg, d0, d1); // Newton: d = -hessian_inv grad (:129-132)
				d0 = -d0, d1 = -d1;
				if (h0) vd[i0] = d0;
				if (h1) vd[i1] = d1;
				__syncwarp();
				float t0, t1;
				matvec(A, vd, t0, t1); // P d, for the line search in difference form
				const float gd = warp_sum(fmaf(g0, d0, g1 * d1));
				const float dPd = warp_sum(fmaf(d0, t0, d1 * t1));
				// line_search :37-63 with armijo :21-35 (alpha0 = 1, c1 = .9, c2 = 1e-4)
				float alpha = 1.0f;
				for (int s = 0; s < NEWTON_MAX_SHRINKS; ++s) {
					const float lhs = fmaf(2.0f * alpha * alpha, dPd, 2.0f * alpha * gd);
					if (lhs <= 1e-4f * alpha * gd)
						break;
					alpha *= 0.9f;
				}
				x0 = fmaf(alpha, d0, x0), x1 = fmaf(alpha, d1, x1); // :136-137
				++steps;
				__syncwarp();
			}
		}

		// ---- results ----------------------------------------------------------------------------
		const float nf = warp_sum(fmaf(x0, 0.0f, x1 * 0.0f));
		if (bad)
			status = QP_STATUS_SINGULAR;
		else if (!(nf == 0.0f))
			status = QP_STATUS_NONFINITE;
		if (h0) a.x[vo + i0] = x0;
		if (h1) a.x[vo + i1] = x1;
		if (a.fval) { // quadratic_form_eval, qp.c:9-27
			if (h0) vx[i0] = x0;
			if (h1) vx[i1] = x1;
			__syncwarp();
			float t0, t1;
			matvec(A, vx, t0, t1);
			const float quad = warp_sum(fmaf(h0 ? x0 : 0.f, t0, (h1 ? x1 : 0.f) * t1));
			const float lin = warp_sum(fmaf(q0, x0, q1 * x1));
			if (lane == 0)
				a.fval[prob] = fmaf(0.5f, quad, lin) + a.r;
		}
		if (lane == 0) {
			if (a.status) a.status[prob] = status;
			if (a.steps) a.steps[prob] = int32_t(steps);
		}
		__syncwarp();
	}
}

} // namespace qpb
