// qp_gpq.cuh -- "warp per QP" kernel for any n <= 64 (n = 40 and 64 in BASELINE.json).
//
// What it replaces: the same reference path as qp_tpq.cuh -- admm()
// (/root/reference/qp_solvers/qp_solvers.c:255-319) with matrix_invert()
// (/root/reference/matrix/matrix_ops/matrix_ops.c:551-630) and matrix_mult()
// (:235-271) underneath.
//
// Mapping:
//   * one warp owns one QP; lane l owns rows l and l+32 of the matrix, held
//     entirely in registers (NP floats per row, NP = n rounded up to 8).
//   * Loading uses the symmetry of P: lane r reads COLUMN r (= row r), so each
//     step of the load loop is one coalesced read of a matrix row.
//   * R = (P + rho I)^-1 is formed explicitly, like the reference, because the
//     per-iteration x = R y is then a register-resident mat-vec with y broadcast
//     from shared memory; triangular solves would serialise across lanes.
//     The inverse is a Gauss-Jordan elimination without pivoting (P + rho I is
//     symmetric positive definite) in "shifting" form: every step eliminates
//     column 0 and appends the new inverse column at the end, so all register
//     indices are compile-time constants while the pivot loop stays a run-time
//     loop.  Row scaling is deferred to one pass at the end, which keeps the
//     pivot step free of divergence.
//   * per iteration: y to shared memory, NP FMAs per owned row, element-wise
//     updates on the owned coordinates, one butterfly all-reduce of the five
//     sums of squares, and a warp-uniform stop test (no divergence).
//   * x, z, lambda are written with lane = coordinate: coalesced.
#pragma once

#include "qp_common.cuh"

namespace qpb {

constexpr int GPQ_WARPS = 4; // warps (QPs in flight) per CTA

template <int NP>
__global__ void __launch_bounds__(GPQ_WARPS * 32) qp_gpq_kernel(const KernelArgs p)
{
	constexpr int RPL = (NP + 31) / 32; // rows per lane
	constexpr unsigned FULL = 0xffffffffu;
	__shared__ __align__(16) float sm[GPQ_WARPS][NP];

	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const int n = p.n;
	float *bc = sm[warp];
	const float rho = p.rho, alpha = p.alpha;
	const bool relax = (alpha != 1.0f);

	const unsigned long long nwarps = (unsigned long long)gridDim.x * GPQ_WARPS;
	for (unsigned long long prob = (unsigned long long)blockIdx.x * GPQ_WARPS + warp;
	     prob < p.count; prob += nwarps) {
		// ---- load: lane r takes column r of the symmetric P -----------------------
		float S[RPL][NP];
		float sigma[RPL];
		const float *Pp = p.P + prob * (unsigned long long)n * n;
#pragma unroll
		for (int s = 0; s < RPL; ++s) {
			const int r = lane + 32 * s;
#pragma unroll
			for (int j = 0; j < NP; ++j) {
				float v = 0.0f;
				if (r < n && j < n)
					v = __ldg(Pp + (unsigned long long)j * n + r);
				if (r == j)
					v = (r < n) ? v + rho : 1.0f; // padding: identity block
				S[s][j] = v;
			}
			sigma[s] = 1.0f;
		}

		// ---- R = (P + rho I)^-1, shifting Gauss-Jordan ------------------------------
		bool bad = false;
		for (int k = 0; k < NP; ++k) {
			const int ks = k >> 5, kl = k & 31;
			// the owner publishes the (unscaled) pivot row
#pragma unroll
			for (int s = 0; s < RPL; ++s) {
				if (s == ks && lane == kl) {
#pragma unroll
					for (int j = 0; j < NP; j += 4)
						*reinterpret_cast<float4 *>(bc + j) = make_float4(
							S[s][j], S[s][j + 1], S[s][j + 2], S[s][j + 3]);
				}
			}
			__syncwarp();
			const float piv = bc[0];
			bad |= !(piv > 0.0f);
			const float dinv = __frcp_rn(piv);
			float g[RPL];
			bool mine[RPL];
#pragma unroll
			for (int s = 0; s < RPL; ++s) {
				mine[s] = (s == ks && lane == kl);
				g[s] = mine[s] ? 0.0f : S[s][0] * dinv;
			}
			// S(:, j-1) <- S(:, j) - g * pivot_row(j), pivot row read 4 at a time
#pragma unroll
			for (int j = 0; j < NP; j += 4) {
				const float4 v = *reinterpret_cast<const float4 *>(bc + j);
				const float pr[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
				for (int e = 0; e < 4; ++e) {
					if (j + e >= 1) {
#pragma unroll
						for (int s = 0; s < RPL; ++s)
							S[s][j + e - 1] = fmaf(-g[s], pr[e], S[s][j + e]);
					}
				}
			}
#pragma unroll
			for (int s = 0; s < RPL; ++s) {
				S[s][NP - 1] = mine[s] ? 1.0f : -g[s];
				if (mine[s])
					sigma[s] = dinv;
			}
			__syncwarp();
		}
#pragma unroll
		for (int s = 0; s < RPL; ++s)
#pragma unroll
			for (int j = 0; j < NP; ++j)
				S[s][j] *= sigma[s];

		// ---- per-coordinate state ---------------------------------------------------
		float q[RPL], lo[RPL], hi[RPL], z[RPL], u[RPL], x[RPL];
#pragma unroll
		for (int s = 0; s < RPL; ++s) {
			const int r = lane + 32 * s;
			const bool in = r < n;
			const unsigned long long o = prob * (unsigned long long)n + r;
			q[s] = in ? __ldg(p.q + o) : 0.0f;
			// lb/ub arrays or the scalar box; z0/u0 or zero (all warp-uniform choices)
			lo[s] = in ? (p.lb ? __ldg(p.lb + o) : p.lb_s) : 0.0f;
			hi[s] = in ? (p.ub ? __ldg(p.ub + o) : p.ub_s) : 0.0f;
			z[s] = (in && p.z0) ? __ldg(p.z0 + o) : 0.0f;
			u[s] = (in && p.u0) ? __ldg(p.u0 + o) : 0.0f;
			x[s] = 0.0f;
		}

		// ---- ADMM iterations (qp_solvers.c:284-310) ---------------------------------
		int status = QP_STATUS_MAX_ITER;
		unsigned it = 0;
		if (!bad) {
			while (it < p.max_iter) {
				// admm_update_x (:146-159)
#pragma unroll
				for (int s = 0; s < RPL; ++s) {
					const int r = lane + 32 * s;
					if (r < NP)
						bc[r] = fmaf(rho, z[s] - u[s], -q[s]);
				}
				__syncwarp();
				float acc[RPL][4];
#pragma unroll
				for (int s = 0; s < RPL; ++s)
					acc[s][0] = acc[s][1] = acc[s][2] = acc[s][3] = 0.0f;
#pragma unroll
				for (int j = 0; j < NP; j += 4) {
					const float4 y = *reinterpret_cast<const float4 *>(bc + j);
#pragma unroll
					for (int s = 0; s < RPL; ++s) {
						acc[s][0] = fmaf(S[s][j], y.x, acc[s][0]);
						acc[s][1] = fmaf(S[s][j + 1], y.y, acc[s][1]);
						acc[s][2] = fmaf(S[s][j + 2], y.z, acc[s][2]);
						acc[s][3] = fmaf(S[s][j + 3], y.w, acc[s][3]);
					}
				}
				__syncwarp();
				float rr = 0.f, ss = 0.f, nx = 0.f, nz = 0.f, nu = 0.f;
#pragma unroll
				for (int s = 0; s < RPL; ++s) {
					const float xs = (acc[s][0] + acc[s][1]) + (acc[s][2] + acc[s][3]);
					x[s] = xs;
					const float xh =
						relax ? fmaf(alpha, xs, (1.0f - alpha) * z[s]) : xs;
					const float zn = fminf(fmaxf(xh + u[s], lo[s]), hi[s]);
					const float dz = zn - z[s];
					ss = fmaf(dz, dz, ss);
					u[s] += xh - zn;
					const float dr = xs - zn;
					rr = fmaf(dr, dr, rr);
					nx = fmaf(xs, xs, nx);
					nz = fmaf(zn, zn, nz);
					nu = fmaf(u[s], u[s], nu);
					z[s] = zn;
				}
				rr = warp_sum(rr);
				ss = warp_sum(ss);
				nx = warp_sum(nx);
				nz = warp_sum(nz);
				nu = warp_sum(nu);
				++it;
				const float r_norm = __fsqrt_rn(rr);
				const float s_norm = fabsf(rho) * __fsqrt_rn(ss);
				const float eps_pri = fmaf(p.reltol, __fsqrt_rn(fmaxf(nx, nz)), p.eps_abs);
				const float eps_dual = fmaf(p.reltol * rho, __fsqrt_rn(nu), p.eps_abs);
				if (r_norm < eps_pri && s_norm < eps_dual) {
					status = QP_STATUS_CONVERGED;
					break;
				}
			}
		}

		// ---- results ---------------------------------------------------------------------
		float nf = 0.0f;
#pragma unroll
		for (int s = 0; s < RPL; ++s)
			nf = fmaf(x[s], 0.0f, nf);
		nf = warp_sum(nf);
		int iters = int(it);
		if (bad) {
			status = QP_STATUS_SINGULAR;
			iters = 0;
		} else if (!(nf == 0.0f)) {
			status = QP_STATUS_NONFINITE;
		}
#pragma unroll
		for (int s = 0; s < RPL; ++s) {
			const int r = lane + 32 * s;
			if (r < n) {
				const unsigned long long o = prob * (unsigned long long)n + r;
				if (p.x)
					p.x[o] = x[s];
				if (p.z)
					p.z[o] = z[s];
				if (p.lam)
					p.lam[o] = rho * u[s];
			}
		}
		if (lane == 0) {
			if (p.status)
				p.status[prob] = status;
			if (p.iters)
				p.iters[prob] = iters;
		}
	}
}

} // namespace qpb
