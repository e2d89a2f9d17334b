// qp_gpq.cuh -- "lane group per QP" kernel for any n <= 64 (n = 40 and 64 in BASELINE.json).
//
// What it replaces: the same reference path as qp_tpq.cuh -- admm()
// (/root/reference/qp_solvers/qp_solvers.c:255-319) with matrix_invert()
// (/root/reference/matrix/matrix_ops/matrix_ops.c:551-630) and matrix_mult()
// (:235-271) underneath.
//
// Mapping:
//   * a group of T lanes (T = 8, 16 or 32; 32/T problems per warp) owns one QP;
//     lane g of the group owns rows g, g+T, g+2T, ... of the matrix, held entirely
//     in registers (NP floats per row, NP = n rounded up to 8; padding rows and
//     columns form an identity block).
//   * Loading uses the symmetry of P: the lane that owns row r reads COLUMN r, so
//     each step of the load loop reads T consecutive floats per group.
//   * R = (P + rho I)^-1 is formed explicitly, like the reference, because the
//     per-iteration x = R y is then a register-resident mat-vec with y broadcast
//     from shared memory; triangular solves would serialise across lanes.
//   * The inverse is a BLOCK Gauss-Jordan elimination without pivoting (P + rho I
//     is symmetric positive definite, so every 4x4 pivot block is too): four pivot
//     rows are published to shared memory at once, every lane inverts the 4x4
//     pivot block redundantly in registers, and each row is updated with four
//     FMAs per element.  The elimination is written in "shifting" form -- each
//     block step drops the four pivot columns and appends the four new inverse
//     columns at the end -- so every register index is a compile-time constant
//     while the block loop stays a run-time loop.  The pivot rows are double
//     buffered: one __syncwarp per four pivots.
//   * per iteration: y to shared memory, NP FMAs per owned row, element-wise
//     updates on the owned coordinates, one butterfly all-reduce (inside the
//     group) of the five sums of squares, and a group-uniform stop test.
//   * x, z, lambda are written with lane = coordinate: coalesced per group.
#pragma once

#include "qp_common.cuh"

namespace qpb {

constexpr int GPQ_WARPS = 4; // warps per CTA

// lanes per QP for a padded dimension
template <int NP> struct GpqLanes {
	static constexpr int T = (NP <= 32) ? 8 : (NP <= 48 ? 16 : 32);
};

template <int NP, int T> struct GpqSmem {
	static constexpr int G = 32 / T; // QPs per warp
	// pivot rows are NP + 4 floats apart so that the four owners' 128-bit stores of
	// one column chunk fall into different bank groups; two buffers of four rows per
	// group, +4 floats so that the groups of a warp are staggered as well
	static constexpr int ROW = NP + 4;
	static constexpr int GROUP = 2 * 4 * ROW + 4;
	static constexpr int WARP = G * GROUP;
};

// In-place inverse of a symmetric positive definite 4x4 block (Gauss-Jordan, no
// pivoting).  Returns true when a pivot is not strictly positive.
__device__ __forceinline__ bool inv4(float (&B)[4][4])
{
	bool bad = false;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const float piv = B[k][k];
		bad |= !(piv > 0.0f);
		const float d = rcp_approx(piv);
#pragma unroll
		for (int j = 0; j < 4; ++j)
			if (j != k)
				B[k][j] *= d;
		B[k][k] = d;
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			if (i == k)
				continue;
			const float f = B[i][k];
#pragma unroll
			for (int j = 0; j < 4; ++j)
				if (j != k)
					B[i][j] = fmaf(-f, B[k][j], B[i][j]);
			B[i][k] = -f * d;
		}
	}
	return bad;
}

// CTAs per SM the register budget is capped for (rows in registers: NP * RPL floats)
template <int NP, int T> struct GpqOcc {
	static constexpr int ROWREGS = NP * ((NP + T - 1) / T);
	static constexpr int MINB = ROWREGS <= 64 ? 6 : (ROWREGS <= 96 ? 4 : 3);
};

template <int NP, int T>
__global__ void __launch_bounds__(GPQ_WARPS * 32, GpqOcc<NP, T>::MINB)
qp_gpq_kernel(const KernelArgs p)
{
	using SM = GpqSmem<NP, T>;
	constexpr int G = SM::G;
	constexpr int RPL = (NP + T - 1) / T; // rows per lane
	constexpr int NC = NP / 4;            // 4-column chunks per row
	constexpr int ROW = SM::ROW;          // floats between published pivot rows
	constexpr unsigned FULL = 0xffffffffu;
	__shared__ __align__(16) float sm[GPQ_WARPS * SM::WARP];

	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const int g = lane & (T - 1); // lane inside its group
	const int grp = lane / T;     // group inside the warp
	const int n = p.n;
	float *bc = sm + warp * SM::WARP + grp * SM::GROUP; // bc[(buf*4+m)*ROW + j]
	const float rho = p.rho, alpha = p.alpha;
	const bool relax = (alpha != 1.0f);
	const float rho2 = rho * rho, relrho = p.reltol * rho;

	const unsigned long long stride = (unsigned long long)gridDim.x * GPQ_WARPS * G;
	for (unsigned long long base = ((unsigned long long)blockIdx.x * GPQ_WARPS + warp) * G;
	     base < p.count; base += stride) {
		const unsigned long long prob = base + grp;
		const bool valid = prob < p.count;

		// ---- load: the owner of row r takes column r of the symmetric P ------------------
		float S[RPL][NP];
		const float *Pp = p.P + (valid ? prob : 0ull) * (unsigned long long)n * n;
#pragma unroll
		for (int s = 0; s < RPL; ++s) {
			const int r = g + T * s;
#pragma unroll
			for (int j = 0; j < NP; ++j) {
				float v = 0.0f;
				if (valid && r < n && j < n)
					v = __ldg(Pp + (unsigned long long)j * n + r);
				if (r == j)
					v = (valid && r < n) ? v + rho : 1.0f; // padding: identity block
				S[s][j] = v;
			}
		}

		// ---- R = (P + rho I)^-1: block Gauss-Jordan, four pivots per step -----------------
		bool bad = false;
		for (int K = 0; K < NC; ++K) {
			const int r0 = 4 * K;
			const int ks = r0 / T, kl = r0 & (T - 1); // slot and first lane of the pivot rows
			const int m0 = g - kl;
			const bool owner = (m0 >= 0) && (m0 < 4);
			float *pb = bc + (K & 1) * 4 * ROW;
			// the four owners publish their (pre-update) rows
#pragma unroll
			for (int s = 0; s < RPL; ++s) {
				if (s == ks && owner) {
#pragma unroll
					for (int c = 0; c < NC; ++c)
						*reinterpret_cast<float4 *>(pb + m0 * ROW + 4 * c) =
							make_float4(S[s][4 * c], S[s][4 * c + 1],
								    S[s][4 * c + 2], S[s][4 * c + 3]);
				}
			}
			__syncwarp();
			// pivot block and its inverse (every lane, redundantly)
			float B[4][4];
#pragma unroll
			for (int m = 0; m < 4; ++m) {
				const float4 v = *reinterpret_cast<const float4 *>(pb + m * ROW);
				B[m][0] = v.x, B[m][1] = v.y, B[m][2] = v.z, B[m][3] = v.w;
			}
			bad |= inv4(B);
			// an owner's new row is (row m0 of B^-1) x (pivot rows): expressed through the
			// common update by replacing its row with -e_m0
#pragma unroll
			for (int s = 0; s < RPL; ++s) {
				if (s == ks && owner) {
#pragma unroll
					for (int j = 0; j < NP; ++j)
						S[s][j] = 0.0f;
#pragma unroll
					for (int m = 0; m < 4; ++m)
						S[s][m] = (m == m0) ? -1.0f : 0.0f;
				}
			}
			float H[RPL][4];
#pragma unroll
			for (int s = 0; s < RPL; ++s) {
#pragma unroll
				for (int m = 0; m < 4; ++m) {
					float h = S[s][0] * B[0][m];
					h = fmaf(S[s][1], B[1][m], h);
					h = fmaf(S[s][2], B[2][m], h);
					h = fmaf(S[s][3], B[3][m], h);
					H[s][m] = -h;
				}
			}
			// S(:, j-4) <- S(:, j) + H * pivot_rows(:, j)
#pragma unroll
			for (int c = 1; c < NC; ++c) {
				float pr[4][4];
#pragma unroll
				for (int m = 0; m < 4; ++m) {
					const float4 v =
						*reinterpret_cast<const float4 *>(pb + m * ROW + 4 * c);
					pr[m][0] = v.x, pr[m][1] = v.y, pr[m][2] = v.z, pr[m][3] = v.w;
				}
#pragma unroll
				for (int s = 0; s < RPL; ++s) {
#pragma unroll
					for (int e = 0; e < 4; ++e) {
						float a = S[s][4 * c + e];
						a = fmaf(H[s][0], pr[0][e], a);
						a = fmaf(H[s][1], pr[1][e], a);
						a = fmaf(H[s][2], pr[2][e], a);
						a = fmaf(H[s][3], pr[3][e], a);
						S[s][4 * (c - 1) + e] = a;
					}
				}
			}
#pragma unroll
			for (int s = 0; s < RPL; ++s)
#pragma unroll
				for (int m = 0; m < 4; ++m)
					S[s][NP - 4 + m] = H[s][m];
		}
		__syncwarp();

		// ---- per-coordinate state -----------------------------------------------------------
		float q[RPL], lo[RPL], hi[RPL], z[RPL], u[RPL], x[RPL];
#pragma unroll
		for (int s = 0; s < RPL; ++s) {
			const int r = g + T * s;
			const bool in = valid && r < n;
			const unsigned long long o = prob * (unsigned long long)n + r;
			q[s] = in ? __ldg(p.q + o) : 0.0f;
			// lb/ub arrays or the scalar box; z0/u0 or zero (group-uniform choices)
			lo[s] = in ? (p.lb ? __ldg(p.lb + o) : p.lb_s) : 0.0f;
			hi[s] = in ? (p.ub ? __ldg(p.ub + o) : p.ub_s) : 0.0f;
			z[s] = (in && p.z0) ? __ldg(p.z0 + o) : 0.0f;
			u[s] = (in && p.u0) ? __ldg(p.u0 + o) : 0.0f;
			x[s] = 0.0f;
		}

		// ---- ADMM iterations (qp_solvers.c:284-310) -----------------------------------------
		int status = QP_STATUS_MAX_ITER;
		unsigned it = 0;
		bool active = valid && !bad && p.max_iter > 0; // uniform inside a group
		float *ybuf = bc;
		for (;;) {
			const unsigned am = __ballot_sync(FULL, active);
			if (am == 0u)
				break;
			if (active) {
				// admm_update_x (:146-159)
#pragma unroll
				for (int s = 0; s < RPL; ++s) {
					const int r = g + T * s;
					if (r < NP)
						ybuf[r] = fmaf(rho, z[s] - u[s], -q[s]);
				}
				__syncwarp(am);
				float acc[RPL][4];
#pragma unroll
				for (int s = 0; s < RPL; ++s)
					acc[s][0] = acc[s][1] = acc[s][2] = acc[s][3] = 0.0f;
#pragma unroll
				for (int j = 0; j < NP; j += 4) {
					const float4 y = *reinterpret_cast<const float4 *>(ybuf + j);
#pragma unroll
					for (int s = 0; s < RPL; ++s) {
						acc[s][0] = fmaf(S[s][j], y.x, acc[s][0]);
						acc[s][1] = fmaf(S[s][j + 1], y.y, acc[s][1]);
						acc[s][2] = fmaf(S[s][j + 2], y.z, acc[s][2]);
						acc[s][3] = fmaf(S[s][j + 3], y.w, acc[s][3]);
					}
				}
				__syncwarp(am);
				float rr = 0.f, ss = 0.f, nx = 0.f, nz = 0.f, nu = 0.f;
#pragma unroll
				for (int s = 0; s < RPL; ++s) {
					const float xs = (acc[s][0] + acc[s][1]) + (acc[s][2] + acc[s][3]);
					x[s] = xs;
					// admm_update_x_hat (:161-174), _z (:176-190), _u (:192-203)
					const float xh = relax ? fmaf(alpha, xs, (1.0f - alpha) * z[s]) : xs;
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
#pragma unroll
				for (int o = T / 2; o > 0; o >>= 1) {
					rr += __shfl_xor_sync(am, rr, o);
					ss += __shfl_xor_sync(am, ss, o);
					nx += __shfl_xor_sync(am, nx, o);
					nz += __shfl_xor_sync(am, nz, o);
					nu += __shfl_xor_sync(am, nu, o);
				}
				++it;
				// stop test (:233-253, :308), squared form as in qp_tpq.cuh
				const float eps_pri = fmaf(p.reltol, sqrt_approx(fmaxf(nx, nz)), p.eps_abs);
				const float eps_dual = fmaf(relrho, sqrt_approx(nu), p.eps_abs);
				if ((rr < eps_pri * eps_pri) && (rho2 * ss < eps_dual * eps_dual)) {
					status = QP_STATUS_CONVERGED;
					active = false;
				} else if (it >= p.max_iter) {
					active = false;
				}
			}
		}
		__syncwarp();

		// ---- results ---------------------------------------------------------------------------
		float nf = 0.0f;
#pragma unroll
		for (int s = 0; s < RPL; ++s)
			nf = fmaf(x[s], 0.0f, nf);
#pragma unroll
		for (int o = T / 2; o > 0; o >>= 1)
			nf += __shfl_xor_sync(FULL, nf, o);
		int iters = int(it);
		if (bad) {
			status = QP_STATUS_SINGULAR;
			iters = 0;
		} else if (!(nf == 0.0f)) {
			status = QP_STATUS_NONFINITE;
		}
#pragma unroll
		for (int s = 0; s < RPL; ++s) {
			const int r = g + T * s;
			if (valid && r < n) {
				const unsigned long long o = prob * (unsigned long long)n + r;
				if (p.x)
					p.x[o] = x[s];
				if (p.z)
					p.z[o] = z[s];
				if (p.lam)
					p.lam[o] = rho * u[s];
			}
		}
		if (valid && g == 0) {
			if (p.status)
				p.status[prob] = status;
			if (p.iters)
				p.iters[prob] = iters;
		}
	}
}

} // namespace qpb
