// qp_cpq.cuh -- one CTA per QP, 64 < n <= 128 (any n <= 128 when forced): the reference's
// N_DIM is a compile-time constant with no upper bound (config.h:5), so dimensions beyond
// the lane-group kernels need a home too.
//
// Same algorithm as the other box kernels -- admm() of /root/reference/qp_solvers/qp_solvers.c:255-319:
//   R = (P + rho I)^-1 once (matrix_invert, matrix_ops.c:551-630), then per iteration
//   x = R (rho (z - u) - q), z = clip(x_hat + u), u += x_hat - z, stop test (:308).
//
// Layout: 256 threads as a 16 x 16 grid (ty = tid / 16, tx = tid % 16).  Thread (ty, tx)
// keeps the TT x TT interleaved tile  A[a][b] = M[ty + 16 a][tx + 16 b]  in REGISTERS
// (TT = ceil(n / 16) <= 8, 64 registers at n = 128); rows / columns past n are the identity.
//
//   * Inverse: in-place Gauss-Jordan without pivoting (P + rho I is positive definite, pivot
//     <= 0 -> QP_STATUS_SINGULAR).  Step k: the 16 owners of column k publish it and put e_k
//     in its place (the classic "identity column takes the pivot column's seat"), the 16
//     owners of row k scale it by 1/pivot and publish it, one __syncthreads, then every
//     thread does its TT x TT rank-1 update from 2 TT shared-memory reads.  The step loop is
//     unrolled over k / 16 so that the pivot row / column are compile-time register indices.
//     Publishing is double buffered: one barrier per step.
//   * Iterations: y goes through shared memory, each thread multiplies its tile by its TT
//     entries of y, the 16 threads of a row group (a half warp) add up with 4 shuffle
//     stages, and lane tx < TT of the half warp owns vector entry i = ty + 16 tx for the
//     element-wise part.  The five norms are reduced warp -> CTA through shared memory.
//     Two barriers per iteration.
#pragma once

#include "qp_common.cuh"

namespace qpb {

constexpr int CPQ_THREADS = 256;
constexpr int CPQ_MAX = 128; // 16 * 8

template <int TT>
__global__ void __launch_bounds__(CPQ_THREADS, 2) qp_cpq_kernel(const KernelArgs p)
{
	__shared__ __align__(16) float colb[2][CPQ_MAX];
	__shared__ __align__(16) float rowb[2][CPQ_MAX];
	__shared__ __align__(16) float yv[CPQ_MAX];
	__shared__ float red[2][CPQ_THREADS / 32][5];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int ty = tid >> 4, tx = tid & 15;
	const int n = p.n;
	const float alpha = p.alpha;
	const bool relax = (alpha != 1.0f);
	// vector entry of this thread (lanes tx < TT of each half warp)
	const int vi = ty + 16 * tx;
	const bool own = tx < TT && vi < n;

	for (unsigned long long prob = blockIdx.x; prob < p.count; prob += gridDim.x) {
		const float rho = p.rho_v ? __ldg(p.rho_v + prob) : p.rho;
		const float rho2 = rho * rho, relrho = p.reltol * rho;

		// ---- load the tile of P + rho I --------------------------------------------------
		float A[TT][TT];
		const float *Pp = p.P + p_index(p, prob) * (unsigned long long)n * n;
#pragma unroll
		for (int a = 0; a < TT; ++a) {
			const int i = ty + 16 * a;
#pragma unroll
			for (int b = 0; b < TT; ++b) {
				const int j = tx + 16 * b;
				float v = (i == j) ? 1.0f : 0.0f;
				if (i < n && j < n) {
					v = __ldg(Pp + i * n + j);
					if (i == j)
						v += rho;
				}
				A[a][b] = v;
			}
		}
		const unsigned long long vo = prob * (unsigned long long)n + vi;
		const float q = own ? __ldg(p.q + vo) : 0.0f;
		const float lo = own ? (p.lb ? __ldg(p.lb + vo) : p.lb_s) : 0.0f;
		const float hi = own ? (p.ub ? __ldg(p.ub + vo) : p.ub_s) : 0.0f;
		float z = (own && p.z0) ? __ldg(p.z0 + vo) : 0.0f;
		float u = (own && p.u0) ? __ldg(p.u0 + vo) : 0.0f;

		// ---- R = (P + rho I)^-1 ------------------------------------------------------------
		bool bad = !(rho > 0.0f && rho < 3.0e38f);
#pragma unroll
		for (int ka = 0; ka < TT; ++ka) {
			const int kend = min(16, n - 16 * ka); // rows past n are already the identity
			for (int kt = 0; kt < kend; ++kt) {
				const int buf = kt & 1;
				// the pivot, read before its column is replaced
				const float piv = __shfl_sync(0xffffffffu, A[ka][ka], (lane & 16) | kt);
				if (tx == kt) {
#pragma unroll
					for (int a = 0; a < TT; ++a) {
						const bool diag = (a == ka) && (ty == kt);
						colb[buf][ty + 16 * a] = diag ? 0.0f : A[a][ka];
						A[a][ka] = diag ? 1.0f : 0.0f;
					}
				}
				if (ty == kt) {
					bad |= !(piv > 0.0f);
					const float inv = 1.0f / piv;
#pragma unroll
					for (int b = 0; b < TT; ++b) {
						A[ka][b] *= inv;
						rowb[buf][tx + 16 * b] = A[ka][b];
					}
				}
				__syncthreads();
				float c[TT], r[TT];
#pragma unroll
				for (int a = 0; a < TT; ++a)
					c[a] = colb[buf][ty + 16 * a];
#pragma unroll
				for (int b = 0; b < TT; ++b)
					r[b] = rowb[buf][tx + 16 * b];
#pragma unroll
				for (int a = 0; a < TT; ++a)
#pragma unroll
					for (int b = 0; b < TT; ++b)
						A[a][b] = fmaf(-c[a], r[b], A[a][b]);
			}
		}
		bad = __syncthreads_or(bad);

		// ---- ADMM iterations (qp_solvers.c:284-310) --------------------------------------
		int status = QP_STATUS_MAX_ITER;
		unsigned it = 0;
		float x = 0.0f;
		if (!bad) {
			while (it < p.max_iter) {
				// admm_update_x (:146-159): y = rho (z - u) - q ; x = R y
				if (own)
					yv[vi] = fmaf(rho, z - u, -q);
				else if (tx < TT)
					yv[vi] = 0.0f; // padding entries (vi < 16 TT <= CPQ_MAX)
				__syncthreads();
				float yr[TT], s[TT];
#pragma unroll
				for (int b = 0; b < TT; ++b)
					yr[b] = yv[tx + 16 * b];
#pragma unroll
				for (int a = 0; a < TT; ++a) {
					float acc = 0.0f;
#pragma unroll
					for (int b = 0; b < TT; ++b)
						acc = fmaf(A[a][b], yr[b], acc);
					s[a] = acc;
				}
#pragma unroll
				for (int o = 8; o > 0; o >>= 1)
#pragma unroll
					for (int a = 0; a < TT; ++a)
						s[a] += __shfl_xor_sync(0xffffffffu, s[a], o);
#pragma unroll
				for (int a = 0; a < TT; ++a)
					if (tx == a)
						x = s[a];

				float rr = 0.f, ss = 0.f, nx = 0.f, nz = 0.f, nu = 0.f;
				if (own) {
					// admm_update_x_hat (:161-174), _z (:176-190), _u (:192-203)
					const float xh = relax ? fmaf(alpha, x, (1.0f - alpha) * z) : x;
					const float zn = fminf(fmaxf(xh + u, lo), hi);
					const float dz = zn - z;
					u += xh - zn;
					const float dr = x - zn;
					ss = dz * dz, rr = dr * dr, nx = x * x, nz = zn * zn, nu = u * u;
					z = zn;
				}
				rr = warp_sum(rr), ss = warp_sum(ss), nx = warp_sum(nx);
				nz = warp_sum(nz), nu = warp_sum(nu);
				const int rb = it & 1;
				if (lane == 0) {
					red[rb][warp][0] = rr, red[rb][warp][1] = ss, red[rb][warp][2] = nx;
					red[rb][warp][3] = nz, red[rb][warp][4] = nu;
				}
				__syncthreads();
				rr = ss = nx = nz = nu = 0.0f;
#pragma unroll
				for (int w = 0; w < CPQ_THREADS / 32; ++w) {
					rr += red[rb][w][0], ss += red[rb][w][1], nx += red[rb][w][2];
					nz += red[rb][w][3], nu += red[rb][w][4];
				}
				++it;
				// admm_eps_pri (:233-243), admm_eps_dual (:245-253), stop test (:308)
				const float eps_pri = fmaf(p.reltol, sqrt_approx(fmaxf(nx, nz)), p.eps_abs);
				const float eps_dual = fmaf(relrho, sqrt_approx(nu), p.eps_abs);
				if ((rr < eps_pri * eps_pri) && (rho2 * ss < eps_dual * eps_dual)) {
					status = QP_STATUS_CONVERGED;
					break;
				}
			}
		}

		// ---- results ------------------------------------------------------------------------
		const bool nonfinite = __syncthreads_or(own && !finite_f(x));
		int iters = int(it);
		if (bad)
			status = QP_STATUS_SINGULAR, iters = 0;
		else if (nonfinite)
			status = QP_STATUS_NONFINITE;
		if (own) {
			if (p.x) p.x[vo] = x;
			if (p.z) p.z[vo] = z;
			if (p.lam) p.lam[vo] = rho * u;
		}
		if (tid == 0) {
			if (p.status) p.status[prob] = status;
			if (p.iters) p.iters[prob] = iters;
		}
	}
}

} // namespace qpb
